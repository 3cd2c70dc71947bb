/* ref_pls_trace.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * LD_PRELOAD copy of the reference's OWN phone-loop look-ahead search,
 * pocketsphinx/src/libpocketsphinx/phone_loop_search.c, compiled from where it lies (the #include
 * below; no source is copied), with one macro hook on the acmod_score call of
 * phone_loop_search_step (:270): it sees the phone HMMs exactly as the previous step left them
 * (nothing else touches them inside an utterance), i.e. the record of frame f+1 is the reference's
 * result for frame f.  With B200_PLS_TRACE=<file> every step appends int32 records:
 *
 *   'M' (once)  n_tmat n_emit | tp[n_tmat][n_emit][n_emit+1]
 *   'P'         frame best_score beam pbeam pip n_phones n_emit |
 *               per phone: score[ne] history[ne] out_score out_history bestscore frame senid[ne] tmatid |
 *               per phone: the frame's senone score of each state [ne]
 * Because the preloaded copy defines phone_loop_search_init, the unmodified libpocketsphinx.so
 * builds its look-ahead search from THIS copy; the code that runs is the reference's.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <sphinxbase/err.h>
#include <sphinxbase/ckd_alloc.h>

#include "pocketsphinx_internal.h"
#include "phone_loop_search.h"
#include "acmod.h"
#include "tmat.h"

static int16 const *b200_pls_trace_score(phone_loop_search_t *pls, acmod_t *acmod, int *frame_idx);

#define acmod_score(a, f) b200_pls_trace_score(pls, a, f)
#include "phone_loop_search.c"
#undef acmod_score

static FILE *pls_fp;
static int pls_state;   /* 0 unknown, 1 tracing, -1 off */

static void pput(int32 v) { fwrite(&v, 4, 1, pls_fp); }

static int16 const *b200_pls_trace_score(phone_loop_search_t *pls, acmod_t *acmod, int *frame_idx)
{
    int16 const *senscr = acmod_score(acmod, frame_idx);
    int ne = hmm_n_emit_state(&pls->phones[0].hmm), i, s;
    if (pls_state == 0) {
        const char *f = getenv("B200_PLS_TRACE");
        pls_state = (f && (pls_fp = fopen(f, "ab")) != NULL) ? 1 : -1;
        if (pls_state == 1) {
            tmat_t *tm = acmod->tmat;
            int t, a, b;
            pput('M'); pput(tm->n_tmat); pput(ne);
            for (t = 0; t < tm->n_tmat; ++t)
                for (a = 0; a < ne; ++a)
                    for (b = 0; b <= ne; ++b) pput(tm->tp[t][a][b]);
        }
    }
    if (pls_state != 1 || senscr == NULL) return senscr;
    pput('P'); pput(*frame_idx); pput(pls->best_score); pput(pls->beam); pput(pls->pbeam); pput(pls->pip);
    pput(pls->n_phones); pput(ne);
    for (i = 0; i < pls->n_phones; ++i) {
        hmm_t *h = &pls->phones[i].hmm;
        for (s = 0; s < ne; ++s) pput(h->score[s]);
        for (s = 0; s < ne; ++s) pput(h->history[s]);
        pput(h->out_score); pput(h->out_history); pput(h->bestscore); pput(h->frame);
        for (s = 0; s < ne; ++s) pput(h->senid[s]);
        pput(h->tmatid);
    }
    for (i = 0; i < pls->n_phones; ++i)
        for (s = 0; s < ne; ++s) pput(senscr[pls->phones[i].hmm.senid[s]]);
    fflush(pls_fp);
    return senscr;
}
