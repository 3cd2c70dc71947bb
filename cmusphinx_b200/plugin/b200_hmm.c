/* b200_hmm.c -- the reference-side binding of the HMM boundary:
 * the three evaluation loops of the forward tree search -- eval_root_chan,
 * eval_nonroot_chan, eval_word_chan (pocketsphinx/src/libpocketsphinx/
 * ngram_search_fwdtree.c:598-691), driven by evaluate_channels (:694-707) -- call
 * hmm_vit_eval once per active channel.  With B200_HMM_PLUGIN=1 this file, preloaded next
 * to an UNMODIFIED libpocketsphinx.so, evaluates all of a frame's channels in ONE batched
 * call of the C ABI instead:
 *
 *   ngram_fwdtree_search (interposed)   remembers the search and the frame, then runs
 *                                       the reference's own function;
 *   hmm_vit_eval (interposed)           the first call of a frame -- made by eval_root_chan,
 *                                       i.e. after the reference has set the frame's senone
 *                                       scores and renormalised -- walks the same channel
 *                                       lists the three loops are about to walk, packs the
 *                                       hmm_t's (b200_hmm_pack), runs b200_hmm_eval_host for one
 *                                       frame and keeps the result; this and every later call of
 *                                       the frame copies its HMM's new state back
 *                                       (b200_hmm_unpack_one) and returns its best score.
 *
 * hmm_vit_eval calls outside a forward-tree frame (fwdflat, FSG, alignment, phone loop) and
 * HMMs that were not in the batch fall through to the reference's own function.
 * No Viterbi arithmetic happens in this file. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <sphinxbase/ckd_alloc.h>
#include <sphinxbase/err.h>

#include "ngram_search.h"
#include "ngram_search_fwdtree.h"
#include "hmm.h"
#include "tmat.h"
#include "bin_mdef.h"

#include "../../include/b200sphinx.h"

/* the ABI's mirror of hmm_t must be the real thing */
_Static_assert(sizeof(b200_ps_hmm_t) == sizeof(hmm_t), "hmm_t size");
_Static_assert(offsetof(b200_ps_hmm_t, score) == offsetof(hmm_t, score), "hmm_t.score");
_Static_assert(offsetof(b200_ps_hmm_t, history) == offsetof(hmm_t, history), "hmm_t.history");
_Static_assert(offsetof(b200_ps_hmm_t, out_score) == offsetof(hmm_t, out_score), "hmm_t.out_score");
_Static_assert(offsetof(b200_ps_hmm_t, senid) == offsetof(hmm_t, senid), "hmm_t.senid");
_Static_assert(offsetof(b200_ps_hmm_t, bestscore) == offsetof(hmm_t, bestscore), "hmm_t.bestscore");
_Static_assert(offsetof(b200_ps_hmm_t, tmatid) == offsetof(hmm_t, tmatid), "hmm_t.tmatid");
_Static_assert(offsetof(b200_ps_hmm_t, mpx) == offsetof(hmm_t, mpx), "hmm_t.mpx");
_Static_assert(offsetof(b200_ps_hmm_t, n_emit_state) == offsetof(hmm_t, n_emit_state), "hmm_t.n_emit_state");

static struct {
    ngram_search_t *ngs;       /* the search whose frame is running, or NULL */
    int frame;
    int batched;               /* this frame's batch has been evaluated */
    b200_hmmctx_t *gpu;        /* built on first use from the search's hmm context */
    hmm_context_t *gpu_for;
    void **ptr; int cap, n;    /* the frame's HMMs, in evaluation order */
    int next;                  /* evaluation order is the call order: the next expected slot */
    b200_hmm_soa_t soa;
    long n_frames, n_hmms, n_fallthrough;
} G;

static int hmm_enabled(void)
{
    const char *e = getenv("B200_HMM_PLUGIN");
    return e && e[0] == '1';
}

static int reserve(int n)
{
    int ne = HMM_MAX_NSTATE;
    if (n <= G.cap) return 0;
    n += n / 2 + 1024;
    G.ptr = ckd_realloc(G.ptr, (size_t)n * sizeof(void *));
    G.soa.score = ckd_realloc(G.soa.score, (size_t)n * ne * sizeof(int32));
    G.soa.history = ckd_realloc(G.soa.history, (size_t)n * ne * sizeof(int32));
    G.soa.senid = ckd_realloc(G.soa.senid, (size_t)n * ne * sizeof(uint16));
    G.soa.out_score = ckd_realloc(G.soa.out_score, (size_t)n * sizeof(int32));
    G.soa.out_history = ckd_realloc(G.soa.out_history, (size_t)n * sizeof(int32));
    G.soa.bestscore = ckd_realloc(G.soa.bestscore, (size_t)n * sizeof(int32));
    G.soa.tmatid = ckd_realloc(G.soa.tmatid, (size_t)n * sizeof(int16));
    G.soa.mpx = ckd_realloc(G.soa.mpx, (size_t)n);
    G.cap = n;
    return 0;
}

static void add(hmm_t *h)
{
    reserve(G.n + 1);
    G.ptr[G.n++] = h;
}

/* The channels evaluate_channels is about to evaluate, in its order
 * (ngram_search_fwdtree.c:598-691). */
static void collect(ngram_search_t *ngs, int frame_idx)
{
    root_chan_t *rhmm;
    chan_t *hmm, **acl;
    int32 i, w, *awl;
    G.n = 0;
    for (i = ngs->n_root_chan, rhmm = ngs->root_chan; i > 0; --i, rhmm++)
        if (hmm_frame(&rhmm->hmm) == frame_idx) add(&rhmm->hmm);
    i = ngs->n_active_chan[frame_idx & 0x1];
    acl = ngs->active_chan_list[frame_idx & 0x1];
    for (; i > 0; --i) { hmm = *(acl++); add(&hmm->hmm); }
    awl = ngs->active_word_list[frame_idx & 0x1];
    for (i = ngs->n_active_word[frame_idx & 0x1]; i > 0; --i) {
        w = *(awl++);
        for (hmm = ngs->word_chan[w]; hmm; hmm = hmm->next) add(&hmm->hmm);
    }
    for (i = 0; i < ngs->n_1ph_words; i++) {
        w = ngs->single_phone_wid[i];
        rhmm = (root_chan_t *)ngs->word_chan[w];
        if (hmm_frame(&rhmm->hmm) < frame_idx) continue;
        add(&rhmm->hmm);
    }
}

static int build_gpu_context(hmm_context_t *ctx, ngram_search_t *ngs)
{
    bin_mdef_t *mdef = ps_search_acmod(ngs)->mdef;
    tmat_t *tmat = ps_search_acmod(ngs)->tmat;
    int ne = ctx->n_emit_state, nt = tmat->n_tmat, ns = bin_mdef_n_sseq(mdef), n_sen = bin_mdef_n_sen(mdef);
    uint8 *tp = ckd_calloc((size_t)nt * ne * (ne + 1), 1);
    uint16 *sseq = ckd_calloc((size_t)ns * ne, sizeof(uint16));
    int t, i, j;
    for (t = 0; t < nt; ++t)
        for (i = 0; i < ne; ++i)
            for (j = 0; j <= ne; ++j) tp[((size_t)t * ne + i) * (ne + 1) + j] = ctx->tp[t][i][j];
    for (i = 0; i < ns; ++i)
        for (j = 0; j < ne; ++j) sseq[(size_t)i * ne + j] = ctx->sseq[i][j];
    if (G.gpu) b200_hmm_ctx_free(G.gpu);
    G.gpu = b200_hmm_ctx_create(ne, tp, nt, sseq, ns, n_sen, getenv("B200_DEVICE") ? atoi(getenv("B200_DEVICE")) : 0);
    ckd_free(tp); ckd_free(sseq);
    if (!G.gpu) { E_ERROR("b200 hmm: %s\n", b200_last_error()); return -1; }
    G.gpu_for = ctx;
    E_INFO("b200 hmm: evaluate_channels runs on the GPU (%d-state HMMs, %d tmats, %d senone sequences)\n", ne, nt, ns);
    return 0;
}

int
ngram_fwdtree_search(ngram_search_t *ngs, int frame_idx)
{
    static int (*next)(ngram_search_t *, int);
    int rc;
    if (!next) next = dlsym(RTLD_NEXT, "ngram_fwdtree_search");
    if (!hmm_enabled()) return next(ngs, frame_idx);
    G.ngs = ngs; G.frame = frame_idx; G.batched = 0;
    rc = next(ngs, frame_idx);
    G.ngs = NULL;
    return rc;
}

int32
hmm_vit_eval(hmm_t *hmm)
{
    static int32 (*next)(hmm_t *);
    if (!next) next = dlsym(RTLD_NEXT, "hmm_vit_eval");
    if (G.ngs == NULL || hmm->ctx != G.ngs->hmmctx) return next(hmm);
    if (!G.batched) {
        int32 best;
        G.batched = 1; G.next = 0;
        collect(G.ngs, G.frame);
        if (G.gpu_for != hmm->ctx && build_gpu_context(hmm->ctx, G.ngs) < 0) { G.n = 0; return next(hmm); }
        if (b200_hmm_pack((const void *const *)G.ptr, G.n, hmm->ctx->n_emit_state, &G.soa) ||
            b200_hmm_eval_host(G.gpu, &G.soa, hmm->ctx->senscore, 1, &best)) {
            E_ERROR("b200 hmm: %s\n", b200_last_error());
            G.n = 0;
            return next(hmm);
        }
        G.n_frames++; G.n_hmms += G.n;
    }
    /* the loops evaluate in the order collect() walked; anything else is looked up */
    {
        int i = G.next;
        if (i >= G.n || G.ptr[i] != hmm)
            for (i = 0; i < G.n && G.ptr[i] != hmm; ++i) ;
        if (i >= G.n) { G.n_fallthrough++; return next(hmm); }
        G.next = i + 1;
        b200_hmm_unpack_one(&G.soa, hmm->ctx->n_emit_state, i, hmm);
        return hmm->bestscore;
    }
}

__attribute__((destructor)) static void b200_hmm_report(void)
{
    if (G.n_frames)
        fprintf(stderr, "b200 hmm: %ld frames, %ld HMM evaluations on the GPU, %ld calls fell through to the reference\n",
                G.n_frames, G.n_hmms, G.n_fallthrough);
}
