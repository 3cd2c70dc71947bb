/* ref_fwdtree_trace.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * An LD_PRELOAD library that is the reference's OWN forward-tree search,
 * pocketsphinx/src/libpocketsphinx/ngram_search_fwdtree.c, compiled from where it lies
 * (the #include at the bottom; no source is copied), with two macro hooks that let us see
 * the state of the lexical tree immediately before prune_root_chan (:714-790) and
 * immediately after prune_nonroot_chan (:792-869) of every frame of a real decode:
 *
 *   ps_search_lookahead(ngs)       used at :730 (prune_root_chan, before anything is
 *                                  modified), :806 (prune_nonroot_chan) and :1257
 *                                  (word_transition);
 *   ngram_search_exit_score(...)   first use after prune_nonroot_chan is :905, the first
 *                                  statement of last_phone_transition's loop that touches a
 *                                  candidate -- the candidate list is still as
 *                                  prune_*_chan left it.
 *
 * Because the preloaded copy defines ngram_fwdtree_init/start/search/finish/..., the
 * unmodified libpocketsphinx.so calls THIS copy; the code that runs is the reference's.
 * With B200_FWDTREE_TRACE=<file> the frames selected by B200_FWDTREE_TRACE_EVERY (default 1)
 * are appended to <file> as int32 records:
 *
 *   'T' topology   n_root n_chan n_edge n_pw n_ci | child_off[n_chan+1] child[n_edge]
 *                  ciphone[n_chan] pw_off[n_chan+1] pw_wid[n_pw] pw_lastphone[n_pw]
 *   'B' before     frame best_score dynamic_beam pbeam lpbeam pip nwpen has_pls n_act |
 *                  pls_pen[n_ci] acl[n_act] state[n_chan][10]
 *   'A' after      frame cand_valid n_nacl n_cand | nacl[n_nacl] cand[n_cand][3] state[n_chan][10]
 *
 * state row = score[0..2] history[0..2] out_score out_history bestscore frame.
 * B200_FWDTREE_TRACE_SHUFFLE=<seed> permutes the frame's active list before prune_root_chan sees it
 * (a real decode always lists a parent ahead of its children; the reference's two functions accept
 * any order and their result depends on it -- this pins that dependence to the reference itself).
 * Channel ids: roots 0..n_root-1 (their index in ngs->root_chan), then the non-root
 * channels in depth-first order of the next/alt links.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>

#include <sphinxbase/ckd_alloc.h>
#include <sphinxbase/listelem_alloc.h>
#include <sphinxbase/err.h>

#include "pocketsphinx_internal.h"
#include "ngram_search.h"
#include "ngram_search_fwdtree.h"
#include "phone_loop_search.h"

static void b200_trace_lookahead(ngram_search_t *ngs, int frame_idx);
static void b200_trace_exit(ngram_search_t *ngs);

#undef ps_search_lookahead
#define ps_search_lookahead(s) (b200_trace_lookahead((ngram_search_t *)(s), frame_idx), ps_search_base(s)->pls)
#define ngram_search_exit_score(a, b, c) (b200_trace_exit(a), (ngram_search_exit_score)(a, b, c))

#include "ngram_search_fwdtree.c"

#undef ngram_search_exit_score

/* ------------------------------------------------------------------ the tracer */
typedef struct { void *p; int32 id; } ptrid_t;

static struct {
    FILE *fp;
    int every;
    int phase;                  /* 0 idle, 1 in prune_root_chan, 2 in prune_nonroot_chan or later, 3 after-snapshot taken */
    int on;                     /* this frame is being recorded */
    unsigned shuffle;           /* LCG state, 0 = keep the list as it is */
    int frame;
    ngram_search_t *tree_of;    /* topology was written for this search ... */
    int32 tree_root, tree_nonroot;
    ptrid_t *map; int32 n_map;
    chan_t **chan; int32 n_chan, n_root;   /* id -> channel (roots cast) */
} TR;

static int cmp_ptr(const void *a, const void *b)
{
    const ptrid_t *x = a, *y = b;
    return x->p < y->p ? -1 : (x->p > y->p ? 1 : 0);
}

static int32 id_of(void *p)
{
    int32 lo = 0, hi = TR.n_map - 1;
    while (lo <= hi) {
        int32 mid = (lo + hi) / 2;
        if (TR.map[mid].p == p) return TR.map[mid].id;
        if (TR.map[mid].p < p) lo = mid + 1; else hi = mid - 1;
    }
    E_FATAL("fwdtree trace: unknown channel %p\n", p);
    return -1;
}

static void put(const void *p, size_t n) { fwrite(p, 4, n, TR.fp); }
static void put1(int32 v) { put(&v, 1); }

static int32 count_sub(chan_t *c)
{
    int32 n = 0;
    for (; c; c = c->alt) n += 1 + count_sub(c->next);
    return n;
}

static void number_sub(chan_t *c, int32 *next_id)
{
    for (; c; c = c->alt) {
        TR.chan[*next_id] = c;
        TR.map[*next_id].p = c; TR.map[*next_id].id = *next_id;
        ++*next_id;
        number_sub(c->next, next_id);
    }
}

static void write_topology(ngram_search_t *ngs)
{
    int32 i, n_non = 0, next_id, n_edge = 0, n_pw = 0, n_ci, w;
    int32 *child_off, *child, *ciph, *pw_off, *pw_wid, *pw_lp;
    dict_t *dict = ps_search_dict(ngs);

    for (i = 0; i < ngs->n_root_chan; i++) n_non += count_sub(ngs->root_chan[i].next);
    TR.n_root = ngs->n_root_chan;
    TR.n_chan = TR.n_root + n_non;
    TR.chan = ckd_realloc(TR.chan, TR.n_chan * sizeof(*TR.chan));
    TR.map = ckd_realloc(TR.map, TR.n_chan * sizeof(*TR.map));
    for (i = 0; i < TR.n_root; i++) {
        TR.chan[i] = (chan_t *)&ngs->root_chan[i];
        TR.map[i].p = &ngs->root_chan[i]; TR.map[i].id = i;
    }
    next_id = TR.n_root;
    for (i = 0; i < TR.n_root; i++) number_sub(ngs->root_chan[i].next, &next_id);
    TR.n_map = TR.n_chan;
    qsort(TR.map, TR.n_map, sizeof(*TR.map), cmp_ptr);

    child_off = ckd_calloc(TR.n_chan + 1, 4); pw_off = ckd_calloc(TR.n_chan + 1, 4); ciph = ckd_calloc(TR.n_chan, 4);
    for (i = 0; i < TR.n_chan; i++) {
        chan_t *c, *first = i < TR.n_root ? ngs->root_chan[i].next : TR.chan[i]->next;
        int32 pw = i < TR.n_root ? ngs->root_chan[i].penult_phn_wid : TR.chan[i]->info.penult_phn_wid;
        ciph[i] = i < TR.n_root ? ngs->root_chan[i].ciphone : TR.chan[i]->ciphone;
        child_off[i] = n_edge; pw_off[i] = n_pw;
        for (c = first; c; c = c->alt) ++n_edge;
        for (w = pw; w >= 0; w = ngs->homophone_set[w]) ++n_pw;
    }
    child_off[TR.n_chan] = n_edge; pw_off[TR.n_chan] = n_pw;
    child = ckd_calloc(n_edge + 1, 4); pw_wid = ckd_calloc(n_pw + 1, 4); pw_lp = ckd_calloc(n_pw + 1, 4);
    n_edge = n_pw = 0;
    for (i = 0; i < TR.n_chan; i++) {
        chan_t *c, *first = i < TR.n_root ? ngs->root_chan[i].next : TR.chan[i]->next;
        int32 pw = i < TR.n_root ? ngs->root_chan[i].penult_phn_wid : TR.chan[i]->info.penult_phn_wid;
        for (c = first; c; c = c->alt) child[n_edge++] = id_of(c);
        for (w = pw; w >= 0; w = ngs->homophone_set[w]) { pw_wid[n_pw] = w; pw_lp[n_pw++] = dict_last_phone(dict, w); }
    }
    n_ci = bin_mdef_n_ciphone(ps_search_acmod(ngs)->mdef);
    put1('T'); put1(TR.n_root); put1(TR.n_chan); put1(n_edge); put1(n_pw); put1(n_ci);
    put(child_off, TR.n_chan + 1); put(child, n_edge); put(ciph, TR.n_chan);
    put(pw_off, TR.n_chan + 1); put(pw_wid, n_pw); put(pw_lp, n_pw);
    ckd_free(child_off); ckd_free(child); ckd_free(ciph); ckd_free(pw_off); ckd_free(pw_wid); ckd_free(pw_lp);
    TR.tree_of = ngs; TR.tree_root = ngs->n_root_chan; TR.tree_nonroot = ngs->n_nonroot_chan;
}

static void put_state(void)
{
    int32 i, row[10];
    for (i = 0; i < TR.n_chan; i++) {
        hmm_t *h = &TR.chan[i]->hmm;
        row[0] = h->score[0]; row[1] = h->score[1]; row[2] = h->score[2];
        row[3] = h->history[0]; row[4] = h->history[1]; row[5] = h->history[2];
        row[6] = h->out_score; row[7] = h->out_history; row[8] = h->bestscore; row[9] = h->frame;
        put(row, 10);
    }
}

static void snapshot_before(ngram_search_t *ngs, int frame_idx)
{
    phone_loop_search_t *pls = (phone_loop_search_t *)ps_search_base(ngs)->pls;
    int32 i, n_ci = bin_mdef_n_ciphone(ps_search_acmod(ngs)->mdef);
    int32 n_act = ngs->n_active_chan[frame_idx & 1];
    chan_t **acl = ngs->active_chan_list[frame_idx & 1];

    if (hmm_n_emit_state(&ngs->root_chan[0].hmm) != 3) { TR.on = 0; return; }
    if (TR.tree_of != ngs || TR.tree_root != ngs->n_root_chan || TR.tree_nonroot != ngs->n_nonroot_chan || frame_idx == 0)
        write_topology(ngs);
    put1('B'); put1(frame_idx); put1(ngs->best_score); put1(ngs->dynamic_beam); put1(ngs->pbeam); put1(ngs->lpbeam);
    put1(ngs->pip); put1(ngs->nwpen); put1(pls != NULL); put1(n_act);
    for (i = 0; i < n_ci; i++) put1(phone_loop_search_score(pls, i));
    for (i = 0; i < n_act; i++) put1(id_of(acl[i]));
    put_state();
}

static void snapshot_after(ngram_search_t *ngs, int cand_valid)
{
    int32 nf = TR.frame + 1, i;
    int32 n = ngs->n_active_chan[nf & 1];
    chan_t **nacl = ngs->active_chan_list[nf & 1];
    put1('A'); put1(TR.frame); put1(cand_valid); put1(n); put1(cand_valid ? ngs->n_lastphn_cand : 0);
    for (i = 0; i < n; i++) put1(id_of(nacl[i]));
    if (cand_valid)
        for (i = 0; i < ngs->n_lastphn_cand; i++) {
            put1(ngs->lastphn_cand[i].wid); put1(ngs->lastphn_cand[i].score); put1(ngs->lastphn_cand[i].bp);
        }
    put_state();
    fflush(TR.fp);
}

static void b200_trace_lookahead(ngram_search_t *ngs, int frame_idx)
{
    if (TR.every == 0) {
        const char *f = getenv("B200_FWDTREE_TRACE"), *e = getenv("B200_FWDTREE_TRACE_EVERY");
        TR.every = -1;
        if (f && (TR.fp = fopen(f, "ab")) != NULL) TR.every = e ? atoi(e) : 1;
        if (TR.every < 1) TR.every = -1;
        if (getenv("B200_FWDTREE_TRACE_SHUFFLE")) TR.shuffle = (unsigned)atoi(getenv("B200_FWDTREE_TRACE_SHUFFLE")) * 2u + 1u;
    }
    if (TR.every < 0) return;
    if (TR.phase == 0 || TR.frame != frame_idx) {            /* prune_root_chan of a new frame */
        TR.frame = frame_idx; TR.phase = 1;
        TR.on = (frame_idx % TR.every) == 0;
        if (TR.shuffle) {
            chan_t **acl = ngs->active_chan_list[frame_idx & 1];
            int32 i, n = ngs->n_active_chan[frame_idx & 1];
            for (i = n - 1; i > 0; --i) {
                int32 j; chan_t *t;
                TR.shuffle = TR.shuffle * 1664525u + 1013904223u;
                j = (int32)((TR.shuffle >> 8) % (unsigned)(i + 1));
                t = acl[i]; acl[i] = acl[j]; acl[j] = t;
            }
        }
        if (TR.on) snapshot_before(ngs, frame_idx);
    }
    else if (TR.phase == 1) TR.phase = 2;                   /* prune_nonroot_chan */
    else {                                                  /* word_transition */
        if (TR.phase == 2 && TR.on) snapshot_after(ngs, ngs->n_lastphn_cand == 0);
        TR.phase = 0;
    }
}

static void b200_trace_exit(ngram_search_t *ngs)
{
    if (TR.every > 0 && TR.phase == 2) {
        if (TR.on) snapshot_after(ngs, 1);
        TR.phase = 3;
    }
}
