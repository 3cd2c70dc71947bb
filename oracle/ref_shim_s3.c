/* oracle/ref_shim_s3.c -- TEST INFRASTRUCTURE ONLY.
 *
 * ctypes-callable marshalling around the reference's OWN sphinx3 functions
 * (mgau_init, mgau_eval, approx_cont_mgau_ci_eval, approx_cont_mgau_frame_eval,
 * fast_gmm_init, mdef_init), compiled against /root/reference by
 * oracle/Makefile (target ref3) into oracle/_ref/libref_shim_s3.so.  No
 * algorithm lives here: it only builds the structs the reference functions
 * take and loops over frames the way S3/libsearch/srch.c does.
 */
#include <string.h>
#include <stdlib.h>

#include <sphinxbase/ckd_alloc.h>
#include <sphinxbase/logmath.h>
#include <sphinxbase/profile.h>

#include "s3types.h"
#include "cont_mgau.h"
#include "approx_cont_mgau.h"
#include "fast_algo_struct.h"
#include "ascr.h"
#include "mdef.h"
#include "logs3.h"
#include "subvq.h"
#include "gs.h"
#include <sphinxbase/cmd_ln.h>

typedef struct {
    logmath_t *lmath;
    mgau_model_t *g;
    mdef_t *mdef;
    int own_mdef;       /* 1: hand-built minimal mdef_t */
    fast_gmm_t *fg;
    ascr_t a;           /* only senscr / sen_active / rec_sen_active used */
    ptmr_t tm;
    subvq_t *svq;       /* -subvq: sub-vector quantised shortlists (S3/libam/subvq.c), or NULL */
    double svqbeam;     /* -subvqbeam (probability) */
    cmd_ln_t *config;   /* carries -vqeval for subvq_init */
    gs_t *gs;           /* -gs: Gaussian selector map (S3/libam/gs.c), or NULL */
} s3h_t;

void *
ref_s3_open(const char *mean, const char *var, const char *mixw, const char *mdef_file,
            const int32 *cd2cisen, int n_sen, int n_ci_sen, double varfloor, double mixwfloor,
            double logbase)
{
    s3h_t *h = ckd_calloc(1, sizeof(*h));
    int i;
    h->lmath = logs3_init(logbase, 0, 1);
    h->g = mgau_init(mean, var, varfloor, mixw, mixwfloor, 1, ".cont.", MIX_INT_FLOAT_COMP, h->lmath);
    if (mdef_file) {
        h->mdef = mdef_init(mdef_file, 0);
    } else {
        h->mdef = ckd_calloc(1, sizeof(mdef_t));
        h->own_mdef = 1;
        h->mdef->n_sen = n_sen;
        h->mdef->n_ci_sen = n_ci_sen;
        h->mdef->cd2cisen = ckd_calloc(n_sen, sizeof(s3senid_t));
        for (i = 0; i < n_sen; ++i)
            h->mdef->cd2cisen[i] = (s3senid_t)cd2cisen[i];
    }
    h->svqbeam = 1e-80;
    h->fg = fast_gmm_init(1, 0, 0, 1, 0, h->svqbeam, 1e-80, 0.5f, 100000, h->mdef->n_ci_sen, h->lmath);
    h->a.senscr = ckd_calloc(h->g->n_mgau, sizeof(int32));
    h->a.sen_active = ckd_calloc(h->g->n_mgau, 1);
    h->a.rec_sen_active = ckd_calloc(h->g->n_mgau, 1);
    ptmr_init(&h->tm);
    return h;
}

void
ref_s3_set_fast(void *vh, double ci_pbeam, int max_cd, int ds_ratio, float tighten)
{
    s3h_t *h = vh;
    int n_ci = h->mdef->n_ci_sen;
    fast_gmm_free(h->fg);
    h->fg = fast_gmm_init(ds_ratio, 0, 0, 1, 0, h->svqbeam, ci_pbeam, tighten, max_cd, n_ci, h->lmath);
}

/* -subvq FILE -subvqbeam P -vqeval N: the reference's own subvq_init (subvq.c:206-373).
 * Call before ref_s3_set_fast (which carries the beam into fast_gmm_init).  0 on success. */
int
ref_s3_open_svq(void *vh, const char *file, double varfloor, int max_sv, int vqeval, double subvqbeam)
{
    static const arg_t defs[] = {
        { "-vqeval", ARG_INT32, "3", "sub-vectors used for the shortlist" },
        { NULL, 0, NULL, NULL }
    };
    s3h_t *h = vh;
    char buf[16];
    snprintf(buf, sizeof buf, "%d", vqeval);
    h->config = cmd_ln_init(NULL, defs, FALSE, "-vqeval", buf, NULL);
    h->svq = subvq_init(file, varfloor, max_sv, h->g, h->config, h->lmath);
    h->svqbeam = subvqbeam;
    return h->svq ? 0 : -1;
}

/* -gs FILE: the reference's own gs_read (gs.c:156-218).  0 on success. */
int
ref_s3_open_gs(void *vh, const char *file)
{
    s3h_t *h = vh;
    h->gs = gs_read(file, h->lmath);
    return h->gs ? 0 : -1;
}

/* gc_compute_closest_cw (gs.c:221-259) for T frames */
void
ref_s3_gs_closest(void *vh, const float *feat, int T, int32 *out)
{
    s3h_t *h = vh;
    int t, L = h->g->veclen;
    for (t = 0; t < T; ++t) out[t] = gc_compute_closest_cw(h->gs, (float32 *)feat + (size_t)t * L);
}

/* d[0..4] = n_sv, vqsize, origsize.r, origsize.c, VQ_EVAL; d[5] = logs3(subvqbeam) as fast_gmm_init stores it */
void
ref_s3_svq_dims(void *vh, int32 *d)
{
    s3h_t *h = vh;
    d[0] = h->svq->n_sv; d[1] = h->svq->vqsize; d[2] = h->svq->origsize.r; d[3] = h->svq->origsize.c;
    d[4] = h->svq->VQ_EVAL; d[5] = h->fg->gaus->subvqbeam;
}

/* precomputed tables of sub-vector sv: featdim[veclen], mean / var [vqsize][veclen] (var = 1/(2 var)), lrd[vqsize];
 * returns veclen; scal[0] = distfloor */
int
ref_s3_svq_tables(void *vh, int sv, int32 *featdim, float *mean, float *var, float *lrd, double *scal)
{
    s3h_t *h = vh;
    vector_gautbl_t *t = &h->svq->gautbl[sv];
    int r, L = t->veclen;
    for (r = 0; r < L; ++r) featdim[r] = h->svq->featdim[sv][r];
    for (r = 0; r < h->svq->vqsize; ++r) {
        memcpy(mean + (size_t)r * L, t->mean[r], L * sizeof(float));
        memcpy(var + (size_t)r * L, t->var[r], L * sizeof(float));
        lrd[r] = t->lrd[r];
    }
    scal[0] = t->distfloor;
    return L;
}

/* the compacted, linearised map [origsize.r][origsize.c][n_sv] */
void
ref_s3_svq_map(void *vh, int32 *out)
{
    s3h_t *h = vh;
    subvq_t *v = h->svq;
    int r, c, s;
    for (r = 0; r < v->origsize.r; ++r)
        for (c = 0; c < v->origsize.c; ++c)
            for (s = 0; s < v->n_sv; ++s)
                out[((size_t)r * v->origsize.c + c) * v->n_sv + s] = v->map[r][c][s];
}

/* subvq_gautbl_eval_logs3 for T frames: out [T][n_sv * vqsize] */
void
ref_s3_svq_vqdist(void *vh, const float *feat, int T, int32 *out)
{
    s3h_t *h = vh;
    subvq_t *v = h->svq;
    int t, n = v->n_sv * v->vqsize, L = h->g->veclen;
    for (t = 0; t < T; ++t) {
        subvq_gautbl_eval_logs3(v, (float32 *)feat + (size_t)t * L, h->lmath);
        memcpy(out + (size_t)t * n, v->vqdist[0], n * sizeof(int32));
    }
}

/* subvq_mgau_shortlist for one senone against the vqdist of the last evaluated frame: out[0..n) flags */
int
ref_s3_svq_shortlist(void *vh, int s, uint8 *out)
{
    s3h_t *h = vh;
    int i, n = mgau_n_comp(h->g, s);
    int ng = subvq_mgau_shortlist(h->svq, s, n, h->fg->gaus->subvqbeam);
    memset(out, 0, n);
    for (i = 0; h->svq->mgau_sl[i] >= 0; ++i) out[h->svq->mgau_sl[i]] = 1;
    return ng;
}

void
ref_s3_dims(void *vh, int32 *d)
{
    s3h_t *h = vh;
    d[0] = h->g->n_mgau; d[1] = h->g->max_comp; d[2] = h->g->veclen;
    d[3] = h->mdef->n_ci_sen; d[4] = h->fg->gmms->ci_pbeam;
}

void
ref_s3_cd2cisen(void *vh, int32 *out)
{
    s3h_t *h = vh;
    int i;
    for (i = 0; i < h->g->n_mgau; ++i) out[i] = h->mdef->cd2cisen[i];
}

/* precomputed parameters, padded to [n_mgau][max_comp][...] */
void
ref_s3_params(void *vh, int32 *n_comp, float *mean, float *var, float *lrd, int32 *mixw, double *scal)
{
    s3h_t *h = vh;
    mgau_model_t *g = h->g;
    int s, c, L = g->veclen, M = g->max_comp;
    for (s = 0; s < g->n_mgau; ++s) {
        n_comp[s] = g->mgau[s].n_comp;
        for (c = 0; c < g->mgau[s].n_comp; ++c) {
            memcpy(mean + ((size_t)s * M + c) * L, g->mgau[s].mean[c], L * sizeof(float));
            memcpy(var + ((size_t)s * M + c) * L, g->mgau[s].var[c], L * sizeof(float));
            lrd[(size_t)s * M + c] = g->mgau[s].lrd[c];
            mixw[(size_t)s * M + c] = g->mgau[s].mixw[c];
        }
    }
    scal[0] = g->distfloor;
    scal[1] = 1.0 / log(logmath_get_base(g->logmath));
}

void
ref_s3_utt_reset(void *vh)
{
    s3h_t *h = vh;
    int i;
    /* S3/libsearch/srch_time_switch_tree.c:484-490 */
    for (i = 0; i < h->g->n_mgau; i++) {
        h->g->mgau[i].bstidx = NO_BSTIDX;
        h->g->mgau[i].updatetime = NOT_UPDATED;
    }
}

void
ref_s3_state(void *vh, int32 *bstidx, int32 *updatetime)
{
    s3h_t *h = vh;
    int i;
    for (i = 0; i < h->g->n_mgau; i++) {
        bstidx[i] = h->g->mgau[i].bstidx;
        updatetime[i] = h->g->mgau[i].updatetime;
    }
}

/* dense: mgau_eval(g, s, NULL, x, t, 1) for every senone and frame */
void
ref_s3_eval_dense(void *vh, const float *feat, int T, int32 *out)
{
    s3h_t *h = vh;
    int t, s, S = h->g->n_mgau, L = h->g->veclen;
    for (t = 0; t < T; ++t)
        for (s = 0; s < S; ++s)
            out[(size_t)t * S + s] = mgau_eval(h->g, s, NULL, (float32 *)feat + (size_t)t * L, t, 1);
}

/* per utterance: CI pass then approx_cont_mgau_frame_eval per frame.
 * sen_active [T][S] in/out or NULL (= all ones each frame); senscr_io [S]. */
void
ref_s3_eval_utt(void *vh, const float *feat, int T, int frame0, uint8 *sen_active,
                int32 *senscr_io, int32 *out, int32 *best)
{
    s3h_t *h = vh;
    int t, S = h->g->n_mgau, L = h->g->veclen;
    int32 *ci = ckd_calloc(h->mdef->n_ci_sen > 0 ? h->mdef->n_ci_sen : 1, sizeof(int32));
    int32 cib;
    memcpy(h->a.senscr, senscr_io, S * sizeof(int32));
    for (t = 0; t < T; ++t) {
        float32 *x = (float32 *)feat + (size_t)t * L;
        if (sen_active) memcpy(h->a.sen_active, sen_active + (size_t)t * S, S);
        else memset(h->a.sen_active, 1, S);
        approx_cont_mgau_ci_eval(h->svq, h->gs, h->g, h->fg, h->mdef, x, ci, &cib, frame0 + t, h->lmath);
        best[t] = approx_cont_mgau_frame_eval(h->mdef, h->svq, h->gs, h->g, h->fg, &h->a, x, frame0 + t,
                                              ci, &h->tm, h->lmath);
        memcpy(out + (size_t)t * S, h->a.senscr, S * sizeof(int32));
        if (sen_active) memcpy(sen_active + (size_t)t * S, h->a.sen_active, S);
    }
    memcpy(senscr_io, h->a.senscr, S * sizeof(int32));
    ckd_free(ci);
}

void
ref_s3_close(void *vh)
{
    s3h_t *h = vh;
    if (!h) return;
    fast_gmm_free(h->fg);
    if (h->svq) subvq_free(h->svq);
    if (h->config) cmd_ln_free_r(h->config);
    mgau_free(h->g);
    if (h->own_mdef) { ckd_free(h->mdef->cd2cisen); ckd_free(h->mdef); }
    else mdef_free(h->mdef);
    ckd_free(h->a.senscr); ckd_free(h->a.sen_active); ckd_free(h->a.rec_sen_active);
    logmath_free(h->lmath);
    ckd_free(h);
}

/* ------------------------------------------------- sphinx3 hmm_vit_eval
 * Batched over HMM-major arrays (score/history/ssid [n_hmm][n_emit]); tp is
 * [n_tmat][n_emit][n_emit + 1] int32 (the hmm_tprob_3st / _5st macros index
 * ctx->tp[tmatid][0] with a row stride of n_emit + 1), sseq [n_sseq][n_emit]. */
#include "hmm.h"
int32
ref_s3hmm_eval_batch(int n_emit, int n_hmm, const int32 *tp, int n_tmat, const int16 *sseq, int n_sseq,
                     const int32 *senscr, int32 *score, int32 *history, int32 *out_score, int32 *out_history,
                     int32 *ssid, const int32 *tmatid, const uint8 *mpx, int32 *bestscore, int repeat)
{
    int32 ***tpp = (int32 ***)ckd_calloc_3d(n_tmat, n_emit + 1, n_emit + 1, sizeof(int32));
    s3senid_t **ss = ckd_calloc(n_sseq, sizeof(*ss));
    hmm_context_t *ctx;
    hmm_t *hm = ckd_calloc(n_hmm, sizeof(*hm));
    int32 best = S3_LOGPROB_ZERO;
    int i, j, k, r;
    for (i = 0; i < n_tmat; ++i)
        for (j = 0; j < n_emit; ++j)
            for (k = 0; k <= n_emit; ++k)
                tpp[i][j][k] = tp[((size_t)i * n_emit + j) * (n_emit + 1) + k];
    for (i = 0; i < n_sseq; ++i) ss[i] = (s3senid_t *)(sseq + (size_t)i * n_emit);
    ctx = hmm_context_init(n_emit, tpp, (int32 *)senscr, ss);
    for (i = 0; i < n_hmm; ++i) {
        hmm_t *h = &hm[i];
        hmm_init(ctx, h, mpx[i], ssid[(size_t)i * n_emit], tmatid[i]);
        for (j = 0; j < n_emit; ++j) {
            hmm_score(h, j) = score[(size_t)i * n_emit + j];
            hmm_history(h, j) = history[(size_t)i * n_emit + j];
            if (mpx[i]) hmm_mpx_ssid(h, j) = ssid[(size_t)i * n_emit + j];
        }
        hmm_out_score(h) = out_score[i];
        hmm_out_history(h) = out_history[i];
        hmm_bestscore(h) = bestscore[i];
    }
    for (r = 0; r < (repeat > 0 ? repeat : 1); ++r) {
        best = S3_LOGPROB_ZERO;
        for (i = 0; i < n_hmm; ++i) {
            int32 b = hmm_vit_eval(&hm[i]);
            if (b > best) best = b;
        }
    }
    for (i = 0; i < n_hmm; ++i) {
        hmm_t *h = &hm[i];
        for (j = 0; j < n_emit; ++j) {
            score[(size_t)i * n_emit + j] = hmm_score(h, j);
            history[(size_t)i * n_emit + j] = (int32)hmm_history(h, j);
            if (mpx[i]) ssid[(size_t)i * n_emit + j] = hmm_mpx_ssid(h, j);
        }
        out_score[i] = hmm_out_score(h);
        out_history[i] = (int32)hmm_out_history(h);
        bestscore[i] = hmm_bestscore(h);
        hmm_deinit(h);
    }
    hmm_context_free(ctx);
    ckd_free(hm);
    ckd_free(ss);
    ckd_free_3d((void ***)tpp);
    return best;
}
