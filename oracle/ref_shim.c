/* oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin ctypes-callable wrapper around the UNMODIFIED reference C code
 * (compiled in place from /root/reference by oracle/Makefile into
 * oracle/_ref/libpocketsphinx.so).  Nothing here re-implements an algorithm:
 * every entry point marshals flat arrays into the reference's own structs and
 * calls the reference's own functions:
 *
 *   ms_mgau_init / ms_cont_mgau_frame_eval    pocketsphinx/src/libpocketsphinx/ms_mgau.c:79,162
 *   acmod_init / ps_mgau_frame_eval           .../acmod.c:224, acmod.h:119
 *   hmm_context_init / hmm_vit_eval           .../hmm.c:55,788
 *   logmath_init / logmath_add / logmath_log  sphinxbase/src/libsphinxbase/util/logmath.c:61,391,446
 *   tmat_init                                 .../tmat.c:191
 *
 * Used (a) to pin oracle/sphinx_oracle.c, (b) to generate tests/golden/,
 * (c) as bench.py's `--impl reference` / cpu_baseline "reference" leg.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <sphinxbase/cmd_ln.h>
#include <sphinxbase/logmath.h>
#include <sphinxbase/ckd_alloc.h>
#include <sphinxbase/err.h>
#include <sphinxbase/feat.h>

#include <pocketsphinx.h>
#include "acmod.h"
#include "hmm.h"
#include "tmat.h"
#include "ms_mgau.h"
#include "ms_gauden.h"
#include "ms_senone.h"

/* ------------------------------------------------------------------ logmath */

/* Dump the log-add table logmath_init(base, shift, TRUE) builds.  Returns the
 * table size; writes min(size, max_out) entries. */
int
ref_logadd_table(double base, int shift, int32 *out, int max_out)
{
    logmath_t *lm = logmath_init(base, shift, TRUE);
    uint32 size, width, tshift, i;
    const void *tbl;
    if (lm == NULL)
        return -1;
    logmath_get_table_shape(lm, &size, &width, &tshift);
    /* No public accessor for the table bytes: probe through logmath_add,
     * which returns r + table[d] for x - y = d (logmath.c:428-435). */
    (void)tbl;
    for (i = 0; i < size && (int)i < max_out; ++i)
        out[i] = logmath_add(lm, 0, -(int)i) - 0;
    logmath_free(lm);
    return (int)size;
}

int
ref_logmath_zero(double base, int shift)
{
    logmath_t *lm = logmath_init(base, shift, FALSE);
    int z = logmath_get_zero(lm);
    logmath_free(lm);
    return z;
}

void
ref_logmath_log(double base, int shift, const double *p, int n, int32 *out)
{
    logmath_t *lm = logmath_init(base, shift, FALSE);
    int i;
    for (i = 0; i < n; ++i)
        out[i] = logmath_log(lm, p[i]);
    logmath_free(lm);
}

void
ref_logmath_add(double base, int shift, const int32 *x, const int32 *y, int n, int32 *out)
{
    logmath_t *lm = logmath_init(base, shift, TRUE);
    int i;
    for (i = 0; i < n; ++i)
        out[i] = logmath_add(lm, x[i], y[i]);
    logmath_free(lm);
}

/* ------------------------------------------------------------ ms_cont_mgau */

typedef struct {
    cmd_ln_t *config;
    logmath_t *lmath;
    ps_mgau_t *mgau;
} ref_ms_t;

void *
ref_ms_init(const char *mean, const char *var, const char *mixw,
            const char *senmgau, double varfloor, double mixwfloor,
            int topn, int aw, double logbase)
{
    ref_ms_t *h = ckd_calloc(1, sizeof(*h));
    char vf[64], mf[64], tn[32], awb[32], lb[64];
    snprintf(vf, sizeof vf, "%.17g", varfloor);
    snprintf(mf, sizeof mf, "%.17g", mixwfloor);
    snprintf(tn, sizeof tn, "%d", topn);
    snprintf(awb, sizeof awb, "%d", aw);
    snprintf(lb, sizeof lb, "%.17g", logbase);
    err_set_logfp(NULL);
    h->config = cmd_ln_init(NULL, ps_args(), TRUE,
                            "-mean", mean, "-var", var, "-mixw", mixw,
                            "-senmgau", senmgau, "-varfloor", vf,
                            "-mixwfloor", mf, "-topn", tn, "-aw", awb,
                            "-logbase", lb, NULL);
    if (h->config == NULL)
        return NULL;
    h->lmath = logmath_init((float64)cmd_ln_float32_r(h->config, "-logbase"), 0, FALSE);
    /* mdef is only consulted for the ".ptm." mapping (ms_senone.c:306-340). */
    h->mgau = ms_mgau_init(h->config, h->lmath, NULL);
    return h;
}

void
ref_ms_free(void *vh)
{
    ref_ms_t *h = vh;
    if (!h) return;
    ps_mgau_free(h->mgau);
    logmath_free(h->lmath);
    cmd_ln_free_r(h->config);
    ckd_free(h);
}

/* dims[0..5] = n_mgau, n_feat, n_density, n_sen, topn, total feature length;
 * featlen (if non-NULL) receives n_feat stream lengths. */
void
ref_ms_dims(void *vh, int32 *dims, int32 *featlen)
{
    ref_ms_t *h = vh;
    ms_mgau_model_t *m = (ms_mgau_model_t *)h->mgau;
    int f, tot = 0;
    dims[0] = m->g->n_mgau;
    dims[1] = m->g->n_feat;
    dims[2] = m->g->n_density;
    dims[3] = m->s->n_sen;
    dims[4] = m->topn;
    for (f = 0; f < m->g->n_feat; ++f) {
        if (featlen) featlen[f] = m->g->featlen[f];
        tot += m->g->featlen[f];
    }
    dims[5] = tot;
}

/* Copy out the arrays gauden_init / senone_init produced (load-time KAT):
 * mean, var(precomputed) as [mgau][feat][density][dim] flat, det as
 * [mgau][feat][density], mixw as logical [sen][feat][cw] uint8. */
void
ref_ms_params(void *vh, float *mean, float *var, float *det, uint8 *mixw)
{
    ref_ms_t *h = vh;
    ms_mgau_model_t *m = (ms_mgau_model_t *)h->mgau;
    gauden_t *g = m->g;
    senone_t *s = m->s;
    int mg, f, d, i, c, sn;
    size_t k = 0, kd = 0;
    for (mg = 0; mg < g->n_mgau; ++mg)
        for (f = 0; f < g->n_feat; ++f)
            for (d = 0; d < g->n_density; ++d) {
                if (det) det[kd] = g->det[mg][f][d];
                ++kd;
                for (i = 0; i < g->featlen[f]; ++i, ++k) {
                    if (mean) mean[k] = g->mean[mg][f][d][i];
                    if (var) var[k] = g->var[mg][f][d][i];
                }
            }
    if (mixw) {
        k = 0;
        for (sn = 0; sn < (int)s->n_sen; ++sn)
            for (f = 0; f < (int)s->n_feat; ++f)
                for (c = 0; c < (int)s->n_cw; ++c, ++k)
                    mixw[k] = (s->n_gauden > 1) ? s->pdf[sn][f][c] : s->pdf[f][c][sn];
    }
}

/* Score T frames with compallsen=1.  feat: [T][featdim] float32 (streams
 * concatenated), out: [T][n_sen] int16. */
int
ref_ms_eval_all(void *vh, const float *feat, int T, int16 *out)
{
    ref_ms_t *h = vh;
    ms_mgau_model_t *m = (ms_mgau_model_t *)h->mgau;
    gauden_t *g = m->g;
    int n_sen = m->s->n_sen, t, f, tot = 0;
    mfcc_t **fp = ckd_calloc(g->n_feat, sizeof(*fp));
    for (f = 0; f < g->n_feat; ++f) tot += g->featlen[f];
    for (t = 0; t < T; ++t) {
        int off = 0;
        for (f = 0; f < g->n_feat; ++f) {
            fp[f] = (mfcc_t *)(feat + (size_t)t * tot + off);
            off += g->featlen[f];
        }
        ps_mgau_frame_eval(h->mgau, out + (size_t)t * n_sen, NULL, 0, fp, t, 1);
    }
    ckd_free(fp);
    return 0;
}

/* Score ONE frame with an active list given as uint8 deltas (acmod.c:1219-1271
 * format).  out must hold n_sen int16; only active entries are written. */
int
ref_ms_eval_active(void *vh, const float *feat, const uint8 *deltas, int n_active,
                   int frame, int16 *out)
{
    ref_ms_t *h = vh;
    ms_mgau_model_t *m = (ms_mgau_model_t *)h->mgau;
    gauden_t *g = m->g;
    int f, off = 0;
    mfcc_t **fp = ckd_calloc(g->n_feat, sizeof(*fp));
    for (f = 0; f < g->n_feat; ++f) {
        fp[f] = (mfcc_t *)(feat + off);
        off += g->featlen[f];
    }
    ps_mgau_frame_eval(h->mgau, out, (uint8 *)deltas, n_active, fp, frame, 0);
    ckd_free(fp);
    return 0;
}

/* --------------------------------------------------- acmod (any back-end) */

typedef struct {
    cmd_ln_t *config;
    logmath_t *lmath;
    acmod_t *acmod;
} ref_acmod_t;

/* Open an acoustic model directory through the reference's acmod_init so the
 * reference's own back-end selection (acmod.c:110-127) runs.  `senmgau` may be
 * NULL/"" (auto: s2_semi -> ptm -> ms) or ".semi."/".ptm."/".cont." to force
 * the generic ms back-end. */
void *
ref_acmod_open_ex(const char *hmmdir, const char *senmgau, int topn, int ds, double logbase,
                  const char *topn_beam)
{
    ref_acmod_t *h = ckd_calloc(1, sizeof(*h));
    char tn[32], dsb[32], lb[64];
    snprintf(tn, sizeof tn, "%d", topn);
    snprintf(dsb, sizeof dsb, "%d", ds);
    snprintf(lb, sizeof lb, "%.17g", logbase);
    err_set_logfp(NULL);
    if (senmgau && senmgau[0])
        h->config = cmd_ln_init(NULL, ps_args(), TRUE, "-hmm", hmmdir,
                                "-senmgau", senmgau, "-topn", tn, "-ds", dsb,
                                "-logbase", lb, "-compallsen", "yes", NULL);
    else
        h->config = cmd_ln_init(NULL, ps_args(), TRUE, "-hmm", hmmdir,
                                "-topn", tn, "-ds", dsb,
                                "-logbase", lb, "-compallsen", "yes", NULL);
    if (h->config == NULL)
        return NULL;
    if (topn_beam && topn_beam[0])
        cmd_ln_set_str_r(h->config, "-topn_beam", topn_beam);
    /* Mirror ps_init_defaults (pocketsphinx.c:146-158): expand -hmm. */
    {
        static const char *const files[][2] = {
            {"-mdef", "mdef"}, {"-mean", "means"}, {"-var", "variances"},
            {"-tmat", "transition_matrices"}, {"-mixw", "mixture_weights"},
            {"-sendump", "sendump"}, {"-featparams", "feat.params"},
            {"-lda", "feature_transform"}, {"-senmgau", "senmgau"},
        };
        size_t i;
        for (i = 0; i < sizeof(files) / sizeof(files[0]); ++i) {
            char path[4096];
            FILE *fp;
            if (cmd_ln_str_r(h->config, files[i][0]) != NULL)
                continue;
            snprintf(path, sizeof path, "%s/%s", hmmdir, files[i][1]);
            if ((fp = fopen(path, "rb")) != NULL) {
                fclose(fp);
                cmd_ln_set_str_r(h->config, files[i][0], path);
            }
        }
    }
    h->lmath = logmath_init((float64)cmd_ln_float32_r(h->config, "-logbase"), 0, FALSE);
    h->acmod = acmod_init(h->config, h->lmath, NULL, NULL);
    if (h->acmod == NULL)
        return NULL;
    return h;
}

void
ref_acmod_close(void *vh)
{
    ref_acmod_t *h = vh;
    if (!h) return;
    acmod_free(h->acmod);
    logmath_free(h->lmath);
    cmd_ln_free_r(h->config);
    ckd_free(h);
}

const char *
ref_acmod_backend(void *vh)
{
    ref_acmod_t *h = vh;
    return h->acmod->mgau->vt->name;
}

/* info[0..3] = n_sen, n_streams, total feature dim, n_emit_state;
 * streamlen gets the per-stream lengths. */
void
ref_acmod_info(void *vh, int32 *info, int32 *streamlen)
{
    ref_acmod_t *h = vh;
    int f, tot = 0;
    info[0] = bin_mdef_n_sen(h->acmod->mdef);
    info[1] = feat_dimension1(h->acmod->fcb);
    for (f = 0; f < info[1]; ++f) {
        if (streamlen) streamlen[f] = feat_dimension2(h->acmod->fcb, f);
        tot += feat_dimension2(h->acmod->fcb, f);
    }
    info[2] = tot;
    info[3] = bin_mdef_n_emit_state(h->acmod->mdef);
}

/* Score T frames of dynamic features (streams concatenated per frame) with
 * compallsen=1 through whichever back-end acmod chose.  Frames are fed in
 * order with frame index t, and mgau->frame_idx kept in step the way
 * acmod_advance does (acmod.c:880), so ptm/s2 top-N history behaves as in a
 * real utterance. */
int
ref_acmod_score_feats(void *vh, const float *feat, int T, int16 *out)
{
    ref_acmod_t *h = vh;
    acmod_t *a = h->acmod;
    int nf = feat_dimension1(a->fcb), n_sen = bin_mdef_n_sen(a->mdef);
    int t, f, tot = 0;
    mfcc_t **fp = ckd_calloc(nf, sizeof(*fp));
    for (f = 0; f < nf; ++f) tot += feat_dimension2(a->fcb, f);
    a->mgau->frame_idx = 0;
    for (t = 0; t < T; ++t) {
        int off = 0;
        for (f = 0; f < nf; ++f) {
            fp[f] = (mfcc_t *)(feat + (size_t)t * tot + off);
            off += feat_dimension2(a->fcb, f);
        }
        ps_mgau_frame_eval(a->mgau, out + (size_t)t * n_sen, NULL, 0, fp, t, 1);
        a->mgau->frame_idx = t + 1;
    }
    ckd_free(fp);
    return 0;
}

/* One ps_mgau_frame_eval call with an explicit active list (uint8 deltas) or
 * compallsen; the caller steps frames in order.  mgau->frame_idx is set to
 * `frame` before the call and frame+1 after, like acmod_score/acmod_advance. */
int
ref_acmod_frame_eval(void *vh, const float *feat, const uint8 *deltas, int n_active,
                     int frame, int compallsen, int16 *out)
{
    ref_acmod_t *h = vh;
    acmod_t *a = h->acmod;
    int nf = feat_dimension1(a->fcb), f, off = 0;
    mfcc_t **fp = ckd_calloc(nf, sizeof(*fp));
    for (f = 0; f < nf; ++f) {
        fp[f] = (mfcc_t *)(feat + off);
        off += feat_dimension2(a->fcb, f);
    }
    a->mgau->frame_idx = frame;
    ps_mgau_frame_eval(a->mgau, out, (uint8 *)deltas, n_active, fp, frame, compallsen);
    a->mgau->frame_idx = frame + 1;
    ckd_free(fp);
    return 0;
}

/* bin_mdef_sen2cimap for every senone (the ptm senone->codebook map,
 * ptm_mgau.c:834-836). */
void
ref_acmod_sen2cimap(void *vh, uint8 *out)
{
    ref_acmod_t *h = vh;
    int i, n = bin_mdef_n_sen(h->acmod->mdef);
    for (i = 0; i < n; ++i)
        out[i] = (uint8)bin_mdef_sen2cimap(h->acmod->mdef, i);
}

/* Feature extraction only: 13-dim cepstra [n_cep_frames][ceplen] -> dynamic
 * features [T][featdim] using the model's own feat_t in whole-utterance mode
 * (feat_s2mfc2feat_live with beginutt=endutt=TRUE, acmod.c:513-540).
 * Returns the number of feature frames written (<= max_T). */
int
ref_acmod_cep2feat(void *vh, const float *cep, int n_frames, int ceplen,
                   float *feat_out, int max_T)
{
    ref_acmod_t *h = vh;
    acmod_t *a = h->acmod;
    int nf = feat_dimension1(a->fcb), f, tot = 0, t, nfr;
    mfcc_t **cepp = ckd_calloc(n_frames, sizeof(*cepp));
    mfcc_t ***fb;
    int32 ncep = n_frames;
    for (f = 0; f < nf; ++f) tot += feat_dimension2(a->fcb, f);
    for (t = 0; t < n_frames; ++t)
        cepp[t] = (mfcc_t *)(cep + (size_t)t * ceplen);
    fb = feat_array_alloc(a->fcb, n_frames + 16);
    nfr = feat_s2mfc2feat_live(a->fcb, cepp, &ncep, TRUE, TRUE, fb);
    if (nfr > max_T) nfr = max_T;
    for (t = 0; t < nfr; ++t)
        memcpy(feat_out + (size_t)t * tot, fb[t][0], tot * sizeof(float));
    feat_array_free(fb);
    ckd_free(cepp);
    return nfr;
}

/* The reference's own feat_t, built directly (no model directory): feat_init
 * (feat.c:851-1037) for `type` / cmn / varnorm / agc, an LDA matrix installed
 * the way feat_read_lda leaves it (lda.c:104-134), -svspec through
 * parse_subvecs + feat_set_subvecs, then one whole utterance through
 * feat_s2mfc2feat_live(beginutt = endutt = TRUE).  out receives rows of
 * *row_len floats (the pre-LDA vector length, as feat_array_alloc lays them
 * out); returns the number of frames, -1 on a configuration error. */
int
ref_feat_compute(const char *type, const char *cmn, int varnorm, const char *agc, int cepsize,
                 const float *lda, int lda_m, int lda_n, int lda_dim, const char *svspec,
                 const float *cep_in, int n_frames, float *out, int *row_len, int *out_dim)
{
    feat_t *fcb;
    mfcc_t **cepp, ***fb;
    float *cep;
    int32 ncep = n_frames;
    int f, tot = 0, t, nfr;
    err_set_logfp(NULL);
    fcb = feat_init(type, cmn_type_from_str(cmn), varnorm, agc_type_from_str(agc), 0, cepsize);
    if (fcb == NULL) return -1;
    if (lda) {
        if (fcb->n_stream != 1 || lda_n != (int)fcb->stream_len[0]) { feat_free(fcb); return -1; }
        fcb->lda = (mfcc_t ***)ckd_calloc_3d(1, lda_m, lda_n, sizeof(float));
        memcpy(fcb->lda[0][0], lda, sizeof(float) * lda_m * lda_n);
        fcb->n_lda = 1;
        if (lda_dim > lda_m || lda_dim <= 0) lda_dim = lda_m;
        fcb->out_dim = lda_dim;
    }
    if (svspec && *svspec) {
        int32 **sv = parse_subvecs(svspec);
        if (sv == NULL || feat_set_subvecs(fcb, sv) < 0) { feat_free(fcb); return -1; }
    }
    for (f = 0; f < fcb->n_stream; ++f) tot += fcb->stream_len[f];
    *row_len = tot;
    *out_dim = fcb->sv_dim ? fcb->sv_dim : feat_dimension(fcb);
    if (n_frames <= 0) { feat_free(fcb); return 0; }
    cep = ckd_calloc((size_t)n_frames * cepsize, sizeof(float));   /* normalised IN PLACE */
    memcpy(cep, cep_in, sizeof(float) * (size_t)n_frames * cepsize);
    cepp = ckd_calloc(n_frames, sizeof(*cepp));
    for (t = 0; t < n_frames; ++t) cepp[t] = (mfcc_t *)(cep + (size_t)t * cepsize);
    fb = feat_array_alloc(fcb, n_frames + 16);
    nfr = feat_s2mfc2feat_live(fcb, cepp, &ncep, TRUE, TRUE, fb);
    for (t = 0; t < nfr; ++t) memcpy(out + (size_t)t * tot, fb[t][0], tot * sizeof(float));
    feat_array_free(fb);
    ckd_free(cepp);
    ckd_free(cep);
    feat_free(fcb);
    return nfr;
}

void *
ref_acmod_open(const char *hmmdir, const char *senmgau, int topn, int ds, double logbase)
{
    return ref_acmod_open_ex(hmmdir, senmgau, topn, ds, logbase, NULL);
}

/* Copy out tmat->tp as [n_tmat][n_state][n_state+1] uint8 and the mdef's
 * sseq table as [n_sseq][n_emit] uint16.  Pass NULL to query sizes only.
 * sizes[0..2] = n_tmat, n_emit_state, n_sseq. */
void
ref_acmod_tables(void *vh, int32 *sizes, uint8 *tp, uint16 *sseq)
{
    ref_acmod_t *h = vh;
    acmod_t *a = h->acmod;
    int n = a->tmat->n_state, nt = a->tmat->n_tmat, i, j, k;
    int ns = bin_mdef_n_sseq(a->mdef), ne = bin_mdef_n_emit_state(a->mdef);
    sizes[0] = nt; sizes[1] = n; sizes[2] = ns;
    if (tp)
        for (i = 0; i < nt; ++i)
            for (j = 0; j < n; ++j)
                for (k = 0; k <= n; ++k)
                    tp[(i * n + j) * (n + 1) + k] = a->tmat->tp[i][j][k];
    if (sseq)
        for (i = 0; i < ns; ++i)
            for (j = 0; j < ne; ++j)
                sseq[i * ne + j] = a->mdef->sseq[i][j];
}

/* ------------------------------------------------------------------- tmat */

/* Run tmat_init on a transition_matrices file; copies tp out.  Returns
 * n_tmat (or <0); *n_state_out gets the emitting state count. */
int
ref_tmat_load(const char *file, double tmatfloor, double logbase,
              uint8 *tp_out, int max_bytes, int32 *n_state_out)
{
    logmath_t *lm = logmath_init(logbase, 0, FALSE);
    tmat_t *t;
    int i, j, k, n, nt;
    err_set_logfp(NULL);
    t = tmat_init(file, lm, tmatfloor, TRUE);
    if (t == NULL) { logmath_free(lm); return -1; }
    n = t->n_state; nt = t->n_tmat;
    *n_state_out = n;
    for (i = 0; i < nt; ++i)
        for (j = 0; j < n; ++j)
            for (k = 0; k <= n; ++k) {
                int idx = (i * n + j) * (n + 1) + k;
                if (idx < max_bytes) tp_out[idx] = t->tp[i][j][k];
            }
    tmat_free(t);
    logmath_free(lm);
    return nt;
}

/* -------------------------------------------------------------------- hmm */

/* Batched hmm_vit_eval over SoA arrays (all in/out unless noted):
 *   score[n_hmm][n_emit], history[n_hmm][n_emit], out_score[n_hmm],
 *   out_history[n_hmm], senid[n_hmm][n_emit] (senone ids, or per-state ssids
 *   for mpx), ssid[n_hmm] (in), tmatid[n_hmm] (in), mpx[n_hmm] (in),
 *   bestscore[n_hmm] (out).
 * tp: [n_tmat][n_emit][n_emit+1] uint8; sseq: [n_sseq][n_emit] uint16;
 * senscr: [n_sen] int16.  Returns max over HMMs of bestscore. */
int32
ref_hmm_eval_batch(int n_emit, int n_hmm, const uint8 *tp, int n_tmat,
                   const uint16 *sseq, int n_sseq, const int16 *senscr,
                   int32 *score, int32 *history, int32 *out_score,
                   int32 *out_history, uint16 *senid, const uint16 *ssid,
                   const int16 *tmatid, const uint8 *mpx, int32 *bestscore,
                   int n_frames_repeat)
{
    uint8 ***tpp = (uint8 ***)ckd_calloc_3d(n_tmat, n_emit, n_emit + 1, sizeof(uint8));
    uint16 **ss = ckd_calloc(n_sseq > 0 ? n_sseq : 1, sizeof(*ss));
    hmm_context_t *ctx;
    hmm_t *hm = ckd_calloc(n_hmm, sizeof(*hm));
    int32 best = WORST_SCORE;
    int i, j, r;

    memcpy(tpp[0][0], tp, (size_t)n_tmat * n_emit * (n_emit + 1));
    for (i = 0; i < n_sseq; ++i)
        ss[i] = (uint16 *)(sseq + (size_t)i * n_emit);
    ctx = hmm_context_init(n_emit, (uint8 ** const *)tpp, senscr, ss);
    for (i = 0; i < n_hmm; ++i) {
        hmm_t *h = &hm[i];
        h->ctx = ctx;
        h->mpx = mpx[i];
        h->n_emit_state = n_emit;
        h->ssid = ssid[i];
        h->tmatid = tmatid[i];
        h->frame = 0;
        for (j = 0; j < n_emit; ++j) {
            h->score[j] = score[i * n_emit + j];
            h->history[j] = history[i * n_emit + j];
            h->senid[j] = senid[i * n_emit + j];
        }
        h->out_score = out_score[i];
        h->out_history = out_history[i];
        h->bestscore = bestscore[i];
    }
    for (r = 0; r < (n_frames_repeat > 0 ? n_frames_repeat : 1); ++r) {
        best = WORST_SCORE;
        for (i = 0; i < n_hmm; ++i) {
            int32 b = hmm_vit_eval(&hm[i]);
            if (b BETTER_THAN best) best = b;
        }
    }
    for (i = 0; i < n_hmm; ++i) {
        hmm_t *h = &hm[i];
        for (j = 0; j < n_emit; ++j) {
            score[i * n_emit + j] = h->score[j];
            history[i * n_emit + j] = h->history[j];
            senid[i * n_emit + j] = h->senid[j];
        }
        out_score[i] = h->out_score;
        out_history[i] = h->out_history;
        bestscore[i] = h->bestscore;
    }
    hmm_context_free(ctx);
    ckd_free(hm);
    ckd_free(ss);
    ckd_free_3d(tpp);
    return best;
}

/* hmm_clear_scores / hmm_normalize / hmm_enter (hmm.c:169-218) on SoA arrays
 * [n_hmm][n_emit] (score, history) + out_score, out_history, bestscore:
 *   op 0: hmm_clear_scores(h) for every HMM with sel[i] != 0
 *   op 1: hmm_normalize(h, arg[i]) for every HMM
 *   op 2: the callers' entry loop (ngram_search_fwdtree.c:757,846): for k in
 *         0..n_list-1: if (lscore[k] BETTER_THAN hmm_in_score(h[lidx[k]]))
 *         hmm_enter(h[lidx[k]], lscore[k], lhist[k], frame) */
int
ref_hmm_maint(int op, int n_emit, int n_hmm, int32 *score, int32 *history, int32 *out_score,
              int32 *out_history, int32 *bestscore, const uint8 *sel, const int32 *arg,
              int n_list, const int32 *lidx, const int32 *lscore, const int32 *lhist)
{
    static uint8 tp1[HMM_MAX_NSTATE * (HMM_MAX_NSTATE + 1)];
    uint8 ***tpp = (uint8 ***)ckd_calloc_3d(1, n_emit, n_emit + 1, sizeof(uint8));
    hmm_context_t *ctx = hmm_context_init(n_emit, (uint8 ** const *)tpp, NULL, NULL);
    hmm_t *hm = ckd_calloc(n_hmm, sizeof(*hm));
    int i, j, k;
    (void)tp1;
    for (i = 0; i < n_hmm; ++i) {
        hmm_t *h = &hm[i];
        h->ctx = ctx; h->mpx = 0; h->n_emit_state = n_emit; h->frame = -1;
        for (j = 0; j < n_emit; ++j) { h->score[j] = score[i * n_emit + j]; h->history[j] = history[i * n_emit + j]; }
        h->out_score = out_score[i]; h->out_history = out_history[i]; h->bestscore = bestscore[i];
    }
    if (op == 0) { for (i = 0; i < n_hmm; ++i) if (sel[i]) hmm_clear_scores(&hm[i]); }
    else if (op == 1) { for (i = 0; i < n_hmm; ++i) hmm_normalize(&hm[i], arg[i]); }
    else if (op == 2) {
        for (k = 0; k < n_list; ++k) {
            hmm_t *h = &hm[lidx[k]];
            if (lscore[k] BETTER_THAN hmm_in_score(h)) hmm_enter(h, lscore[k], lhist[k], 1);
        }
    }
    for (i = 0; i < n_hmm; ++i) {
        hmm_t *h = &hm[i];
        for (j = 0; j < n_emit; ++j) { score[i * n_emit + j] = h->score[j]; history[i * n_emit + j] = h->history[j]; }
        out_score[i] = h->out_score; out_history[i] = h->out_history; bestscore[i] = h->bestscore;
    }
    hmm_context_free(ctx);
    ckd_free(hm);
    ckd_free_3d(tpp);
    return 0;
}
