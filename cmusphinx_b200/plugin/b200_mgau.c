/* b200_mgau.c -- the reference-side binding: a ps_mgau_t back-end
 * (pocketsphinx/src/libpocketsphinx/acmod.h:95-124) whose frame_eval runs on
 * the GPU through the C ABI of libb200sphinx.so (include/b200sphinx.h).
 *
 * Built as libb200_ps_plugin.so against the reference's own headers and used
 * with an UNMODIFIED libpocketsphinx.so:
 *
 *     LD_PRELOAD=libb200_ps_plugin.so pocketsphinx_batch -hmm ... -ctl ...
 *
 * It interposes the three back-end constructors that acmod_init_am tries in
 * turn (acmod.c:110-127) -- s2_semi_mgau_init, ptm_mgau_init, ms_mgau_init --
 * keeping their selection rules (one codebook -> s2_semi; <= 256 codebooks ->
 * ptm; otherwise / -senmgau given -> ms), so ps_init / ps_process_raw /
 * ps_get_hyp stay drop-in.  acmod_start_utt is interposed only to learn which
 * acmod_t owns an ms back-end (ms_mgau_init is not given the acmod).
 *
 * Scoring is utterance-batched: in batch mode acmod has the whole utterance's
 * features before the first acmod_score (acmod.c:513-540), so the first
 * frame_eval of an utterance ships every buffered frame to the GPU in one
 * call (b200_mgau_utt_begin); later calls are served from the device-resident
 * result with the caller's active list (b200_mgau_utt_frame).  In live mode
 * the same code degrades to however many frames are buffered (>= 1).
 *
 * No scoring arithmetic happens in this file.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <sphinxbase/ckd_alloc.h>
#include <sphinxbase/cmd_ln.h>
#include <sphinxbase/err.h>
#include <sphinxbase/feat.h>
#include <sphinxbase/logmath.h>

#include "acmod.h"
#include "bin_mdef.h"
#include "ms_mgau.h"
#include "ptm_mgau.h"
#include "s2_semi_mgau.h"

#include "../../include/b200sphinx.h"

typedef struct b200_ps_mgau_s {
    ps_mgau_t base;            /* must be first (acmod.h:113-116) */
    b200_mgau_t *gpu;
    cmd_ln_t *config;
    acmod_t *acmod;            /* NULL until known */
    int kind;                  /* 0 ms, 1 ptm, 2 s2_semi */
    int n_feat, n_density, n_mgau, featdim;
    int32 veclen[B200_MAX_STREAMS];
    double logbase;
    /* utterance cache bookkeeping */
    int cache_first, cache_n;  /* frames [first, first+n) are on the device */
    float *stage;              /* host staging for non-contiguous feature buffers */
    int stage_cap;
} b200_ps_mgau_t;

static int b200_frame_eval(ps_mgau_t *mg, int16 *senscr, uint8 *senone_active, int32 n_senone_active,
                           mfcc_t **feat, int32 frame, int32 compallsen);
static int b200_transform(ps_mgau_t *mg, ps_mllr_t *mllr);
static void b200_free(ps_mgau_t *mg);

static ps_mgaufuncs_t b200_funcs[3] = {
    { "b200_ms", &b200_frame_eval, &b200_transform, &b200_free },
    { "b200_ptm", &b200_frame_eval, &b200_transform, &b200_free },
    { "b200_semi", &b200_frame_eval, &b200_transform, &b200_free },
};

/* registry: back-ends created without an acmod (ms) learn it at acmod_start_utt */
#define MAX_REG 64
static b200_ps_mgau_t *g_reg[MAX_REG];
static int g_nreg;

static int enabled(void)
{
    const char *e = getenv("B200_PLUGIN_DISABLE");
    return !(e && e[0] == '1');
}

/* ---- parameter loading through the C ABI's own S3 readers / precompute ---- */
typedef struct { float *mean, *var, *det; int32 dims[4]; int32 veclen[64]; } gau_t;

static int load_gauden(cmd_ln_t *config, double logbase, gau_t *g, ps_mllr_t *mllr)
{
    const char *mf = cmd_ln_str_r(config, "-mean"), *vf = cmd_ln_str_r(config, "-var");
    int32 dv[4], vl2[64];
    int m, f, i;
    memset(g, 0, sizeof(*g));
    if (!mf || !vf) return -1;
    if (b200_s3_read_gauden(mf, g->dims, g->veclen, NULL) || b200_s3_read_gauden(vf, dv, vl2, NULL)) {
        E_ERROR("b200: %s\n", b200_last_error());
        return -1;
    }
    if (g->dims[0] != dv[0] || g->dims[1] != dv[1] || g->dims[2] != dv[2] || g->dims[1] > B200_MAX_STREAMS) {
        E_ERROR("b200: mean/variance dimensions differ or too many streams\n");
        return -1;
    }
    g->mean = ckd_calloc(g->dims[3], sizeof(float));
    g->var = ckd_calloc(g->dims[3], sizeof(float));
    g->det = ckd_calloc((size_t)g->dims[0] * g->dims[1] * g->dims[2], sizeof(float));
    if (b200_s3_read_gauden(mf, g->dims, g->veclen, g->mean) || b200_s3_read_gauden(vf, dv, vl2, g->var)) {
        E_ERROR("b200: %s\n", b200_last_error());
        return -1;
    }
    {
        int n_mgau = g->dims[0], n_feat = g->dims[1], n_density = g->dims[2], veclen = 0;
        for (f = 0; f < n_feat; ++f) veclen += g->veclen[f];
        for (m = 0; m < n_mgau; ++m) {
            int off = 0;
            for (f = 0; f < n_feat; ++f) {
                int len = g->veclen[f], d, l, k;
                float *mp = g->mean + (size_t)m * n_density * veclen + (size_t)n_density * off;
                float *vp = g->var + (size_t)m * n_density * veclen + (size_t)n_density * off;
                float *dp = g->det + ((size_t)m * n_feat + f) * n_density;
                if (mllr) {
                    /* gauden_mllr_transform (ms_gauden.c:551-605): one class,
                     * mean <- A mean + b (float64 accumulate), var <- var * h */
                    double *tmp = ckd_calloc(len, sizeof(double));
                    for (d = 0; d < n_density; ++d) {
                        for (l = 0; l < len; ++l) {
                            tmp[l] = 0.0;
                            for (k = 0; k < len; ++k)
                                tmp[l] += mllr->A[f][0][l][k] * mp[d * len + k];
                            tmp[l] += mllr->b[f][0][l];
                        }
                        for (l = 0; l < len; ++l) {
                            mp[d * len + l] = (float32)tmp[l];
                            vp[d * len + l] *= mllr->h[f][0][l];
                        }
                    }
                    ckd_free(tmp);
                }
                if (b200_gauden_precompute(vp, dp, n_density, len, cmd_ln_float32_r(config, "-varfloor"), logbase))
                    return -1;
                off += len;
            }
        }
        (void)i;
    }
    return 0;
}

static void free_gauden(gau_t *g)
{
    ckd_free(g->mean); ckd_free(g->var); ckd_free(g->det);
}

static b200_ps_mgau_t *wrap(b200_mgau_t *gpu, int kind, cmd_ln_t *config, acmod_t *acmod, const gau_t *g,
                            double logbase)
{
    b200_ps_mgau_t *s;
    int f;
    if (gpu == NULL) {
        E_ERROR("b200: GPU back-end creation failed: %s\n", b200_last_error());
        return NULL;
    }
    s = ckd_calloc(1, sizeof(*s));
    s->base.vt = &b200_funcs[kind];
    s->gpu = gpu; s->kind = kind; s->config = config; s->acmod = acmod; s->logbase = logbase;
    s->n_mgau = g->dims[0]; s->n_feat = g->dims[1]; s->n_density = g->dims[2];
    for (f = 0; f < s->n_feat; ++f) { s->veclen[f] = g->veclen[f]; s->featdim += g->veclen[f]; }
    s->cache_first = 0; s->cache_n = 0;
    /* the registry is how acmod_start_utt finds the back-end to invalidate its utterance cache:
     * a back-end that is not in it would re-serve the previous utterance's scores */
    if (g_nreg >= MAX_REG)
        E_FATAL("b200: more than %d live back-ends in one process; raise MAX_REG in plugin/b200_mgau.c\n", MAX_REG);
    g_reg[g_nreg++] = s;
    E_INFO("b200: %s back-end on GPU %d: %d codebooks x %d streams x %d densities, %d senones\n",
           s->base.vt->name, 0, s->n_mgau, s->n_feat, s->n_density, b200_mgau_n_sen(gpu));
    return s;
}

static void fill_cfg(b200_mgau_cfg_t *c, cmd_ln_t *config, const gau_t *g, int n_sen, double logbase)
{
    int f;
    memset(c, 0, sizeof(*c));
    c->n_mgau = g->dims[0]; c->n_feat = g->dims[1]; c->n_density = g->dims[2]; c->n_sen = n_sen;
    for (f = 0; f < c->n_feat; ++f) c->featlen[f] = g->veclen[f];
    c->topn = cmd_ln_int32_r(config, "-topn");
    c->aw = cmd_ln_int32_r(config, "-aw");
    c->ds_ratio = cmd_ln_int32_r(config, "-ds");
    c->logbase = logbase;
    {   /* -topn_beam as split_topn reads it (s2_semi_mgau.c:1204-1231): comma list of uint8,
         * streams past the end of the list get the largest value given */
        const char *str = cmd_ln_str_r(config, "-topn_beam");
        int i = 0, maxn = 0;
        while (str && *str && i < c->n_feat) {
            int v = (uint8)atoi(str);
            const char *comma = strchr(str, ',');
            c->topn_beam[i++] = v;
            if (v > maxn) maxn = v;
            if (!comma) break;
            str = comma + 1;
        }
        while (i < c->n_feat) c->topn_beam[i++] = maxn;
    }
    c->device = getenv("B200_DEVICE") ? atoi(getenv("B200_DEVICE")) : 0;
}

/* tied (ptm / s2_semi) mixture weights: sendump or mixture_weights */
static int load_tied_mixw(cmd_ln_t *config, const gau_t *g, int n_sen, double logbase, uint8 **rows,
                          int *n_clust, uint8 cb[16], int *n_sen_out)
{
    const char *sd = cmd_ln_str_r(config, "-sendump");
    if (sd) {
        int32 dims[5] = { g->dims[1], g->dims[2], n_sen, 0, 0 };
        if (b200_s3_read_sendump(sd, dims, NULL, NULL)) { E_ERROR("b200: %s\n", b200_last_error()); return -1; }
        if (dims[0] != g->dims[1] || dims[1] != g->dims[2] || dims[2] != n_sen) {
            E_ERROR("b200: sendump dimensions do not match the model\n");
            return -1;
        }
        *rows = ckd_calloc((size_t)dims[0] * dims[1] * dims[4], 1);
        dims[0] = g->dims[1]; dims[1] = g->dims[2]; dims[2] = n_sen;
        if (b200_s3_read_sendump(sd, dims, *rows, cb)) { E_ERROR("b200: %s\n", b200_last_error()); return -1; }
        *n_clust = dims[3];
        *n_sen_out = n_sen;
        return 0;
    }
    else {
        const char *mw = cmd_ln_str_r(config, "-mixw");
        int32 dims[4];
        float *raw;
        if (!mw || b200_s3_read_mixw(mw, dims, NULL)) { E_ERROR("b200: %s\n", b200_last_error()); return -1; }
        if (dims[1] != g->dims[1] || dims[2] != g->dims[2]) { E_ERROR("b200: mixw dimensions mismatch\n"); return -1; }
        raw = ckd_calloc(dims[3], sizeof(float));
        *rows = ckd_calloc(dims[3], 1);
        if (b200_s3_read_mixw(mw, dims, raw) ||
            b200_mixw_quantize_tied(raw, *rows, dims[0], dims[1], dims[2], cmd_ln_float32_r(config, "-mixwfloor"),
                                    logbase)) {
            E_ERROR("b200: %s\n", b200_last_error());
            ckd_free(raw);
            return -1;
        }
        ckd_free(raw);
        *n_clust = 0;
        *n_sen_out = dims[0];
        return 0;
    }
}

/* ------------------------------------------------ interposed constructors */
ps_mgau_t *
s2_semi_mgau_init(acmod_t *acmod)
{
    gau_t g;
    b200_mgau_cfg_t cfg;
    uint8 *rows = NULL, cb[16];
    int n_clust = 0, n_sen = 0, f;
    double lb = logmath_get_base(acmod->lmath);
    b200_ps_mgau_t *s = NULL;

    if (!enabled()) {
        ps_mgau_t *(*next)(acmod_t *) = dlsym(RTLD_NEXT, "s2_semi_mgau_init");
        return next ? next(acmod) : NULL;
    }
    if (load_gauden(acmod->config, lb, &g, NULL) < 0) goto out;
    if (g.dims[0] != 1) goto out;                              /* s2_semi_mgau.c:1264 */
    if (g.dims[1] != feat_dimension1(acmod->fcb)) { E_ERROR("Number of streams does not match\n"); goto out; }
    for (f = 0; f < g.dims[1]; ++f)
        if (g.veclen[f] != (int32)feat_dimension2(acmod->fcb, f)) { E_ERROR("Stream dimension mismatch\n"); goto out; }
    if (load_tied_mixw(acmod->config, &g, bin_mdef_n_sen(acmod->mdef), lb, &rows, &n_clust, cb, &n_sen) < 0) goto out;
    fill_cfg(&cfg, acmod->config, &g, n_sen, lb);
    s = wrap(b200_semi_create(&cfg, g.mean, g.var, g.det, rows, n_clust, cb), 2, acmod->config, acmod, &g, lb);
out:
    free_gauden(&g);
    ckd_free(rows);
    return (ps_mgau_t *)s;
}

ps_mgau_t *
ptm_mgau_init(acmod_t *acmod)
{
    gau_t g;
    b200_mgau_cfg_t cfg;
    uint8 *rows = NULL, *s2c = NULL, cb[16];
    int n_clust = 0, n_sen = 0, f, i;
    double lb = logmath_get_base(acmod->lmath);
    b200_ps_mgau_t *s = NULL;

    if (!enabled()) {
        ps_mgau_t *(*next)(acmod_t *) = dlsym(RTLD_NEXT, "ptm_mgau_init");
        return next ? next(acmod) : NULL;
    }
    if (load_gauden(acmod->config, lb, &g, NULL) < 0) goto out;
    if (g.dims[0] > 256) { E_INFO("Number of codebooks exceeds 256: %d\n", g.dims[0]); goto out; }   /* ptm_mgau.c:799 */
    /* Returning NULL here would make acmod_init_am fall through to the ms back-end (other scoring
     * semantics, or a failure when only a sendump exists): refuse loudly instead. */
    if (cmd_ln_int32_r(acmod->config, "-ds") > 1)
        E_FATAL("b200: -ds %d with a ptm model: the reference's ptm back-end leaves un-normalised scores in its "
                "lists on skipped frames (ptm_mgau.c:247-248) and indexes its log-add table with them -- undefined "
                "output, not reproduced; use -ds 1 or B200_PLUGIN_DISABLE=1\n", (int)cmd_ln_int32_r(acmod->config, "-ds"));
    if (g.dims[1] != feat_dimension1(acmod->fcb)) { E_ERROR("Number of streams does not match\n"); goto out; }
    for (f = 0; f < g.dims[1]; ++f)
        if (g.veclen[f] != (int32)feat_dimension2(acmod->fcb, f)) { E_ERROR("Stream dimension mismatch\n"); goto out; }
    if (load_tied_mixw(acmod->config, &g, bin_mdef_n_sen(acmod->mdef), lb, &rows, &n_clust, cb, &n_sen) < 0) goto out;
    s2c = ckd_calloc(n_sen, 1);
    for (i = 0; i < n_sen; ++i)
        s2c[i] = (uint8)bin_mdef_sen2cimap(acmod->mdef, i);   /* ptm_mgau.c:834-836 */
    fill_cfg(&cfg, acmod->config, &g, n_sen, lb);
    s = wrap(b200_ptm_create(&cfg, g.mean, g.var, g.det, rows, n_clust, cb, s2c), 1, acmod->config, acmod, &g, lb);
out:
    free_gauden(&g);
    ckd_free(rows);
    ckd_free(s2c);
    return (ps_mgau_t *)s;
}

ps_mgau_t *
ms_mgau_init(cmd_ln_t *config, logmath_t *lmath, bin_mdef_t *mdef)
{
    const char *senmgau = cmd_ln_str_r(config, "-senmgau");
    uint8 *s2c = NULL;
    double lb = logmath_get_base(lmath);
    b200_ps_mgau_t *s = NULL;
    b200_mgau_t *gpu;
    gau_t g;
    int i, n_sen;

    if (!enabled()) {
        ps_mgau_t *(*next)(cmd_ln_t *, logmath_t *, bin_mdef_t *) = dlsym(RTLD_NEXT, "ms_mgau_init");
        return next ? next(config, lmath, mdef) : NULL;
    }
    memset(&g, 0, sizeof(g));
    if (b200_s3_read_gauden(cmd_ln_str_r(config, "-mean"), g.dims, g.veclen, NULL)) {
        E_ERROR("b200: %s\n", b200_last_error());
        return NULL;
    }
    /* senone -> codebook map selection of ms_senone.c:297-340 */
    if (senmgau == NULL && mdef != NULL && g.dims[0] > 1 && g.dims[0] == bin_mdef_n_ciphone(mdef))
        senmgau = ".ptm.";
    if (senmgau && strcmp(senmgau, ".ptm.") == 0) {
        if (mdef == NULL) { E_ERROR("b200: .ptm. mapping needs the model definition\n"); return NULL; }
        n_sen = bin_mdef_n_sen(mdef);
        s2c = ckd_calloc(n_sen, 1);
        for (i = 0; i < n_sen; ++i) s2c[i] = (uint8)bin_mdef_sen2cimap(mdef, i);
    }
    gpu = b200_ms_load(cmd_ln_str_r(config, "-mean"), cmd_ln_str_r(config, "-var"), cmd_ln_str_r(config, "-mixw"),
                       senmgau, s2c, cmd_ln_float32_r(config, "-varfloor"), cmd_ln_float32_r(config, "-mixwfloor"),
                       cmd_ln_int32_r(config, "-topn"), cmd_ln_int32_r(config, "-aw"), lb,
                       getenv("B200_DEVICE") ? atoi(getenv("B200_DEVICE")) : 0);
    ckd_free(s2c);
    /* B200_MS_PATH=0 forces the exact CUDA-core kernels, 1 the tensor-core GEMM */
    if (gpu && getenv("B200_MS_PATH") && b200_mgau_set_path(gpu, atoi(getenv("B200_MS_PATH"))))
        E_WARN("b200: %s\n", b200_last_error());
    s = wrap(gpu, 0, config, NULL, &g, lb);
    return (ps_mgau_t *)s;
}

/* Learn the owning acmod of back-ends that were created without one. */
int
acmod_start_utt(acmod_t *acmod)
{
    static int (*next)(acmod_t *);
    int i;
    if (!next) next = dlsym(RTLD_NEXT, "acmod_start_utt");
    for (i = 0; i < g_nreg; ++i)
        if ((ps_mgau_t *)g_reg[i] == acmod->mgau) {
            g_reg[i]->acmod = acmod;
            g_reg[i]->cache_n = 0;
        }
    return next(acmod);
}

/* ------------------------------------------------------------- vtable */
static int
b200_frame_eval(ps_mgau_t *mg, int16 *senscr, uint8 *senone_active, int32 n_senone_active, mfcc_t **feat,
                int32 frame, int32 compallsen)
{
    b200_ps_mgau_t *s = (b200_ps_mgau_t *)mg;
    acmod_t *a = s->acmod;
    int rc;

    if (frame < s->cache_first || frame >= s->cache_first + s->cache_n) {
        /* Not on the device yet: ship this frame and every frame buffered behind it.
         * feat_buf is one contiguous [n_feat_alloc][sum of stream lengths] block
         * (sphinxbase feat.c:519-549) used as a ring in live mode. */
        int n = 1;
        const float *src = (const float *)feat[0];
        if (a && a->feat_buf && feat == a->feat_buf[a->feat_outidx] && a->n_feat_frame > 0 &&
            (a->n_feat_alloc < 2 || a->feat_buf[1][0] - a->feat_buf[0][0] == s->featdim)) {
            /* (with -lda the buffer rows are wider than the model's streams:
             * then frames are not contiguous and we go one frame at a time) */
            n = a->n_feat_frame;
            if (a->feat_outidx + n > a->n_feat_alloc) n = a->n_feat_alloc - a->feat_outidx;   /* up to the wrap */
        }
        else if (s->n_feat > 1) {
            /* streams of a foreign feature array may not be contiguous: stage one frame */
            int f, off = 0;
            if (s->stage_cap < s->featdim) { s->stage = ckd_realloc(s->stage, s->featdim * sizeof(float)); s->stage_cap = s->featdim; }
            for (f = 0; f < s->n_feat; ++f) { memcpy(s->stage + off, feat[f], s->veclen[f] * sizeof(float)); off += s->veclen[f]; }
            src = s->stage;
        }
        rc = b200_mgau_utt_begin_at(s->gpu, src, n, frame);
        if (rc) { E_ERROR("b200: utt_begin failed: %s\n", b200_last_error()); return -1; }
        s->cache_first = frame;
        s->cache_n = n;
    }
    rc = b200_mgau_utt_frame(s->gpu, senscr, senone_active, n_senone_active, frame - s->cache_first, compallsen);
    if (rc) { E_ERROR("b200: frame_eval failed: %s\n", b200_last_error()); return -1; }
    return 0;
}

static int
b200_transform(ps_mgau_t *mg, ps_mllr_t *mllr)
{
    b200_ps_mgau_t *s = (b200_ps_mgau_t *)mg;
    gau_t g;
    int rc = -1;
    if (s->base.vt == &b200_funcs[2] && !getenv("B200_SEMI_MLLR")) {
        /* The reference's s2_semi back-end transforms s->g (s2_semi_mgau.c:1338-1343) but scores
         * with the aliases s->means / s->vars / s->dets taken at init (:1267-1269), which
         * gauden_mllr_transform leaves pointing at the untransformed arrays: -mllr has no
         * effect on its scores.  Same here, so that results stay identical; B200_SEMI_MLLR=1
         * applies the transform that was intended. */
        E_INFO("b200: s2_semi ignores the MLLR transform, as the reference does\n");
        return 0;
    }
    if (load_gauden(s->config, s->logbase, &g, mllr) == 0)
        rc = b200_mgau_update_params(s->gpu, g.mean, g.var, g.det);
    if (rc) E_ERROR("b200: transform failed: %s\n", b200_last_error());
    free_gauden(&g);
    s->cache_n = 0;
    return rc ? -1 : 0;
}

static void
b200_free(ps_mgau_t *mg)
{
    b200_ps_mgau_t *s = (b200_ps_mgau_t *)mg;
    int i;
    if (!s) return;
    for (i = 0; i < g_nreg; ++i)
        if (g_reg[i] == s) { g_reg[i] = g_reg[--g_nreg]; break; }
    b200_mgau_free(s->gpu);
    ckd_free(s->stage);
    ckd_free(s);
}
