/* b200_s3_mgau.c -- the reference-side binding of the sphinx3 boundary:
 * srch_funcs_t.gmm_compute_lv2 -> s3_cd_gmm_compute_sen -> approx_cont_mgau_frame_eval
 * (sphinx3/src/libs3decoder/libsearch/gmm_wrap.c:103-171, libam/approx_cont_mgau.c:433-616).
 * Preloaded next to an UNMODIFIED libs3decoder.so this file interposes
 *
 *   mgau_init                     (libam/cont_mgau.c:900-958, called from kbcore_init,
 *                                  libsearch/kbcore.c:301) -- only to learn the model files and floors,
 *                                  then runs the reference's own function;
 *   subvq_init / gs_read          (libam/subvq.c:206, libam/gs.c:156) -- the same for -subvq / -gs;
 *   approx_cont_mgau_frame_eval   the per-frame CD scoring call -- served by b200_s3_frame_eval: the CI pass,
 *                                  the CI / dynamic beam, -ds, sub-VQ or Gaussian-selector shortlists, back-off
 *                                  and normalisation all happen on the GPU; ascr_t.senscr / sen_active /
 *                                  rec_sen_active are updated in place and the frame's best score (senscale) is
 *                                  returned, exactly as the reference function does.
 *
 * approx_cont_mgau_ci_eval (the look-ahead CI pass of gmm_compute_lv1) is left to the reference: its
 * scores feed the phoneme look-ahead only, the CI scores that go into senscr are re-derived on the GPU.
 * Settings the GPU path does not implement (-cond_ds, -svq4svq, full covariances) fall through to the
 * reference's own function.  B200_S3_PLUGIN_DISABLE=1 forwards everything.  No scoring arithmetic
 * happens in this file. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <sphinxbase/ckd_alloc.h>
#include <sphinxbase/err.h>
#include <sphinxbase/logmath.h>
#include <sphinxbase/profile.h>

#include "s3types.h"
#include "cont_mgau.h"
#include "approx_cont_mgau.h"
#include "fast_algo_struct.h"
#include "ascr.h"
#include "mdef.h"
#include "subvq.h"
#include "gs.h"

#include "../../include/b200sphinx.h"

static struct {
    char *mean, *var, *mixw, *svq_file, *gs_file;
    double varfloor, mixwfloor, svq_varfloor;
    int svq_max_sv, svq_vqeval;
    mgau_model_t *g;            /* the model the captured files belong to */
    b200_s3mgau_t *gpu;
    int failed;
} S;

static int disabled(void) { const char *e = getenv("B200_S3_PLUGIN_DISABLE"); return e && atoi(e); }
static char *dupstr(const char *s) { return s ? ckd_salloc(s) : NULL; }

mgau_model_t *
mgau_init(const char *meanfile, const char *varfile, float64 varfloor, const char *mixwfile, float64 mixwfloor,
          int32 precomp, const char *senmgau, int32 comp_type, logmath_t *logmath)
{
    static mgau_model_t *(*real)(const char *, const char *, float64, const char *, float64, int32, const char *, int32,
                                 logmath_t *);
    mgau_model_t *g;
    if (!real) real = dlsym(RTLD_NEXT, "mgau_init");
    g = real(meanfile, varfile, varfloor, mixwfile, mixwfloor, precomp, senmgau, comp_type, logmath);
    if (g && !disabled() && precomp && comp_type == MIX_INT_FLOAT_COMP) {
        ckd_free(S.mean); ckd_free(S.var); ckd_free(S.mixw);
        S.mean = dupstr(meanfile); S.var = dupstr(varfile); S.mixw = dupstr(mixwfile);
        S.varfloor = varfloor; S.mixwfloor = mixwfloor; S.g = g;
        if (S.gpu) { b200_s3_free(S.gpu); S.gpu = NULL; }
        S.failed = 0;
    }
    return g;
}

subvq_t *
subvq_init(const char *file, float64 varfloor, int32 max_sv, mgau_model_t *g, cmd_ln_t *config, logmath_t *logmath)
{
    static subvq_t *(*real)(const char *, float64, int32, mgau_model_t *, cmd_ln_t *, logmath_t *);
    subvq_t *v;
    if (!real) real = dlsym(RTLD_NEXT, "subvq_init");
    v = real(file, varfloor, max_sv, g, config, logmath);
    if (v && !disabled()) {
        ckd_free(S.svq_file);
        S.svq_file = dupstr(file); S.svq_varfloor = varfloor; S.svq_max_sv = max_sv;
        S.svq_vqeval = cmd_ln_int32_r(config, "-vqeval");
    }
    return v;
}

gs_t *
gs_read(const char *file, logmath_t *logmath)
{
    static gs_t *(*real)(const char *, logmath_t *);
    gs_t *v;
    if (!real) real = dlsym(RTLD_NEXT, "gs_read");
    v = real(file, logmath);
    if (v && !disabled()) { ckd_free(S.gs_file); S.gs_file = dupstr(file); }
    return v;
}

static b200_s3mgau_t *
gpu_model(mdef_t *mdef, subvq_t *svq, gs_t *gs, mgau_model_t *g, logmath_t *logmath)
{
    int32 *cd2ci;
    int i, dev = getenv("B200_DEVICE") ? atoi(getenv("B200_DEVICE")) : 0;
    if (S.gpu || S.failed) return S.gpu;
    S.failed = 1;                                   /* until everything below has worked */
    if (g != S.g || !S.mean) { E_WARN("b200: this mgau_model_t did not come through mgau_init; using the reference\n"); return NULL; }
    cd2ci = ckd_calloc(mdef->n_sen, sizeof(int32));
    for (i = 0; i < mdef->n_sen; ++i) cd2ci[i] = mdef->cd2cisen[i];
    S.gpu = b200_s3_load(S.mean, S.var, S.mixw, S.varfloor, S.mixwfloor, logmath_get_base(logmath), cd2ci, mdef->n_ci_sen, dev);
    ckd_free(cd2ci);
    if (!S.gpu) { E_ERROR("b200: %s; using the reference\n", b200_last_error()); return NULL; }
    if (svq && (!S.svq_file || b200_s3_set_subvq(S.gpu, S.svq_file, S.svq_varfloor, S.svq_max_sv, S.svq_vqeval, 1.0) != 0)) {
        E_ERROR("b200: sub-VQ model: %s; using the reference\n", b200_last_error());
        b200_s3_free(S.gpu); S.gpu = NULL; return NULL;
    }
    if (gs && (!S.gs_file || b200_s3_set_gs(S.gpu, S.gs_file) != 0)) {
        E_ERROR("b200: Gaussian selector: %s; using the reference\n", b200_last_error());
        b200_s3_free(S.gpu); S.gpu = NULL; return NULL;
    }
    S.failed = 0;
    E_INFO("b200: approx_cont_mgau_frame_eval is served by libb200sphinx (%d senones, %d CI)\n", mdef->n_sen, mdef->n_ci_sen);
    return S.gpu;
}

int32
approx_cont_mgau_frame_eval(mdef_t *mdef, subvq_t *svq, gs_t *gs, mgau_model_t *g, fast_gmm_t *fastgmm, ascr_t *a,
                            float32 *feat, int32 frame, int32 *cache_ci_senscr, ptmr_t *tm_ovrhd, logmath_t *logmath)
{
    static int32 (*real)(mdef_t *, subvq_t *, gs_t *, mgau_model_t *, fast_gmm_t *, ascr_t *, float32 *, int32, int32 *,
                         ptmr_t *, logmath_t *);
    b200_s3mgau_t *m;
    int32 best = 0;
    if (!real) real = dlsym(RTLD_NEXT, "approx_cont_mgau_frame_eval");
    if (disabled() || fastgmm->downs->cond_ds > 0 || fastgmm->downs->dist_ds > 0 || (svq && fastgmm->svq4svq) ||
        (gs && !fastgmm->gs4gs) || !(m = gpu_model(mdef, svq, gs, g, logmath)))
        return real(mdef, svq, gs, g, fastgmm, a, feat, frame, cache_ci_senscr, tm_ovrhd, logmath);
    if (frame == 0) b200_s3_utt_reset(m);           /* srch_time_switch_tree.c:484-490 / srch_flat_fwd.c: per-utterance reset */
    b200_s3_set_fast_log(m, fastgmm->gmms->ci_pbeam, fastgmm->gmms->max_cd, fastgmm->downs->ds_ratio,
                         fastgmm->gmms->tighten_factor, fastgmm->gaus->subvqbeam);
    if (b200_s3_frame_eval(m, feat, frame, a->sen_active, a->senscr, &best) != 0)
        E_FATAL("b200_s3_frame_eval: %s\n", b200_last_error());
    memcpy(a->rec_sen_active, a->sen_active, mdef->n_sen);
    g->frm_sen_eval = 0; g->frm_gau_eval = 0;       /* (statistics of the host loop; nothing was evaluated there) */
    return best;
}
