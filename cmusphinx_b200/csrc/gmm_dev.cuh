// gmm_dev.cuh -- device-side view of one acoustic model (passed by value to
// kernels) and the launcher prototypes shared by gmm_exact.cu / mahal_tc.cu /
// abi.cu.
#pragma once
#include "dev_common.cuh"

namespace b200 {

struct GmmDev {
    int n_mgau, n_feat, n_density, n_sen, topn, aw;
    int veclen, maxlen;
    int featlen[B200_MAX_STREAMS], featoff[B200_MAX_STREAMS];
    const float *mean;          // [mgau][feat][density][featlen f] (reference layout)
    const float *var;           // precomputed scaled 1/(2 var), same layout
    const float *det;           // [mgau][feat][density]
    // ms: uint8 [feat][cw][sen] (transposed for coalescing over senones)
    // tied: [feat][cw][row_bytes] raw rows (8-bit, or 4-bit packed when n_clust)
    const uint8_t *mixw_t;
    const uint32_t *sen2mgau;   // ms
    const uint8_t *sen2cb;      // ptm
    int n_clust, row_bytes;
    uint8_t mixw_cb[16];
    uint8_t logadd[256];        // logmath_init(base, 10, 1) byte table
};

int gmm_launch_topn(const GmmDev &g, int mode, const float *d_feat, int T, int t0, int tn, int2 *lists,
                    int16_t *raw, int fused, cudaStream_t st);
int gmm_launch_ms_senone(const GmmDev &g, const int2 *lists, int T, int t0, int tn, int16_t *raw,
                         cudaStream_t st);
int gmm_launch_normalize(int16_t *scr, int T, int n_sen, cudaStream_t st);
int gmm_launch_tied_senone(const GmmDev &g, const int2 *lists, int T, int t0, int tn, int semi,
                           const uint8_t *d_active, int n_active, int16_t *out, cudaStream_t st);

size_t gmm_tied_smem(const GmmDev &g, int n_active);
int gmm_launch_ms_active_normalize(const int16_t *raw, const uint8_t *d_active, int n_active,
                                   int16_t *out, cudaStream_t st);

// mahal_tc.cu: tensor-core path for single-stream .cont. models.
struct TcPlan;
TcPlan *tc_plan_create(const GmmDev &g, const float *h_mean, const float *h_var, const float *h_det,
                       const uint8_t *h_mixw_sfc /* [sen][feat][cw] */, int device);
void tc_plan_free(TcPlan *p);
bool tc_shape_supported(const GmmDev &g);
// Stage 1: operand prep + scoring kernel -> tile-major raw scores inside the
// plan (ev_prep, if given, is recorded between the two kernels).  Stage 2:
// transpose to row-major d_out[T][n_sen], minus the frame best if asked.
int tc_score_raw(TcPlan *p, const float *d_feat, int T, cudaStream_t st, cudaEvent_t *ev_prep, int *T_pad_out);
int tc_finish(TcPlan *p, int T, int T_pad, int subtract_best, int16_t *d_out, cudaStream_t st);

}  // namespace b200
