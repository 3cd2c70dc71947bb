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
    int topn_beam[B200_MAX_STREAMS];   // s2_semi -topn_beam per stream, 0 = off
};

// Integer top-N list of the tied back-ends (shared by the exact kernel and the
// tensor-core path's fallback).
template <int N>
struct TopI {  // ptm / s2_semi: int32 scores, codewords
    int32_t s[N];
    int cw[N];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < N; ++j) { s[j] = kWorstDistI; cw[j] = j; }
    }
    // eval_topn's insertion_sort_topn (ptm_mgau.c:82-96): entry i receives
    // score d and bubbles up past strictly smaller scores.
    __device__ __forceinline__ void seed(int i, int32_t d) {
        int c = cw[i];
#pragma unroll
        for (int j = N - 1; j >= 0; --j) {
            if (j > i) continue;
            if (j > 0 && d > s[j - 1]) { s[j] = s[j - 1]; cw[j] = cw[j - 1]; }
            else { s[j] = d; cw[j] = c; break; }
        }
    }
    __device__ __forceinline__ bool has(int c) const {
        bool h = false;
#pragma unroll
        for (int j = 0; j < N; ++j) h |= (cw[j] == c);
        return h;
    }
    // eval_cb's insertion (ptm_mgau.c:146-157): ahead of equal scores.
    __device__ __forceinline__ void insert(int32_t d, int c) {
#pragma unroll
        for (int j = N - 1; j >= 0; --j) {
            if (j > 0 && d >= s[j - 1]) { s[j] = s[j - 1]; cw[j] = cw[j - 1]; }
            else { s[j] = d; cw[j] = c; break; }
        }
    }
};

int gmm_launch_topn(const GmmDev &g, int mode, const float *d_feat, int T, int t0, int tn, int2 *lists,
                    int16_t *raw, int fused, cudaStream_t st);
int gmm_launch_tied_ds(const GmmDev &g, const float *d_feat, int t0, int tn, int frame0, int ds_ratio, int2 *lists,
                       const int2 *carry, cudaStream_t st);
int gmm_launch_ms_senone(const GmmDev &g, const int2 *lists, int T, int t0, int tn, int16_t *raw,
                         cudaStream_t st);
int gmm_launch_normalize(int16_t *scr, int T, int n_sen, cudaStream_t st);
int gmm_launch_tied_senone(const GmmDev &g, const int2 *lists, int T, int t0, int tn, int semi,
                           const uint8_t *d_active, int n_active, int16_t *out, cudaStream_t st);

size_t gmm_tied_smem(const GmmDev &g, int n_active);
int gmm_launch_ms_active_normalize(const int16_t *raw, const uint8_t *d_active, int n_active,
                                   int16_t *out, cudaStream_t st);

// mahal_tc.cu: tensor-core codebook stage of the tied back-ends (ptm / s2_semi):
// lists [tn][n_mgau*n_feat][topn] {codeword, (int32)score} for frames
// [t0, t0+tn) of d_feat, bit-identical to gmm_launch_topn(mode 1|2).
struct TcTied;
TcTied *tc_tied_create(const GmmDev &g, int mode, const float *h_mean, const float *h_var, const float *h_det,
                       int device);
void tc_tied_free(TcTied *p);
int tc_tied_lists(TcTied *p, const GmmDev &g, const float *d_feat, int t0, int tn, int2 *lists, cudaStream_t st);
// statistics of the last call: {lists, lists sent to the exact fallback, largest
// |GEMM - exact| distance seen on a best candidate (raw log units, 0 if <= 4)}
void tc_tied_stats(TcTied *p, long long out[3]);

// mahal_tc.cu: tensor-core path for single-stream .cont. models.
struct TcPlan;
TcPlan *tc_plan_create(const GmmDev &g, const float *h_mean, const float *h_var, const float *h_det,
                       const uint8_t *h_mixw_sfc /* [sen][feat][cw] */, int device);
void tc_plan_free(TcPlan *p);
bool tc_shape_supported(const GmmDev &g);
// Stage 1: operand prep + scoring kernel -> tile-major raw scores inside the
// plan (ev_prep, if given, is recorded between the two kernels).  Stage 2:
// transpose to row-major d_out[T][n_sen], minus the frame best if asked.
int tc_score_raw(TcPlan *p, const float *d_feat, int T, cudaStream_t st, cudaEvent_t *ev_prep, int *T_pad_out, cudaEvent_t *ev_fix = nullptr);
int tc_last_format(TcPlan *p);   // 1 all tiles fp16, 0 all TF32, 2 mixed (synchronises)
// {pairs scored, pairs on the hard path, queue A items, queue B items, overflow flag,
//  largest |GEMM - reference| distance seen} of the last tc_score_raw (synchronises)
int tc_last_stats(TcPlan *p, long long out[7]);
int tc_finish(TcPlan *p, int T, int T_pad, int subtract_best, int16_t *d_out, cudaStream_t st);

}  // namespace b200
