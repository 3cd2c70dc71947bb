// feat_kernels.cu -- the feature stage next to the scorer: full-utterance
// cepstra -> 1s_c_d_dd dynamic features with -cmn current, on the device, so a
// batch of utterances uploads 13 floats per frame instead of 39 and the
// scorer's input never leaves HBM.
//
// Reference (SB = sphinxbase/src/libsphinxbase):
//   feat_s2mfc2feat_block_utt  SB/feat/feat.c:1241-1265  pad with 3 copies of the
//                              first / last frame, THEN normalise (the padding
//                              is part of the mean)
//   cmn                        SB/feat/cmn.c:150-186     float32 running sum in
//                              frame order, mean = sum / n, subtract
//   feat_1s_c_d_dd_cep2feat    SB/feat/feat.c:726-769    c | c[+2]-c[-2] |
//                              (c[+3]-c[-1]) - (c[+1]-c[-3])
// Bit-exact: every float operation is the reference's, in its order.  The
// hub4wsj_sc_8k models use the same features split into three 13-dim streams
// (-svspec 0-12/13-25/26-38), which is the same memory layout.
#include "dev_common.cuh"

using namespace b200;

namespace {

constexpr int kWin = 3;   // feat_window_size of 1s_c_d_dd (FEAT_DCEP_WIN + 1)

// one thread per (utterance, cepstral dimension): the sequential float32 sum
__global__ void feat_cmn_mean_kernel(const float *__restrict__ cep, const int32_t *__restrict__ utt_off, int n_utt,
                                     int cepsize, float *__restrict__ mean) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_utt * cepsize) return;
    const int u = id / cepsize, i = id % cepsize;
    const int t0 = utt_off[u], T = utt_off[u + 1] - t0;
    if (T <= 0) { mean[id] = 0.f; return; }
    const float *c = cep + (size_t)t0 * cepsize + i;
    const float first = c[0], last = c[(size_t)(T - 1) * cepsize];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) s = __fadd_rn(s, first);
    for (int t = 0; t < T; ++t) s = __fadd_rn(s, c[(size_t)t * cepsize]);
#pragma unroll
    for (int k = 0; k < kWin; ++k) s = __fadd_rn(s, last);
    mean[id] = __fdiv_rn(s, (float)(T + 2 * kWin));
}

// one thread per (frame, cepstral dimension)
__global__ void feat_dyn_kernel(const float *__restrict__ cep, const int32_t *__restrict__ utt_off, int n_utt,
                                int T_total, int cepsize, const float *__restrict__ mean /* null: no cmn */,
                                float *__restrict__ feat) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)T_total * cepsize) return;
    const int t = (int)(id / cepsize), i = (int)(id % cepsize);
    // utterance of frame t (binary search over the offsets)
    int lo = 0, hi = n_utt - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (utt_off[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int t0 = utt_off[lo], t1 = utt_off[lo + 1] - 1;
    const float mu = mean ? mean[lo * cepsize + i] : 0.f;
    auto at = [&](int k) {
        const int tt = min(max(t + k, t0), t1);
        const float v = cep[(size_t)tt * cepsize + i];
        return mean ? __fsub_rn(v, mu) : v;
    };
    float *f = feat + (size_t)t * 3 * cepsize;
    f[i] = at(0);
    f[cepsize + i] = __fsub_rn(at(2), at(-2));
    f[2 * cepsize + i] = __fsub_rn(__fsub_rn(at(3), at(-1)), __fsub_rn(at(1), at(-3)));
}

}  // namespace

extern "C" {

int b200_feat_1s_c_d_dd_dev(const float *d_cep, const int32_t *d_utt_off, int n_utt, int T_total, int cepsize,
                            int cmn, float *d_mean_scratch, float *d_feat, void *stream) {
    if (!d_cep || !d_utt_off || !d_feat || n_utt < 1 || T_total < 0 || cepsize < 1 || (cmn && !d_mean_scratch)) {
        set_error("b200_feat_1s_c_d_dd_dev: bad argument");
        return B200_ERR_ARG;
    }
    if (cmn != 0 && cmn != 1) { set_error("only -cmn none|current are computed on the device"); return B200_ERR_UNSUP; }
    if (T_total == 0) return B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (cmn) {
        feat_cmn_mean_kernel<<<(n_utt * cepsize + 127) / 128, 128, 0, st>>>(d_cep, d_utt_off, n_utt, cepsize, d_mean_scratch);
        B200_LAUNCH_CHECK();
    }
    const long long n = (long long)T_total * cepsize;
    feat_dyn_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_cep, d_utt_off, n_utt, T_total, cepsize,
                                                                cmn ? d_mean_scratch : nullptr, d_feat);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int b200_feat_1s_c_d_dd_host(const float *cep, const int32_t *utt_off, int n_utt, int cepsize, int cmn, float *feat,
                             int device) {
    if (!cep || !utt_off || !feat || n_utt < 1 || cepsize < 1) { set_error("b200_feat_1s_c_d_dd_host: bad argument"); return B200_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: libb200sphinx has no CPU fallback"); return B200_ERR_CUDA; }
    B200_CUDA_OK(cudaSetDevice(device));
    const int T = utt_off[n_utt];
    if (T <= 0) return B200_OK;
    float *d_cep = nullptr, *d_feat = nullptr, *d_mean = nullptr; int32_t *d_off = nullptr;
    int rc = B200_OK;
    if (cudaMalloc((void **)&d_cep, (size_t)T * cepsize * 4) != cudaSuccess ||
        cudaMalloc((void **)&d_feat, (size_t)T * cepsize * 12) != cudaSuccess ||
        cudaMalloc((void **)&d_mean, (size_t)n_utt * cepsize * 4) != cudaSuccess ||
        cudaMalloc((void **)&d_off, (size_t)(n_utt + 1) * 4) != cudaSuccess ||
        cudaMemcpy(d_cep, cep, (size_t)T * cepsize * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d_off, utt_off, (size_t)(n_utt + 1) * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("feature stage: device allocation/upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = B200_ERR_CUDA;
    }
    if (!rc) rc = b200_feat_1s_c_d_dd_dev(d_cep, d_off, n_utt, T, cepsize, cmn, d_mean, d_feat, nullptr);
    if (!rc && cudaMemcpy(feat, d_feat, (size_t)T * cepsize * 12, cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("feature stage: download failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = B200_ERR_CUDA;
    }
    cudaFree(d_cep); cudaFree(d_feat); cudaFree(d_mean); cudaFree(d_off);
    return rc;
}

}  // extern "C"
