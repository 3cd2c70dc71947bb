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


// ---------------------------------------------------------------- general form
// Output element e of a frame = one of three shapes over the normalised cepstra
// (feat.c:559-849): copy c[i], difference c[+a][i] - c[-a][i], or the second
// difference (c[+3][i] - c[-1][i]) - (c[+1][i] - c[-3][i]).
constexpr int kMaxFeatLen = 256;
struct FeatMap {
    uint16_t e[kMaxFeatLen];   // kind << 12 | a << 8 | i
    int n;
};

// one thread per (utterance, cepstral dimension): mean, 1/std and (dimension 0)
// the AGC maximum, each the reference's sequential float32 loop over the PADDED
// utterance (win copies of the first and of the last frame, feat.c:1253-1259).
__global__ void feat_stats_kernel(const float *__restrict__ cep, const int32_t *__restrict__ utt_off, int n_utt,
                                  int cepsize, int win, int cmn, int varnorm, int agc, float *__restrict__ mean,
                                  float *__restrict__ scale, float *__restrict__ agcmax) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_utt * cepsize) return;
    const int u = id / cepsize, i = id % cepsize;
    const int t0 = utt_off[u], T = utt_off[u + 1] - t0;
    float mu = 0.f, sc = 1.f;
    if (T > 0) {
        const float *c = cep + (size_t)t0 * cepsize + i;
        const float first = c[0], last = c[(size_t)(T - 1) * cepsize];
        const int n = T + 2 * win;
        if (cmn) {
            float s = 0.f;
            for (int k = 0; k < win; ++k) s = __fadd_rn(s, first);
            for (int t = 0; t < T; ++t) s = __fadd_rn(s, c[(size_t)t * cepsize]);
            for (int k = 0; k < win; ++k) s = __fadd_rn(s, last);
            mu = __fdiv_rn(s, (float)n);
            if (varnorm) {
                float v = 0.f, d;
                d = __fsub_rn(first, mu);
                for (int k = 0; k < win; ++k) v = __fadd_rn(v, __fmul_rn(d, d));
                for (int t = 0; t < T; ++t) {
                    d = __fsub_rn(c[(size_t)t * cepsize], mu);
                    v = __fadd_rn(v, __fmul_rn(d, d));
                }
                d = __fsub_rn(last, mu);
                for (int k = 0; k < win; ++k) v = __fadd_rn(v, __fmul_rn(d, d));
                sc = (float)__dsqrt_rn(__ddiv_rn((double)n, (double)v));
            }
        }
        if (agc && i == 0) {
            auto norm = [&](float x) {
                if (cmn) { x = __fsub_rn(x, mu); if (varnorm) x = __fmul_rn(x, sc); }
                return x;
            };
            float m = norm(c[0]);
            for (int t = 1; t < T; ++t) {
                const float x = norm(c[(size_t)t * cepsize]);
                if (x > m) m = x;
            }
            agcmax[u] = m;
        }
    }
    mean[id] = mu;
    scale[id] = sc;
}

// one thread per (frame, output element)
__global__ void feat_map_kernel(const float *__restrict__ cep, const int32_t *__restrict__ utt_off, int n_utt,
                                int T_total, int cepsize, int cmn, int varnorm, int agc,
                                const float *__restrict__ mean, const float *__restrict__ scale,
                                const float *__restrict__ agcmax, const FeatMap map, float *__restrict__ out) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)T_total * map.n) return;
    const int t = (int)(id / map.n), e = (int)(id % map.n);
    int lo = 0, hi = n_utt - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (utt_off[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int t0 = utt_off[lo], t1 = utt_off[lo + 1] - 1;
    const int code = map.e[e], kind = code >> 12, a = (code >> 8) & 15, i = code & 255;
    const float mu = cmn ? mean[lo * cepsize + i] : 0.f;
    const float sc = (cmn && varnorm) ? scale[lo * cepsize + i] : 1.f;
    const float am = (agc && i == 0) ? agcmax[lo] : 0.f;
    auto at = [&](int k) {
        const int tt = min(max(t + k, t0), t1);
        float v = cep[(size_t)tt * cepsize + i];
        if (cmn) { v = __fsub_rn(v, mu); if (varnorm) v = __fmul_rn(v, sc); }
        if (agc && i == 0) v = __fsub_rn(v, am);
        return v;
    };
    float r;
    if (kind == 0) r = at(0);
    else if (kind == 3) r = at(a - 8);            // plain copy of a neighbouring frame (offset biased by 8)
    else if (kind == 1) r = __fsub_rn(at(a), at(-a));
    else r = __fsub_rn(__fsub_rn(at(3), at(-1)), __fsub_rn(at(1), at(-3)));
    out[id] = r;
}

// feat_lda_transform (lda.c:141-160): out[j] = sum_k in[k] * lda[j][k], float32,
// product and sum rounded separately, k ascending.  One thread per (frame, j).
__global__ void feat_lda_kernel(const float *__restrict__ in, int T_total, int cols, int rows,
                                const float *__restrict__ lda, float *__restrict__ out) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)T_total * rows) return;
    const int t = (int)(id / rows), j = (int)(id % rows);
    const float *x = in + (size_t)t * cols, *w = lda + (size_t)j * cols;
    float s = 0.f;
    for (int k = 0; k < cols; ++k) s = __fadd_rn(s, __fmul_rn(x[k], w[k]));
    out[id] = s;
}

// feat_subvec_project (feat.c:334-355): a gather; entries past the vector's
// valid length read the zeros feat_lda_transform leaves there.
__global__ void feat_subvec_kernel(const float *__restrict__ in, int T_total, int in_len, int n_sv,
                                   const int32_t *__restrict__ idx, float *__restrict__ out) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)T_total * n_sv) return;
    const int t = (int)(id / n_sv), d = idx[id % n_sv];
    out[id] = d < in_len ? in[(size_t)t * in_len + d] : 0.f;
}

inline uint16_t fm(int kind, int a, int i) { return (uint16_t)(kind << 12 | a << 8 | i); }

// window, map and pre-LDA length of a -feat type; false if the reference's
// feat_init would reject it
bool feat_layout(const b200_feat_cfg_t *cfg, int &win, FeatMap &m) {
    const int type = cfg->type, cs = cfg->cepsize;
    m.n = 0;
    if (cs < 1 || cs > 64) return false;
    auto add = [&](int kind, int a, int i) { m.e[m.n++] = fm(kind, a, i); };
    switch (type) {
    case B200_FEAT_1S_C_D_DD:
        win = 3;
        for (int i = 0; i < cs; ++i) add(0, 0, i);
        for (int i = 0; i < cs; ++i) add(1, 2, i);
        for (int i = 0; i < cs; ++i) add(2, 0, i);
        return true;
    case B200_FEAT_S3_1X39:
        if (cs != 13) return false;
        win = 3;
        for (int i = 1; i < cs; ++i) add(0, 0, i);
        for (int i = 1; i < cs; ++i) add(1, 2, i);
        add(0, 0, 0); add(1, 2, 0); add(2, 0, 0);
        for (int i = 1; i < cs; ++i) add(2, 0, i);
        return true;
    case B200_FEAT_S2_4X:
        if (cs != 13) return false;
        win = 4;
        for (int i = 1; i < cs; ++i) add(0, 0, i);
        for (int i = 1; i < cs; ++i) add(1, 2, i);
        for (int i = 1; i < cs; ++i) add(1, 4, i);
        add(0, 0, 0); add(1, 2, 0); add(2, 0, 0);
        for (int i = 1; i < cs; ++i) add(2, 0, i);
        return true;
    case B200_FEAT_1S_C_D_LD_DD:
        win = 4;
        for (int i = 0; i < cs; ++i) add(0, 0, i);
        for (int i = 0; i < cs; ++i) add(1, 2, i);
        for (int i = 0; i < cs; ++i) add(1, 4, i);
        for (int i = 0; i < cs; ++i) add(2, 0, i);
        return true;
    case B200_FEAT_1S_C:
        win = 0;
        for (int i = 0; i < cs; ++i) add(0, 0, i);
        return true;
    case B200_FEAT_COPY: {
        // feat_copy (feat.c:828-849): per stream, the window's frames one after the other
        const int w = cfg->copy_window, ns = cfg->copy_streams;
        if (w < 0 || w > 7 || ns < 0 || ns > B200_MAX_STREAMS) return false;
        int tot = 0;
        for (int j = 0; j < ns; ++j) { if (cfg->copy_len[j] < 1) return false; tot += cfg->copy_len[j]; }
        if (ns && tot > cs) return false;
        if ((ns ? tot : cs) * (2 * w + 1) > kMaxFeatLen) return false;
        win = w;
        int spos = 0;
        for (int j = 0; j < (ns ? ns : 1); ++j) {
            const int len = ns ? cfg->copy_len[j] : cs;
            for (int i = -w; i <= w; ++i)
                for (int d = 0; d < len; ++d) add(3, i + 8, spos + d);
            spos += len;
        }
        return true;
    }
    case B200_FEAT_1S_C_D:
        win = 2;
        for (int i = 0; i < cs; ++i) add(0, 0, i);
        for (int i = 0; i < cs; ++i) add(1, 2, i);
        return true;
    }
    return false;
}

struct FeatPlan {
    int win, k, lda_dim, out_len;
    FeatMap map;
};

int feat_plan(const b200_feat_cfg_t *c, FeatPlan &p) {
    if (!c) { set_error("feature stage: null configuration"); return B200_ERR_ARG; }
    if (!feat_layout(c, p.win, p.map)) {
        set_error("feature stage: unknown -feat type %d or cepsize %d not valid for it", c->type, c->cepsize);
        return B200_ERR_ARG;
    }
    if (c->cmn != 0 && c->cmn != 1) { set_error("feature stage: only -cmn none|current exist in whole-utterance mode"); return B200_ERR_UNSUP; }
    if (c->agc != 0 && c->agc != 1) { set_error("feature stage: only -agc none|max exist in whole-utterance mode"); return B200_ERR_UNSUP; }
    p.k = p.map.n;
    p.lda_dim = 0;
    p.out_len = p.k;
    if (c->lda_rows > 0) {
        if (c->type == B200_FEAT_S2_4X || (c->type == B200_FEAT_COPY && c->copy_streams > 1)) { set_error("LDA incompatible with multi-stream features (lda.c:69-73)"); return B200_ERR_ARG; }
        if (!c->lda || c->lda_cols != p.k) {
            set_error("LDA matrix dimension %d doesn't match feature stream size %d (lda.c:127-128)", c->lda_cols, p.k);
            return B200_ERR_ARG;
        }
        p.lda_dim = (c->lda_dim > c->lda_rows || c->lda_dim <= 0) ? c->lda_rows : c->lda_dim;   // lda.c:131-134
        if (p.lda_dim > p.k) { set_error("LDA output dimension %d exceeds the stream size %d", p.lda_dim, p.k); return B200_ERR_ARG; }
        p.out_len = p.lda_dim;
    }
    if (c->n_subvec > 0) {
        if (c->type == B200_FEAT_S2_4X || (c->type == B200_FEAT_COPY && c->copy_streams > 1)) { set_error("subvector specifications require single-stream features (feat.c:294-297)"); return B200_ERR_ARG; }
        if (!c->subvec || c->n_subvec > p.out_len) {
            set_error("total dimensionality of subvector specification %d > feature dimensionality %d (feat.c:309-313)",
                      c->n_subvec, p.out_len);
            return B200_ERR_ARG;
        }
        for (int d = 0; d < c->n_subvec; ++d)
            if (c->subvec[d] < 0 || c->subvec[d] >= p.k) { set_error("subvector index %d outside the feature vector", c->subvec[d]); return B200_ERR_ARG; }
        p.out_len = c->n_subvec;
    }
    return B200_OK;
}

inline size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }

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

int b200_feat_dims(const b200_feat_cfg_t *cfg, int32_t dims[3]) {
    FeatPlan p;
    const int rc = feat_plan(cfg, p);
    if (rc) return rc;
    if (dims) { dims[0] = p.win; dims[1] = p.k; dims[2] = p.out_len; }
    return B200_OK;
}

size_t b200_feat_scratch_bytes(const b200_feat_cfg_t *cfg, int n_utt, int T_total) {
    FeatPlan p;
    if (feat_plan(cfg, p) || n_utt < 1 || T_total < 0) return 0;
    size_t b = up256((size_t)(2 * cfg->cepsize + 1) * n_utt * 4);
    if (p.lda_dim || cfg->n_subvec > 0) b += up256((size_t)T_total * p.k * 4);
    if (p.lda_dim && cfg->n_subvec > 0) b += up256((size_t)T_total * p.lda_dim * 4);
    if (p.lda_dim) b += up256((size_t)p.lda_dim * p.k * 4);
    if (cfg->n_subvec > 0) b += up256((size_t)cfg->n_subvec * 4);
    return b;
}

int b200_feat_compute_dev(const b200_feat_cfg_t *cfg, const float *d_cep, const int32_t *d_utt_off, int n_utt,
                          int T_total, void *d_scratch, float *d_feat, void *stream) {
    FeatPlan p;
    const int rc = feat_plan(cfg, p);
    if (rc) return rc;
    if (!d_cep || !d_utt_off || !d_feat || !d_scratch || n_utt < 1 || T_total < 0) {
        set_error("b200_feat_compute_dev: bad argument");
        return B200_ERR_ARG;
    }
    if (T_total == 0) return B200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int cs = cfg->cepsize, cmn = cfg->cmn, vn = cmn && cfg->varnorm, agc = cfg->agc;
    const bool lda = p.lda_dim > 0, sv = cfg->n_subvec > 0;
    char *s = (char *)d_scratch;
    float *mean = (float *)s, *scale = mean + (size_t)n_utt * cs, *agcmax = scale + (size_t)n_utt * cs;
    s += up256((size_t)(2 * cs + 1) * n_utt * 4);
    float *rowsA = nullptr, *rowsB = nullptr, *d_lda = nullptr; int32_t *d_sv = nullptr;
    if (lda || sv) { rowsA = (float *)s; s += up256((size_t)T_total * p.k * 4); }
    if (lda && sv) { rowsB = (float *)s; s += up256((size_t)T_total * p.lda_dim * 4); }
    if (lda) {
        d_lda = (float *)s; s += up256((size_t)p.lda_dim * p.k * 4);
        B200_CUDA_OK(cudaMemcpyAsync(d_lda, cfg->lda, (size_t)p.lda_dim * p.k * 4, cudaMemcpyHostToDevice, st));
    }
    if (sv) {
        d_sv = (int32_t *)s;
        B200_CUDA_OK(cudaMemcpyAsync(d_sv, cfg->subvec, (size_t)cfg->n_subvec * 4, cudaMemcpyHostToDevice, st));
    }
    if (cmn || agc) {
        feat_stats_kernel<<<(n_utt * cs + 127) / 128, 128, 0, st>>>(d_cep, d_utt_off, n_utt, cs, p.win, cmn, vn, agc, mean,
                                                                   scale, agcmax);
        B200_LAUNCH_CHECK();
    }
    {
        const long long n = (long long)T_total * p.k;
        feat_map_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_cep, d_utt_off, n_utt, T_total, cs, cmn, vn, agc, mean,
                                                                    scale, agcmax, p.map, (lda || sv) ? rowsA : d_feat);
        B200_LAUNCH_CHECK();
    }
    const float *cur = rowsA; int cur_len = p.k;
    if (lda) {
        const long long n = (long long)T_total * p.lda_dim;
        float *dst = sv ? rowsB : d_feat;
        feat_lda_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cur, T_total, p.k, p.lda_dim, d_lda, dst);
        B200_LAUNCH_CHECK();
        cur = dst; cur_len = p.lda_dim;
    }
    if (sv) {
        const long long n = (long long)T_total * cfg->n_subvec;
        feat_subvec_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cur, T_total, cur_len, cfg->n_subvec, d_sv, d_feat);
        B200_LAUNCH_CHECK();
    }
    return B200_OK;
}

int b200_feat_compute_host(const b200_feat_cfg_t *cfg, const float *cep, const int32_t *utt_off, int n_utt,
                           float *feat, int device) {
    FeatPlan p;
    int rc = feat_plan(cfg, p);
    if (rc) return rc;
    if (!cep || !utt_off || !feat || n_utt < 1) { set_error("b200_feat_compute_host: bad argument"); return B200_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: libb200sphinx has no CPU fallback"); return B200_ERR_CUDA; }
    B200_CUDA_OK(cudaSetDevice(device));
    const int T = utt_off[n_utt], cs = cfg->cepsize;
    if (T <= 0) return B200_OK;
    const size_t sb = b200_feat_scratch_bytes(cfg, n_utt, T);
    float *d_cep = nullptr, *d_feat = nullptr; void *d_scr = nullptr; int32_t *d_off = nullptr;
    if (cudaMalloc((void **)&d_cep, (size_t)T * cs * 4) != cudaSuccess ||
        cudaMalloc((void **)&d_feat, (size_t)T * p.out_len * 4) != cudaSuccess ||
        cudaMalloc(&d_scr, sb) != cudaSuccess ||
        cudaMalloc((void **)&d_off, (size_t)(n_utt + 1) * 4) != cudaSuccess ||
        cudaMemcpy(d_cep, cep, (size_t)T * cs * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(d_off, utt_off, (size_t)(n_utt + 1) * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("feature stage: device allocation/upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = B200_ERR_CUDA;
    }
    if (!rc) rc = b200_feat_compute_dev(cfg, d_cep, d_off, n_utt, T, d_scr, d_feat, nullptr);
    if (!rc && cudaMemcpy(feat, d_feat, (size_t)T * p.out_len * 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("feature stage: download failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = B200_ERR_CUDA;
    }
    cudaFree(d_cep); cudaFree(d_feat); cudaFree(d_scr); cudaFree(d_off);
    return rc;
}

}  // extern "C"
