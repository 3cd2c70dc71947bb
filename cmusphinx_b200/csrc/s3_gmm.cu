// s3_gmm.cu -- sphinx3's flavour of the acoustic-scoring path on the GPU.
//
// Reference (S3 = sphinx3/src/libs3decoder):
//   approx_cont_mgau_frame_eval   S3/libam/approx_cont_mgau.c:433-616
//   approx_cont_mgau_ci_eval      S3/libam/approx_cont_mgau.c:368-431
//   approx_compute_dyn_ci_pbeam   S3/libam/approx_cont_mgau.c:302-358
//   approx_isskip (-ds only)      S3/libam/approx_cont_mgau.c:93-143
//   mgau_eval / _all / _active    S3/libam/cont_mgau.c:1033-1205
//   mgau_init, mgau_precomp,
//   mgau_uninit_compact           S3/libam/cont_mgau.c:700-958
//   fast_gmm_init                 S3/libam/fast_algo_struct.c:420-467
//
// Arithmetic contract (bit-exact with the reference): differences x-m in
// float32, widened to float64; dval -= (diff*diff)*v with separately rounded
// float64 multiplies and subtract, sequential over dimensions; gauscr =
// (int32)(f*dval) + mixw; integer table log-add (shift 0) in component order.
//
// Device decomposition of one utterance chunk of T frames:
//   K1 s3_eval_kernel   (CI senones, all frames)  -> raw CI scores
//   K2 s3_decide_kernel (block per frame)         -> CI best, (dynamic) beam,
//                                                    per (t,s) flag {0 inactive,1 full,2 back-off}
//   K1 s3_eval_kernel   (CD senones, flagged)     -> raw scores + best component
//   K3 s3_backoff_kernel (thread per (t, CD senone)): best-index / update-time
//                        state resolved by a bounded walk back over the flags, CI or
//                        best-Gaussian back-off; s3_state_kernel: state after the chunk
//   K4 s3_best_kernel   (block per frame)         -> frame best over active
//   K5 s3_norm_kernel   (thread per senone, sequential in t): subtract best on
//                        active entries, carry stale entries forward
// The frame axis is sequential only in K3/K5, which touch a few bytes per
// (t,s); all Gaussian arithmetic (K1) is parallel over frames and senones.
#include "dev_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

using namespace b200;

namespace {

constexpr int32_t kS3Zero = (int32_t)0xc8000000;   // s3types.h:192
constexpr int kNoBst = -1;                         // cont_mgau.h:134
constexpr int kNotUpdated = -100;                  // cont_mgau.h:135
constexpr int kFB = 32;                            // frames per block in K1
constexpr int kEvalThreads = 128;
constexpr int kFR = 8;                             // frames sharing one parameter pass in K1

struct S3Dev {
    int n_sen, n_ci, veclen, cpt;   // cpt = padded components per senone
    const float *mean;     // [s][i][cpt]
    const double *var;     // [s][i][cpt]  1/(2 var) widened at load
    const float *lrd;      // [s][cpt]
    const int32_t *mixw;   // [s][cpt]
    const int32_t *ncomp;  // [s]
    const int32_t *cd2ci;  // [s]
    const uint16_t *tab16; const uint32_t *tab32; uint32_t tab_size;
    int32_t lzero;         // logmath zero (MIN_INT32 >> 2)
    double distfloor, f;
    // sub-vector quantised shortlists (S3/libam/subvq.c); svq_n_sv == 0: none
    int svq_n_sv, svq_size, svq_eval; int32_t svq_beam;
    const int32_t *svq_map;    // [s][cpt][n_sv] compacted + linearised (sub-vector * size + codeword)
    // Gaussian selector (S3/libam/gs.c); gs_n_code == 0: none
    int gs_n_code;
    const uint32_t *gs_map;    // [s][n_code] bit c = component c is short-listed for that codeword
};

// logmath_add, SB/util/logmath.c:391-436 (shift 0 table, 16 or 32 bit wide)
__device__ __forceinline__ int32_t s3_logadd(const S3Dev &g, int32_t x, int32_t y) {
    if (x <= g.lzero) return y;
    if (y <= g.lzero) return x;
    int32_t d, r;
    if (x > y) { d = x - y; r = x; } else { d = y - x; r = y; }
    if (d < 0 || (uint32_t)d >= g.tab_size) return r;
    return r + (g.tab16 ? (int32_t)__ldg(g.tab16 + d) : (int32_t)__ldg(g.tab32 + d));
}

// one density against one frame: cont_mgau.c:1062-1068
__device__ __forceinline__ double s3_dval(const S3Dev &g, int s, int c, const float *x) {
    const float *mp = g.mean + (size_t)s * g.veclen * g.cpt + c;
    const double *vp = g.var + (size_t)s * g.veclen * g.cpt + c;
    double dval = (double)__ldg(g.lrd + (size_t)s * g.cpt + c);
#pragma unroll 3
    for (int i = 0; i < g.veclen; ++i) {
        float diff = __fsub_rn(x[i], __ldg(mp + (size_t)i * g.cpt));
        double dd = (double)diff;
        dval = __dsub_rn(dval, __dmul_rn(__dmul_rn(dd, dd), __ldg(vp + (size_t)i * g.cpt)));
    }
    return dval;
}

__device__ __forceinline__ int32_t s3_gauscr(const S3Dev &g, int s, int c, double dval) {
    if (dval < g.distfloor) dval = g.distfloor;
    return __double2int_rz(__dmul_rn(g.f, dval)) + __ldg(g.mixw + (size_t)s * g.cpt + c);
}

// K1: CP lanes per senone (one component each, KC rounds when a senone has
// more than 32 components).  flags == nullptr: every frame of every senone in
// [s_lo, s_hi).  Writes raw[t][s] and (bst != nullptr) the arg-max component.
template <int CP, int KC>
__global__ void __launch_bounds__(kEvalThreads, 8)
s3_eval_kernel(S3Dev g, const float *__restrict__ feat, int T, int s_lo, int s_hi,
               const uint8_t *__restrict__ flags, int32_t *__restrict__ raw, int16_t *__restrict__ bst,
               const int32_t *__restrict__ vqd /* [T][n_sv * size] sub-VQ scores of every frame, or null */,
               const int32_t *__restrict__ gscw /* [T] closest Gaussian-selector codeword of every frame, or null */) {
    extern __shared__ float xs[];   // [kFB][veclen] | int32 stage[G][kFR][CP*KC]
    constexpr int G = kEvalThreads / CP;
    const int t0 = blockIdx.y * kFB;
    const int nf = min(kFB, T - t0);
    for (int i = threadIdx.x; i < nf * g.veclen; i += kEvalThreads) xs[i] = feat[(size_t)t0 * g.veclen + i];
    __syncthreads();
    const int grp = threadIdx.x / CP, lc = threadIdx.x % CP;
    int32_t *gstage = reinterpret_cast<int32_t *>(xs + kFB * g.veclen);
    const int s = s_lo + blockIdx.x * G + grp;
    if (s >= s_hi) return;
    const unsigned gmask = CP == 32 ? 0xffffffffu : (((1u << CP) - 1u) << ((threadIdx.x % 32) / CP * CP));
    const int lane0 = (threadIdx.x % 32) / CP * CP;
    // frames of this block that need this senone
    uint32_t mask = 0;
    if (flags) {
        for (int k = lc; k < nf; k += CP)
            if (flags[(size_t)(t0 + k) * g.n_sen + s] == 1) mask |= 1u << k;
#pragma unroll
        for (int o = CP / 2; o > 0; o >>= 1) mask |= __shfl_xor_sync(gmask, mask, o);
    } else {
        mask = nf == 32 ? 0xffffffffu : ((1u << nf) - 1u);
    }
    const int nc = __ldg(g.ncomp + s);
    const int D = g.veclen;
    while (mask) {
        // up to kFR flagged frames share one pass over this senone's parameters
        const uint32_t mask0 = mask;
        int k[kFR], nv = 0;
#pragma unroll
        for (int j = 0; j < kFR; ++j) {
            if (mask) { k[j] = __ffs(mask) - 1; mask &= mask - 1; nv = j + 1; }
            else k[j] = k[0];
        }
        int32_t *stage = gstage + (size_t)grp * kFR * CP * KC;   // [kFR][CP*KC] of this lane group
#pragma unroll
        for (int r = 0; r < KC; ++r) {
            const int c = lc + r * CP;
            if (c < nc) {
                const float *mp = g.mean + (size_t)s * D * g.cpt + c;
                const double *vp = g.var + (size_t)s * D * g.cpt + c;
                double dv[kFR];
                const float *xp[kFR];
                const double lrd = (double)__ldg(g.lrd + (size_t)s * g.cpt + c);
#pragma unroll
                for (int j = 0; j < kFR; ++j) { dv[j] = lrd; xp[j] = xs + k[j] * D; }
#pragma unroll 3
                for (int i = 0; i < D; ++i) {
                    const float mu = __ldg(mp + (size_t)i * g.cpt);
                    const double va = __ldg(vp + (size_t)i * g.cpt);
#pragma unroll
                    for (int j = 0; j < kFR; ++j) {
                        const double dd = (double)__fsub_rn(xp[j][i], mu);
                        dv[j] = __dsub_rn(dv[j], __dmul_rn(__dmul_rn(dd, dd), va));
                    }
                }
#pragma unroll
                for (int j = 0; j < kFR; ++j) stage[j * CP * KC + c] = s3_gauscr(g, s, c, dv[j]);
            }
        }
        __syncwarp(gmask);
        // sequential log-add in component order (mgau_eval_all), one lane per
        // frame; with update_best_id == 1 the best component is the first
        // strict maximum
        for (int j = lc; j < nv; j += CP) {
            const int32_t *sj = stage + j * CP * KC;
            const int kj = __fns(mask0, 0, j + 1);      // j-th flagged frame of this pass
            // approx_mgau_eval (approx_cont_mgau.c:187-284) with a sub-VQ model: the components whose
            // quantised score is within the beam of the senone's best one (subvq_mgau_shortlist,
            // subvq.c:383-468) -- a mask over the dense result -- and the whole mixture again when
            // the shortlist's score is hopeless (< S3_LOGPROB_ZERO + 100000, :256-281)
            const int32_t *vq = vqd ? vqd + (size_t)(t0 + kj) * g.svq_n_sv * g.svq_size : nullptr;
            const int32_t *mp0 = g.svq_map + (size_t)s * g.cpt * g.svq_n_sv;
            auto quant = [&](int c) -> int32_t {
                const int32_t *mp = mp0 + (size_t)c * g.svq_n_sv;
                uint32_t v;                              // (the reference's int32 sums wrap)
                if (g.svq_n_sv == 3) {
                    if (g.svq_eval == 1) v = (uint32_t)vq[mp[0]];
                    else if (g.svq_eval == 2) v = (uint32_t)vq[mp[0]] + 2u * (uint32_t)vq[mp[1]];
                    else v = (uint32_t)vq[mp[0]] + (uint32_t)vq[mp[1]] + (uint32_t)vq[mp[2]];
                } else {
                    v = 0;
                    for (int k2 = 0; k2 < g.svq_n_sv; ++k2) v += (uint32_t)vq[mp[k2]];
                }
                return (int32_t)v;
            };
            int32_t th = (int32_t)0x80000000;
            // the Gaussian selector goes first (gs4gs): the bit map of (senone, closest codeword), every
            // component when it is empty (gs_mgau_shortlist, gs.c:263-300)
            uint32_t gsbits = 0xffffffffu;
            if (gscw) {
                gsbits = g.gs_map[(size_t)s * g.gs_n_code + gscw[t0 + kj]];
                if (nc < 32) gsbits &= (1u << nc) - 1u;
                if (gsbits == 0) gsbits = 0xffffffffu;
                vq = nullptr;
            }
            bool shortlist = vq != nullptr || gscw != nullptr;
            if (vq) {
                int32_t bv = (int32_t)0x80000000;
                for (int c = 0; c < nc; ++c) bv = max(bv, quant(c));
                th = (int32_t)((uint32_t)bv + (uint32_t)g.svq_beam);
            }
            int32_t score, bscr; int bidx;
            while (true) {
                score = kS3Zero; bscr = kS3Zero; bidx = kNoBst;
                for (int c = 0; c < nc; ++c) {
                    if (shortlist && (vq ? quant(c) < th : !((gsbits >> (c & 31)) & 1u))) continue;
                    const int32_t v = sj[c];
                    score = s3_logadd(g, score, v);
                    if (v > bscr) { bscr = v; bidx = c; }
                }
                if (score <= kS3Zero) score = kS3Zero;
                if (!(shortlist && score < kS3Zero + 100000)) break;
                shortlist = false;
            }
            raw[(size_t)(t0 + kj) * g.n_sen + s] = score;
            if (bst) bst[(size_t)(t0 + kj) * g.n_sen + s] = (int16_t)bidx;
        }
        __syncwarp(gmask);
    }
}

// subvq_gautbl_eval_logs3 (subvq.c:488-506) / vector_gautbl_eval_logs3 (S3/libcommon/vector.c:590-650) for every
// frame: one block per frame, one thread per (sub-vector, codeword); the same float32-difference /
// float64-accumulate arithmetic as the densities.  Sub-vectors >= VQ_EVAL are never evaluated by the reference
// and keep the 0 its calloc left there.
struct S3Vq {
    int n_sv, size, n_eval;
    const int32_t *veclen, *off_dim, *off_par;   // [n_sv]: length, offset into featdim, offset into mean / var
    const int32_t *featdim;
    const float *mean, *lrd;                     // mean [off_par + r * L + i], lrd [sv * size + r]
    const double *var;
    double distfloor, f;
};
__global__ void __launch_bounds__(128)
s3_vq_kernel(S3Vq q, const float *__restrict__ feat, int T, int veclen, int32_t *__restrict__ out) {
    const int t = blockIdx.x;
    const float *x = feat + (size_t)t * veclen;
    for (int e = threadIdx.x; e < q.n_sv * q.size; e += blockDim.x) {
        const int sv = e / q.size, r = e % q.size;
        int32_t v = 0;
        if (sv < q.n_eval) {
            const int L = q.veclen[sv];
            const int32_t *fd = q.featdim + q.off_dim[sv];
            const float *mu = q.mean + q.off_par[sv] + (size_t)r * L;
            const double *va = q.var + q.off_par[sv] + (size_t)r * L;
            double dval = (double)q.lrd[e];
            for (int i = 0; i < L; ++i) {
                const double dd = (double)__fsub_rn(x[fd[i]], mu[i]);
                dval = __dsub_rn(dval, __dmul_rn(__dmul_rn(dd, dd), va[i]));
            }
            if (dval < q.distfloor) dval = q.distfloor;
            v = __double2int_rz(__dmul_rn(q.f, dval));
        }
        out[(size_t)t * q.n_sv * q.size + e] = v;
    }
}

// gc_compute_closest_cw (gs.c:221-259) for every frame: squared Euclidean distance with float32 differences
// summed in float64 in dimension order; the FIRST minimum wins.  One block per frame, one thread per codeword.
__global__ void __launch_bounds__(128)
s3_gs_kernel(const float *__restrict__ cw, int n_code, int L, const float *__restrict__ feat, int veclen, int32_t *__restrict__ out) {
    __shared__ double s_d[128];
    __shared__ int s_i[128];
    const int t = blockIdx.x;
    const float *x = feat + (size_t)t * veclen;
    double best = 1.7976931348623157e308; int bi = 0x7fffffff;
    for (int c = threadIdx.x; c < n_code; c += blockDim.x) {
        double tmp = 0.0;
        for (int i = 0; i < L; ++i) {
            const double dd = (double)__fsub_rn(x[i], cw[(size_t)c * L + i]);
            tmp = __dadd_rn(tmp, __dmul_rn(dd, dd));
        }
        if (tmp < best) { best = tmp; bi = c; }        // ascending c within the thread: first minimum
    }
    s_d[threadIdx.x] = best; s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double d2 = s_d[threadIdx.x + o]; const int i2 = s_i[threadIdx.x + o];
            if (d2 < s_d[threadIdx.x] || (d2 == s_d[threadIdx.x] && i2 < s_i[threadIdx.x])) { s_d[threadIdx.x] = d2; s_i[threadIdx.x] = i2; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[t] = s_i[0] == 0x7fffffff ? 0 : s_i[0];
}

// K2: one block per frame.
__global__ void __launch_bounds__(256)
s3_decide_kernel(S3Dev g, int T, int frame0, int32_t ci_pbeam, int max_cd, int ds_ratio, float tighten,
                 const int32_t *__restrict__ raw, uint8_t *__restrict__ act /* may be null */,
                 uint8_t *__restrict__ flags, int32_t *__restrict__ beam_out) {
    extern __shared__ int32_t sm[];      // ci[n_ci] | occ[n_ci] | order[n_ci]
    int32_t *ci = sm, *occ = sm + g.n_ci, *order = sm + 2 * g.n_ci;
    __shared__ int32_t s_pbest, s_beam;
    const int t = blockIdx.x;
    const int32_t *row = raw + (size_t)t * g.n_sen;
    uint8_t *arow = act ? act + (size_t)t * g.n_sen : nullptr;
    if (threadIdx.x == 0) s_pbest = INT32_MIN;
    for (int i = threadIdx.x; i < g.n_ci; i += blockDim.x) { ci[i] = row[i]; occ[i] = 0; }
    __syncthreads();
    int32_t m = INT32_MIN;
    for (int i = threadIdx.x; i < g.n_ci; i += blockDim.x) m = max(m, ci[i]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x % 32 == 0) atomicMax(&s_pbest, m);
    __syncthreads();
    const int32_t pbest = s_pbest;
    int32_t beam = ci_pbeam;
    if (max_cd < g.n_sen - g.n_ci) {
        // approx_compute_dyn_ci_pbeam: occupancy of every CI senone by active
        // CD senones, CI scores sorted descending, cumulative occupancy.
        for (int s = g.n_ci + threadIdx.x; s < g.n_sen; s += blockDim.x)
            if (!arow || arow[s]) atomicAdd(&occ[g.cd2ci[s]], 1);
        __syncthreads();
        for (int i = threadIdx.x; i < g.n_ci; i += blockDim.x) {
            const int32_t v = ci[i];
            int r = 0;
            for (int j = 0; j < g.n_ci; ++j) r += (ci[j] > v) || (ci[j] == v && j < i);
            order[r] = i;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int total = 0;
            for (int k = 0; k < g.n_ci && ci[order[k]] > pbest + ci_pbeam; ++k) {
                total += occ[order[k]];
                if (total > max_cd) { beam = ci[order[k]] - pbest; break; }
            }
            s_beam = beam;
        }
        __syncthreads();
        beam = s_beam;
    }
    const bool skip = ((frame0 + t) % ds_ratio) != 0;
    if (skip) beam = (int32_t)__fmul_rn((float)beam, tighten);
    if (threadIdx.x == 0) beam_out[t] = beam;
    const int32_t thr = pbest + beam;
    uint8_t *frow = flags + (size_t)t * g.n_sen;
    for (int s = threadIdx.x; s < g.n_sen; s += blockDim.x) {
        if (s < g.n_ci) { frow[s] = 1; if (arow) arow[s] = 1; continue; }
        uint8_t f = 0;
        if (!arow || arow[s]) f = ci[g.cd2ci[s]] >= thr ? 1 : 2;
        frow[s] = f;
    }
}

// K3a: back-off entries (flag 2), one thread per (frame, CD senone), parallel in
// time.  The reference's per-senone state machine (bstidx, updatetime) only
// lets a best-Gaussian back-off happen at frame t when the state was written at
// frame t-1: by a full evaluation (flag 1), or -- on down-sampled ("skipped")
// frames, where mgau_eval runs with update_best_id = 1 -- by a previous
// back-off.  Such a chain is at most ds_ratio-1 frames long, so every entry
// resolves its own state by walking back over the flags.
__global__ void __launch_bounds__(128)
s3_backoff_kernel(S3Dev g, const float *__restrict__ feat, int T, int frame0, int ds_ratio,
                  const uint8_t *__restrict__ flags, int32_t *__restrict__ raw, const int16_t *__restrict__ bst,
                  const int32_t *__restrict__ st_bstidx, const int32_t *__restrict__ st_update) {
    const int n_cd = g.n_sen - g.n_ci;
    const int s = g.n_ci + blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (s >= g.n_sen || n_cd <= 0) return;
    if (flags[(size_t)t * g.n_sen + s] != 2) return;
    // walk back over chain-continuing frames (flag 2 on a skipped frame)
    int u = t - 1;
    while (u >= 0 && flags[(size_t)u * g.n_sen + s] == 2 && ((frame0 + u) % ds_ratio) != 0) --u;
    int bidx = kNoBst;
    if (u >= 0) {
        if (flags[(size_t)u * g.n_sen + s] == 1) bidx = bst[(size_t)u * g.n_sen + s];
    } else if (st_update[s] == frame0 - 1) {
        bidx = st_bstidx[s];
    }
    // replay the chain: a back-off on a skipped frame keeps the index only while
    // its Gaussian score stays above S3_LOGPROB_ZERO
    int32_t v = 0;
    for (int w = u + 1; w <= t && bidx != kNoBst; ++w) {
        v = s3_gauscr(g, s, bidx, s3_dval(g, s, bidx, feat + (size_t)w * g.veclen));
        if (w < t && !(v > kS3Zero)) bidx = kNoBst;
    }
    int32_t score;
    if (bidx == kNoBst) {
        score = raw[(size_t)t * g.n_sen + g.cd2ci[s]];
    } else {
        score = s3_logadd(g, kS3Zero, v);
        if (score <= kS3Zero) score = kS3Zero;
    }
    raw[(size_t)t * g.n_sen + s] = score;
}

// K3b: state after the chunk, one thread per CD senone: the last frame that
// wrote (bstidx, updatetime).
__global__ void __launch_bounds__(128)
s3_state_kernel(S3Dev g, const float *__restrict__ feat, int T, int frame0, int ds_ratio,
                const uint8_t *__restrict__ flags, const int16_t *__restrict__ bst,
                int32_t *__restrict__ st_bstidx, int32_t *__restrict__ st_update) {
    const int s = g.n_ci + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.n_sen) return;
    // last full evaluation
    int r = T - 1;
    while (r >= 0 && flags[(size_t)r * g.n_sen + s] != 1) --r;
    int bidx, upd;
    if (r >= 0) { bidx = bst[(size_t)r * g.n_sen + s]; upd = frame0 + r; }
    else { bidx = st_bstidx[s]; upd = st_update[s]; }
    // a chain of back-offs on skipped frames directly after it keeps updating
    for (int w = r + 1; w < T; ++w) {
        const int frame = frame0 + w;
        if (flags[(size_t)w * g.n_sen + s] != 2 || (frame % ds_ratio) == 0 || bidx == kNoBst || upd != frame - 1) break;
        const int32_t v = s3_gauscr(g, s, bidx, s3_dval(g, s, bidx, feat + (size_t)w * g.veclen));
        bidx = v > kS3Zero ? bidx : kNoBst;
        upd = frame;
    }
    st_bstidx[s] = bidx; st_update[s] = upd;
}

// CI senones are re-evaluated with update_best_id = 1 every frame
// (approx_cont_mgau_ci_eval): their state after the chunk is that of the last frame.
__global__ void s3_ci_state_kernel(S3Dev g, int T, int frame0, const int16_t *__restrict__ bst,
                                   int32_t *__restrict__ st_bstidx, int32_t *__restrict__ st_update) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.n_ci) return;
    st_bstidx[s] = bst[(size_t)(T - 1) * g.n_sen + s];
    st_update[s] = frame0 + T - 1;
}

// K4: frame best over active senones.
__global__ void __launch_bounds__(256)
s3_best_kernel(int n_sen, const uint8_t *__restrict__ flags, const int32_t *__restrict__ raw,
               int32_t *__restrict__ best) {
    __shared__ int32_t s_best;
    const int t = blockIdx.x;
    if (threadIdx.x == 0) s_best = INT32_MIN;
    __syncthreads();
    int32_t m = INT32_MIN;
    for (int s = threadIdx.x; s < n_sen; s += blockDim.x)
        if (flags[(size_t)t * n_sen + s]) m = max(m, raw[(size_t)t * n_sen + s]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x % 32 == 0) atomicMax(&s_best, m);
    __syncthreads();
    if (threadIdx.x == 0) best[t] = s_best;
}

// K5: normalise active entries, carry inactive ones forward.
__global__ void __launch_bounds__(128)
s3_norm_kernel(int n_sen, int T, const uint8_t *__restrict__ flags, const int32_t *__restrict__ raw,
               const int32_t *__restrict__ best, int32_t *__restrict__ prev, int32_t *__restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sen) return;
    int32_t p = prev[s];
    for (int t = 0; t < T; ++t) {
        if (flags[(size_t)t * n_sen + s]) p = raw[(size_t)t * n_sen + s] - best[t];
        out[(size_t)t * n_sen + s] = p;
    }
    prev[s] = p;
}

// Measurement utility: issue rate of the FP64 pipe with the instruction mix the
// scoring kernel is bound by (independent DMUL / DADD chains, no FMA).
__global__ void __launch_bounds__(256) s3_fp64_probe_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = __dmul_rn(x0, a); x1 = __dadd_rn(x1, b); x2 = __dmul_rn(x2, a); x3 = __dadd_rn(x3, b);
        x4 = __dmul_rn(x4, a); x5 = __dadd_rn(x5, b); x6 = __dmul_rn(x6, a); x7 = __dadd_rn(x7, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace

struct b200_s3mgau {
    int device = 0;
    int n_sen = 0, n_ci = 0, max_comp = 0, veclen = 0, cp = 1, kc = 1, cpt = 1;
    double logbase = 0, distfloor = 0, f = 0;
    int32_t ci_pbeam = 0; int max_cd = 100000, ds_ratio = 1; float tighten = 0.5f;
    LogMath *lm = nullptr;
    // host copies (reference layout, padded to max_comp) for b200_s3_params
    std::vector<int32_t> h_ncomp, h_mixw, h_cd2ci;
    std::vector<float> h_mean, h_var, h_lrd;
    // device
    float *d_mean = nullptr, *d_lrd = nullptr; double *d_var = nullptr;
    int32_t *d_mixw = nullptr, *d_ncomp = nullptr, *d_cd2ci = nullptr;
    uint16_t *d_tab16 = nullptr; uint32_t *d_tab32 = nullptr;
    int32_t *d_bstidx = nullptr, *d_update = nullptr, *d_prev = nullptr;
    // scratch (chunk of frames)
    size_t capT = 0;
    float *d_feat = nullptr; uint8_t *d_act = nullptr, *d_flags = nullptr;
    int32_t *d_raw = nullptr, *d_out = nullptr, *d_beam = nullptr, *d_best = nullptr; int16_t *d_bst = nullptr;
    // sub-VQ model (b200_s3_set_subvq), optional
    int svq_n_sv = 0, svq_size = 0, svq_eval = 0; int32_t svq_beam = 0;
    int32_t *d_svq_map = nullptr, *d_svq_i = nullptr; float *d_svq_f = nullptr; double *d_svq_var = nullptr;
    int32_t *d_vqd = nullptr; size_t vqd_cap = 0;
    S3Vq vq{};
    // Gaussian selector (b200_s3_set_gs), optional
    int gs_n_code = 0, gs_featlen = 0;
    float *d_gs_cw = nullptr; uint32_t *d_gs_map = nullptr; int32_t *d_gscw = nullptr; size_t gscw_cap = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    float last_ms = 0.f;
    S3Dev dev() const {
        S3Dev g{};
        g.n_sen = n_sen; g.n_ci = n_ci; g.veclen = veclen; g.cpt = cpt;
        g.mean = d_mean; g.var = d_var; g.lrd = d_lrd; g.mixw = d_mixw; g.ncomp = d_ncomp; g.cd2ci = d_cd2ci;
        g.tab16 = d_tab16; g.tab32 = d_tab32; g.tab_size = (uint32_t)lm->table.size();
        g.lzero = lm->zero; g.distfloor = distfloor; g.f = f;
        g.svq_n_sv = svq_n_sv; g.svq_size = svq_size; g.svq_eval = svq_eval; g.svq_beam = svq_beam; g.svq_map = d_svq_map;
        g.gs_n_code = gs_n_code; g.gs_map = d_gs_map;
        return g;
    }
};

namespace {

constexpr int kChunkT = 2048;

int s3_reserve(b200_s3mgau *m, int T) {
    if ((size_t)T <= m->capT) return B200_OK;
    void *ptrs[] = {m->d_feat, m->d_act, m->d_flags, m->d_raw, m->d_out, m->d_beam, m->d_best, m->d_bst};
    for (void *p : ptrs) if (p) cudaFree(p);
    m->d_feat = nullptr; m->d_act = m->d_flags = nullptr; m->d_raw = m->d_out = m->d_beam = m->d_best = nullptr; m->d_bst = nullptr;
    m->capT = 0;
    const size_t S = m->n_sen, t = T;
    B200_CUDA_OK(cudaMalloc((void **)&m->d_feat, t * m->veclen * sizeof(float)));
    B200_CUDA_OK(cudaMalloc((void **)&m->d_act, t * S));
    B200_CUDA_OK(cudaMalloc((void **)&m->d_flags, t * S));
    B200_CUDA_OK(cudaMalloc((void **)&m->d_raw, t * S * sizeof(int32_t)));
    B200_CUDA_OK(cudaMalloc((void **)&m->d_out, t * S * sizeof(int32_t)));
    B200_CUDA_OK(cudaMalloc((void **)&m->d_bst, t * S * sizeof(int16_t)));
    B200_CUDA_OK(cudaMalloc((void **)&m->d_beam, t * sizeof(int32_t)));
    B200_CUDA_OK(cudaMalloc((void **)&m->d_best, t * sizeof(int32_t)));
    m->capT = T;
    return B200_OK;
}

template <int CP, int KC>
int launch_eval_t(const b200_s3mgau *m, const float *d_feat, int T, int s_lo, int s_hi, const uint8_t *flags,
                  int32_t *raw, int16_t *bst, cudaStream_t st, const int32_t *vqd, const int32_t *gscw) {
    if (s_hi <= s_lo || T <= 0) return B200_OK;
    constexpr int G = kEvalThreads / CP;
    dim3 grid((s_hi - s_lo + G - 1) / G, (T + kFB - 1) / kFB);
    size_t smem = (size_t)kFB * m->veclen * sizeof(float) + (size_t)kEvalThreads * KC * kFR * sizeof(int32_t);
    s3_eval_kernel<CP, KC><<<grid, kEvalThreads, smem, st>>>(m->dev(), d_feat, T, s_lo, s_hi, flags, raw, bst, vqd, gscw);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int launch_eval(const b200_s3mgau *m, const float *d_feat, int T, int s_lo, int s_hi, const uint8_t *flags,
                int32_t *raw, int16_t *bst, cudaStream_t st, const int32_t *vqd = nullptr, const int32_t *gscw = nullptr) {
    switch (m->cp * 8 + m->kc) {
    case 1 * 8 + 1: return launch_eval_t<1, 1>(m, d_feat, T, s_lo, s_hi, flags, raw, bst, st, vqd, gscw);
    case 2 * 8 + 1: return launch_eval_t<2, 1>(m, d_feat, T, s_lo, s_hi, flags, raw, bst, st, vqd, gscw);
    case 4 * 8 + 1: return launch_eval_t<4, 1>(m, d_feat, T, s_lo, s_hi, flags, raw, bst, st, vqd, gscw);
    case 8 * 8 + 1: return launch_eval_t<8, 1>(m, d_feat, T, s_lo, s_hi, flags, raw, bst, st, vqd, gscw);
    case 16 * 8 + 1: return launch_eval_t<16, 1>(m, d_feat, T, s_lo, s_hi, flags, raw, bst, st, vqd, gscw);
    case 32 * 8 + 1: return launch_eval_t<32, 1>(m, d_feat, T, s_lo, s_hi, flags, raw, bst, st, vqd, gscw);
    case 32 * 8 + 2: return launch_eval_t<32, 2>(m, d_feat, T, s_lo, s_hi, flags, raw, bst, st, vqd, gscw);
    case 32 * 8 + 4: return launch_eval_t<32, 4>(m, d_feat, T, s_lo, s_hi, flags, raw, bst, st, vqd, gscw);
    }
    set_error("unsupported component count");
    return B200_ERR_UNSUP;
}

// One chunk, everything device-resident.  d_act may be null (all active).
int s3_chunk_dev(b200_s3mgau *m, const float *d_feat, int T, int frame0, uint8_t *d_act, int32_t *d_out,
                 int32_t *d_best, cudaStream_t st) {
    const S3Dev g = m->dev();
    int rc;
    const int32_t *vqd = nullptr;
    if (m->svq_n_sv) {
        const size_t need = (size_t)T * m->svq_n_sv * m->svq_size;
        if (need > m->vqd_cap) {
            cudaFree(m->d_vqd); m->d_vqd = nullptr; m->vqd_cap = 0;
            B200_CUDA_OK(cudaMalloc((void **)&m->d_vqd, need * sizeof(int32_t)));
            m->vqd_cap = need;
        }
        s3_vq_kernel<<<T, 128, 0, st>>>(m->vq, d_feat, T, m->veclen, m->d_vqd);
        B200_LAUNCH_CHECK();
        vqd = m->d_vqd;
    }
    const int32_t *gscw = nullptr;
    if (m->gs_n_code) {
        if ((size_t)T > m->gscw_cap) {
            cudaFree(m->d_gscw); m->d_gscw = nullptr; m->gscw_cap = 0;
            B200_CUDA_OK(cudaMalloc((void **)&m->d_gscw, (size_t)T * sizeof(int32_t)));
            m->gscw_cap = T;
        }
        s3_gs_kernel<<<T, 128, 0, st>>>(m->d_gs_cw, m->gs_n_code, m->gs_featlen, d_feat, m->veclen, m->d_gscw);
        B200_LAUNCH_CHECK();
        gscw = m->d_gscw;
    }
    if ((rc = launch_eval(m, d_feat, T, 0, m->n_ci, nullptr, m->d_raw, m->d_bst, st, vqd, gscw))) return rc;
    s3_decide_kernel<<<T, 256, (size_t)3 * std::max(m->n_ci, 1) * sizeof(int32_t), st>>>(
        g, T, frame0, m->ci_pbeam, m->max_cd, m->ds_ratio, m->tighten, m->d_raw, d_act, m->d_flags, m->d_beam);
    B200_LAUNCH_CHECK();
    if ((rc = launch_eval(m, d_feat, T, m->n_ci, m->n_sen, m->d_flags, m->d_raw, m->d_bst, st, vqd, gscw))) return rc;
    const int n_cd = m->n_sen - m->n_ci;
    if (n_cd > 0) {
        s3_backoff_kernel<<<dim3((n_cd + 127) / 128, T), 128, 0, st>>>(g, d_feat, T, frame0, m->ds_ratio, m->d_flags,
                                                                     m->d_raw, m->d_bst, m->d_bstidx, m->d_update);
        B200_LAUNCH_CHECK();
        s3_state_kernel<<<(n_cd + 127) / 128, 128, 0, st>>>(g, d_feat, T, frame0, m->ds_ratio, m->d_flags, m->d_bst,
                                                            m->d_bstidx, m->d_update);
        B200_LAUNCH_CHECK();
    }
    if (m->n_ci > 0) {
        s3_ci_state_kernel<<<(m->n_ci + 127) / 128, 128, 0, st>>>(g, T, frame0, m->d_bst, m->d_bstidx, m->d_update);
        B200_LAUNCH_CHECK();
    }
    s3_best_kernel<<<T, 256, 0, st>>>(m->n_sen, m->d_flags, m->d_raw, d_best);
    B200_LAUNCH_CHECK();
    s3_norm_kernel<<<(m->n_sen + 127) / 128, 128, 0, st>>>(m->n_sen, T, m->d_flags, m->d_raw, d_best, m->d_prev, d_out);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

}  // namespace

extern "C" {

b200_s3mgau_t *b200_s3_create(int n_sen, int n_comp, int veclen, const float *mean, const float *var,
                              const float *mixw, double varfloor, double mixwfloor, double logbase,
                              const int32_t *cd2cisen, int n_ci_sen, int device) {
    if (!mean || !var || !mixw || !cd2cisen || n_sen <= 0 || n_comp <= 0 || veclen <= 0 || n_ci_sen < 0 ||
        n_ci_sen > n_sen) { set_error("b200_s3_create: bad argument"); return nullptr; }
    if (n_comp > 128) { set_error("more than 128 components per senone is not supported"); return nullptr; }
    if (n_ci_sen > 4096) { set_error("more than 4096 CI senones is not supported"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: libb200sphinx has no CPU fallback");
        return nullptr;
    }
    if (device < 0 || device >= ndev || cudaSetDevice(device) != cudaSuccess) { set_error("bad device %d", device); return nullptr; }
    b200_s3mgau *m = new (std::nothrow) b200_s3mgau();
    if (!m) return nullptr;
    m->device = device; m->n_sen = n_sen; m->n_ci = n_ci_sen; m->max_comp = n_comp; m->veclen = veclen;
    m->logbase = logbase;
    m->lm = new LogMath(logbase, 0, true);
    const size_t S = n_sen, M = n_comp, D = veclen;
    m->h_mean.assign(mean, mean + S * M * D); m->h_var.assign(var, var + S * M * D);
    m->h_lrd.assign(S * M, 0.f); m->h_mixw.assign(S * M, 0); m->h_ncomp.assign(S, 0);
    m->h_cd2ci.assign(cd2cisen, cd2cisen + S);
    // ---- mgau_mixw_read (cont_mgau.c:624-668): floor non-zero, normalise, logs3
    std::vector<float> pdf(M);
    for (size_t s = 0; s < S; ++s) {
        std::copy(mixw + s * M, mixw + (s + 1) * M, pdf.begin());
        bool zero = true;
        for (size_t c = 0; c < M; ++c) if (pdf[c] != 0.0f) { zero = false; break; }
        if (zero) { for (size_t c = 0; c < M; ++c) m->h_mixw[s * M + c] = kS3Zero; continue; }
        for (size_t c = 0; c < M; ++c) if (pdf[c] != 0.0 && pdf[c] < mixwfloor) pdf[c] = (float)mixwfloor;
        double sum = 0.0;
        for (size_t c = 0; c < M; ++c) sum += pdf[c];
        if (sum != 0.0) { double f = 1.0 / sum; for (size_t c = 0; c < M; ++c) pdf[c] = (float)((double)pdf[c] * f); }
        for (size_t c = 0; c < M; ++c)
            m->h_mixw[s * M + c] = (pdf[c] != 0.0 && pdf[c] > 0.0) ? m->lm->log(pdf[c]) : kS3Zero;
    }
    // ---- mgau_uninit_compact (:700-790), mgau_var_floor (:798-825), mgau_precomp (:852-894)
    auto is_nan = [&](const float *v) { for (size_t i = 0; i < D; ++i) if (std::isnan(v[i])) return true; return false; };
    auto is_zero = [&](const float *v) { for (size_t i = 0; i < D; ++i) if (v[i] != 0.0f) return false; return true; };
    for (size_t s = 0; s < S; ++s) {
        size_t c2 = 0;
        for (size_t c = 0; c < M; ++c) {
            float *mu = &m->h_mean[(s * M + c) * D], *va = &m->h_var[(s * M + c) * D];
            if (is_nan(mu) || is_nan(va) || is_zero(va)) continue;
            if (c2 != c) {
                std::memcpy(&m->h_mean[(s * M + c2) * D], mu, D * sizeof(float));
                std::memcpy(&m->h_var[(s * M + c2) * D], va, D * sizeof(float));
                m->h_mixw[s * M + c2] = m->h_mixw[s * M + c];
            }
            ++c2;
        }
        m->h_ncomp[s] = (int32_t)c2;
        for (size_t c = 0; c < c2; ++c) {
            float *va = &m->h_var[(s * M + c) * D];
            double lrd = 0.0;
            if (varfloor > 0.0) for (size_t i = 0; i < D; ++i) if (va[i] < varfloor) va[i] = (float)varfloor;
            for (size_t i = 0; i < D; ++i) {
                lrd += std::log((double)va[i]);
                va[i] = (float)(1.0 / (va[i] * 2.0));
            }
            lrd += (double)veclen * std::log(2.0 * M_PI);
            m->h_lrd[s * M + c] = (float)(-0.5 * lrd);
        }
    }
    m->distfloor = (double)kS3Zero * m->lm->log_of_base;   // logmath_log_to_ln
    m->f = 1.0 / std::log(logbase);
    m->ci_pbeam = m->lm->log(1e-80);
    // ---- device layout: components innermost, padded to cp*kc
    m->cp = std::min(32, pow2ceil(n_comp));
    m->kc = n_comp <= 32 ? 1 : (n_comp <= 64 ? 2 : 4);
    m->cpt = m->cp * m->kc;
    const size_t CPT = m->cpt;
    std::vector<float> t_mean(S * D * CPT, 0.f), t_lrd(S * CPT, 0.f);
    std::vector<double> t_var(S * D * CPT, 0.0);
    std::vector<int32_t> t_mixw(S * CPT, kS3Zero);
    for (size_t s = 0; s < S; ++s)
        for (size_t c = 0; c < (size_t)m->h_ncomp[s]; ++c) {
            for (size_t i = 0; i < D; ++i) {
                t_mean[(s * D + i) * CPT + c] = m->h_mean[(s * M + c) * D + i];
                t_var[(s * D + i) * CPT + c] = (double)m->h_var[(s * M + c) * D + i];
            }
            t_lrd[s * CPT + c] = m->h_lrd[s * M + c];
            t_mixw[s * CPT + c] = m->h_mixw[s * M + c];
        }
    auto up = [&](void **dst, const void *src, size_t bytes) {
        return cudaMalloc(dst, bytes) == cudaSuccess && cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    bool ok = up((void **)&m->d_mean, t_mean.data(), t_mean.size() * sizeof(float)) &&
              up((void **)&m->d_var, t_var.data(), t_var.size() * sizeof(double)) &&
              up((void **)&m->d_lrd, t_lrd.data(), t_lrd.size() * sizeof(float)) &&
              up((void **)&m->d_mixw, t_mixw.data(), t_mixw.size() * sizeof(int32_t)) &&
              up((void **)&m->d_ncomp, m->h_ncomp.data(), S * sizeof(int32_t)) &&
              up((void **)&m->d_cd2ci, m->h_cd2ci.data(), S * sizeof(int32_t));
    if (ok) {
        if (m->lm->width <= 2) {
            std::vector<uint16_t> t16(m->lm->table.begin(), m->lm->table.end());
            ok = up((void **)&m->d_tab16, t16.data(), t16.size() * sizeof(uint16_t));
        } else {
            ok = up((void **)&m->d_tab32, m->lm->table.data(), m->lm->table.size() * sizeof(uint32_t));
        }
    }
    ok = ok && cudaMalloc((void **)&m->d_bstidx, S * 4) == cudaSuccess && cudaMalloc((void **)&m->d_update, S * 4) == cudaSuccess &&
         cudaMalloc((void **)&m->d_prev, S * 4) == cudaSuccess && cudaMemset(m->d_prev, 0, S * 4) == cudaSuccess &&
         cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreate(&m->ev[0]) == cudaSuccess && cudaEventCreate(&m->ev[1]) == cudaSuccess;
    if (!ok) { set_error("b200_s3_create: device allocation/upload failed: %s", cudaGetErrorString(cudaGetLastError())); b200_s3_free(m); return nullptr; }
    if (b200_s3_utt_reset(m) != B200_OK) { b200_s3_free(m); return nullptr; }
    return m;
}

b200_s3mgau_t *b200_s3_load(const char *meanfile, const char *varfile, const char *mixwfile, double varfloor,
                            double mixwfloor, double logbase, const int32_t *cd2cisen, int n_ci_sen, int device) {
    int32_t dm[4], dv[4], dw[4], vl[64];
    if (b200_s3_read_gauden(meanfile, dm, vl, nullptr) || b200_s3_read_gauden(varfile, dv, vl, nullptr) ||
        b200_s3_read_mixw(mixwfile, dw, nullptr)) return nullptr;
    if (dm[1] != 1) { set_error("#Features streams(%d) != 1 in continuous HMM", dm[1]); return nullptr; }   // cont_mgau.c:561-565
    if (dm[0] != dv[0] || dm[2] != dv[2] || dm[3] != dv[3]) { set_error("means/variances dimensions differ (full covariances are not supported)"); return nullptr; }
    if (dw[0] != dm[0] || dw[1] != 1 || dw[2] != dm[2]) { set_error("mixture weights do not match the Gaussians"); return nullptr; }
    std::vector<float> mean(dm[3]), var(dv[3]), mixw(dw[3]);
    if (b200_s3_read_gauden(meanfile, dm, vl, mean.data()) || b200_s3_read_gauden(varfile, dv, vl, var.data()) ||
        b200_s3_read_mixw(mixwfile, dw, mixw.data())) return nullptr;
    return b200_s3_create(dm[0], dm[2], vl[0], mean.data(), var.data(), mixw.data(), varfloor, mixwfloor, logbase,
                          cd2cisen, n_ci_sen, device);
}

void b200_s3_free(b200_s3mgau_t *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    void *ptrs[] = {m->d_mean, m->d_var, m->d_lrd, m->d_mixw, m->d_ncomp, m->d_cd2ci, m->d_tab16, m->d_tab32, m->d_bstidx,
                    m->d_update, m->d_prev, m->d_feat, m->d_act, m->d_flags, m->d_raw, m->d_out, m->d_beam, m->d_best, m->d_bst,
                    m->d_svq_map, m->d_svq_i, m->d_svq_f, m->d_svq_var, m->d_vqd, m->d_gs_cw, m->d_gs_map, m->d_gscw};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (m->st) cudaStreamDestroy(m->st);
    for (auto &e : m->ev) if (e) cudaEventDestroy(e);
    delete m->lm;
    delete m;
}

int b200_s3_dims(const b200_s3mgau_t *m, int32_t dims[5]) {
    if (!m || !dims) { set_error("null argument"); return B200_ERR_ARG; }
    dims[0] = m->n_sen; dims[1] = m->max_comp; dims[2] = m->veclen; dims[3] = m->n_ci; dims[4] = m->ci_pbeam;
    return B200_OK;
}

int b200_s3_set_fast(b200_s3mgau_t *m, double ci_pbeam, int max_cd, int ds_ratio, float tighten_factor) {
    if (!m || ds_ratio < 1) { set_error("b200_s3_set_fast: bad argument"); return B200_ERR_ARG; }
    m->ci_pbeam = ci_pbeam <= 0.0 ? kS3Zero : m->lm->log(ci_pbeam);   // logs3(), S3/libcommon/logs3.c:110-118
    m->max_cd = max_cd; m->ds_ratio = ds_ratio; m->tighten = tighten_factor;
    return B200_OK;
}

// The same from the values fast_gmm_t already holds (fast_algo_struct.h:204-262: the beams are stored as
// logs3 integers) -- what a binding inside the decoder has at hand.  svq_beam_log <= 0 is kept as is.
int b200_s3_set_fast_log(b200_s3mgau_t *m, int32_t ci_pbeam_log, int max_cd, int ds_ratio, float tighten_factor,
                         int32_t subvqbeam_log) {
    if (!m || ds_ratio < 1) { set_error("b200_s3_set_fast_log: bad argument"); return B200_ERR_ARG; }
    m->ci_pbeam = ci_pbeam_log; m->max_cd = max_cd; m->ds_ratio = ds_ratio; m->tighten = tighten_factor;
    m->svq_beam = subvqbeam_log;
    return B200_OK;
}

// -subvq FILE, -svmax, -vqeval, -subvqbeam: subvq_init (S3/libam/subvq.c:206-373) -- the text file gausubvq
// writes -- then subvq_maha_precomp (:106-123: variance floor, vector_maha_precomp), subvq_map_compact
// (:127-181) and subvq_map_linearize (:191-203); the beam goes through logs3 as fast_gmm_init does
// (fast_algo_struct.c:454).  file == NULL switches the layer off again.
int b200_s3_set_subvq(b200_s3mgau_t *m, const char *file, double varfloor, int max_sv, int vqeval, double subvqbeam) {
    if (!m) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(m->device));
    B200_CUDA_OK(cudaStreamSynchronize(m->st));
    cudaFree(m->d_svq_map); cudaFree(m->d_svq_i); cudaFree(m->d_svq_f); cudaFree(m->d_svq_var);
    m->d_svq_map = m->d_svq_i = nullptr; m->d_svq_f = nullptr; m->d_svq_var = nullptr; m->svq_n_sv = 0;
    if (!file) return B200_OK;
    FILE *fp = fopen(file, "r");
    if (!fp) { set_error("cannot open sub-VQ file %s", file); return B200_ERR_IO; }
    std::vector<char> line(1 << 20);
    auto next = [&]() { return fgets(line.data(), (int)line.size(), fp) != nullptr; };
    auto fail = [&](const char *what) { fclose(fp); set_error("sub-VQ file %s: %s", file, what); return B200_ERR_IO; };
    int R = 0, Cc = 0, n_sv_file = 0, size = 0;
    for (;;) {
        if (!next()) return fail("no VQParam header");
        if (sscanf(line.data(), "VQParam %d %d -> %d %d", &R, &Cc, &n_sv_file, &size) == 4) break;
    }
    if (R != m->n_sen || Cc != m->max_comp) { fclose(fp); set_error("Model size conflict: %d x %d (SubVQ) vs %d x %d (Original)", R, Cc, m->n_sen, m->max_comp); return B200_ERR_ARG; }
    if (n_sv_file < 1 || size < 1) return fail("bad VQParam header");
    int n_sv = (max_sv < 0 || max_sv > n_sv_file) ? n_sv_file : max_sv;
    if (n_sv < 1) return fail("no sub-vector left (-svmax)");
    std::vector<int32_t> veclen(n_sv_file), off_dim(n_sv_file + 1, 0), off_par(n_sv_file + 1, 0), featdim;
    for (int sv = 0; sv < n_sv_file; ++sv) {
        int k = -1, l = 0, n = 0;
        if (!next() || sscanf(line.data(), "Subvector %d length %d%n", &k, &l, &n) != 2 || k != sv || l < 1) return fail("sub-vector header");
        veclen[sv] = l;
        const char *sp = line.data() + n;
        for (int c = 0; c < l; ++c) {
            int d = 0, adv = 0;
            if (sscanf(sp, "%d%n", &d, &adv) != 1 || d < 0 || d >= m->veclen) return fail("sub-vector dimension");
            featdim.push_back(d); sp += adv;
        }
        off_dim[sv + 1] = off_dim[sv] + l; off_par[sv + 1] = off_par[sv] + l * size;
    }
    std::vector<float> mean(off_par[n_sv_file]), lrd((size_t)n_sv_file * size);
    std::vector<float> varf(off_par[n_sv_file]);
    std::vector<int32_t> map((size_t)R * Cc * n_sv_file);
    for (int sv = 0; sv < n_sv_file; ++sv) {
        int k = -1;
        if (!next() || sscanf(line.data(), "Codebook %d", &k) != 1 || k != sv) return fail("codebook header");
        for (int r = 0; r < size; ++r) {
            if (!next()) return fail("codebook row");
            const char *sp = line.data();
            for (int c = 0; c < veclen[sv]; ++c) {
                int adv = 0;
                if (sscanf(sp, "%f %f%n", &mean[off_par[sv] + (size_t)r * veclen[sv] + c], &varf[off_par[sv] + (size_t)r * veclen[sv] + c], &adv) != 2) return fail("codebook entry");
                sp += adv;
            }
        }
        if (!next() || sscanf(line.data(), "Map %d", &k) != 1 || k != sv) return fail("map header");
        for (int r = 0; r < R; ++r) {
            if (!next()) return fail("map row");
            const char *sp = line.data();
            for (int c = 0; c < Cc; ++c) {
                int v = 0, adv = 0;
                if (sscanf(sp, "%d%n", &v, &adv) != 1 || v >= size) return fail("map entry");
                map[((size_t)r * Cc + c) * n_sv_file + sv] = v; sp += adv;
            }
        }
    }
    char tok[64] = "";
    if (fscanf(fp, "%63s", tok) != 1 || strcmp(tok, "End") != 0) return fail("no End token");
    fclose(fp);
    // precompute (only the sub-vectors in use matter)
    std::vector<double> var(off_par[n_sv]);
    for (int sv = 0; sv < n_sv; ++sv)
        for (int r = 0; r < size; ++r) {
            float *v = &varf[off_par[sv] + (size_t)r * veclen[sv]];
            double det = 0.0;
            for (int i = 0; i < veclen[sv]; ++i) if (v[i] < varfloor) v[i] = (float)varfloor;
            for (int i = 0; i < veclen[sv]; ++i) { det -= std::log((double)v[i]); v[i] = (float)(1.0 / (v[i] * 2.0)); }
            det -= std::log(2.0 * M_PI) * veclen[sv];
            lrd[(size_t)sv * size + r] = (float)(det * 0.5);
            for (int i = 0; i < veclen[sv]; ++i) var[off_par[sv] + (size_t)r * veclen[sv] + i] = (double)v[i];
        }
    // compact + linearise, in the device layout [s][cpt][n_sv]
    std::vector<int32_t> dmap((size_t)R * m->cpt * n_sv, -1);
    for (int r = 0; r < R; ++r) {
        int c2 = 0;
        for (int c = 0; c < Cc; ++c) {
            const int32_t *src = &map[((size_t)r * Cc + c) * n_sv_file];
            if (src[0] < 0) {
                for (int sv = 1; sv < n_sv; ++sv) if (src[sv] >= 0) { set_error("Partially undefined map[%d][%d]", r, c); return B200_ERR_ARG; }
                continue;
            }
            for (int sv = 0; sv < n_sv; ++sv) {
                if (src[sv] < 0) { set_error("Partially undefined map[%d][%d]", r, c); return B200_ERR_ARG; }
                dmap[((size_t)r * m->cpt + c2) * n_sv + sv] = sv * size + src[sv];
            }
            ++c2;
        }
        if (c2 != m->h_ncomp[r]) { set_error("Mixture %d: #Valid components conflict: %d (SubVQ) vs %d (Original)", r, c2, m->h_ncomp[r]); return B200_ERR_ARG; }
    }
    std::vector<int32_t> ints;
    ints.insert(ints.end(), veclen.begin(), veclen.begin() + n_sv);
    ints.insert(ints.end(), off_dim.begin(), off_dim.begin() + n_sv);
    ints.insert(ints.end(), off_par.begin(), off_par.begin() + n_sv);
    ints.insert(ints.end(), featdim.begin(), featdim.begin() + off_dim[n_sv]);
    std::vector<float> flt(mean.begin(), mean.begin() + off_par[n_sv]);
    flt.insert(flt.end(), lrd.begin(), lrd.begin() + (size_t)n_sv * size);
    auto up = [&](void **dst, const void *src, size_t bytes) {
        return cudaMalloc(dst, bytes) == cudaSuccess && cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    if (!(up((void **)&m->d_svq_map, dmap.data(), dmap.size() * 4) && up((void **)&m->d_svq_i, ints.data(), ints.size() * 4) &&
          up((void **)&m->d_svq_f, flt.data(), flt.size() * 4) && up((void **)&m->d_svq_var, var.data(), var.size() * 8))) {
        set_error("sub-VQ upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return B200_ERR_CUDA;
    }
    m->svq_n_sv = n_sv; m->svq_size = size; m->svq_eval = std::min(n_sv, vqeval);
    m->svq_beam = subvqbeam <= 0.0 ? kS3Zero : m->lm->log(subvqbeam);
    S3Vq &q = m->vq;
    q.n_sv = n_sv; q.size = size; q.n_eval = m->svq_eval;
    q.veclen = m->d_svq_i; q.off_dim = m->d_svq_i + n_sv; q.off_par = m->d_svq_i + 2 * n_sv; q.featdim = m->d_svq_i + 3 * n_sv;
    q.mean = m->d_svq_f; q.lrd = m->d_svq_f + off_par[n_sv]; q.var = m->d_svq_var;
    q.distfloor = m->distfloor; q.f = m->f;
    return B200_OK;
}

// -gs FILE: gs_read (S3/libam/gs.c:156-218) -- five int32 (n_mgau, n_feat, n_density, n_code, featlen), then per
// codeword its featlen float32 and one bit vector per mixture, of which the reference keeps the first 32-bit word.
// file == NULL switches the layer off.  With both layers set the selector wins (approx_mgau_eval, gs4gs).
int b200_s3_set_gs(b200_s3mgau_t *m, const char *file) {
    if (!m) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(m->device));
    B200_CUDA_OK(cudaStreamSynchronize(m->st));
    cudaFree(m->d_gs_cw); cudaFree(m->d_gs_map); m->d_gs_cw = nullptr; m->d_gs_map = nullptr; m->gs_n_code = 0;
    if (!file) return B200_OK;
    FILE *fp = fopen(file, "rb");
    if (!fp) { set_error("cannot open Gaussian-selector map %s", file); return B200_ERR_IO; }
    int32_t hd[5];
    if (fread(hd, 4, 5, fp) != 5) { fclose(fp); set_error("%s: short header", file); return B200_ERR_IO; }
    const int n_mgau = hd[0], n_feat = hd[1], n_density = hd[2], n_code = hd[3], featlen = hd[4];
    if (n_mgau != m->n_sen || n_feat != 1 || n_density != m->max_comp || featlen != m->veclen || n_code < 1) {
        fclose(fp);
        set_error("%s: %d mixtures x %d streams x %d densities, feature length %d do not match the model", file, n_mgau, n_feat, n_density, featlen);
        return B200_ERR_ARG;
    }
    if (n_density > 32) { fclose(fp); set_error("Gaussian selector: more than 32 densities (the reference keeps one 32-bit word per map)"); return B200_ERR_UNSUP; }
    if (n_code & 1) { fclose(fp); set_error("Gaussian selector: odd codeword count (gc_compute_closest_cw walks the codewords in pairs)"); return B200_ERR_UNSUP; }
    const size_t words = (size_t)(n_density + 31) / 32;
    std::vector<float> cw((size_t)n_code * featlen);
    std::vector<uint32_t> map((size_t)n_mgau * n_code), row((size_t)n_mgau * words);
    for (int k = 0; k < n_code; ++k) {
        if (fread(&cw[(size_t)k * featlen], 4, featlen, fp) != (size_t)featlen || fread(row.data(), 4, row.size(), fp) != row.size()) {
            fclose(fp); set_error("%s: truncated", file); return B200_ERR_IO;
        }
        for (int s = 0; s < n_mgau; ++s) map[(size_t)s * n_code + k] = row[(size_t)s * words];
    }
    fclose(fp);
    if (cudaMalloc((void **)&m->d_gs_cw, cw.size() * 4) != cudaSuccess || cudaMalloc((void **)&m->d_gs_map, map.size() * 4) != cudaSuccess ||
        cudaMemcpy(m->d_gs_cw, cw.data(), cw.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(m->d_gs_map, map.data(), map.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("Gaussian selector upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return B200_ERR_CUDA;
    }
    m->gs_n_code = n_code; m->gs_featlen = featlen;
    return B200_OK;
}

int b200_s3_utt_reset(b200_s3mgau_t *m) {
    if (!m) { set_error("null argument"); return B200_ERR_ARG; }
    cudaSetDevice(m->device);
    std::vector<int32_t> a(m->n_sen, kNoBst), b(m->n_sen, kNotUpdated);
    B200_CUDA_OK(cudaMemcpy(m->d_bstidx, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(m->d_update, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    return B200_OK;
}

int b200_s3_params(const b200_s3mgau_t *m, int32_t *n_comp, float *mean, float *var, float *lrd, int32_t *mixw,
                   double scal[2]) {
    if (!m) { set_error("null argument"); return B200_ERR_ARG; }
    if (n_comp) std::copy(m->h_ncomp.begin(), m->h_ncomp.end(), n_comp);
    if (mean) std::copy(m->h_mean.begin(), m->h_mean.end(), mean);
    if (var) std::copy(m->h_var.begin(), m->h_var.end(), var);
    if (lrd) std::copy(m->h_lrd.begin(), m->h_lrd.end(), lrd);
    if (mixw) std::copy(m->h_mixw.begin(), m->h_mixw.end(), mixw);
    if (scal) { scal[0] = m->distfloor; scal[1] = m->f; }
    return B200_OK;
}

int b200_s3_state(b200_s3mgau_t *m, int32_t *bstidx, int32_t *updatetime) {
    if (!m || !bstidx || !updatetime) { set_error("null argument"); return B200_ERR_ARG; }
    cudaSetDevice(m->device);
    B200_CUDA_OK(cudaStreamSynchronize(m->st));
    B200_CUDA_OK(cudaMemcpy(bstidx, m->d_bstidx, (size_t)m->n_sen * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(updatetime, m->d_update, (size_t)m->n_sen * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}

int b200_s3_dense_dev(b200_s3mgau_t *m, const float *d_feat, int T, int32_t *d_out, void *stream) {
    if (!m || !d_feat || !d_out || T < 0) { set_error("b200_s3_dense_dev: bad argument"); return B200_ERR_ARG; }
    cudaSetDevice(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    B200_CUDA_OK(cudaEventRecord(m->ev[0], st));
    int rc = launch_eval(m, d_feat, T, 0, m->n_sen, nullptr, d_out, nullptr, st);
    if (rc) return rc;
    B200_CUDA_OK(cudaEventRecord(m->ev[1], st));
    return B200_OK;
}

int b200_s3_dense_host(b200_s3mgau_t *m, const float *feat, int T, int32_t *out) {
    if (!m || !feat || !out || T < 0) { set_error("b200_s3_dense_host: bad argument"); return B200_ERR_ARG; }
    cudaSetDevice(m->device);
    for (int t0 = 0; t0 < T; t0 += kChunkT) {
        const int n = std::min(kChunkT, T - t0);
        int rc = s3_reserve(m, n);
        if (rc) return rc;
        B200_CUDA_OK(cudaMemcpyAsync(m->d_feat, feat + (size_t)t0 * m->veclen, (size_t)n * m->veclen * 4, cudaMemcpyHostToDevice, m->st));
        if ((rc = b200_s3_dense_dev(m, m->d_feat, n, m->d_out, m->st))) return rc;
        B200_CUDA_OK(cudaMemcpyAsync(out + (size_t)t0 * m->n_sen, m->d_out, (size_t)n * m->n_sen * 4, cudaMemcpyDeviceToHost, m->st));
        B200_CUDA_OK(cudaStreamSynchronize(m->st));
    }
    return B200_OK;
}

int b200_s3_score_utt_dev(b200_s3mgau_t *m, const float *d_feat, int T, int frame0, uint8_t *d_sen_active,
                          int32_t *d_out, int32_t *d_best, void *stream) {
    if (!m || !d_feat || !d_out || !d_best || T < 0) { set_error("b200_s3_score_utt_dev: bad argument"); return B200_ERR_ARG; }
    cudaSetDevice(m->device);
    cudaStream_t st = (cudaStream_t)stream;
    B200_CUDA_OK(cudaEventRecord(m->ev[0], st));
    for (int t0 = 0; t0 < T; t0 += kChunkT) {
        const int n = std::min(kChunkT, T - t0);
        int rc = s3_reserve(m, n);
        if (rc) return rc;
        rc = s3_chunk_dev(m, d_feat + (size_t)t0 * m->veclen, n, frame0 + t0,
                          d_sen_active ? d_sen_active + (size_t)t0 * m->n_sen : nullptr,
                          d_out + (size_t)t0 * m->n_sen, d_best + t0, st);
        if (rc) return rc;
    }
    B200_CUDA_OK(cudaEventRecord(m->ev[1], st));
    return B200_OK;
}

int b200_s3_score_utt_host(b200_s3mgau_t *m, const float *feat, int T, int frame0, uint8_t *sen_active,
                           int32_t *senscr_io, int32_t *out, int32_t *best) {
    if (!m || !feat || !out || !best || T < 0) { set_error("b200_s3_score_utt_host: bad argument"); return B200_ERR_ARG; }
    cudaSetDevice(m->device);
    const size_t S = m->n_sen;
    if (senscr_io) B200_CUDA_OK(cudaMemcpyAsync(m->d_prev, senscr_io, S * 4, cudaMemcpyHostToDevice, m->st));
    for (int t0 = 0; t0 < T; t0 += kChunkT) {
        const int n = std::min(kChunkT, T - t0);
        int rc = s3_reserve(m, n);
        if (rc) return rc;
        B200_CUDA_OK(cudaMemcpyAsync(m->d_feat, feat + (size_t)t0 * m->veclen, (size_t)n * m->veclen * 4, cudaMemcpyHostToDevice, m->st));
        if (sen_active)
            B200_CUDA_OK(cudaMemcpyAsync(m->d_act, sen_active + (size_t)t0 * S, (size_t)n * S, cudaMemcpyHostToDevice, m->st));
        if ((rc = s3_chunk_dev(m, m->d_feat, n, frame0 + t0, sen_active ? m->d_act : nullptr, m->d_out, m->d_best, m->st))) return rc;
        B200_CUDA_OK(cudaMemcpyAsync(out + (size_t)t0 * S, m->d_out, (size_t)n * S * 4, cudaMemcpyDeviceToHost, m->st));
        B200_CUDA_OK(cudaMemcpyAsync(best + t0, m->d_best, (size_t)n * 4, cudaMemcpyDeviceToHost, m->st));
        if (sen_active)
            B200_CUDA_OK(cudaMemcpyAsync(sen_active + (size_t)t0 * S, m->d_act, (size_t)n * S, cudaMemcpyDeviceToHost, m->st));
        B200_CUDA_OK(cudaStreamSynchronize(m->st));
    }
    if (senscr_io) {
        B200_CUDA_OK(cudaMemcpyAsync(senscr_io, m->d_prev, S * 4, cudaMemcpyDeviceToHost, m->st));
        B200_CUDA_OK(cudaStreamSynchronize(m->st));
    }
    return B200_OK;
}

int b200_s3_frame_eval(b200_s3mgau_t *m, const float *feat, int32_t frame, uint8_t *sen_active, int32_t *senscr,
                       int32_t *best) {
    if (!m || !feat || !sen_active || !senscr || !best) { set_error("b200_s3_frame_eval: null argument"); return B200_ERR_ARG; }
    std::vector<int32_t> row(m->n_sen);
    int rc = b200_s3_score_utt_host(m, feat, 1, frame, sen_active, senscr, row.data(), best);
    return rc;
}

double b200_fp64_issue_rate(int device) {
    if (cudaSetDevice(device) != cudaSuccess) { set_error("bad device %d", device); return -1.0; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    const int blocks = prop.multiProcessorCount * 8, iters = 1 << 14;
    double *d = nullptr;
    if (cudaMalloc((void **)&d, (size_t)blocks * 256 * sizeof(double)) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        s3_fp64_probe_kernel<<<blocks, 256>>>(d, iters, 1.0000001, 1e-9);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double rate = (double)blocks * 256 * iters * 8 / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return best;
}

float b200_s3_last_ms(b200_s3mgau_t *m) {
    if (!m) return -1.f;
    float ms = 0.f;
    if (cudaEventSynchronize(m->ev[1]) != cudaSuccess || cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]) != cudaSuccess) return -1.f;
    m->last_ms = ms;
    return ms;
}

}  // extern "C"
