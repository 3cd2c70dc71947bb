// gmm_exact.cu -- exact CUDA-core GMM kernels (sm_100a).
//
// These kernels evaluate every Gaussian with the reference's own arithmetic:
// sequential float32 `dval -= diff*diff*var` with each operation rounded
// separately (x86-64 SSE semantics of pocketsphinx/src/libpocketsphinx/
// ms_gauden.c:433-449, ptm_mgau.c:98-231, s2_semi_mgau.c:80-173), so the
// resulting (int32)dval are bit-identical to the CPU's.  They serve
//   * the ms back-end when the tensor-core path does not apply (multi-stream,
//     shared codebooks) and as the on-device exactness check of that path;
//   * the codebook stage of the ptm / s2_semi back-ends (integer top-N lists).
//
// Mapping: one thread per frame, a block of 128 frames x one codebook; the
// codebook's (mean, var, det) are staged through shared memory in chunks and
// read as warp-wide broadcasts; the frame's features sit in shared memory
// transposed ([dim][frame]) so the per-dimension read is conflict free.  The
// per-thread top-N list lives in registers (N is a template parameter).
#include "gmm_dev.cuh"

namespace b200 {

constexpr int kFramesPerBlock = 128;
constexpr int kDensChunk = 64;

template <int N>
struct TopF {  // ms: float distances, ids
    float v[N];
    int id[N];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < N; ++j) { v[j] = (float)kWorstDistI; id[j] = 0; }
    }
    // ms_gauden.c:506-520: insert if dval >= worst; the new entry goes ahead
    // of equal ones.
    __device__ __forceinline__ void offer(float d, int idx) {
        if (!(d >= v[N - 1])) return;
#pragma unroll
        for (int j = N - 1; j >= 0; --j) {
            if (j > 0 && d >= v[j - 1]) { v[j] = v[j - 1]; id[j] = id[j - 1]; }
            else { v[j] = d; id[j] = idx; break; }
        }
    }
};

// Exact sequential distance of frame (column `tid` of xs) to one density.
__device__ __forceinline__ float seq_dist(const float *xs, int tid, const float *m,
                                          const float *v, float det, int len) {
    float d = det;
    for (int i = 0; i < len; ++i) {
        float diff = __fsub_rn(xs[i * kFramesPerBlock + tid], m[i]);
        d = __fsub_rn(d, __fmul_rn(__fmul_rn(diff, diff), v[i]));
    }
    return d;
}

// MODE 0: ms (float lists; fused senone score when `fused`), 1: ptm, 2: s2_semi.
template <int N, int MODE>
__global__ void __launch_bounds__(kFramesPerBlock)
gmm_topn_kernel(GmmDev g, const float *__restrict__ feat, int T, int t0,
                int2 *__restrict__ lists,      // [Tchunk][n_mgau*n_feat][N] (ids/cw, score bits)
                int16_t *__restrict__ raw_out, // fused: [T][n_sen]
                int fused) {
    extern __shared__ float smem[];
    const int tid = threadIdx.x;
    const int mg = blockIdx.x;
    const int t = t0 + blockIdx.y * kFramesPerBlock + tid;
    const bool live = t < T;
    float *xs = smem;                                    // [maxlen][128]
    float *pm = xs + g.maxlen * kFramesPerBlock;         // [chunk][len]
    float *pv = pm + kDensChunk * g.maxlen;
    float *pd = pv + kDensChunk * g.maxlen;              // [chunk]
    __shared__ uint8_t s_tab[256];
    if (fused)
        for (int i = tid; i < 256; i += blockDim.x) s_tab[i] = g.logadd[i];

    int32_t scr = 0;  // fused senone score accumulator
    for (int f = 0; f < g.n_feat; ++f) {
        const int len = g.featlen[f];
        __syncthreads();
        // stage this stream's features, transposed
        for (int i = 0; i < len; ++i)
            xs[i * kFramesPerBlock + tid] = live ? feat[(size_t)t * g.veclen + g.featoff[f] + i] : 0.f;
        const size_t pbase = (size_t)mg * g.n_density * g.veclen + (size_t)g.n_density * g.featoff[f];
        const size_t dbase = ((size_t)mg * g.n_feat + f) * g.n_density;

        TopF<N> tf;
        TopI<N> ti;
        if (MODE == 0) tf.init(); else ti.init();

        for (int c0 = 0; c0 < g.n_density; c0 += kDensChunk) {
            const int nc = min(kDensChunk, g.n_density - c0);
            __syncthreads();
            for (int i = tid; i < nc * len; i += blockDim.x) {
                pm[i] = g.mean[pbase + (size_t)c0 * len + i];
                pv[i] = g.var[pbase + (size_t)c0 * len + i];
            }
            for (int i = tid; i < nc; i += blockDim.x) pd[i] = g.det[dbase + c0 + i];
            __syncthreads();
            if (MODE != 0 && c0 == 0) {
                // eval_topn on the seed codewords 0..N-1 (frame-0 semantics)
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    float d = seq_dist(xs, tid, pm + i * len, pv + i * len, pd[i], len);
                    ti.seed(i, (int32_t)d);
                }
            }
            for (int c = 0; c < nc; ++c) {
                const int cw = c0 + c;
                float d = seq_dist(xs, tid, pm + c * len, pv + c * len, pd[c], len);
                if (MODE == 0) {
                    tf.offer(d, cw);
                } else {
                    if (cw < N) continue;  // still present from the seed pass
                    const int32_t worst = ti.s[N - 1];
                    if (MODE == 1) { if (d < (float)worst) continue; }
                    else { if ((int32_t)d < worst) continue; }
                    if (ti.has(cw)) continue;
                    ti.insert((int32_t)d, cw);
                }
            }
        }

        if (MODE == 0 && g.n_density <= N) {
            // compute_dist_all (ms_gauden.c:417-452): every density, index order
            // -- rebuild the list unsorted.  Only reachable when topn == n_density.
            TopF<N> all;
            all.init();
#pragma unroll
            for (int j = 0; j < N; ++j) {
                all.id[j] = j;
#pragma unroll
                for (int k = 0; k < N; ++k)
                    if (tf.id[k] == j && tf.v[k] != (float)kWorstDistI) all.v[j] = tf.v[k];
            }
            tf = all;
        }

        if (MODE == 0 && fused) {
            // senone_eval (ms_senone.c:372-421) for senone == codebook mg
            const uint8_t *mw = g.mixw_t + (size_t)f * g.n_density * g.n_sen + mg;
            int32_t fden = ((int32_t)tf.v[0] + ((1 << kShift) - 1)) >> kShift;
            int32_t fscr = fden - (int32_t)mw[(size_t)tf.id[0] * g.n_sen];
#pragma unroll
            for (int j = 1; j < N; ++j) {
                fden = ((int32_t)tf.v[j] + ((1 << kShift) - 1)) >> kShift;
                int32_t fw = fden - (int32_t)mw[(size_t)tf.id[j] * g.n_sen];
                fscr = logadd_tab(s_tab, fscr, fw);
            }
            scr -= fscr;
        } else if (live) {
            int2 *o = lists + ((size_t)(t - t0) * g.n_mgau * g.n_feat + (size_t)mg * g.n_feat + f) * N;
#pragma unroll
            for (int j = 0; j < N; ++j)
                o[j] = (MODE == 0) ? make_int2(tf.id[j], __float_as_int(tf.v[j]))
                                   : make_int2(ti.cw[j], ti.s[j]);
        }
    }
    if (fused) {
        scr /= g.aw;
        scr = clamp16(scr);
        if (live) raw_out[(size_t)t * g.n_sen + mg] = (int16_t)scr;
    }
}

// senone_eval for shared-codebook ms models: one thread per senone.
__global__ void __launch_bounds__(128)
ms_senone_kernel(GmmDev g, const int2 *__restrict__ lists, int T, int t0, int N,
                 int16_t *__restrict__ raw_out) {
    __shared__ uint8_t s_tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tab[i] = g.logadd[i];
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = t0 + blockIdx.y;
    if (t >= T) return;
    if (s < g.n_sen) {
        const int mg = g.sen2mgau[s];
        int32_t scr = 0;
        for (int f = 0; f < g.n_feat; ++f) {
            const int2 *l = lists + ((size_t)(t - t0) * g.n_mgau * g.n_feat + (size_t)mg * g.n_feat + f) * N;
            const uint8_t *mw = g.mixw_t + (size_t)f * g.n_density * g.n_sen + s;
            int2 e = l[0];
            int32_t fden = ((int32_t)__int_as_float(e.y) + ((1 << kShift) - 1)) >> kShift;
            int32_t fscr = fden - (int32_t)mw[(size_t)e.x * g.n_sen];
            for (int j = 1; j < N; ++j) {
                e = l[j];
                fden = ((int32_t)__int_as_float(e.y) + ((1 << kShift) - 1)) >> kShift;
                fscr = logadd_tab(s_tab, fscr, fden - (int32_t)mw[(size_t)e.x * g.n_sen]);
            }
            scr -= fscr;
        }
        scr /= g.aw;
        scr = clamp16(scr);
        raw_out[(size_t)t * g.n_sen + s] = (int16_t)scr;
    }
}

// ms_mgau.c:188-204: best = min over senones of the frame; senscr[s] =
// clamp16(senscr[s] - best), in place.  One block per frame: the row is read
// once into shared memory (vectorised), reduced, and written back.
__global__ void __launch_bounds__(256)
ms_normalize_kernel(int16_t *__restrict__ scr, int T, int n_sen) {
    extern __shared__ int16_t s_row[];
    __shared__ int s_best;
    const int tid = threadIdx.x;
    for (int t = blockIdx.x; t < T; t += gridDim.x) {
        int16_t *row = scr + (size_t)t * n_sen;
        if (tid == 0) s_best = 0x7fffffff;
        __syncthreads();
        int32_t b = 0x7fffffff;
        if (((n_sen & 7) == 0) && ((((size_t)row) & 15) == 0)) {
            const int4 *r4 = reinterpret_cast<const int4 *>(row);
            int4 *s4 = reinterpret_cast<int4 *>(s_row);
            for (int i = tid; i < n_sen / 8; i += blockDim.x) {
                int4 v = r4[i];
                s4[i] = v;
                const int16_t *h = reinterpret_cast<const int16_t *>(&v);
#pragma unroll
                for (int k = 0; k < 8; ++k) b = min(b, (int32_t)h[k]);
            }
        } else {
            for (int i = tid; i < n_sen; i += blockDim.x) { int16_t v = row[i]; s_row[i] = v; b = min(b, (int32_t)v); }
        }
        for (int o = 16; o > 0; o >>= 1) b = min(b, __shfl_xor_sync(0xffffffffu, b, o));
        if ((tid & 31) == 0) atomicMin(&s_best, b);
        __syncthreads();
        const int32_t best = s_best;
        if (((n_sen & 7) == 0) && ((((size_t)row) & 15) == 0)) {
            const int4 *s4 = reinterpret_cast<const int4 *>(s_row);
            int4 *r4 = reinterpret_cast<int4 *>(row);
            for (int i = tid; i < n_sen / 8; i += blockDim.x) {
                int4 v = s4[i];
                int16_t *h = reinterpret_cast<int16_t *>(&v);
#pragma unroll
                for (int k = 0; k < 8; ++k) h[k] = (int16_t)clamp16((int32_t)h[k] - best);
                r4[i] = v;
            }
        } else {
            for (int i = tid; i < n_sen; i += blockDim.x) row[i] = (int16_t)clamp16((int32_t)s_row[i] - best);
        }
        __syncthreads();
    }
}

// acmod.c:1219-1271 delta list -> senone ids, decoded by the whole block:
// every thread sums a contiguous slice, a block scan gives slice bases.
// ids must hold n_active entries of shared memory.  Block size 256.
__device__ void decode_active(const uint8_t *__restrict__ active, int n_active, uint16_t *ids) {
    __shared__ int32_t s_part[256];
    const int tid = threadIdx.x;
    const int per = (n_active + 255) / 256;
    const int b = min(n_active, tid * per), e = min(n_active, b + per);
    int32_t sum = 0;
    for (int i = b; i < e; ++i) sum += active[i];
    s_part[tid] = sum;
    __syncthreads();
    if (tid < 32) {
        // 256 partials: each lane scans 8 of them serially, then a warp scan
        int32_t loc[8], run = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { loc[k] = run; run += s_part[tid * 8 + k]; }
        int32_t x = run;
        for (int o = 1; o < 32; o <<= 1) { int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (tid >= o) x += y; }
        const int32_t base = x - run;
#pragma unroll
        for (int k = 0; k < 8; ++k) s_part[tid * 8 + k] = base + loc[k];
    }
    __syncthreads();
    int32_t sen = s_part[tid];
    for (int i = b; i < e; ++i) { sen += active[i]; ids[i] = (uint16_t)sen; }
    __syncthreads();
}

// ms_mgau.c:226-248 for one frame served from a dense raw row: best over the
// ACTIVE senones, out[s] = clamp16(raw[s] - best) for active s (other entries
// of out are left untouched, as the reference leaves them).
__global__ void __launch_bounds__(256)
ms_active_normalize_kernel(const int16_t *__restrict__ raw, const uint8_t *__restrict__ active,
                           int n_active, int16_t *__restrict__ out) {
    extern __shared__ uint16_t s_ids[];
    __shared__ int s_best;
    const int tid = threadIdx.x;
    if (tid == 0) s_best = 0x7fffffff;
    decode_active(active, n_active, s_ids);
    int32_t b = 0x7fffffff;
    for (int i = tid; i < n_active; i += 256) b = min(b, (int32_t)raw[s_ids[i]]);
    for (int o = 16; o > 0; o >>= 1) b = min(b, __shfl_xor_sync(0xffffffffu, b, o));
    if ((tid & 31) == 0) atomicMin(&s_best, b);
    __syncthreads();
    const int32_t best = s_best;
    for (int i = tid; i < n_active; i += 256) {
        const int sidx = s_ids[i];
        out[sidx] = (int16_t)clamp16((int32_t)raw[sidx] - best);
    }
}

// ptm_mgau_codebook_eval tail + ptm_mgau_senone_eval (ptm_mgau.c:259-400) and
// s2_semi's mgau_norm + get_scores_* (s2_semi_mgau.c:189-835) for one frame per
// block.  Optional active list (uint8 deltas) reproduces the non-compallsen
// call: codebook activity, per-stream norm over active codebooks, best over
// active senones.  `semi` selects the s2_semi flavour (no best subtraction).
__global__ void __launch_bounds__(256)
tied_senone_kernel(GmmDev g, const int2 *__restrict__ lists, int T, int t0, int N, int semi,
                   const uint8_t *__restrict__ active, int n_active,
                   int16_t *__restrict__ out) {
    extern __shared__ int32_t sm[];
    const int t = t0 + blockIdx.x;
    if (t >= T) return;
    const int C = g.n_mgau, F = g.n_feat;
    int32_t *l_cw = sm;                       // [C*F*N]
    int32_t *l_sc = l_cw + C * F * N;         // [C*F*N]
    int32_t *norm = l_sc + C * F * N;         // [F]
    int32_t *cb_act = norm + F;               // [C]
    int16_t *row = (int16_t *)(cb_act + C);   // [n_sen]
    uint16_t *ids = (uint16_t *)(row + ((g.n_sen + 1) & ~1));  // [n_active]
    __shared__ uint8_t s_tab[256];
    __shared__ uint8_t s_cb16[16];
    __shared__ int s_best;
    __shared__ int s_neff[B200_MAX_STREAMS];   // mgau_norm's return value per stream (s2_semi -topn_beam)
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += blockDim.x) s_tab[i] = g.logadd[i];
    if (tid < 16) s_cb16[tid] = g.mixw_cb[tid];
    if (tid == 0) s_best = 0x7fffffff;
    const int2 *l = lists + (size_t)(t - t0) * C * F * N;
    for (int i = tid; i < C * F * N; i += blockDim.x) { int2 e = l[i]; l_cw[i] = e.x; l_sc[i] = e.y; }
    for (int i = tid; i < C; i += blockDim.x) cb_act[i] = active ? 0 : 1;
    for (int i = tid; i < g.n_sen; i += blockDim.x) row[i] = 0;
    for (int i = tid; i < F; i += blockDim.x) norm[i] = 0x7fffffff;
    __syncthreads();
    if (active) {
        // ptm_mgau_calc_cb_active (ptm_mgau.c:291-316)
        decode_active(active, n_active, ids);
        for (int i = tid; i < n_active; i += blockDim.x) cb_act[semi ? 0 : g.sen2cb[ids[i]]] = 1;
        __syncthreads();
    }
    // per-stream normaliser: min over active codebooks of top-1 >> 10
    for (int i = tid; i < C * F; i += blockDim.x) {
        const int c = i / F, f = i % F;
        if (cb_act[c]) atomicMin(&norm[f], l_sc[(c * F + f) * N] >> kShift);
    }
    __syncthreads();
    for (int i = tid; i < C * F * N; i += blockDim.x) {
        const int c = i / (F * N), f = (i / N) % F;
        int32_t v;
        if (cb_act[c]) {
            v = -((l_sc[i] >> kShift) - norm[f]);
            if (v > 96) v = 96;
        } else v = 96;  // ptm_mgau.c:352-358
        l_sc[i] = v;
    }
    __syncthreads();
    if (tid < F) {
        // s2_semi_mgau.c:199-206: the list ends at the first score above the stream's beam
        int n = N;
        const int beam = semi ? g.topn_beam[tid] : 0;
        if (beam)
            for (int j = 0; j < N; ++j)
                if (l_sc[tid * N + j] > beam) { n = j; break; }
        s_neff[tid] = n;
    }
    __syncthreads();

    int32_t mybest = 0x7fffffff;
    const int rb = g.row_bytes;
    auto eval = [&](int s) {
        const int c = semi ? 0 : g.sen2cb[s];
        int32_t ascore = 0;
        for (int f = 0; f < F; ++f) {
            const int32_t *cw = l_cw + (c * F + f) * N;
            const int32_t *sc = l_sc + (c * F + f) * N;
            int32_t fden = 0;
            const int nf = s_neff[f];
            for (int j = 0; j < nf; ++j) {
                int32_t mw;
                const uint8_t *r = g.mixw_t + ((size_t)f * g.n_density + cw[j]) * rb;
                if (g.n_clust) {
                    int b = r[s >> 1];
                    mw = s_cb16[(s & 1) ? (b >> 4) : (b & 0x0f)];
                } else mw = r[s];
                fden = (j == 0) ? mw + sc[j] : fast_logadd_neg(s_tab, fden, mw + sc[j]);
            }
            ascore += fden;
        }
        row[s] = (int16_t)ascore;
        mybest = min(mybest, ascore);
    };
    if (active) {
        for (int i = tid; i < n_active; i += blockDim.x) eval(ids[i]);
    } else {
        for (int s = tid; s < g.n_sen; s += blockDim.x) eval(s);
    }
    if (!semi) {
        for (int o = 16; o > 0; o >>= 1) mybest = min(mybest, __shfl_xor_sync(0xffffffffu, mybest, o));
        if ((tid & 31) == 0) atomicMin(&s_best, mybest);
    }
    __syncthreads();
    const int sub = semi ? 0 : s_best;
    for (int i = tid; i < g.n_sen; i += blockDim.x)
        out[(size_t)t * g.n_sen + i] = (int16_t)(row[i] - sub);
}

// s2_semi -ds (s2_semi_mgau.c:176-186): on a frame whose number is not a multiple
// of ds_ratio mgau_dist stops after eval_topn (:80-118) -- the previous frame's N
// codewords re-scored on this frame, each bubbling up past strictly smaller scores.
// The chain back to the last fully evaluated frame is at most ds_ratio - 1 frames
// long, so every down-sampled frame replays it on its own: one thread per (frame,
// codebook, stream).  lists[] holds the batch's frames [t0, t0 + tn); frame t of
// the batch has number frame0 + t.  `carry` is the finished list of frame
// frame0 - 1 for chains that start before the batch.
__global__ void tied_ds_kernel(GmmDev g, const float *__restrict__ feat, int t0, int tn, int frame0, int R,
                               int2 *__restrict__ lists, const int2 *__restrict__ carry) {
    const int CF = g.n_mgau * g.n_feat, N = g.topn;
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (long long)tn * CF) return;
    const int t = t0 + (int)(id / CF), cf = (int)(id % CF);
    const int r = (frame0 + t) % R;
    if (r == 0) return;
    const int c = cf / g.n_feat, f = cf % g.n_feat, len = g.featlen[f];
    const int base = t - r;   // batch index of the last full frame (may precede the batch)
    int2 L[B200_MAX_TOPN];
    const int2 *src = base >= t0 ? lists + ((size_t)(base - t0) * CF + cf) * N : carry + (size_t)cf * N;
    for (int i = 0; i < N; ++i) L[i] = src[i];
    const size_t pbase = (size_t)c * g.n_density * g.veclen + (size_t)g.n_density * g.featoff[f];
    const size_t dbase = ((size_t)c * g.n_feat + f) * g.n_density;
    for (int u = (base >= t0 ? base + 1 : t0); u <= t; ++u) {
        const float *x = feat + (size_t)u * g.veclen + g.featoff[f];
        for (int i = 0; i < N; ++i) {
            const int cw = L[i].x;
            const float *m = g.mean + pbase + (size_t)cw * len, *v = g.var + pbase + (size_t)cw * len;
            float d = g.det[dbase + cw];
            for (int k = 0; k < len; ++k) {
                const float diff = __fsub_rn(x[k], m[k]);
                d = __fsub_rn(d, __fmul_rn(__fmul_rn(diff, diff), v[k]));
            }
            const int2 e = make_int2(cw, (int32_t)d);
            int j = i - 1;
            for (; j >= 0 && e.y > L[j].y; --j) L[j + 1] = L[j];
            L[j + 1] = e;
        }
    }
    int2 *dst = lists + ((size_t)(t - t0) * CF + cf) * N;
    for (int i = 0; i < N; ++i) dst[i] = L[i];
}

// ------------------------------------------------------------ host launchers
int gmm_launch_tied_ds(const GmmDev &g, const float *d_feat, int t0, int tn, int frame0, int ds_ratio, int2 *lists,
                       const int2 *carry, cudaStream_t st) {
    if (ds_ratio <= 1 || tn <= 0) return B200_OK;
    if (t0 != 0 && (frame0 + t0) % ds_ratio != 0) { set_error("-ds: list chunk does not start on a fully evaluated frame"); return B200_ERR_ARG; }
    if (t0 == 0 && frame0 % ds_ratio != 0 && !carry) { set_error("-ds %d: frame %d needs the list of frame %d, which was not scored by this back-end", ds_ratio, frame0, frame0 - 1); return B200_ERR_ARG; }
    const long long n = (long long)tn * g.n_mgau * g.n_feat;
    tied_ds_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(g, d_feat, t0, tn, frame0, ds_ratio, lists, carry);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

template <int MODE>
static int launch_topn(const GmmDev &g, const float *d_feat, int T, int t0, int tn, int2 *lists,
                       int16_t *raw, int fused, cudaStream_t st) {
    dim3 grid(g.n_mgau, (tn + kFramesPerBlock - 1) / kFramesPerBlock);
    size_t sh = ((size_t)g.maxlen * kFramesPerBlock + (size_t)kDensChunk * (2 * g.maxlen + 1)) * sizeof(float);
#define B200_TOPN_CASE(NN)                                                                      \
    case NN:                                                                                    \
        gmm_topn_kernel<NN, MODE><<<grid, kFramesPerBlock, sh, st>>>(g, d_feat, min(T, t0 + tn), t0, lists, \
                                                                     raw, fused);               \
        break;
    switch (g.topn) {
        B200_TOPN_CASE(1) B200_TOPN_CASE(2) B200_TOPN_CASE(3) B200_TOPN_CASE(4)
        B200_TOPN_CASE(5) B200_TOPN_CASE(6) B200_TOPN_CASE(7) B200_TOPN_CASE(8)
        default: set_error("topn %d unsupported (1..%d)", g.topn, B200_MAX_TOPN); return B200_ERR_UNSUP;
    }
#undef B200_TOPN_CASE
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int gmm_launch_topn(const GmmDev &g, int mode, const float *d_feat, int T, int t0, int tn, int2 *lists,
                    int16_t *raw, int fused, cudaStream_t st) {
    if (mode == 0) return launch_topn<0>(g, d_feat, T, t0, tn, lists, raw, fused, st);
    if (mode == 1) return launch_topn<1>(g, d_feat, T, t0, tn, lists, raw, fused, st);
    return launch_topn<2>(g, d_feat, T, t0, tn, lists, raw, fused, st);
}

int gmm_launch_ms_senone(const GmmDev &g, const int2 *lists, int T, int t0, int tn, int16_t *raw,
                         cudaStream_t st) {
    dim3 grid((g.n_sen + 127) / 128, tn);
    ms_senone_kernel<<<grid, 128, 0, st>>>(g, lists, min(T, t0 + tn), t0, g.topn, raw);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int gmm_launch_normalize(int16_t *scr, int T, int n_sen, cudaStream_t st) {
    if (T <= 0) return B200_OK;
    size_t sh = ((size_t)n_sen * sizeof(int16_t) + 15) & ~(size_t)15;
    if (sh > 200 * 1024) { set_error("n_sen %d too large for the normalise kernel", n_sen); return B200_ERR_UNSUP; }
    static AttrOnce attr_set;
    if (attr_set.need()) {
        B200_CUDA_OK(cudaFuncSetAttribute(ms_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    int blocks = std::min(T, 148 * 32);
    ms_normalize_kernel<<<blocks, 256, sh, st>>>(scr, T, n_sen);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

size_t gmm_tied_smem(const GmmDev &g, int n_active) {
    size_t n = (size_t)g.n_mgau * g.n_feat * g.topn;
    return (2 * n + g.n_feat + g.n_mgau) * sizeof(int32_t) + (size_t)((g.n_sen + 1) & ~1) * sizeof(int16_t) +
           (size_t)n_active * sizeof(uint16_t) + 16;
}

int gmm_launch_tied_senone(const GmmDev &g, const int2 *lists, int T, int t0, int tn, int semi,
                           const uint8_t *d_active, int n_active, int16_t *out, cudaStream_t st) {
    size_t sh = gmm_tied_smem(g, d_active ? n_active : 0);
    if (sh > 200 * 1024) { set_error("tied senone kernel needs %zu B of shared memory", sh); return B200_ERR_UNSUP; }
    static AttrOnce attr_set;
    if (attr_set.need()) {
        B200_CUDA_OK(cudaFuncSetAttribute(tied_senone_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    tied_senone_kernel<<<tn, 256, sh, st>>>(g, lists, min(T, t0 + tn), t0, g.topn, semi, d_active, n_active, out);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int gmm_launch_ms_active_normalize(const int16_t *raw, const uint8_t *d_active, int n_active,
                                   int16_t *out, cudaStream_t st) {
    if (n_active <= 0) return B200_OK;
    ms_active_normalize_kernel<<<1, 256, (size_t)n_active * 2 + 16, st>>>(raw, d_active, n_active, out);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

}  // namespace b200
