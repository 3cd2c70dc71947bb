// mahal_tc.cu -- tensor-core Mahalanobis scoring for fully-continuous models
// (single stream, one codebook per senone): the K1+K2 kernel of SURVEY.md
// section 2.4 written for sm_100a with tcgen05 / TMEM / bulk-TMA.
//
// Reference arithmetic being reproduced (pocketsphinx/src/libpocketsphinx):
//   ms_gauden.c:417-523  d[t,g] = det[g] - sum_i (x[t,i]-mu[g,i])^2 * v[g,i]
//   ms_senone.c:372-421  top-N of the M densities of a senone,
//                        fden = ((int32)d + 1023) >> 10, table log-add of
//                        (fden - mixw) in descending-d order, negate, /aw,
//                        clamp to int16
// (the per-frame best-score subtraction of ms_mgau.c:188-204 runs as a
// separate streaming pass, gmm_launch_normalize-style, fused with the
// tile-major -> row-major transpose of the scores).
//
// GEMM restatement.  d[t,g] = sum_k A[t,k] * B[g,k] with K = 2D+2 columns
//     k = 0,1      : A = 1             B = (det - sum_i mu^2 v) in two pieces
//     k = 2+2i     : A = x_i^2         B = -v_i
//     k = 3+2i     : A = x_i           B = 2 mu_i v_i
// padded to a multiple of 8.  Scores must be right to about one raw log unit
// in 10^5..10^6, i.e. fp32-class operands: every operand is split into a
// hi + lo pair (2 x 11 significant bits) and the product is formed as
//     Ahi*Bhi + Ahi*Blo + Alo*Bhi          
// with fp32 accumulation in TMEM.  The pairs are fp16 numbers with a
// power-of-two scale per n-tile and K column (kind::f16, K = 16 per MMA, the
// default) or TF32 numbers (kind::tf32, K = 8; the tiles fp16 cannot hold) --
// see tc_score_kernel.  The constant comes first in K so partial sums stay
// near the final magnitude.
//
// Data layout in HBM (all pre-tiled so that every shared-memory stage is ONE
// contiguous bulk copy, in the UMMA K-major no-swizzle canonical layout:
// 16-byte K-chunks, 8-row core matrices, SBO = 128 B, LBO = rows*16 B):
//   B  [n_tile][kstep][hi|lo][chunk 0|1][256 rows][4 f32 | 8 f16]   16 KB / kstep, static
//   X  [m_tile][40 dims][128 rows] f32 -- the RAW features, tiled + transposed
//      (20 KB / frame tile).  The 4x larger [1,1,x^2,x..] hi/lo A operand is
//      built inside the SM: streaming it pre-expanded made the kernel L2-bound
//      (40 GB of L2->SM traffic per 100k-frame step)
//   raw scores, tile-major [n_tile][T_pad][256/M] int16  (coalesced 16 B stores)
//
// Kernel (persistent, 1 CTA / SM, 704 threads):
//   warp 0      bulk-TMA producer: the unit's B tile once (resident, 80 KB fp16 /
//               160 KB TF32), then one raw feature tile per frame tile
//   warps 2..5  A builders: thread = frame row; x^2, scale, hi/lo split
//               written k-step by k-step into a 5-deep A ring
//   warp 1      single-thread tcgen05.mma issuer, 128x256x16 kind::f16 (or
//               128x256x8 kind::tf32), 3 MMAs per k-step, two 256-column TMEM
//               accumulators
//   warps 6..21 epilogue: tcgen05.ld of 64 columns (= two 32-density senones)
//               per thread, accumulator released immediately, integer keys (trunc(d) << log2 M | density id), top-4 by
//               a sort4 + bitonic-merge network in registers, table log-add
//               from shared memory, int16 stores
// A work unit is (n_tile, frame-range); units are ordered so that concurrently
// running CTAs stream the same A tiles (L2 reuse) against different B tiles.
#include "gmm_dev.cuh"

#include <cuda_fp16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace b200 {

namespace {

constexpr int kTileM = 128;           // frames per tile (UMMA M)
constexpr int kTileN = 256;           // Gaussians per tile (UMMA N)
// A ring depth (k-steps).  Chosen so that a tile's k-steps map to ring slots
// statically (slot = j % depth): the MMA issue loop is then fully unrolled with
// compile-time descriptor offsets -- the issuing thread is a serial resource.
constexpr int ring_depth(int ks) { return ks == 10 ? 5 : ks; }
constexpr int kAStageBytes = 2 * 2 * kTileM * 16;   // hi|lo x 2 chunks x 128 rows x 16 B = 8 KB
constexpr int kBStageBytes = 2 * 2 * kTileN * 16;   // 16 KB per k-step
constexpr int kMaxKSteps = 10;        // K <= 80  (D <= 39)
constexpr int kBuildWarps = 4;        // A-operand builders: one thread per frame row
constexpr int kEpiWarps = 16;         // 4 per TMEM lane quarter, 64 accumulator columns each
constexpr int kEpiThreads = kEpiWarps * 32;
// Warp roles.  The SM's issue arbiter prefers the HIGHEST warp id among the eligible warps of a
// scheduler, so the latency-critical roles (A builders -> MMA issuer -> producer) get the top
// ids and the throughput work (epilogue) the low ones.  TMEM lane quarter of a warp = id % 4.
constexpr int kFirstEpiWarp = 0;
constexpr int kFirstBuildWarp = kEpiWarps;
constexpr int kMmaWarp = kEpiWarps + kBuildWarps;
constexpr int kProdWarp = kMmaWarp + 1;
constexpr int kThreads = (2 + kBuildWarps) * 32 + kEpiThreads;
constexpr int kColsPerEpiThread = kTileN / (kEpiWarps / 4);   // 64
constexpr uint32_t kTmemCols = 512;
// MODE 0 (fully continuous) epilogue, round 2: the GEMM delivers w = -32 (d - 1024 mixw)
// and most (frame, senone) pairs are decided by one cheap certificate (see
// tc_score_kernel).  The rest are compacted: every epilogue warp appends the
// accumulator rows of its undecided pairs to a PRIVATE ring in shared memory and,
// whenever the ring holds a warp's worth, runs the full top-4 network on 32 of
// them at once (one item per lane).  No warp waits for another one: the 16 epilogue
// warps are symmetric (an earlier version with 8 dedicated hard-path warps was bound
// by the latency of those few warps: 11.0 ms against 6.2 ms with every pair easy).
constexpr int kQCap = 40;                       // items per warp ring
constexpr int kQPass = 26;                      // run a pass when the ring holds this many (a pass takes up to 32)
constexpr int epi_warps(int) { return kEpiWarps; }
constexpr int score_threads(int) { return (2 + kBuildWarps + kEpiWarps) * 32; }
constexpr int queue_bytes(int M) { return kEpiWarps * kQCap * (M * 4 + 4) + 128; }
constexpr int kWinFden = 42;                    // 29 (log-add table reach) + 11 (three other terms can add up to that) + 2
constexpr float kWinV = (float)kWinFden * 1024.f * 32.f;     // the same window on the -32-scaled accumulator
constexpr float kBig = 1099511627776.f;         // 2^40: sat((th - x) * 2^40) is a crisp 0/1 indicator on the FMA pipe
constexpr float kDeltaV = 64.f * 32.f;          // slack of the "at most 3 better densities" count (64 raw units)
constexpr int kFixRegionsMax = 160;             // one fix-up queue region per CTA
constexpr int kEpsCap = 400;                    // beyond this bound a pair goes to the literal scan
constexpr int kFixChunk = 64;                   // queue A slots a warp reserves at a time

struct TcParams {
    const float *gB;        // pre-tiled B operand
    const float *gX;        // features, tiled + transposed: [m_tile][4*ksteps dims][128 rows]
    const uint8_t *gMixw;   // [n_tiles_n][256] mixture weights in tile row order
    int16_t *raw;           // [n_tiles_n][T_pad][spt]
    int T, T_pad, n_sen, n_tiles_m, n_tiles_n, ksteps, m_chunks, tiles_per_chunk, n_units, aw;
    int m31;                // the constant 31 (63 for tied lists), kept opaque to the compiler (see make_key)
    const float *scaleA;    // fp16 operands: per-tile, per-column power-of-two scale of the A operand [n_tiles_n][16 * ksteps]
    const uint8_t *fmt;     // [n_tiles_n] operand format of every n-tile for THIS batch (1 fp16, 0 TF32); null: TF32 everywhere
    uint4 *part;            // tied mode: [n_tiles_n][T_pad][4] sorted top-4 keys of each 64-column group
    int dbg;                // development knobs (B200_TC_DBG): 1 = skip epilogue math, 2 = one MMA per k-step, 4 = every pair takes the hard path
    // MODE 0, round 2
    const float *gCw;       // [n_tiles_n][256] 32*1024*mixw of every tile row (the offset folded into the B constant)
    uint4 *qa;              // fix-up queue A: [regions][capA] {t, senone, slots 0|1, slots 2|3}
    uint2 *qb;              // fix-up queue B: [capB] {t, senone}
    unsigned *qcnt;         // [1] n_B, [2] overflow, [3] max |GEMM - exact| seen, [4] hard pairs, [5] of those finished in place (tile queue full), [8 + region] n_A
    unsigned capA, capB;
    int eps0, eps_shift;    // bound on |GEMM distance - reference distance|: eps0 + (|d| >> eps_shift) raw units
    float epsA, epsB;       // the same bound for cert_eval, as a fraction of one fden step (1024 raw units)
    uint8_t logadd[256];
};

// ------------------------------------------------------------------ PTX glue
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Makes a value opaque to the compiler: it then stays in its register instead of being
// re-derived (shared-window base + offset, three instructions) at every use in a hot loop.
__device__ __forceinline__ uint32_t keep(uint32_t x) { asm volatile("" : "+r"(x)); return x; }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Spin with a watchdog: a protocol bug must trap, not hang the GPU box.  The loop is three
// instructions (poll, count, branch): polling warps share the issue slots of the warps they wait for.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try(bar, parity)) {
        if (++spins > (1u << 24)) __trap();             // > 1 s
    }
}
// The same for waits that can be long (an epilogue warp that is ahead of the MMA pipeline):
// back off between polls.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try(bar, parity)) {
        __nanosleep(64);
        if (++spins > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred)::"memory");
    return pred != 0;
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor):
// [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout=0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 (1<<4), a=b=TF32
// (2<<7, 2<<10), K-major A and B, N>>3 at bit 17, M>>4 at bit 24.
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileN >> 3) << 17) |
                            ((uint32_t)(kTileM >> 4) << 24);
// same with a = b = F16 (format code 0): kind::f16, K = 16 per instruction
constexpr uint32_t kIdescF16 = (1u << 4) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// ------------------------------------------------------------- top-4 network
// Keys are ints where SMALLER is better (see make_key); lists are ascending.
__device__ __forceinline__ void ce(int32_t &a, int32_t &b) {   // a <= b afterwards
    const int32_t lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}
__device__ __forceinline__ void sort4(int32_t &a, int32_t &b, int32_t &c, int32_t &d) {
    ce(a, b); ce(c, d); ce(a, c); ce(b, d); ce(b, c);
}
// top[0..3] (ascending) <- the 4 smallest of top U {a,b,c,d}
__device__ __forceinline__ void merge4(int32_t (&top)[4], int32_t a, int32_t b, int32_t c, int32_t d) {
    sort4(a, b, c, d);
    int32_t m0 = min(top[0], d), m1 = min(top[1], c), m2 = min(top[2], b), m3 = min(top[3], a);
    ce(m0, m2); ce(m1, m3); ce(m0, m1); ce(m2, m3);   // bitonic merge
    top[0] = m0; top[1] = m1; top[2] = m2; top[3] = m3;
}

// The B operand is stored scaled by -32, so the accumulator holds -32*d (an
// exact power-of-two scaling).  One saturating F2I gives J = trunc(-32 d); its
// five fractional bits are replaced by (31 - density id).  The key then sorts
// ascending by floor(-d) -- for d <= 0 that is -(int32)d, the integer the
// reference keeps (ms_senone.c:392) -- and on equal integers the LATER density
// comes first, as in ms_gauden.c:510-514.  Saturation makes far-away
// Gaussians (|d| > 6.7e7) compare as "worst" instead of wrapping.
// (For d > 0 with a fractional part >= 1/32 the recovered integer is
// trunc(d)+1; it changes fden only when that integer is a multiple of 1024.)
constexpr float kAccScale = 32.0f;
// `m31` is the constant 31 passed as a run-time value so that the compiler
// emits ONE three-input LOP3 ((j | m31) ^ id, id immediate) instead of two.
__device__ __forceinline__ int32_t make_key(uint32_t bits, int id, int32_t m31) {
    return (__float2int_rz(__uint_as_float(bits)) | m31) ^ id;
}

// sphinxbase logmath_add on values that can never reach the table's "zero"
// (fden >= -65536 here): r = max + table[min(|x-y|, 255)], table[255] == 0.
__device__ __forceinline__ int32_t logadd_fast(const uint8_t *tab, int32_t x, int32_t y) {
    const int32_t d = min(abs(x - y), 255);
    return max(x, y) + (int32_t)tab[d];
}

// (Building the key on the FMA pipe instead -- IMAD.HI + IMAD -- saves one ALU op per
// value and was measured slower: 8.95 vs 8.2 ms, see profiles/README.md.)

// senone_eval for one senone from its 4 best keys; mixw_rev[j] = mixw[31 - j].
__device__ __forceinline__ int32_t senone_from_keys(const int32_t (&top)[4], const uint8_t *mixw_rev,
                                                    const uint8_t *tab, int aw) {
    int32_t fw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        // fden = ((int32)d + 1023) >> 10 with (int32)d = -(key >> 5)
        const int32_t fden = (((1 << kShift) - 1) - (top[j] >> 5)) >> kShift;
        fw[j] = fden - (int32_t)mixw_rev[top[j] & 31];
    }
    int32_t fscr = logadd_fast(tab, fw[0], fw[1]);
    fscr = logadd_fast(tab, fscr, fw[2]);
    fscr = logadd_fast(tab, fscr, fw[3]);
    int32_t scr = -fscr;
    if (aw != 1) scr /= aw;
    return clamp16(scr);
}

// ------------------------------------------------ MODE 0 round-2 epilogue parts
// Exactness contract.  The GEMM distance d~ differs from the reference's
// sequentially rounded float32 d (ms_gauden.c:417-523) by at most
// eps(d) = eps0 + (|d| >> eps_shift) raw units (measured, and monitored at run
// time: TcParams::qcnt[3]).  A score is emitted by this kernel only when every
// decision it rests on holds for ALL d within eps of d~:
//   * the set and the order of the top 4 densities (gaps > 2 eps),
//   * fden = ((int32)d + 1023) >> 10 of every density that can change the
//     log-add chain: the chain is evaluated at the lower and at the upper end
//     of every density's interval -- it is monotone in each term, so equal ends
//     pin the true value between them.
// Everything else is queued for tc_fix_a_kernel (re-scores the uncertain
// densities with the reference's float32 arithmetic) or tc_fix_b_kernel
// (near-ties: the literal scan of the senone's M densities).
// slots: 16 bits per top-4 density -- 0|fden - mixw (15 bits) when settled, 1|10 low bits of floor(-d~)|5-bit id when open
struct FixOut { uint32_t kind, s01, s23; };   // kind 0 settled, 1 queue A, 2 queue B

__device__ __forceinline__ void fix_append(const TcParams &p, const FixOut &fx, uint32_t t, uint32_t sen) {
    if (fx.kind == 1) {
        const unsigned region = blockIdx.x < kFixRegionsMax ? blockIdx.x : 0;
        const unsigned i = atomicAdd(p.qcnt + 8 + region, 1u);
        if (i < p.capA) p.qa[(size_t)region * p.capA + i] = make_uint4(t, sen, fx.s01, fx.s23);
        else p.qcnt[2] = 1u;
    } else if (fx.kind == 2) {
        const unsigned i = atomicAdd(p.qcnt + 1, 1u);
        if (i < p.capB) p.qb[i] = make_uint2(t, sen);
        else p.qcnt[2] = 1u;
    }
}

__device__ __forceinline__ int32_t chain4(const uint8_t *tab, const int32_t (&f)[4]) {
    int32_t r = logadd_fast(tab, f[0], f[1]);
    r = logadd_fast(tab, r, f[2]);
    return logadd_fast(tab, r, f[3]);
}

// The full selection for one (frame, senone): w[j] = accumulator bits of density
// j (= -32 (d - 1024 mixw)), cw[j] = 32*1024*mixw_j (16-byte aligned), mixw_rev as
// in senone_from_keys.  Returns the score and what (if anything) must be re-done
// exactly.
template <int M>
__device__ __forceinline__ int32_t hard_eval(const uint32_t (&w)[M], const float *cw, const uint8_t *mixw_rev,
                                             const uint8_t *tab, int aw, int eps0, int eps_shift, int32_t m31,
                                             FixOut &fx) {
    int32_t k[M];
#pragma unroll
    for (int j = 0; j < M; j += 4) {
        const float4 c = *reinterpret_cast<const float4 *>(cw + j);
        k[j] = make_key(__float_as_uint(__fadd_rn(__uint_as_float(w[j]), -c.x)), j, m31);
        k[j + 1] = make_key(__float_as_uint(__fadd_rn(__uint_as_float(w[j + 1]), -c.y)), j + 1, m31);
        k[j + 2] = make_key(__float_as_uint(__fadd_rn(__uint_as_float(w[j + 2]), -c.z)), j + 2, m31);
        k[j + 3] = make_key(__float_as_uint(__fadd_rn(__uint_as_float(w[j + 3]), -c.w)), j + 3, m31);
    }
    int32_t top[4] = {k[0], k[1], k[2], k[3]};
    sort4(top[0], top[1], top[2], top[3]);
#pragma unroll
    for (int g = 4; g < M; g += 4) merge4(top, k[g], k[g + 1], k[g + 2], k[g + 3]);
    // the smallest key above the fourth: keys below it wrap to the top of the unsigned range
    const uint32_t base = (uint32_t)top[3] + 1u;
    uint32_t r5 = min(min((uint32_t)k[0] - base, (uint32_t)k[1] - base), (uint32_t)k[2] - base);
#pragma unroll
    for (int j = 3; j + 1 < M; j += 2) r5 = min(min(r5, (uint32_t)k[j] - base), (uint32_t)k[j + 1] - base);
    if ((M & 1) == 0) r5 = min(r5, (uint32_t)k[M - 1] - base);
    int32_t J[4], lo[4], hi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) J[j] = top[j] >> 5;            // floor(-d~): (int32)d = -J for d <= 0
    // one bound for the four (they are sorted: the largest |d| is at one of the ends)
    const int32_t ep_raw = eps0 + (max(abs(J[0]), abs(J[3])) >> eps_shift);
    const int32_t ep = min(ep_raw, kEpsCap);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int32_t mw = (int32_t)mixw_rev[top[j] & 31];
        lo[j] = ((((1 << kShift) - 1) - J[j] - ep) >> kShift) - mw;
        hi[j] = ((((1 << kShift) - 1) - J[j] + ep) >> kShift) - mw;
    }
    // the chain at both ends of every density's fden interval; equal ends pin the true value
    const int32_t Flo = chain4(tab, lo), Fhi = chain4(tab, hi);
    const int32_t gap = 2 * ep + 2;
    bool close = (J[1] - J[0]) <= gap || (J[2] - J[1]) <= gap || (J[3] - J[2]) <= gap;
    close |= r5 < (uint32_t)((gap + 2) << 5);                   // a fifth density within reach of the fourth
    close |= ep_raw > kEpsCap;                                  // too far out for the interval logic: literal scan
    fx.kind = close ? 2u : (Flo != Fhi ? 1u : 0u);
    uint32_t sl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool sure = lo[j] == hi[j] && lo[j] > -16000 && lo[j] < 16000;
        const uint32_t id = 31u - (uint32_t)(top[j] & 31);
        sl[j] = sure ? ((uint32_t)lo[j] & 0x7fffu) : (0x8000u | (((uint32_t)J[j] & 0x3ffu) << 5) | id);
    }
    fx.s01 = sl[0] | (sl[1] << 16); fx.s23 = sl[2] | (sl[3] << 16);
    int32_t scr = -Fhi;
    if (__builtin_expect(aw != 1, 0)) scr /= aw;
    return clamp16(scr);
}

__device__ __forceinline__ float fmin3(float a, float b, float c) { return fminf(fminf(a, b), c); }

// The cheap certificate.  w[j] as above; c2[j] = c_j * 2^40 with
// c_j = max(kWinV, 32*1024*(mixw_j - min mixw of the senone) + kDeltaV).
// Let j* be the density with the smallest w (the largest d - 1024 mixw), W1 = w_j*.  If for
// every other density w_j >= W1 + c_j, then
//   (1) it sits > 40 log-add units below j*; whichever of them are in the top 4,
//       their chain cannot reach the table's 29-unit window of j* even when three
//       of them add up (x + 7 + 4), and
//   (2) d_j = S_j + 1024 mixw_j <= S_j* - 1024 (mixw_j - min mixw) - delta + 1024 mixw_j
//       <= d_j* - delta: j* is the top-1 density by distance (an exact tie in w fails
//       the test too), so it is in the top 4 whatever the others are, and
//   (3) when fden of j* is the same over the whole eps interval,
// the senone score is exactly -(fden(j*) - mixw(j*)) = floor(W1 / 32768) -- read off W1 alone
// (fden = ((int32)d + 1023) >> 10 with (int32)d = -floor(W1 / 32) for d <= 0).
// Instruction budget per density: FMNMX3 (two per instruction), one FFMA for the row's
// threshold K_j = (W1 + c_j) 2^40, one FFMA.SAT for the crisp 0/1 indicator
// sat(K_j - w_j 2^40) (both with an immediate operand), and one IADD3 per two indicators
// on their float bit patterns (N * 0x3f800000 cannot alias 0x3f800000 for N <= 512).
// epsA = (eps0 + 1.5) / 1024, epsB = 2^-(15 + eps_shift): the eps interval as a fraction of one
// fden step.
// UNI: every c_j of the senone is kWinV (at most three densities could violate (2), so j* is
// among the top 4 by distance whatever they do): one threshold for the whole row.
template <int M, bool UNI>
__device__ __forceinline__ bool cert_eval(const uint32_t *w /* M registers */, const float *c2, int aw, float epsA,
                                          float epsB, int32_t &score) {
    // four independent min chains (the whole certificate hangs on this value)
    float W1;
    {
        float m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            constexpr int Q = M / 4;
            m[k] = __uint_as_float(w[k * Q]);
#pragma unroll
            for (int j = 1; j + 1 < Q; j += 2) m[k] = fmin3(m[k], __uint_as_float(w[k * Q + j]), __uint_as_float(w[k * Q + j + 1]));
            if ((Q & 1) == 0) m[k] = fminf(m[k], __uint_as_float(w[k * Q + Q - 1]));
        }
        W1 = fminf(fmin3(m[0], m[1], m[2]), m[3]);
    }
    uint32_t n0 = 0, n1 = 0;
    if (UNI) {
        const float Ku = fmaf(W1, kBig, kWinV * kBig);
#pragma unroll
        for (int j = 0; j < M; j += 4) {
            const float i0 = __saturatef(fmaf(__uint_as_float(w[j]), -kBig, Ku));
            const float i1 = __saturatef(fmaf(__uint_as_float(w[j + 1]), -kBig, Ku));
            const float i2 = __saturatef(fmaf(__uint_as_float(w[j + 2]), -kBig, Ku));
            const float i3 = __saturatef(fmaf(__uint_as_float(w[j + 3]), -kBig, Ku));
            n0 += __float_as_uint(i0) + __float_as_uint(i1);
            n1 += __float_as_uint(i2) + __float_as_uint(i3);
        }
    } else {
#pragma unroll
        for (int j = 0; j < M; j += 4) {
            const float4 c = *reinterpret_cast<const float4 *>(c2 + j);
            const float i0 = __saturatef(fmaf(__uint_as_float(w[j]), -kBig, fmaf(W1, kBig, c.x)));
            const float i1 = __saturatef(fmaf(__uint_as_float(w[j + 1]), -kBig, fmaf(W1, kBig, c.y)));
            const float i2 = __saturatef(fmaf(__uint_as_float(w[j + 2]), -kBig, fmaf(W1, kBig, c.z)));
            const float i3 = __saturatef(fmaf(__uint_as_float(w[j + 3]), -kBig, fmaf(W1, kBig, c.w)));
            n0 += __float_as_uint(i0) + __float_as_uint(i1);
            n1 += __float_as_uint(i2) + __float_as_uint(i3);
        }
    }
    const float u = W1 * (1.0f / (kAccScale * 1024.f));                 // exact (power of two)
    const float fl = floorf(u);
    const float g = fabsf((u - fl) - 0.5f);                             // exact; 0.5 = the middle of a fden step
    const float lim = fmaf(W1, -epsB, 0.5f - epsA);                     // <= 0 far out: never certain there
    int32_t scr = __float2int_rd(u);
    if (__builtin_expect(aw != 1, 0)) scr /= aw;
    score = clamp16(scr);
    return W1 > 0.f && (n0 + n1) == 0x3f800000u && g < lim;
}

// ------------------------------------------------------------------- kernels
// Feature rows [T][D] -> per frame tile, transposed and zero padded:
// gX[m_tile][dim 0..Dp)[128 rows].  20 KB per tile for D = 39; the x^2 / hi / lo
// expansion (4x the bytes) happens inside the SM so it never crosses L2.
__global__ void __launch_bounds__(kTileM)
tc_prep_kernel(const float *__restrict__ feat, int T, int stride, int off, int D, int Dp, float *__restrict__ gX,
               unsigned int *__restrict__ xmax /* [D] max |x| as float bits, or null */,
               float *__restrict__ xrow /* [T_pad][(D+3)&~3] 16-byte aligned zero-padded rows for the exact fix-ups, or null */) {
    __shared__ float tile[kTileM][41];
    const int mt = blockIdx.x, r = threadIdx.x;
    const int t0 = mt * kTileM;
    // coalesced read of the tile's rows (stream `off`..`off+D` of each frame vector)
    const int nrow = min(kTileM, T - t0);
    __shared__ unsigned int s_max[41];
    if (xmax && r < D) s_max[r] = 0u;
    if (xmax) __syncthreads();
    for (int e = r; e < nrow * D; e += kTileM) {
        const float v = feat[(size_t)(t0 + e / D) * stride + off + e % D];
        tile[e / D][e % D] = v;
        if (xmax) {
            // |x| as float bits orders like an unsigned int; NaN (exponent 255, mantissa != 0) sorts above +inf
            atomicMax(&s_max[e % D], __float_as_uint(fabsf(v)));
        }
    }
    __syncthreads();
    for (int i = 0; i < Dp; ++i)
        gX[((size_t)mt * Dp + i) * kTileM + r] = (r < nrow && i < D) ? tile[r][i] : 0.f;
    if (xrow) {
        const int D4 = (D + 3) & ~3;
        for (int e = r; e < kTileM * D4; e += kTileM) {
            const int rr = e / D4, i = e % D4;
            xrow[(size_t)t0 * D4 + e] = (rr < nrow && i < D) ? tile[rr][i] : 0.f;
        }
    }
    if (xmax && r < D) atomicMax(xmax + r, s_max[r]);
}

// Operand format of every n-tile for this batch: fp16 iff no feature exceeds the
// tile's limits (and the tile's own range fits at all: lim > 0).
__global__ void tc_tile_format_kernel(const unsigned int *__restrict__ xmax, const float *__restrict__ lim, int D,
                                      int n_tiles_n, uint8_t *__restrict__ fmt, int *__restrict__ n_half) {
    const int nt = blockIdx.x * blockDim.x + threadIdx.x;
    if (nt >= n_tiles_n) return;
    bool ok = true;
    for (int i = 0; i < D; ++i) ok = ok && (__uint_as_float(xmax[i]) <= lim[(size_t)nt * D + i]);   // false for NaN
    fmt[nt] = ok ? 1 : 0;
    if (ok) atomicAdd(n_half, 1);
}

// Veltkamp split on the FMA pipe: hi carries the top 11 significant bits of a
// (a valid TF32), lo = a - hi exactly (the tensor core reads its top 11 bits).
__device__ __forceinline__ void split_tf32(float a, float &hi, float &lo) {
    const float c = __fmul_rn(a, 8193.0f);
    hi = __fsub_rn(c, __fsub_rn(c, a));
    lo = __fsub_rn(a, hi);
}

// MODE 0: fully-continuous senones (top-4 of each senone's M densities, log-add,
//         int16 score).  MODE 1: tied codebooks (ptm / s2_semi): every thread
//         reduces its 64 accumulator columns to their 4 best keys and stores
//         them; tied_select_kernel merges the groups of a codebook and rescores
//         the survivors exactly.
// HALF 1: the operands are fp16 hi/lo pairs (22 significant bits, the same
//         three-product scheme) issued as kind::f16 MMAs with K = 16: half the
//         MMA instructions of the TF32 form.  fp16's 5-bit exponent is handled by
//         power-of-two scales per n-tile and K column (A column k times 2^e,
//         B column k times 2^-e, chosen at load so that the tile's B fills the
//         fp16 range).  Per batch, tc_tile_format_kernel compares the largest
//         |feature| of every dimension with each tile's limits and writes
//         p.fmt[n_tile]; both kernels are launched and each takes the units of
//         its own format, so one sharp Gaussian or one large feature only moves
//         the affected tiles to TF32.  KS counts 16-column steps then.
// smem bytes of one instantiation (host and device agree on it)
constexpr int score_smem_bytes(int KS, int HALF, int HW, int M = 32) {
    return KS * kBStageBytes + ring_depth(KS) * kAStageBytes + (HALF ? 8 : 4) * KS * kTileM * 4 + 512 + 320 +
           40 * 8 + 16 + (kTileN + 128) * 4 + kTileN * 4 + 128 + (HW ? queue_bytes(M) : 0);
}

// HW: 1 = every epilogue warp keeps a ring of undecided pairs in shared memory (MODE 0
// only); 0 = undecided pairs are finished in place by the lane that found them, e.g. the
// 160 KB-B TF32 instantiation that has no room for the rings.
template <int M, int KS, int MODE, int HALF, int HW>
__global__ void __launch_bounds__(score_threads(HW), 1)
tc_score_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int SPT = kTileN / M;       // senones per tile
    constexpr int kStages = ring_depth(KS);
    constexpr int DP = (HALF ? 8 : 4) * KS;   // padded dims per frame in the X tile
    constexpr int kXBytes = DP * kTileM * 4;
    constexpr int NT = score_threads(HW);
    constexpr int EW = epi_warps(HW);             // epilogue warps
    uint8_t *sB = smem;                                         // KS * 16 KB
    uint8_t *sA = sB + KS * kBStageBytes;                       // kStages * 8 KB
    uint8_t *sX = sA + kStages * kAStageBytes;                  // DP * 128 * 4
    uint8_t *sMixw = sX + kXBytes;                              // 256 B
    uint8_t *sTab = sMixw + 256;                                // 256 B
    float *sScale = reinterpret_cast<float *>(sTab + 256);      // 16 * KS floats (HALF only; <= 320 B reserved)
    uint64_t *bars = reinterpret_cast<uint64_t *>(sTab + 256 + 320);
    // barrier map: b_full, b_empty, x_full, x_empty, a_full[S], a_empty[S], tmem_full[2], tmem_empty[2]
    constexpr int B_FULL = 0, B_EMPTY = 1, X_FULL = 2, X_EMPTY = 3, A_FULL = 4, A_EMPTY = 4 + kStages,
                  T_FULL = 4 + 2 * kStages, T_EMPTY = T_FULL + 2, N_BARS = T_EMPTY + 2;
    static_assert(N_BARS <= 36, "barrier area");
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 40);
    // 32*1024*mixw per tile row, senone stride M + 4 (16-byte rows; the hard-path lanes of a warp read different senones)
    float *sCw = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(tmem_slot) + 16);
    float *sC2 = sCw + kTileN + 128;                                                        // [256] certificate constants
    // HW: per epilogue warp a ring of [kQCap][M] accumulator rows (128-byte aligned), then all the headers
    uint32_t *sUni = reinterpret_cast<uint32_t *>(sC2 + kTileN);   // [SPT <= 32] uniform-window flag per senone of the tile
    unsigned *sTick = reinterpret_cast<unsigned *>(bars + 36);     // chunk ticket counter per lane quarter (the barrier area's spare 32 bytes)
    uint32_t *sQ = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(sUni + 32) + 127) & ~(uintptr_t)127);
    uint32_t *sQh = sQ + kEpiWarps * kQCap * M;
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(BAR(B_FULL), 1);
        mbar_init(BAR(B_EMPTY), 1);
        mbar_init(BAR(X_FULL), 1);
        mbar_init(BAR(X_EMPTY), kBuildWarps);
        for (int s = 0; s < kStages; ++s) { mbar_init(BAR(A_FULL + s), kBuildWarps); mbar_init(BAR(A_EMPTY + s), 1); }
        // accumulator releases: one per epilogue warp (MODE 1) or one per 32-column chunk and lane quarter (MODE 0)
        for (int a = 0; a < 2; ++a) { mbar_init(BAR(T_FULL + a), 1); mbar_init(BAR(T_EMPTY + a), MODE == 0 ? 4 * (kTileN / 32) : EW); }
        for (int a = 0; a < 4; ++a) sTick[a] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 256; i += NT) sTab[i] = p.logadd[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    constexpr int ksteps = KS;
    // a unit is skipped by the kernel of the other operand format
#define OTHER_FORMAT(nt_) ((p.fmt ? (int)p.fmt[(nt_)] : 0) != HALF)

    if (warp == kProdWarp) {
        // ===================== producer =====================
        // Bulk-TMA: the unit's B tile (resident for the whole unit), then one raw
        // feature tile (DP x 128 fp32) per frame tile.
        if (lane == 0) {
            uint32_t xphase = 0, bphase = 0;
            for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
                const int nt = u % p.n_tiles_n, mc = u / p.n_tiles_n;
                if (OTHER_FORMAT(nt)) continue;
                const int mt0 = mc * p.tiles_per_chunk, mt1 = min(p.n_tiles_m, mt0 + p.tiles_per_chunk);
                mbar_wait(BAR(B_EMPTY), bphase ^ 1);
                mbar_expect_tx(BAR(B_FULL), (uint32_t)ksteps * kBStageBytes);
                const uint8_t *gb = reinterpret_cast<const uint8_t *>(p.gB) + (size_t)nt * ksteps * kBStageBytes;
                for (int j = 0; j < ksteps; ++j)
                    bulk_g2s(smem_u32(sB + j * kBStageBytes), gb + (size_t)j * kBStageBytes, kBStageBytes, BAR(B_FULL));
                bphase ^= 1;
                for (int mt = mt0; mt < mt1; ++mt) {
                    mbar_wait(BAR(X_EMPTY), xphase ^ 1);
                    mbar_expect_tx(BAR(X_FULL), kXBytes);
                    bulk_g2s(smem_u32(sX), reinterpret_cast<const uint8_t *>(p.gX) + (size_t)mt * kXBytes, kXBytes,
                             BAR(X_FULL));
                    xphase ^= 1;
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer =====================
        // The whole warp walks the (fully unrolled) loop so control flow stays
        // uniform; one elected lane issues.  Ring slot and barrier parity of
        // k-step j are compile-time, descriptors are base + constant.
        uint32_t bphase = 0, accphase = 0, tilepar = 0;
        int acc = 0;
        const uint64_t dA0 = make_desc(smem_u32(sA), kTileM * 16, 128);
        const uint64_t dB0 = make_desc(smem_u32(sB), kTileN * 16, 128);
        for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
            const int mc = u / p.n_tiles_n;
            if (OTHER_FORMAT(u % p.n_tiles_n)) continue;
            const int mt0 = mc * p.tiles_per_chunk, mt1 = min(p.n_tiles_m, mt0 + p.tiles_per_chunk);
            mbar_wait(BAR(B_FULL), bphase);
            bphase ^= 1;
            for (int mt = mt0; mt < mt1; ++mt) {
                mbar_wait(BAR(T_EMPTY + acc), accphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * kTileN;
#pragma unroll
                for (int j = 0; j < KS; ++j) {
                    constexpr int uses = KS / kStages;           // ring passes per tile (1 or 2)
                    const int stage = j % kStages;
                    const uint32_t par = (uses & 1) ? tilepar : (uint32_t)((j / kStages) & 1);
                    mbar_wait(BAR(A_FULL + stage), par);
                    tc_fence_after();
                    if (elect_one()) {
                        // descriptor address fields are in 16-byte units
                        const uint64_t dAhi = dA0 + (uint64_t)((stage * kAStageBytes) >> 4);
                        const uint64_t dAlo = dAhi + (uint64_t)((kAStageBytes / 2) >> 4);
                        const uint64_t dBhi = dB0 + (uint64_t)((j * kBStageBytes) >> 4);
                        const uint64_t dBlo = dBhi + (uint64_t)((kBStageBytes / 2) >> 4);
                        if (HALF) {
                            tc_mma_f16(d_tmem, dAhi, dBhi, kIdescF16, j > 0 ? 1u : 0u);
                            if (!(p.dbg & 2)) {
                                tc_mma_f16(d_tmem, dAhi, dBlo, kIdescF16, 1u);
                                tc_mma_f16(d_tmem, dAlo, dBhi, kIdescF16, 1u);
                            }
                        } else {
                            tc_mma_tf32(d_tmem, dAhi, dBhi, kIdesc, j > 0 ? 1u : 0u);
                            if (!(p.dbg & 2)) {
                                tc_mma_tf32(d_tmem, dAhi, dBlo, kIdesc, 1u);
                                tc_mma_tf32(d_tmem, dAlo, dBhi, kIdesc, 1u);
                            }
                        }
                        tc_commit(BAR(A_EMPTY + stage));
                        if (j == KS - 1) tc_commit(BAR(T_FULL + acc));
                    }
                    __syncwarp();
                }
                if ((KS / kStages) & 1) tilepar ^= 1;
                if (++acc == 2) { acc = 0; accphase ^= 1; }
            }
            if (elect_one()) tc_commit(BAR(B_EMPTY));
            __syncwarp();
        }
    } else if (warp >= kFirstBuildWarp) {
        // ===================== A-operand builders (4 warps) =====================
        // Thread r owns frame row r of the tile: it keeps the row's DP features in
        // registers and, k-step by k-step, writes [1,1,x^2,x,...] split into TF32
        // hi/lo straight into the UMMA K-major layout of the A ring.
        const int r = threadIdx.x - kFirstBuildWarp * 32; // 0..127
        const float *xs = reinterpret_cast<const float *>(sX);
        int stage = 0; uint32_t phase = 0, xphase = 0;
        for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
            const int mc = u / p.n_tiles_n;
            if (OTHER_FORMAT(u % p.n_tiles_n)) continue;
            const int mt0 = mc * p.tiles_per_chunk, mt1 = min(p.n_tiles_m, mt0 + p.tiles_per_chunk);
            if (HALF) {
                // this tile's scale table: everybody is done with the previous one, then refill
                asm volatile("bar.sync 2, %0;" ::"n"(kBuildWarps * 32) : "memory");
                for (int k = r; k < 16 * KS; k += kBuildWarps * 32) sScale[k] = p.scaleA[(size_t)(u % p.n_tiles_n) * (16 * KS) + k];
                asm volatile("bar.sync 2, %0;" ::"n"(kBuildWarps * 32) : "memory");
            }
            for (int mt = mt0; mt < mt1; ++mt) {
                mbar_wait(BAR(X_FULL), xphase);
                xphase ^= 1;
                float x[DP];
#pragma unroll
                for (int i = 0; i < DP; ++i) x[i] = xs[i * kTileM + r];
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(X_EMPTY));     // staging buffer may be refilled
#pragma unroll
                for (int j = 0; j < KS; ++j) {
                    if (HALF) {
                        // 16 columns: scaled value -> fp16 hi + fp16 lo (exact remainder, rounded once)
                        uint32_t hp[8], lp[8];
                        float sc[16];                     // this k-step's 16 column scales: four 16-byte broadcast loads
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const float4 q4 = *reinterpret_cast<const float4 *>(sScale + j * 16 + k4 * 4);
                            sc[k4 * 4] = q4.x; sc[k4 * 4 + 1] = q4.y; sc[k4 * 4 + 2] = q4.z; sc[k4 * 4 + 3] = q4.w;
                        }
#pragma unroll
                        for (int kk = 0; kk < 16; kk += 2) {
                            float a[2];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int k = j * 16 + kk + e;
                                if (k < 2) a[e] = sc[kk + e];
                                else {
                                    const int i = (k - 2) >> 1;
                                    a[e] = ((k - 2) & 1) ? __fmul_rn(x[i], sc[kk + e]) : __fmul_rn(__fmul_rn(x[i], sc[kk + e]), x[i]);
                                }
                            }
                            const __half2 h = __floats2half2_rn(a[0], a[1]);
                            const float2 hf = __half22float2(h);
                            const __half2 l = __floats2half2_rn(__fsub_rn(a[0], hf.x), __fsub_rn(a[1], hf.y));
                            hp[kk >> 1] = *reinterpret_cast<const uint32_t *>(&h);
                            lp[kk >> 1] = *reinterpret_cast<const uint32_t *>(&l);
                        }
                        mbar_wait(BAR(A_EMPTY + stage), phase ^ 1);
                        uint4 *dsth = reinterpret_cast<uint4 *>(sA + stage * kAStageBytes);
                        dsth[0 * kTileM + r] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                        dsth[1 * kTileM + r] = make_uint4(hp[4], hp[5], hp[6], hp[7]);
                        dsth[2 * kTileM + r] = make_uint4(lp[0], lp[1], lp[2], lp[3]);
                        dsth[3 * kTileM + r] = make_uint4(lp[4], lp[5], lp[6], lp[7]);
                    } else {
                    float hi[8], lo[8];
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const int k = j * 8 + kk;
                        if (k < 2) { hi[kk] = 1.f; lo[kk] = 0.f; }
                        else {
                            const int i = (k - 2) >> 1;              // i < DP by construction
                            const float a = ((k - 2) & 1) ? x[i] : __fmul_rn(x[i], x[i]);
                            split_tf32(a, hi[kk], lo[kk]);
                        }
                    }
                    mbar_wait(BAR(A_EMPTY + stage), phase ^ 1);
                    float4 *dst = reinterpret_cast<float4 *>(sA + stage * kAStageBytes);
                    dst[0 * kTileM + r] = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    dst[1 * kTileM + r] = make_float4(hi[4], hi[5], hi[6], hi[7]);
                    dst[2 * kTileM + r] = make_float4(lo[0], lo[1], lo[2], lo[3]);
                    dst[3 * kTileM + r] = make_float4(lo[4], lo[5], lo[6], lo[7]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic -> async proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(A_FULL + stage));
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue (16 warps) =====================
        // Warp w may read TMEM lanes 32*(w%4)..+31 only; the four warps of a lane
        // quarter split the 256 accumulator columns into 64-column groups.  Each
        // thread pulls 32 columns at a time into registers; after the last load the
        // accumulator is released (the MMA warp can start the tile after next).
        constexpr int CPT = kTileN / (EW / 4);           // columns per thread (64)
        constexpr int SPE = CPT / M;                     // senones per thread
        const int q = warp & 3;
        const int cg = (warp - kFirstEpiWarp) >> 2;      // column group 0..3
        const int ew = warp - kFirstEpiWarp;
        const int et = threadIdx.x - kFirstEpiWarp * 32; // 0..511
        const int row = q * 32 + lane;                   // frame row in the tile
        int acc = 0; uint32_t accphase = 0;
        const int32_t m31 = p.m31;
        // MODE 0: this warp's ring of undecided pairs and its reserved run of queue A slots
        // [kQCap][M] accumulator rows; 16-byte chunk c of slot s sits at chunk c ^ swz(s) of the item, so that
        // eight lanes with consecutive slots touch eight different 16-byte bank groups
        const uint32_t wq = keep(smem_u32(sQ + ew * (kQCap * M)));
        const uint32_t wqh = keep(smem_u32(sQh + ew * kQCap)); // [kQCap] headers: (frame - first frame of the unit) << 5 | senone in the tile
        const uint32_t bT = keep(BAR(T_FULL));           // T_FULL[a] = bT + 8 a, T_EMPTY[a] = bT + 16 + 8 a
        const uint32_t tmq = keep(tmem_base + ((uint32_t)(q * 32) << 16));
        const uint32_t tkq = keep(smem_u32(sTick + q));
        auto item_addr = [&](int slot) {                 // address of chunk 0 of the slot, swizzle bits folded in
            const int swz = M == 32 ? (slot & 7) : (M == 16 ? ((slot >> 1) & 3) : ((slot >> 2) & 1));
            return wq + (uint32_t)slot * (M * 4) + ((uint32_t)swz << 4);
        };
        int qhead = 0, qpend = 0;
        unsigned n_hard = 0, n_inplace = 0;
        uint32_t tick = 0; bool have_tick = false;       // chunk ticket of this warp's lane quarter (MODE 0)
        int tile_base = 0;                               // running tile number of the current unit's first tile
        const unsigned region = blockIdx.x < kFixRegionsMax ? blockIdx.x : 0;
        unsigned fbase = 0; int fleft = 0;
        const bool all_hard = (p.dbg & 4) != 0, all_easy = (p.dbg & 1) != 0;
        for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
            const int nt = u % p.n_tiles_n, mc = u / p.n_tiles_n;
            if (OTHER_FORMAT(nt)) continue;
            const int mt0 = mc * p.tiles_per_chunk, mt1 = min(p.n_tiles_m, mt0 + p.tiles_per_chunk);
            if (MODE == 0) {
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");      // every ring is empty: the previous unit's tables are free
                if (et < kTileN) {
                    // stored reversed inside each senone so that key & 31 indexes it directly
                    const int sl = et / M, dens = et % M;
                    sMixw[sl * M + (M - 1 - dens)] = p.gMixw[(size_t)nt * kTileN + et];
                    const float c = p.gCw[(size_t)nt * kTileN + et];
                    sCw[sl * (M + 4) + dens] = c;
                    // the senone's smallest mixw offset (its M rows are M consecutive lanes)
                    float cmin = c;
#pragma unroll
                    for (int o = M / 2; o; o >>= 1) cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
                    // condition (2) asks for more than the window only from a density whose mixw offset exceeds
                    // the senone's smallest by ~42 units; with at most three of those nobody has to ask (cert_eval)
                    const float cd = (c - cmin) + kDeltaV;
                    const unsigned bm = __ballot_sync(0xffffffffu, cd > kWinV);
                    const unsigned gmask = M == 32 ? 0xffffffffu : (((1u << (M & 31)) - 1u) << ((lane / M) * M));
                    const bool uni = __popc(bm & gmask) <= 3;
                    sC2[et] = (uni ? kWinV : fmaxf(kWinV, cd)) * kBig;
                    if (dens == 0) sUni[sl] = uni ? 1u : 0u;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
            }
            int16_t *rawt = MODE == 0 ? p.raw + (size_t)nt * p.T_pad * SPT : nullptr;
            uint4 *partt = MODE == 1 ? p.part + (size_t)nt * p.T_pad * 4 : nullptr;
            if (MODE == 1) {
                for (int mt = mt0; mt < mt1; ++mt) {
                    mbar_wait(BAR(T_FULL + acc), accphase);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kTileN + cg * CPT);
                    uint32_t v0[32], v1[32];
                    tmem_ld32(taddr, v0);
                    tmem_ld32(taddr + 32, v1);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(T_EMPTY + acc));
                    // keys carry a 6-bit id (63 - column within the thread's 64)
                    int32_t ta[4], tb[4];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint32_t (&v)[32] = c ? v1 : v0;
                        int32_t (&top)[4] = c ? tb : ta;
                        top[0] = make_key(v[0], c * 32 + 0, m31); top[1] = make_key(v[1], c * 32 + 1, m31);
                        top[2] = make_key(v[2], c * 32 + 2, m31); top[3] = make_key(v[3], c * 32 + 3, m31);
                        sort4(top[0], top[1], top[2], top[3]);
#pragma unroll
                        for (int g = 4; g < 32; g += 4)
                            merge4(top, make_key(v[g], c * 32 + g, m31), make_key(v[g + 1], c * 32 + g + 1, m31),
                                   make_key(v[g + 2], c * 32 + g + 2, m31), make_key(v[g + 3], c * 32 + g + 3, m31));
                    }
                    int32_t m0 = min(ta[0], tb[3]), m1 = min(ta[1], tb[2]), m2 = min(ta[2], tb[1]), m3 = min(ta[3], tb[0]);
                    ce(m0, m2); ce(m1, m3); ce(m0, m1); ce(m2, m3);
                    partt[(size_t)(mt * kTileM + row) * 4 + cg] = make_uint4((uint32_t)m0, (uint32_t)m1, (uint32_t)m2, (uint32_t)m3);
                    if (++acc == 2) { acc = 0; accphase ^= 1; }
                }
                continue;
            }
            // ---- MODE 0: certificate per senone; undecided pairs go to the warp's ring (or are
            // finished in place when there is no ring / it is full).
            // Chunks (32 accumulator columns of one lane quarter) are handed out by a ticket counter
            // per quarter: a warp that is busy with a pass simply takes fewer chunks and the other
            // three warps of its quarter take more, so the accumulator is released at the AVERAGE
            // pace of the quarter's warps, not at the pace of the slowest one (with a fixed
            // warp -> column mapping one warp in a pass held up the release, the MMA warp and
            // through it all other warps: 9.6 ms against 6.1 ms without undecided pairs).
            constexpr int SPC = 32 / M;                                         // senones per chunk
            constexpr int NCH = kTileN / 32;                                    // chunks per tile and quarter
            const int ntile = mt1 - mt0;
            const int n_valid = min(SPT, p.n_sen - nt * SPT);                   // real senones of this n-tile
            // One chunk per iteration, NOT unrolled: the loop body (certificate, append, pass)
            // exists once and must stay in the instruction cache next to the other warp roles' code.
#pragma unroll 1
            while (true) {
                if (!have_tick) {
                    if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(tick) : "r"(tkq) : "memory");
                    tick = __shfl_sync(0xffffffffu, tick, 0);
                    have_tick = true;
                }
                const int tseq = (int)(tick / NCH) - tile_base;                 // tile of this unit
                bool drain = tseq >= ntile;                                     // the ticket belongs to the next unit: keep it
                if (drain && !(HW && qpend)) break;
                const int c = (int)(tick % NCH);
                const int mt = mt0 + tseq;
                const int t = mt * kTileM + row;
                const uint32_t trel = (uint32_t)(tseq * kTileM + row);
                {
                    if (!drain) {
                        const uint32_t gseq = tick / NCH;                       // running tile number of this CTA
                        const uint32_t acc = gseq & 1u;
                        mbar_wait(bT + 8u * acc, (gseq >> 1) & 1u);
                        tc_fence_after();
                        const uint32_t taddr = tmq + acc * kTileN + (uint32_t)(c * 32);
                        uint32_t v[32];
                        tmem_ld32(taddr, v);
                        tmem_ld_wait();
                        // the chunk is in registers: one of the accumulator's 4 * NCH releases
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bT + 16u + 8u * acc);
                        have_tick = false;
#pragma unroll
                        for (int s = 0; s < SPC; ++s) {      // senones inside the 32-column chunk
                            const int st = c * SPC + s;                       // senone within the tile
                            int32_t sc;
                            bool easy = sUni[st] ? cert_eval<M, true>(&v[s * M], sC2 + st * M, p.aw, p.epsA, p.epsB, sc)
                                                 : cert_eval<M, false>(&v[s * M], sC2 + st * M, p.aw, p.epsA, p.epsB, sc);
                            if (all_hard) easy = false;
                            if (all_easy || st >= n_valid) easy = true;                  // (padding senones of the last tile)
                            const unsigned hm = __ballot_sync(0xffffffffu, !easy);
                            if (hm) {
                                const int rank = __popc(hm & ((1u << lane) - 1u));
                                const int pos = qpend + rank;                  // position in the ring, if it fits
                                if (HW) qpend = min(qpend + __popc(hm), kQCap);
                                n_hard += (unsigned)__popc(hm);
                                if (!easy) {
                                    if (HW && pos < kQCap) {
                                        int slot = qhead + pos;
                                        slot -= slot >= kQCap ? kQCap : 0;
                                        const uint32_t it = item_addr(slot);
#pragma unroll
                                        for (int j = 0; j < M; j += 4)
                                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(it ^ (uint32_t)(j << 2)),
                                                         "r"(v[s * M + j]), "r"(v[s * M + j + 1]), "r"(v[s * M + j + 2]), "r"(v[s * M + j + 3]) : "memory");
                                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(wqh + 4u * (uint32_t)slot), "r"((trel << 5) | (uint32_t)st) : "memory");
                                    } else {
                                        // no ring (or it is full): finish here
                                        uint32_t w[M];
#pragma unroll
                                        for (int j = 0; j < M; ++j) w[j] = v[s * M + j];
                                        FixOut fx;
                                        sc = hard_eval<M>(w, sCw + st * (M + 4), sMixw + st * M - (32 - M), sTab, p.aw, p.eps0,
                                                          p.eps_shift, m31, fx);
                                        if (fx.kind && t < p.T) fix_append(p, fx, (uint32_t)t, (uint32_t)(nt * SPT + st));
                                        ++n_inplace;
                                        easy = true;
                                    }
                                }
                            }
                            if (easy && t < p.T) rawt[(size_t)t * SPT + st] = (int16_t)sc;
                        }
                    }
                    // ---- the ring: 32 undecided pairs at a time, one per lane, through the full
                    // top-4 network + the interval check of hard_eval
                    while (HW && (qpend >= kQPass || (drain && qpend > 0))) {
                        __syncwarp();                                          // the appends above are visible
                        const int n = min(qpend, 32);
                        int slot = qhead + min(lane, n - 1);
                        slot -= slot >= kQCap ? kQCap : 0;
                        const uint32_t it = item_addr(slot);
                        uint32_t w[M];
#pragma unroll
                        for (int j = 0; j < M; j += 4)
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[j]), "=r"(w[j + 1]), "=r"(w[j + 2]), "=r"(w[j + 3])
                                         : "r"(it ^ (uint32_t)(j << 2)) : "memory");
                        uint32_t hdr;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(hdr) : "r"(wqh + 4u * (uint32_t)slot) : "memory");
                        const int sl = (int)(hdr & 31u);                       // senone in the tile
                        const int th = mt0 * kTileM + (int)(hdr >> 5);
                        FixOut fx;
                        const int32_t sc = hard_eval<M>(w, sCw + sl * (M + 4), sMixw + sl * M - (32 - M), sTab, p.aw, p.eps0,
                                                        p.eps_shift, m31, fx);
                        const bool live = lane < n && th < p.T;
                        if (live) rawt[(size_t)th * SPT + sl] = (int16_t)sc;
                        // queue A appends: slots are reserved kFixChunk at a time (one global atomic per
                        // chunk instead of one round trip per pass)
                        const unsigned ma = __ballot_sync(0xffffffffu, live && fx.kind == 1);
                        if (ma) {
                            const int cnt = __popc(ma), rank = __popc(ma & ((1u << lane) - 1u));
                            unsigned nbase = 0;
                            if (cnt > fleft) {
                                if (lane == 0) nbase = atomicAdd(p.qcnt + 8 + region, (unsigned)kFixChunk);
                                nbase = __shfl_sync(0xffffffffu, nbase, 0);
                            }
                            if (live && fx.kind == 1) {
                                const unsigned idx = rank < fleft ? fbase + rank : nbase + (rank - fleft);
                                if (idx < p.capA) p.qa[(size_t)region * p.capA + idx] = make_uint4((uint32_t)th, (uint32_t)(nt * SPT + sl), fx.s01, fx.s23);
                                else p.qcnt[2] = 1u;
                            }
                            if (cnt > fleft) { fbase = nbase + (cnt - fleft); fleft = kFixChunk - (cnt - fleft); }
                            else { fbase += cnt; fleft -= cnt; }
                        }
                        if (live && fx.kind == 2) fix_append(p, fx, (uint32_t)th, (uint32_t)(nt * SPT + sl));
                        qhead += n; qhead -= qhead >= kQCap ? kQCap : 0;
                        qpend -= n;
                        __syncwarp();                                          // the slots may be overwritten
                    }
                }
            }
            tile_base += ntile;
        }
        if (MODE == 0) {
            // the unused rest of the last reservation: null items (tc_fix_a_kernel skips them)
            for (int k = lane; k < fleft; k += 32)
                if (fbase + k < p.capA) p.qa[(size_t)region * p.capA + fbase + k] = make_uint4(0xffffffffu, 0u, 0u, 0u);
            n_hard = __reduce_add_sync(0xffffffffu, lane == 0 ? n_hard : 0u);
            n_inplace = __reduce_add_sync(0xffffffffu, n_inplace);
            if (lane == 0 && n_hard) atomicAdd(p.qcnt + 4, n_hard);
            if (lane == 0 && n_inplace) atomicAdd(p.qcnt + 5, n_inplace);
        }
    }

#undef OTHER_FORMAT
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// Tile-major raw scores -> row-major [T][n_sen], optionally minus the frame's
// best (ms_mgau.c:188-204).  One block per kFinFrames frames; every thread owns
// a fixed (frame, 16-byte column group) and walks the n-tiles, so reads are
// contiguous runs of kFinFrames*spt int16 and all traffic is 16-byte vectors.
// Requires spt % 8 == 0 and n_sen % 8 == 0 (else the generic kernel below).
constexpr int kFinFrames = 2;
__global__ void __launch_bounds__(256)
tc_finish_vec_kernel(const int16_t *__restrict__ raw, int T, int T_pad, int n_sen, int spt, int n_tiles_n,
                     int subtract_best, int16_t *__restrict__ out) {
    extern __shared__ __align__(16) int16_t s_rows[];   // [kFinFrames][stride]
    __shared__ int s_best[kFinFrames];
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * kFinFrames;
    const int stride = n_tiles_n * spt;                  // padded senone count (multiple of 8)
    const int q = spt >> 3;                              // uint4 per (tile, frame)
    const int group = kFinFrames * q;                    // threads covering one n-tile
    const int f = (tid % group) / q, k = tid % q;
    const int nt_step = 256 / group;
    if (tid < kFinFrames) s_best[tid] = 0x7fffffff;
    __syncthreads();
    int32_t best = 0x7fffffff;
    if (t0 + f < T) {
        for (int nt = tid / group; nt < n_tiles_n; nt += nt_step) {
            const uint4 v = *reinterpret_cast<const uint4 *>(raw + ((size_t)nt * T_pad + t0 + f) * spt + 8 * k);
            *reinterpret_cast<uint4 *>(s_rows + (size_t)f * stride + nt * spt + 8 * k) = v;
            if (nt * spt + 8 * k + 8 <= n_sen) {
                uint32_t m = __vmins2(__vmins2(v.x, v.y), __vmins2(v.z, v.w));
                best = min(best, min((int32_t)(int16_t)(m & 0xffff), (int32_t)(int16_t)(m >> 16)));
            } else {
                const int16_t *h = reinterpret_cast<const int16_t *>(&v);
                for (int e = 0; e < 8; ++e)
                    if (nt * spt + 8 * k + e < n_sen) best = min(best, (int32_t)h[e]);
            }
        }
        if (subtract_best) atomicMin(&s_best[f], best);
    }
    __syncthreads();
    const int n8 = n_sen >> 3;
    for (int ff = 0; ff < kFinFrames && t0 + ff < T; ++ff) {
        const int32_t b = subtract_best ? s_best[ff] : 0;
        const uint32_t b2 = ((uint32_t)(uint16_t)(int16_t)b) * 0x10001u;
        const uint4 *src = reinterpret_cast<const uint4 *>(s_rows + (size_t)ff * stride);
        uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)(t0 + ff) * n_sen);
        for (int i = tid; i < n8; i += 256) {
            uint4 v = src[i];
            v.x = __vsubss2(v.x, b2); v.y = __vsubss2(v.y, b2); v.z = __vsubss2(v.z, b2); v.w = __vsubss2(v.w, b2);
            dst[i] = v;
        }
    }
}

// Generic (any spt / n_sen) version of the same pass.
constexpr int kNormFrames = 8;
__global__ void __launch_bounds__(256)
tc_finish_kernel(const int16_t *__restrict__ raw, int T, int T_pad, int n_sen, int spt, int n_tiles_n,
                 int subtract_best, int16_t *__restrict__ out) {
    extern __shared__ __align__(16) int16_t s_rows[];   // [8][stride]
    __shared__ int s_best[kNormFrames];
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * kNormFrames;
    const int nf = min(kNormFrames, T - t0);
    const int stride = n_tiles_n * spt;
    if (tid < kNormFrames) s_best[tid] = 0x7fffffff;
    __syncthreads();
    const int per_tile = kNormFrames * spt;
    for (int e = tid; e < n_tiles_n * per_tile; e += blockDim.x) {
        const int nt = e / per_tile, rem = e - nt * per_tile;
        const int f = rem / spt, k = rem - f * spt;
        const int s = nt * spt + k;
        if (f < nf && s < n_sen) {
            const int16_t v = raw[((size_t)nt * T_pad + t0 + f) * spt + k];
            s_rows[f * stride + s] = v;
            if (subtract_best) atomicMin(&s_best[f], (int32_t)v);
        }
    }
    __syncthreads();
    for (int f = 0; f < nf; ++f) {
        const int32_t b = subtract_best ? s_best[f] : 0;
        int16_t *o = out + (size_t)(t0 + f) * n_sen;
        for (int s = tid; s < n_sen; s += blockDim.x) o[s] = (int16_t)clamp16((int32_t)s_rows[f * stride + s] - b);
    }
}


// ------------------------------------------------------------- exact fix-ups
// The reference's float32 distance of density `id` of codebook `s` (single stream):
// ms_gauden.c:417-445, sequential in i, every operation rounded separately.
// `rows`: 16-byte aligned copy of the parameters, [codebook * n_density + id][mean Dp | scaled 1/(2 var) Dp]
// with Dp = D rounded up to 4 (one 16-byte gather per 4 dimensions).
template <int DC>
__device__ __forceinline__ float exact_dist_n(const GmmDev &g, const float4 *__restrict__ rows, const float4 *__restrict__ x4,
                                              int s, int id, int D) {
    const int q = (D + 3) >> 2;
    const float4 *__restrict__ rp = rows + ((size_t)s * g.n_density + id) * (2 * q);
    float d = __ldg(g.det + (size_t)s * g.n_density + id);
    constexpr int QC = (DC + 3) / 4;
    if (DC > 0) {
#pragma unroll
        for (int i4 = 0; i4 < QC; ++i4) {
            const float4 m4 = __ldg(rp + i4), v4 = __ldg(rp + QC + i4), xx = __ldg(x4 + i4);
            const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, xs[4] = {xx.x, xx.y, xx.z, xx.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i4 * 4 + e < DC) {
                    const float diff = __fsub_rn(xs[e], mm[e]);
                    d = __fsub_rn(d, __fmul_rn(__fmul_rn(diff, diff), vv[e]));
                }
        }
    } else {
        for (int i4 = 0; i4 < q; ++i4) {
            const float4 m4 = __ldg(rp + i4), v4 = __ldg(rp + q + i4), xx = __ldg(x4 + i4);
            const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, xs[4] = {xx.x, xx.y, xx.z, xx.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i4 * 4 + e < D) {
                    const float diff = __fsub_rn(xs[e], mm[e]);
                    d = __fsub_rn(d, __fmul_rn(__fmul_rn(diff, diff), vv[e]));
                }
        }
    }
    return d;
}
// `xrow`: the batch's features as 16-byte aligned zero-padded rows [T_pad][(D+3)&~3] (tc_prep_kernel)
__device__ __forceinline__ float exact_dist(const GmmDev &g, const float4 *rows, const float *xrow, int t, int s, int id) {
    const int D = g.featlen[0];
    const float4 *x4 = reinterpret_cast<const float4 *>(xrow) + (size_t)t * ((D + 3) >> 2);
    if (D == 39) return exact_dist_n<39>(g, rows, x4, s, id, D);
    return exact_dist_n<0>(g, rows, x4, s, id, D);
}

__device__ __forceinline__ int32_t exact_chain(const GmmDev &g, const int32_t (&fw)[4], int n) {
    int32_t f = fw[0];
    for (int j = 1; j < n; ++j) f = logadd_tab(g.logadd, f, fw[j]);
    int32_t scr = -f;
    if (g.aw != 1) scr /= g.aw;
    return clamp16(scr);
}

// Queue A: the set and the order of the top 4 are certain, but some density's
// fden = ((int32)d + 1023) >> 10 is not.  A settled slot carries fden - mixw, an
// open one the density id.  One item per lane; the open slots of the warp's 32
// items (about one per item) are compacted through shared memory so that every
// lane of the exact-distance loop has work.
__global__ void __launch_bounds__(256)
tc_fix_a_kernel(const __grid_constant__ GmmDev g, const float4 *__restrict__ rows, const float *__restrict__ feat, int T_pad, int spt,
                const uint4 *__restrict__ qa, unsigned capA, unsigned *__restrict__ qcnt, int16_t *__restrict__ raw,
                int eps0, int eps_shift) {
    if (qcnt[2]) return;                                   // overflow: queue B's kernel redoes everything
    __shared__ uint16_t s_work[8][128];                    // per warp: (lane << 2) | slot of every open slot
    __shared__ int32_t s_fw[8][32][4];
    const int region = blockIdx.y;
    const unsigned n = min(qcnt[8 + region], capA);
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    unsigned maxerr = 0;
    // 28 items per warp and round: their open slots (about 1.1 per item) then mostly fit one 32-lane pass
    constexpr unsigned kBatch = 28;
    for (unsigned i0 = (blockIdx.x * 8 + wp) * kBatch; i0 < n; i0 += gridDim.x * 8 * kBatch) {
        const unsigned i = i0 + lane;
        uint4 it = (lane < kBatch && i < n) ? qa[(size_t)region * capA + i] : make_uint4(0xffffffffu, 0, 0, 0);
        const bool live = it.x != 0xffffffffu;          // (null items pad the warps' reservations)
        // compact the open slots
        int n_open = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t slot = ((c < 2 ? it.z : it.w) >> (16 * (c & 1))) & 0xffffu;
            const bool open = live && (slot & 0x8000u);
            const unsigned m = __ballot_sync(0xffffffffu, open);
            if (open) s_work[wp][n_open + __popc(m & ((1u << lane) - 1u))] = (uint16_t)((lane << 2) | c);
            if (!open) s_fw[wp][lane][c] = (int32_t)(slot << 17) >> 17;                // 15-bit two's complement
            n_open += __popc(m);
        }
        __syncwarp();
        // Four lanes per open slot: lane ql of a quad loads 16-byte piece 4 i + ql of the density's
        // mean / variance rows and of the frame's feature row (the quad reads 64 contiguous bytes,
        // the warp 8 rows per instruction instead of 32) and forms its four terms
        // (x - mean)^2 * v; the subtraction chain -- the reference's order, one rounding per step --
        // runs over the terms in dimension order, handed round the quad by shuffles.
        const int D = g.featlen[0], q4 = (D + 3) >> 2;
        const int ql = lane & 3;
        for (int k0 = 0; k0 < n_open; k0 += 8) {
            const int k = k0 + (lane >> 2);
            const int wk = s_work[wp][min(k, n_open - 1)];
            const int src = wk >> 2, c = wk & 3;
            const int t = (int)__shfl_sync(0xffffffffu, it.x, src), sn = (int)__shfl_sync(0xffffffffu, it.y, src);
            const uint32_t z = __shfl_sync(0xffffffffu, it.z, src), w = __shfl_sync(0xffffffffu, it.w, src);
            const uint32_t slot = ((c < 2 ? z : w) >> (16 * (c & 1))) & 0xffffu;
            const int id = slot & 31;
            const float4 *__restrict__ rp = rows + ((size_t)sn * g.n_density + id) * (2 * q4);
            const float4 *__restrict__ x4 = reinterpret_cast<const float4 *>(feat) + (size_t)t * q4;
            float d = __ldg(g.det + (size_t)sn * g.n_density + id);
            for (int i0 = 0; i0 < q4; i0 += 4) {
                const int idx = i0 + ql;
                float tm[4] = {0.f, 0.f, 0.f, 0.f};
                if (idx < q4) {
                    // (pad dimensions: mean, variance and feature are all 0 there -- the term is +0 and d - 0 == d)
                    const float4 m4 = __ldg(rp + idx), v4 = __ldg(rp + q4 + idx), xx = __ldg(x4 + idx);
                    float df;
                    df = __fsub_rn(xx.x, m4.x); tm[0] = __fmul_rn(__fmul_rn(df, df), v4.x);
                    df = __fsub_rn(xx.y, m4.y); tm[1] = __fmul_rn(__fmul_rn(df, df), v4.y);
                    df = __fsub_rn(xx.z, m4.z); tm[2] = __fmul_rn(__fmul_rn(df, df), v4.z);
                    df = __fsub_rn(xx.w, m4.w); tm[3] = __fmul_rn(__fmul_rn(df, df), v4.w);
                }
#pragma unroll
                for (int sq = 0; sq < 4; ++sq)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        d = __fsub_rn(d, __shfl_sync(0xffffffffu, tm[e], sq, 4));
            }
            if (k < n_open && ql == 0) {
                const int32_t di = (int32_t)d;
                s_fw[wp][src][c] = ((di + ((1 << kShift) - 1)) >> kShift) - (int32_t)g.mixw_t[(size_t)id * g.n_sen + sn];
                // |GEMM - reference| on this density, from the 10 low bits the item kept of floor(-d~)
                const int32_t err = abs(((((-di) - (int32_t)((slot >> 5) & 0x3ff) + 512) & 1023) - 512));
                maxerr = max(maxerr, (unsigned)err);
                // Safety net: the certificates assumed |error| <= eps0 + (|d| >> eps_shift).  These
                // re-scored densities are a 1-2 % sample of all; when one of them uses more than
                // three quarters of that bound the whole batch is redone by the literal scan.
                if (4 * err > 3 * (eps0 + (abs(di) >> eps_shift))) qcnt[2] = 1u;
            }
        }
        __syncwarp();
        if (live) {
            int32_t f4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) f4[j] = s_fw[wp][lane][j];
            const int t = (int)it.x, sn = (int)it.y;
            raw[((size_t)(sn / spt) * T_pad + t) * spt + sn % spt] = (int16_t)exact_chain(g, f4, 4);
        }
        __syncwarp();
    }
    maxerr = __reduce_max_sync(0xffffffffu, maxerr);
    if (lane == 0 && maxerr > 2 && maxerr > qcnt[3]) atomicMax(qcnt + 3, maxerr);
}

// Queue B: near-ties among the top 5 (or a duplicated density): the literal
// evaluation of the senone -- every density exactly, the top N in the reference's
// order (descending d, the later density first on equal d: ms_gauden.c:505-520).
// One warp per item, lane = density.  After a queue overflow: every (frame, senone).
__global__ void __launch_bounds__(256)
tc_fix_b_kernel(const __grid_constant__ GmmDev g, const float4 *__restrict__ rows, const float *__restrict__ feat, int T, int T_pad, int spt,
                const uint2 *__restrict__ qb, unsigned capB, const unsigned *__restrict__ qcnt,
                int16_t *__restrict__ raw) {
    const bool all = qcnt[2] != 0;
    const unsigned long long n = all ? (unsigned long long)T * g.n_sen : min(qcnt[1], capB);
    const int lane = threadIdx.x & 31;
    const unsigned long long nw = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    for (unsigned long long i = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nw) {
        int t, s;
        if (all) { t = (int)(i / g.n_sen); s = (int)(i % g.n_sen); }
        else { const uint2 it = qb[i]; t = (int)it.x; s = (int)it.y; }
        if (s >= g.n_sen || t >= T) continue;
        float d = lane < g.n_density ? exact_dist(g, rows, feat, t, s, lane) : 0.f;
        bool in = lane < g.n_density;
        int32_t fw[4];
        for (int r = 0; r < 4; ++r) {
            // warp arg-max over (d, id): larger d, then larger id
            float bd = d; int bi = in ? lane : -1;
            for (int o = 16; o; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oi >= 0 && (bi < 0 || od > bd || (od == bd && oi > bi))) { bd = od; bi = oi; }
            }
            const int32_t di = (int32_t)bd;
            fw[r] = ((di + ((1 << kShift) - 1)) >> kShift) - (int32_t)g.mixw_t[(size_t)max(bi, 0) * g.n_sen + s];
            if (lane == bi) in = false;
        }
        if (lane == 0) raw[((size_t)(s / spt) * T_pad + t) * spt + s % spt] = (int16_t)exact_chain(g, fw, 4);
    }
}

// host-side TF32 rounding (round to nearest, ties away, like cvt.rna.tf32.f32)
float tf32_round(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u = (u + 0x1000u) & 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

// Columns of one Gaussian's B row, times -32 (see make_key), in double:
// col[0] = the whole constant det - sum mu^2 v (- 1024 mixw for the fully
// continuous path, whose accumulator is w = -32 (d - 1024 mixw)), col[2+2i] = -v_i,
// col[3+2i] = 2 mu_i v_i.  mu == nullptr: a padding Gaussian far below anything real.
void b_row_cols(std::vector<double> &col, int KP, int D, const float *mu, const float *v, float det, int mixw = 0) {
    col.assign(KP, 0.0);
    if (!mu) { col[0] = 3.0e7 * kAccScale; return; }
    double c = (double)det - 1024.0 * (double)mixw;
    for (int i = 0; i < D; ++i) {
        c -= (double)mu[i] * (double)mu[i] * (double)v[i];
        col[2 + 2 * i] = -(double)v[i];
        col[3 + 2 * i] = 2.0 * (double)mu[i] * (double)v[i];
    }
    for (int k = 0; k < KP; ++k) col[k] *= -(double)kAccScale;
    col[0] = c * -(double)kAccScale;
}

// TF32 hi/lo row in the [kstep(8 cols)][hi|lo][chunk][256 rows][4 f32] layout;
// the constant is spread over columns 0 and 1 (4 x 11 bits).
void build_b_row(float *tile, int r, int ksteps, int D, const float *mu, const float *v, float det, int mixw = 0) {
    const int KP = ksteps * 8;
    std::vector<double> col;
    b_row_cols(col, KP, D, mu, v, det, mixw);
    const double c = col[0];
    const double hi1 = tf32_round((float)c), lo1 = mu ? tf32_round((float)(c - hi1)) : 0.0;
    const double c2 = mu ? c - hi1 - lo1 : 0.0;
    const double hi2 = tf32_round((float)c2), lo2 = tf32_round((float)(c2 - hi2));
    for (int k = 0; k < KP; ++k) {
        float hi, lo;
        if (k == 0) { hi = (float)hi1; lo = (float)lo1; }
        else if (k == 1) { hi = (float)hi2; lo = (float)lo2; }
        else { hi = tf32_round((float)col[k]); lo = tf32_round((float)(col[k] - (double)hi)); }
        const int j = k / 8, c4 = (k % 8) / 4, e = k % 4;
        float *st = tile + (size_t)j * (kBStageBytes / 4);
        st[((0 * 2 + c4) * kTileN + r) * 4 + e] = hi;
        st[((1 * 2 + c4) * kTileN + r) * 4 + e] = lo;
    }
}

// fp16 hi/lo row in the [kstep(16 cols)][hi|lo][chunk][256 rows][8 f16] layout
// (the same bytes per stage); column k is stored times 2^-e[k].
void build_b_row_half(__half *tile, int r, int ksteps, int D, const float *mu, const float *v, float det,
                      const int *e, int mixw = 0) {
    const int KP = ksteps * 16;
    std::vector<double> col;
    b_row_cols(col, KP, D, mu, v, det, mixw);
    auto put = [&](int k, double val) {
        const __half hi = __float2half_rn((float)val);
        const __half lo = __float2half_rn((float)(val - (double)__half2float(hi)));
        const int j = k / 16, c8 = (k % 16) / 8, el = k % 8;
        __half *st = tile + (size_t)j * (kBStageBytes / 2);
        st[((0 * 2 + c8) * kTileN + r) * 8 + el] = hi;
        st[((1 * 2 + c8) * kTileN + r) * 8 + el] = lo;
        return val - (double)__half2float(hi) - (double)__half2float(lo);
    };
    // constant: columns 0 and 1 share the scale e[0]; column 1 takes what column 0 left over
    const double c = std::ldexp(col[0], -e[0]);
    const double rest = put(0, c);
    put(1, rest);
    for (int k = 2; k < KP; ++k) put(k, std::ldexp(col[k], -e[k]));
}

// fp16 form of the B operand: scales per n-tile and K column, the feature limits
// that go with them, and the per-batch format decision.
struct HalfOperand {
    __half *dB = nullptr;
    float *dScale = nullptr;     // [n_tiles_n][16 * ksteps] A-column factors 2^e
    float *dLim = nullptr;       // [n_tiles_n][D] largest |x_i| the tile's scaled A operand can hold (0: tile never fp16)
    unsigned int *dXmax = nullptr;   // [D] per batch
    uint8_t *dFmt = nullptr;     // [n_tiles_n] per batch
    int *dNHalf = nullptr;       // tiles on fp16 in the last batch
    int ksteps = 0, n_tiles = 0, D = 0;
    void release() {
        cudaFree(dB); cudaFree(dScale); cudaFree(dLim); cudaFree(dXmax); cudaFree(dFmt); cudaFree(dNHalf);
        dB = nullptr; dScale = dLim = nullptr; dXmax = nullptr; dFmt = nullptr; dNHalf = nullptr; ksteps = 0;
    }
};

int half_ksteps(int D) {
    const int need = (2 * D + 2 + 15) / 16;
    return need <= 2 ? 2 : (need <= 4 ? 4 : (need <= 5 ? 5 : 0));
}

bool half_enabled() {
    const char *e = getenv("B200_TC_F16");
    return !(e && atoi(e) == 0);
}

// rows(tile, r, mu, v, det, mixw) -> false for a padding row.
template <typename RowFn>
bool build_half_operand(HalfOperand &h, int n_tiles_n, int D, RowFn rows) {
    h.ksteps = half_ksteps(D);
    if (!h.ksteps || !half_enabled()) { h.ksteps = 0; return false; }
    h.n_tiles = n_tiles_n; h.D = D;
    const int KP = h.ksteps * 16;
    const size_t tile_halves = (size_t)h.ksteps * (kBStageBytes / 2);
    std::vector<__half> B((size_t)n_tiles_n * tile_halves);
    std::fill(B.begin(), B.end(), __float2half_rn(0.f));
    std::vector<float> scale((size_t)n_tiles_n * KP), lim((size_t)n_tiles_n * D);
    std::vector<double> colmax(KP), col;
    std::vector<int> e(KP);
    int n_usable = 0;
    for (int nt = 0; nt < n_tiles_n; ++nt) {
        std::fill(colmax.begin(), colmax.end(), 0.0);
        for (int r = 0; r < kTileN; ++r) {
            const float *mu, *v; float det; int mw = 0;
            const bool real = rows(nt, r, mu, v, det, mw);
            b_row_cols(col, KP, D, real ? mu : nullptr, v, det, mw);
            for (int k = 0; k < KP; ++k) colmax[k] = std::max(colmax[k], std::fabs(col[k]));
        }
        // B column k times 2^-e[k] peaks in (2^14, 2^15]; the A column carries 2^e[k],
        // itself an fp16-representable power of two
        bool fits = true;
        for (int k = 0; k < KP; ++k) {
            e[k] = 0;
            if (colmax[k] > 0.0) { int ex; std::frexp(colmax[k], &ex); e[k] = ex - 15; }
            e[k] = std::max(-14, std::min(15, e[k]));
        }
        e[1] = e[0];
        for (int k = 0; k < KP; ++k) if (std::ldexp(colmax[k], -e[k]) > 60000.0) fits = false;   // beyond fp16 even when scaled
        for (int k = 0; k < KP; ++k) scale[(size_t)nt * KP + k] = std::ldexp(1.0f, e[k]);
        for (int i = 0; i < D; ++i) {
            // x^2 * 2^e[2+2i] and |x| * 2^e[3+2i] must stay below the fp16 maximum (with margin)
            const double l2 = std::sqrt(60000.0 / std::ldexp(1.0, e[2 + 2 * i])), l1 = 60000.0 / std::ldexp(1.0, e[3 + 2 * i]);
            lim[(size_t)nt * D + i] = fits ? (float)std::min(l1, l2) : 0.f;
        }
        n_usable += fits ? 1 : 0;
        for (int r = 0; r < kTileN; ++r) {
            const float *mu, *v; float det; int mw = 0;
            const bool real = rows(nt, r, mu, v, det, mw);
            if (fits) build_b_row_half(B.data() + (size_t)nt * tile_halves, r, h.ksteps, D, real ? mu : nullptr, v, det, e.data(), mw);
        }
    }
    if (n_usable == 0) { h.ksteps = 0; return false; }
    auto up = [](void **dst, const void *src, size_t bytes) {
        return cudaMalloc(dst, bytes) == cudaSuccess && (!src || cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess);
    };
    const bool ok = up((void **)&h.dB, B.data(), B.size() * 2) && up((void **)&h.dScale, scale.data(), scale.size() * 4) &&
                    up((void **)&h.dLim, lim.data(), lim.size() * 4) && up((void **)&h.dXmax, nullptr, (size_t)D * 4) &&
                    up((void **)&h.dFmt, nullptr, (size_t)n_tiles_n) && up((void **)&h.dNHalf, nullptr, 4) &&
                    cudaMemset(h.dNHalf, 0, 4) == cudaSuccess;
    if (!ok) { cudaGetLastError(); h.release(); return false; }
    return true;
}

}  // namespace

struct TcPlan {
    int device = 0;
    int M = 0, D = 0, S = 0, ksteps = 0, spt = 0, n_tiles_n = 0;
    float *dB = nullptr;
    uint8_t *dMixw = nullptr;
    float *dA = nullptr; size_t a_cap = 0;       // tiled/transposed features (bytes)
    float *dAh = nullptr; size_t ah_cap = 0;     // same in the fp16 kernel's padding, when it differs
    float *dXrow = nullptr; size_t xrow_cap = 0; // 16-byte aligned zero-padded feature rows for the exact fix-up kernels
    HalfOperand half;                            // fp16 hi/lo form of the B operand (ksteps == 0: not available)
    int16_t *dRaw = nullptr; size_t raw_cap = 0; // bytes
    int n_sm = 148;
    int aw = 1;
    uint8_t logadd[256];
    // round 2: exact fix-up of the pairs the score kernel could not settle
    GmmDev g{};                                  // device pointers of the reference-layout parameters
    float *dCw = nullptr;                        // [n_tiles_n][256] 32*1024*mixw per tile row
    float4 *dRows = nullptr;                     // [S*M][mean Dp | var Dp] 16-byte aligned rows for the exact re-scoring gathers
    uint4 *dQa = nullptr; size_t qa_cap = 0;     // items per region
    uint2 *dQb = nullptr; size_t qb_cap = 0;     // items
    unsigned *dQcnt = nullptr;                   // [8 + kFixRegionsMax]
    int eps0 = 5, eps_shift = 18;
    long long last_T = 0;
};

bool tc_shape_supported(const GmmDev &g) {
    if (g.n_feat != 1 || g.topn != 4 || g.n_mgau != g.n_sen) return false;
    if (!(g.n_density == 8 || g.n_density == 16 || g.n_density == 32)) return false;
    const int K = 2 * g.featlen[0] + 2;
    if ((K + 7) / 8 > kMaxKSteps) return false;
    return true;
}

void tc_plan_free(TcPlan *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaFree(p->dB); cudaFree(p->dMixw); cudaFree(p->dA); cudaFree(p->dAh); cudaFree(p->dRaw); cudaFree(p->dXrow);
    cudaFree(p->dCw); cudaFree(p->dRows); cudaFree(p->dQa); cudaFree(p->dQb); cudaFree(p->dQcnt);
    p->half.release();
    delete p;
}

TcPlan *tc_plan_create(const GmmDev &g, const float *h_mean, const float *h_var, const float *h_det,
                       const uint8_t *h_mixw, int device) {
    if (!tc_shape_supported(g)) return nullptr;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        cudaGetLastError();
        return nullptr;   // tcgen05 needs sm_100; the exact path still runs
    }
    TcPlan *p = new TcPlan();
    p->device = device; p->n_sm = prop.multiProcessorCount;
    p->M = g.n_density; p->D = g.featlen[0]; p->S = g.n_sen; p->aw = g.aw;
    const int K = 2 * p->D + 2;
    const int need = (K + 7) / 8;
    p->ksteps = need <= 4 ? 4 : (need <= 7 ? 7 : 10);   // instantiated k-step counts
    p->spt = kTileN / p->M;
    p->n_tiles_n = (p->S + p->spt - 1) / p->spt;
    memcpy(p->logadd, g.logadd, 256);
    const int M = p->M, D = p->D;
    p->g = g;
    { const char *e = getenv("B200_TC_EPS0"); if (e) p->eps0 = std::max(2, atoi(e)); }
    { const char *e = getenv("B200_TC_EPS_SHIFT"); if (e) p->eps_shift = std::max(8, std::min(30, atoi(e))); }
    const size_t tile_floats = (size_t)p->ksteps * (kBStageBytes / 4);
    std::vector<float> B((size_t)p->n_tiles_n * tile_floats, 0.f);
    std::vector<uint8_t> mw((size_t)p->n_tiles_n * kTileN, 0);
    std::vector<float> cw((size_t)p->n_tiles_n * kTileN, 0.f);
    for (int nt = 0; nt < p->n_tiles_n; ++nt)
        for (int r = 0; r < kTileN; ++r) {
            const int s = nt * p->spt + r / M, dens = r % M;
            float *tile = B.data() + (size_t)nt * tile_floats;
            if (s < p->S) {
                const int q = h_mixw[(size_t)s * M + dens];
                build_b_row(tile, r, p->ksteps, D, h_mean + ((size_t)s * M + dens) * D, h_var + ((size_t)s * M + dens) * D,
                            h_det[(size_t)s * M + dens], q);
                mw[(size_t)nt * kTileN + r] = (uint8_t)q;
                cw[(size_t)nt * kTileN + r] = kAccScale * 1024.f * (float)q;
            } else {
                build_b_row(tile, r, p->ksteps, D, nullptr, nullptr, 0.f);
            }
        }
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaMalloc((void **)&p->dB, B.size() * 4) != cudaSuccess ||
        cudaMalloc((void **)&p->dMixw, mw.size()) != cudaSuccess ||
        cudaMalloc((void **)&p->dCw, cw.size() * 4) != cudaSuccess ||
        cudaMalloc((void **)&p->dQcnt, (8 + kFixRegionsMax) * sizeof(unsigned)) != cudaSuccess ||
        cudaMemset(p->dQcnt, 0, (8 + kFixRegionsMax) * sizeof(unsigned)) != cudaSuccess ||
        cudaMemcpy(p->dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p->dCw, cw.data(), cw.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(p->dMixw, mw.data(), mw.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("tensor-core plan allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        tc_plan_free(p);
        return nullptr;
    }
    {
        const int Dp = (D + 3) & ~3;
        std::vector<float> R((size_t)p->S * M * 2 * Dp, 0.f);
        for (size_t r = 0; r < (size_t)p->S * M; ++r) {
            memcpy(R.data() + r * 2 * Dp, h_mean + r * D, D * sizeof(float));
            memcpy(R.data() + r * 2 * Dp + Dp, h_var + r * D, D * sizeof(float));
        }
        if (cudaMalloc((void **)&p->dRows, R.size() * 4) != cudaSuccess ||
            cudaMemcpy(p->dRows, R.data(), R.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("tensor-core plan allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            tc_plan_free(p);
            return nullptr;
        }
    }
    build_half_operand(p->half, p->n_tiles_n, D, [&](int nt, int r, const float *&mu, const float *&v, float &det, int &q) {
        const int s = nt * p->spt + r / M, dens = r % M;
        if (s >= p->S) { mu = v = nullptr; det = 0.f; q = 0; return false; }
        mu = h_mean + ((size_t)s * M + dens) * D; v = h_var + ((size_t)s * M + dens) * D; det = h_det[(size_t)s * M + dens];
        q = h_mixw[(size_t)s * M + dens];
        return true;
    });
    return p;
}

template <int M, int KS, int MODE, int HALF>
static int launch_score(const TcParams &prm, int grid, cudaStream_t st) {
    // the epilogue warps' rings of undecided pairs, where the B tile leaves room for them
    constexpr int HW = (MODE == 0 && score_smem_bytes(KS, HALF, 1, M) <= 227 * 1024) ? 1 : 0;
    constexpr size_t smem = score_smem_bytes(KS, HALF, HW, M);
    static AttrOnce attr;
    if (attr.need()) {
        B200_CUDA_OK(cudaFuncSetAttribute(tc_score_kernel<M, KS, MODE, HALF, HW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    tc_score_kernel<M, KS, MODE, HALF, HW><<<grid, score_threads(HW), smem, st>>>(prm);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

template <int M, int MODE>
static int launch_score_ks(const TcParams &prm, int ks, int grid, cudaStream_t st) {
    switch (ks) {
        case 4: return launch_score<M, 4, MODE, 0>(prm, grid, st);
        case 7: return launch_score<M, 7, MODE, 0>(prm, grid, st);
        case 10: return launch_score<M, 10, MODE, 0>(prm, grid, st);
    }
    set_error("tensor-core path: %d k-steps unsupported", ks);
    return B200_ERR_UNSUP;
}

template <int M, int MODE>
static int launch_score_ks_half(const TcParams &prm, int ks, int grid, cudaStream_t st) {
    switch (ks) {
        case 2: return launch_score<M, 2, MODE, 1>(prm, grid, st);
        case 4: return launch_score<M, 4, MODE, 1>(prm, grid, st);
        case 5: return launch_score<M, 5, MODE, 1>(prm, grid, st);
    }
    set_error("tensor-core path (fp16 operands): %d k-steps unsupported", ks);
    return B200_ERR_UNSUP;
}

// Feature tiles for one launch pair: the TF32 layout always, the fp16 layout
// too when its padded dimension count differs; reduces max |x_i| per dimension and
// decides every n-tile's operand format for this batch.
static int prep_features(const float *d_feat, int T, int stride, int off, int D, int ks_tf32, const HalfOperand &h,
                         float **dX, size_t *x_cap, float **dXh, size_t *xh_cap, const float **gx_half, cudaStream_t st,
                         float *xrow = nullptr) {
    const int n_tiles_m = (T + kTileM - 1) / kTileM;
    const int Dp = 4 * ks_tf32, Dph = 8 * h.ksteps;
    const size_t bytes = (size_t)n_tiles_m * Dp * kTileM * sizeof(float);
    if (*x_cap < bytes) {
        cudaFree(*dX); *dX = nullptr; *x_cap = 0;
        B200_CUDA_OK(cudaMalloc((void **)dX, bytes));
        *x_cap = bytes;
    }
    if (h.ksteps) {
        B200_CUDA_OK(cudaMemsetAsync(h.dXmax, 0, (size_t)D * 4, st));
        B200_CUDA_OK(cudaMemsetAsync(h.dNHalf, 0, 4, st));
    }
    tc_prep_kernel<<<n_tiles_m, kTileM, 0, st>>>(d_feat, T, stride, off, D, Dp, *dX, h.ksteps ? h.dXmax : nullptr, xrow);
    B200_LAUNCH_CHECK();
    if (h.ksteps) {
        tc_tile_format_kernel<<<(h.n_tiles + 255) / 256, 256, 0, st>>>(h.dXmax, h.dLim, D, h.n_tiles, h.dFmt, h.dNHalf);
        B200_LAUNCH_CHECK();
    }
    *gx_half = *dX;
    if (h.ksteps && Dph != Dp) {
        const size_t hb = (size_t)n_tiles_m * Dph * kTileM * sizeof(float);
        if (*xh_cap < hb) {
            cudaFree(*dXh); *dXh = nullptr; *xh_cap = 0;
            B200_CUDA_OK(cudaMalloc((void **)dXh, hb));
            *xh_cap = hb;
        }
        tc_prep_kernel<<<n_tiles_m, kTileM, 0, st>>>(d_feat, T, stride, off, D, Dph, *dXh, nullptr, nullptr);
        B200_LAUNCH_CHECK();
        *gx_half = *dXh;
    }
    return B200_OK;
}

int tc_score_raw(TcPlan *p, const float *d_feat, int T, cudaStream_t st, cudaEvent_t *ev_prep, int *T_pad_out, cudaEvent_t *ev_fix) {
    const int n_tiles_m = (T + kTileM - 1) / kTileM;
    const int T_pad = n_tiles_m * kTileM;
    const size_t raw_bytes = (size_t)p->n_tiles_n * T_pad * p->spt * sizeof(int16_t);
    if (p->raw_cap < raw_bytes) {
        cudaFree(p->dRaw); p->dRaw = nullptr; p->raw_cap = 0;
        B200_CUDA_OK(cudaMalloc((void **)&p->dRaw, raw_bytes));
        p->raw_cap = raw_bytes;
    }
    const float *gx_half = nullptr;
    const size_t xrow_bytes = (size_t)T_pad * ((p->D + 3) & ~3) * sizeof(float);
    if (p->xrow_cap < xrow_bytes) {
        cudaFree(p->dXrow); p->dXrow = nullptr; p->xrow_cap = 0;
        B200_CUDA_OK(cudaMalloc((void **)&p->dXrow, xrow_bytes));
        p->xrow_cap = xrow_bytes;
    }
    int rc0 = prep_features(d_feat, T, p->D, 0, p->D, p->ksteps, p->half, &p->dA, &p->a_cap, &p->dAh, &p->ah_cap, &gx_half, st, p->dXrow);
    if (rc0) return rc0;
    if (ev_prep) cudaEventRecord(*ev_prep, st);

    TcParams prm;
    prm.gB = p->dB; prm.gX = p->dA; prm.gMixw = p->dMixw; prm.raw = p->dRaw;
    prm.T = T; prm.T_pad = T_pad; prm.n_sen = p->S; prm.n_tiles_m = n_tiles_m; prm.n_tiles_n = p->n_tiles_n;
    prm.ksteps = p->ksteps; prm.aw = p->aw;
    { const char *e = getenv("B200_TC_DBG"); prm.dbg = e ? atoi(e) : 0; }
    prm.m31 = 31; prm.part = nullptr; prm.scaleA = nullptr; prm.fmt = nullptr;
    // split the frame axis so that there are >= ~16 units per CTA, but never
    // less than 8 frame tiles per unit (B reload amortisation)
    int m_chunks = 1;
    while ((long long)p->n_tiles_n * m_chunks < 16LL * p->n_sm && (n_tiles_m + m_chunks) / (m_chunks + 1) >= 8) ++m_chunks;
    prm.m_chunks = m_chunks;
    prm.tiles_per_chunk = (n_tiles_m + m_chunks - 1) / m_chunks;
    prm.m_chunks = (n_tiles_m + prm.tiles_per_chunk - 1) / prm.tiles_per_chunk;
    prm.n_units = p->n_tiles_n * prm.m_chunks;
    memcpy(prm.logadd, p->logadd, 256);
    const int grid = std::min(prm.n_units, p->n_sm);
    *T_pad_out = T_pad;
    // fix-up queues: A holds up to 1/8 of all pairs (split into one region per CTA), B 1/64
    const size_t pairs = (size_t)T_pad * p->S;
    const size_t capA = std::max<size_t>(4096, pairs / 8 / grid), capB = std::max<size_t>(65536, pairs / 64);
    if (p->qa_cap < capA) {
        cudaFree(p->dQa); p->dQa = nullptr; p->qa_cap = 0;
        B200_CUDA_OK(cudaMalloc((void **)&p->dQa, capA * kFixRegionsMax * sizeof(uint4)));
        p->qa_cap = capA;
    }
    if (p->qb_cap < capB) {
        cudaFree(p->dQb); p->dQb = nullptr; p->qb_cap = 0;
        B200_CUDA_OK(cudaMalloc((void **)&p->dQb, capB * sizeof(uint2)));
        p->qb_cap = capB;
    }
    B200_CUDA_OK(cudaMemsetAsync(p->dQcnt, 0, (8 + kFixRegionsMax) * sizeof(unsigned), st));
    prm.gCw = p->dCw; prm.qa = p->dQa; prm.qb = p->dQb; prm.qcnt = p->dQcnt;
    prm.capA = (unsigned)p->qa_cap; prm.capB = (unsigned)p->qb_cap;
    prm.eps0 = p->eps0; prm.eps_shift = p->eps_shift;
    prm.epsA = ((float)p->eps0 + 1.5f) / 1024.f;
    prm.epsB = std::ldexp(1.0f, -(15 + p->eps_shift));
    p->last_T = T;
    int rc = B200_OK;
    if (p->half.ksteps) {
        // fp16 operands first; the TF32 kernel below returns at once unless a feature overflowed fp16
        TcParams ph = prm;
        ph.gB = reinterpret_cast<const float *>(p->half.dB); ph.gX = gx_half; ph.ksteps = p->half.ksteps;
        ph.scaleA = p->half.dScale; ph.fmt = p->half.dFmt;
        switch (p->M) {
            case 8: rc = launch_score_ks_half<8, 0>(ph, ph.ksteps, grid, st); break;
            case 16: rc = launch_score_ks_half<16, 0>(ph, ph.ksteps, grid, st); break;
            case 32: rc = launch_score_ks_half<32, 0>(ph, ph.ksteps, grid, st); break;
        }
        if (rc) return rc;
        prm.fmt = p->half.dFmt;
    }
    switch (p->M) {
        case 8: rc = launch_score_ks<8, 0>(prm, p->ksteps, grid, st); break;
        case 16: rc = launch_score_ks<16, 0>(prm, p->ksteps, grid, st); break;
        case 32: rc = launch_score_ks<32, 0>(prm, p->ksteps, grid, st); break;
        default: set_error("tensor-core path: n_density %d unsupported", p->M); return B200_ERR_UNSUP;
    }
    if (rc) return rc;
    if (prm.dbg & 8) return B200_OK;      // development: leave the queued pairs un-fixed
    if (ev_fix) cudaEventRecord(*ev_fix, st);   // score kernel(s) | exact fix-up kernels
    tc_fix_a_kernel<<<dim3(8, grid), 256, 0, st>>>(p->g, p->dRows, p->dXrow, T_pad, p->spt, p->dQa, prm.capA, p->dQcnt, p->dRaw, p->eps0, p->eps_shift);
    B200_LAUNCH_CHECK();
    tc_fix_b_kernel<<<p->n_sm * 4, 256, 0, st>>>(p->g, p->dRows, p->dXrow, T, T_pad, p->spt, p->dQb, prm.capB, p->dQcnt, p->dRaw);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// {frames x senones of the last call, pairs that took the hard path, queue A items,
//  queue B items, queue overflow (everything redone exactly), largest |GEMM - reference|
//  distance seen on a re-scored density (raw log units; 0 if <= 2)}   (synchronises)
int tc_last_stats(TcPlan *p, long long out[7]) {
    for (int i = 0; i < 7; ++i) out[i] = 0;
    if (!p) return 0;
    std::vector<unsigned> c(8 + kFixRegionsMax);
    cudaSetDevice(p->device);
    if (cudaMemcpy(c.data(), p->dQcnt, c.size() * sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    out[0] = p->last_T * p->S; out[1] = c[4]; out[3] = c[1]; out[4] = c[2]; out[5] = c[3]; out[6] = c[5];
    for (int r = 0; r < kFixRegionsMax; ++r) out[2] += c[8 + r];
    return 0;
}

// 1: every n-tile of the last tc_score_raw ran on fp16 operands, 0: none did (no fp16
// operand, or the batch's features exceed every tile's range), 2: some did
int tc_last_format(TcPlan *p) {
    if (!p || !p->half.ksteps) return 0;
    int n = 0;
    cudaSetDevice(p->device);
    if (cudaMemcpy(&n, p->half.dNHalf, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return n == 0 ? 0 : (n == p->n_tiles_n ? 1 : 2);
}

int tc_finish(TcPlan *p, int T, int T_pad, int subtract_best, int16_t *d_out, cudaStream_t st) {
    const size_t stride = (size_t)p->n_tiles_n * p->spt;
    static AttrOnce attr;
    if (attr.need()) {
        B200_CUDA_OK(cudaFuncSetAttribute(tc_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        B200_CUDA_OK(cudaFuncSetAttribute(tc_finish_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    const bool vec = (p->spt % 8 == 0) && (p->S % 8 == 0) && ((reinterpret_cast<size_t>(d_out) & 15) == 0) &&
                     (256 % (kFinFrames * (p->spt / 8)) == 0);
    if (vec) {
        const size_t sh = (size_t)kFinFrames * stride * sizeof(int16_t);
        if (sh > 200 * 1024) { set_error("n_sen too large for the finish kernel"); return B200_ERR_UNSUP; }
        tc_finish_vec_kernel<<<(T + kFinFrames - 1) / kFinFrames, 256, sh, st>>>(p->dRaw, T, T_pad, p->S, p->spt,
                                                                                 p->n_tiles_n, subtract_best, d_out);
    } else {
        const size_t sh = (size_t)kNormFrames * stride * sizeof(int16_t);
        if (sh > 200 * 1024) { set_error("n_sen too large for the finish kernel"); return B200_ERR_UNSUP; }
        tc_finish_kernel<<<(T + kNormFrames - 1) / kNormFrames, 256, sh, st>>>(p->dRaw, T, T_pad, p->S, p->spt,
                                                                               p->n_tiles_n, subtract_best, d_out);
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}


// ====================================================================
// Tied codebooks (ptm_mgau / s2_semi_mgau): the codebook stage
// (PS/ptm_mgau.c:98-231 eval_topn + eval_cb, PS/s2_semi_mgau.c:80-187) on
// the tensor cores, bit-identical to the exact kernel (gmm_topn_kernel<N,1|2>).
//
//   1. tc_score_kernel<32,KS,1>: the same TF32x3 GEMM; every epilogue thread
//      keeps the 4 best keys of its 64 accumulator columns       (approximate d)
//   2. tied_select_kernel: per (frame, codebook) the 8 best keys over all
//      column groups -> EXACT sequential float32 distances of those 8
//      candidates (the reference's arithmetic) -> top-N.  Two cheap sufficient
//      conditions prove that the result equals the reference's full scan:
//        (a) the 5 best candidates have pairwise different (int32) scores
//            (then no tie / same-integer-bucket rule of eval_cb can matter), and
//        (b) the 5th best exact distance beats  max(d~(8th candidate), best d~ any
//            64-column group may have dropped) + eps, an upper bound on every
//            density that is not a candidate.
//   3. tied_fallback_kernel: the few pairs where (a) or (b) fails (~1e-3) get
//      the reference's literal scan over the whole codebook.
// ====================================================================
namespace {

constexpr int kCand = 8;
constexpr float kTiedEps = 32.f;      // bound on |approximate - exact| distance (raw log units)

// Stage 2a: per (frame, codebook) the 8 best keys over the codebook's tiles x 4
// column groups -> candidate density indices (+ the bound of condition (b)).
// block = 128 frames x one codebook (blockIdx.y); reads are 64 B per thread and
// tile, contiguous across the warp.
__global__ void __launch_bounds__(128)
tied_merge_kernel(int n_density, int tn, int T_pad, int tpc, const uint4 *__restrict__ part,
                  int32_t *__restrict__ cand /* [n_mgau][tn][8] */, float *__restrict__ bound /* [n_mgau][tn][3]: bound, approximate best, lane mask */) {
    const int tl = blockIdx.x * 128 + threadIdx.x, mg = blockIdx.y;
    if (tl >= tn) return;
    int32_t k8[kCand]; int p8[kCand];
#pragma unroll
    for (int j = 0; j < kCand; ++j) { k8[j] = 0x7fffffff; p8[j] = 0; }
    // Values a 64-column group dropped (its 5th best and worse) never reach this
    // kernel; each of them is >= the group's 4th kept key, so the smallest such
    // key over all groups bounds them all.
    int32_t gmin = 0x7fffffff;
    // two tiles (8 independent 16-byte loads) in flight; the insertion is branch-free
    // inside (static register indices) and compiled once per key slot, not once per tile
#pragma unroll 1
    for (int j0 = 0; j0 < tpc; j0 += 2) {
        uint4 q4[2][4];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int j = min(j0 + jj, tpc - 1);
            const uint4 *src = part + ((size_t)(mg * tpc + j) * T_pad + tl) * 4;
#pragma unroll
            for (int cg = 0; cg < 4; ++cg) q4[jj][cg] = src[cg];
        }
#pragma unroll 1
        for (int u = 0; u < 8; ++u) {
            const int jj = u >> 2, cg = u & 3;
            if (j0 + jj >= tpc) break;
            uint4 q = q4[0][0];
#pragma unroll
            for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                for (int c2 = 0; c2 < 4; ++c2) if (a2 == jj && c2 == cg) q = q4[a2][c2];
            const int32_t key[4] = {(int32_t)q.x, (int32_t)q.y, (int32_t)q.z, (int32_t)q.w};
            gmin = min(gmin, key[3]);
            const int pos = (j0 + jj) * 4 + cg;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int32_t kv = key[e];
                if (kv < k8[kCand - 1]) {
#pragma unroll
                    for (int r = kCand - 1; r >= 1; --r) {
                        const bool sh = kv < k8[r - 1];
                        const bool here = !sh && kv < k8[r];
                        k8[r] = sh ? k8[r - 1] : (here ? kv : k8[r]);
                        p8[r] = sh ? p8[r - 1] : (here ? pos : p8[r]);
                    }
                    if (kv < k8[0]) { k8[0] = kv; p8[0] = pos; }
                }
            }
        }
    }
    int32_t *o = cand + ((size_t)mg * tn + tl) * kCand;
    int4 lo, hi;
    int32_t idx[kCand];
#pragma unroll
    for (int c = 0; c < kCand; ++c) {
        const int i = (p8[c] >> 2) * kTileN + (p8[c] & 3) * 64 + (63 - (k8[c] & 63));
        idx[c] = (k8[c] != 0x7fffffff && i < n_density) ? i : -1;
    }
    lo = make_int4(idx[0], idx[1], idx[2], idx[3]); hi = make_int4(idx[4], idx[5], idx[6], idx[7]);
    reinterpret_cast<int4 *>(o)[0] = lo; reinterpret_cast<int4 *>(o)[1] = hi;
    // every density that is not a candidate -- kept by its group but not among the
    // 8 best, or dropped inside its group -- has approximate distance <= this
    bound[((size_t)mg * tn + tl) * 3] = -(float)min(k8[kCand - 1], gmin) * (1.0f / kAccScale);
    bound[((size_t)mg * tn + tl) * 3 + 1] = -(float)k8[0] * (1.0f / kAccScale);   // approximate best, for the error statistic
    // Candidates 6..8 only have to be re-scored when they could still reach the 5 best:
    // one whose approximate distance is more than 3 eps below the 5th candidate's cannot
    // (|approximate - exact| <= eps on both), so its lane skips the gather.
    const float thr = (float)k8[4] + 3.f * kAccScale * (kTiedEps + 1.5e-5f * fabsf((float)k8[4]) * (1.0f / kAccScale));
    int need = 0x1f;
#pragma unroll
    for (int c = 5; c < kCand; ++c) need |= ((float)k8[c] <= thr) ? (1 << c) : 0;
    bound[((size_t)mg * tn + tl) * 3 + 2] = __int_as_float(need);
}

// Stage 2b: 8 lanes per (frame, codebook), one candidate each: exact sequential
// float32 distance (the reference's arithmetic) from a padded parameter copy
// (row = [mean[lenp] | var[lenp]], 16-byte gathers), rank by shuffles, prove
// conditions (a) and (b), write the top-N list or flag the pair.
template <int N>
__global__ void __launch_bounds__(128)
tied_rescore_kernel(GmmDev g, int f, const float *__restrict__ feat, int t0, int tn,
                    const int32_t *__restrict__ cand, const float *__restrict__ bound,
                    const float4 *__restrict__ rows, int lenp, int2 *__restrict__ lists,
                    int2 *__restrict__ flagged, int *__restrict__ n_flagged) {
    const int mg = blockIdx.y;
    const int c = threadIdx.x & 7;
    const int tl = blockIdx.x * 16 + (threadIdx.x >> 3);
    if (tl >= tn) return;                      // whole 8-lane groups leave together
    const unsigned gm = 0xffu << (threadIdx.x & 24);
    const int len = g.featlen[f], q = lenp >> 2;
    const int idx0 = cand[((size_t)mg * tn + tl) * kCand + c];
    const int idx = idx0 >= 0 ? idx0 : 0;
    const bool need = (__float_as_int(bound[((size_t)mg * tn + tl) * 3 + 2]) >> c) & 1;
    const float4 *rp = rows + ((size_t)mg * g.n_density + idx) * (2 * q);
    const float *x = feat + (size_t)(t0 + tl) * g.veclen + g.featoff[f];
    float d = __ldg(g.det + ((size_t)mg * g.n_feat + f) * g.n_density + idx);
    if (need)
#pragma unroll 5
    for (int i4 = 0; i4 < q; ++i4) {
        const float4 m4 = __ldg(rp + i4), v4 = __ldg(rp + q + i4);
        const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (i4 * 4 + e < len) {
                const float diff = __fsub_rn(__ldg(x + i4 * 4 + e), mm[e]);
                d = __fsub_rn(d, __fmul_rn(__fmul_rn(diff, diff), vv[e]));
            }
        }
    }
    if (idx0 < 0 || !need) d = -3.0e38f;      // a skipped candidate is provably outside the N+1 best
    // rank among the 8 (descending, earlier candidate first on equal distances)
    int rank = 0;
#pragma unroll
    for (int o = 1; o < 8; ++o) {
        const int src = (c + o) & 7;
        const float od = __shfl_sync(gm, d, (threadIdx.x & 24) + src, 32);
        rank += (od > d || (od == d && src < c)) ? 1 : 0;
    }
    // (a): the N+1 best have pairwise different integer scores
    const int32_t di = (int32_t)d;
    bool clash = false;
#pragma unroll
    for (int o = 1; o < 8; ++o) {
        const int src = (threadIdx.x & 24) + ((c + o) & 7);
        const int orank = __shfl_sync(gm, rank, src, 32);
        const int32_t odi = __shfl_sync(gm, di, src, 32);
        clash |= (rank <= N && orank <= N && odi == di);
    }
    // (b): the (N+1)-th best beats everything that is not a candidate
    const float bd = bound[((size_t)mg * tn + tl) * 3] + kTiedEps + 1.5e-5f * fabsf(d);
    // statistic: largest |GEMM distance - exact distance| seen on a best candidate (+2 for the id bits)
    if (c == 0 && idx0 >= 0) {
        const float err = fabsf(bound[((size_t)mg * tn + tl) * 3 + 1] - d);
        if (err > 4.f) atomicMax(n_flagged + 2, (int)fminf(err, 1.0e9f));
        // Safety net: the completeness proof (b) assumes |GEMM - exact| <= eps for the densities that
        // were NOT re-scored.  The best candidates are the sample we can check; when one of them uses
        // more than half of the bound, the whole launch is redone by the literal scan.
        if (err > 0.5f * (kTiedEps + 1.5e-5f * fabsf(d))) n_flagged[3] = 1;
    }
    const bool bad = clash || idx0 < 0 || (rank == N && !(d > bd));
    const bool ok = (__ballot_sync(gm, bad) & gm) == 0;
    if (ok) {
        if (rank < N)
            lists[((size_t)tl * g.n_mgau * g.n_feat + (size_t)mg * g.n_feat + f) * N + rank] = make_int2(idx0, di);
    } else if (c == 0) {
        flagged[atomicAdd(n_flagged, 1)] = make_int2(tl, mg);
    }
}

// The reference's literal scan for the flagged (frame, codebook) pairs: all
// distances in parallel into shared memory, then one thread replays eval_topn +
// eval_cb in index order (same code as gmm_topn_kernel).
template <int N, int MODE>
__global__ void __launch_bounds__(256)
tied_fallback_kernel(GmmDev g, int f, const float *__restrict__ feat, int t0, int tn, const int2 *__restrict__ flagged,
                     const int *__restrict__ n_flagged, const float4 *__restrict__ rows, int lenp,
                     int2 *__restrict__ lists) {
    extern __shared__ float sh[];   // x[lenp] | d[n_density]
    const int len = g.featlen[f], q = lenp >> 2;
    float *x = sh, *dd = sh + lenp;
    const bool all = n_flagged[3] != 0;                    // the error monitor tripped: every (frame, codebook) pair
    const int n = all ? tn * g.n_mgau : *n_flagged;
    for (int w = blockIdx.x; w < n; w += gridDim.x) {
        const int tl = all ? w / g.n_mgau : flagged[w].x, mg = all ? w % g.n_mgau : flagged[w].y;
        __syncthreads();
        for (int i = threadIdx.x; i < lenp; i += blockDim.x)
            x[i] = i < len ? feat[(size_t)(t0 + tl) * g.veclen + g.featoff[f] + i] : 0.f;
        __syncthreads();
        const size_t dbase = ((size_t)mg * g.n_feat + f) * g.n_density;
        for (int c = threadIdx.x; c < g.n_density; c += blockDim.x) {
            const float4 *rp = rows + ((size_t)mg * g.n_density + c) * (2 * q);
            float d = g.det[dbase + c];
#pragma unroll 5
            for (int i4 = 0; i4 < q; ++i4) {
                const float4 m4 = __ldg(rp + i4), v4 = __ldg(rp + q + i4);
                const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (i4 * 4 + e < len) {
                        const float diff = __fsub_rn(x[i4 * 4 + e], mm[e]);
                        d = __fsub_rn(d, __fmul_rn(__fmul_rn(diff, diff), vv[e]));
                    }
                }
            }
            dd[c] = d;
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp 0 replays the scan: the list is replicated in every lane; 32
            // densities are tested at once against the current worst score (it
            // only ever rises, so a density that fails now would fail later too)
            // and the survivors are inserted one by one in index order.
            const int lane = threadIdx.x;
            TopI<N> ti;
            ti.init();
#pragma unroll
            for (int i = 0; i < N; ++i) ti.seed(i, (int32_t)dd[i]);
            for (int c0 = 0; c0 < g.n_density; c0 += 32) {
                const int c = c0 + lane;
                const float d = c < g.n_density ? dd[c] : 0.f;
                bool pend = c >= N && c < g.n_density;
                for (;;) {
                    const int32_t worst = ti.s[N - 1];
                    if (MODE == 1) pend = pend && !(d < (float)worst);
                    else pend = pend && !((int32_t)d < worst);
                    const unsigned m = __ballot_sync(0xffffffffu, pend);
                    if (!m) break;
                    const int src = __ffs(m) - 1;
                    const float ds = __shfl_sync(0xffffffffu, d, src);
                    ti.insert((int32_t)ds, c0 + src);
                    if (lane == src) pend = false;
                }
            }
            if (lane == 0) {
                int2 *o = lists + ((size_t)tl * g.n_mgau * g.n_feat + (size_t)mg * g.n_feat + f) * N;
#pragma unroll
                for (int j = 0; j < N; ++j) o[j] = make_int2(ti.cw[j], ti.s[j]);
            }
        }
    }
}

__global__ void tied_count_roll(int *c, int pairs) { c[1] += c[3] ? pairs : c[0]; c[0] = 0; c[3] = 0; }

}  // namespace

struct TcTied {
    int device = 0, mode = 1, n_feat = 0, tpc = 0, n_tiles_n = 0, n_sm = 148, topn = 4;
    int ksteps[B200_MAX_STREAMS] = {0, 0, 0, 0};
    float *dB[B200_MAX_STREAMS] = {nullptr, nullptr, nullptr, nullptr};
    float4 *dRows[B200_MAX_STREAMS] = {nullptr, nullptr, nullptr, nullptr};   // [mgau][density][mean lenp | var lenp]
    HalfOperand half[B200_MAX_STREAMS];
    float *dXh = nullptr; size_t xh_cap = 0;
    int lenp[B200_MAX_STREAMS] = {0, 0, 0, 0};
    float *dX = nullptr; size_t x_cap = 0;
    uint4 *dPart = nullptr; size_t part_cap = 0;
    int2 *dFlag = nullptr; size_t flag_cap = 0;
    int32_t *dCand = nullptr; float *dBound = nullptr; size_t cand_cap = 0;   // pairs
    int *dCount = nullptr;          // [1 + n_feat]: running total per stream launch
    long long pairs = 0;
    int chunk = 0;                  // frames per internal chunk
};

void tc_tied_free(TcTied *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (auto b : p->dB) cudaFree(b);
    for (auto b : p->dRows) cudaFree(b);
    for (auto &h : p->half) h.release();
    cudaFree(p->dXh);
    cudaFree(p->dX); cudaFree(p->dPart); cudaFree(p->dFlag); cudaFree(p->dCount); cudaFree(p->dCand); cudaFree(p->dBound);
    delete p;
}

TcTied *tc_tied_create(const GmmDev &g, int mode, const float *h_mean, const float *h_var, const float *h_det,
                       int device) {
    if (g.n_density % kTileN != 0 || g.topn > 4 || g.topn + 1 > kCand || g.n_density < 8) return nullptr;
    for (int f = 0; f < g.n_feat; ++f)
        if ((2 * g.featlen[f] + 2 + 7) / 8 > kMaxKSteps) return nullptr;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) { cudaGetLastError(); return nullptr; }
    TcTied *p = new TcTied();
    p->device = device; p->mode = mode; p->n_feat = g.n_feat; p->topn = g.topn; p->n_sm = prop.multiProcessorCount;
    p->tpc = g.n_density / kTileN;
    p->n_tiles_n = g.n_mgau * p->tpc;
    bool ok = cudaSetDevice(device) == cudaSuccess;
    for (int f = 0; f < g.n_feat && ok; ++f) {
        const int D = g.featlen[f];
        const int need = (2 * D + 2 + 7) / 8;
        p->ksteps[f] = need <= 4 ? 4 : (need <= 7 ? 7 : 10);
        const size_t tile_floats = (size_t)p->ksteps[f] * (kBStageBytes / 4);
        std::vector<float> B((size_t)p->n_tiles_n * tile_floats, 0.f);
        for (int mg = 0; mg < g.n_mgau; ++mg) {
            const size_t pbase = (size_t)mg * g.n_density * g.veclen + (size_t)g.n_density * g.featoff[f];
            const size_t dbase = ((size_t)mg * g.n_feat + f) * g.n_density;
            for (int c = 0; c < g.n_density; ++c)
                build_b_row(B.data() + (size_t)(mg * p->tpc + c / kTileN) * tile_floats, c % kTileN, p->ksteps[f], D,
                            h_mean + pbase + (size_t)c * D, h_var + pbase + (size_t)c * D, h_det[dbase + c]);
        }
        ok = cudaMalloc((void **)&p->dB[f], B.size() * 4) == cudaSuccess &&
             cudaMemcpy(p->dB[f], B.data(), B.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        // padded parameter rows for the exact re-scoring gathers
        const int lenp = (D + 3) & ~3;
        p->lenp[f] = lenp;
        std::vector<float> R((size_t)g.n_mgau * g.n_density * 2 * lenp, 0.f);
        for (int mg = 0; mg < g.n_mgau; ++mg) {
            const size_t pbase = (size_t)mg * g.n_density * g.veclen + (size_t)g.n_density * g.featoff[f];
            for (int c = 0; c < g.n_density; ++c) {
                float *row = R.data() + ((size_t)mg * g.n_density + c) * 2 * lenp;
                memcpy(row, h_mean + pbase + (size_t)c * D, D * sizeof(float));
                memcpy(row + lenp, h_var + pbase + (size_t)c * D, D * sizeof(float));
            }
        }
        ok = ok && cudaMalloc((void **)&p->dRows[f], R.size() * 4) == cudaSuccess &&
             cudaMemcpy(p->dRows[f], R.data(), R.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
        if (ok)
            build_half_operand(p->half[f], p->n_tiles_n, D, [&](int nt, int r, const float *&mu, const float *&v, float &det, int &) {
                const int mg = nt / p->tpc, c = (nt % p->tpc) * kTileN + r;
                const size_t pbase = (size_t)mg * g.n_density * g.veclen + (size_t)g.n_density * g.featoff[f];
                mu = h_mean + pbase + (size_t)c * D; v = h_var + pbase + (size_t)c * D;
                det = h_det[((size_t)mg * g.n_feat + f) * g.n_density + c];
                return true;
            });
    }
    ok = ok && cudaMalloc((void **)&p->dCount, 16 * sizeof(int)) == cudaSuccess;
    if (!ok) {
        set_error("tensor-core plan (tied) allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        tc_tied_free(p);
        return nullptr;
    }
    // internal frame chunk: partial lists are 64 B per (tile, frame); keep them under ~2 GB
    long long c = (2LL << 30) / ((long long)p->n_tiles_n * 64);
    c = std::max(1024LL, std::min(32768LL, c)) / kTileM * kTileM;
    p->chunk = (int)c;
    return p;
}

void tc_tied_stats(TcTied *p, long long out[3]) {
    out[0] = p ? p->pairs : 0;
    out[1] = out[2] = 0;
    if (!p) return;
    int c[3] = {0, 0, 0};
    cudaSetDevice(p->device);
    if (cudaMemcpy(c, p->dCount, sizeof(c), cudaMemcpyDeviceToHost) == cudaSuccess) { out[1] = c[1]; out[2] = c[2]; }
}

template <int N>
static int tied_select_launch(TcTied *p, const GmmDev &g, int f, const float *d_feat, int t0, int tn, int T_pad,
                              int2 *lists, cudaStream_t st) {
    const int len = g.featlen[f];
    tied_merge_kernel<<<dim3((tn + 127) / 128, g.n_mgau), 128, 0, st>>>(g.n_density, tn, T_pad, p->tpc, p->dPart,
                                                                       p->dCand, p->dBound);
    B200_LAUNCH_CHECK();
    tied_rescore_kernel<N><<<dim3((tn + 15) / 16, g.n_mgau), 128, 0, st>>>(
        g, f, d_feat, t0, tn, p->dCand, p->dBound, p->dRows[f], p->lenp[f], lists, p->dFlag, p->dCount);
    B200_LAUNCH_CHECK();
    const size_t sh = ((size_t)p->lenp[f] + g.n_density) * 4;
    if (sh > 200 * 1024) { set_error("codebook too large for the fallback kernel"); return B200_ERR_UNSUP; }
    auto kern = p->mode == 1 ? tied_fallback_kernel<N, 1> : tied_fallback_kernel<N, 2>;
    static AttrOnce attr[2];
    if (attr[p->mode - 1].need()) {
        B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    kern<<<p->n_sm * 2, 256, sh, st>>>(g, f, d_feat, t0, tn, p->dFlag, p->dCount, p->dRows[f], p->lenp[f], lists);
    B200_LAUNCH_CHECK();
    (void)len;
    return B200_OK;
}

int tc_tied_lists(TcTied *p, const GmmDev &g, const float *d_feat, int t0, int tn, int2 *lists, cudaStream_t st) {
    p->pairs = 0;
    B200_CUDA_OK(cudaMemsetAsync(p->dCount, 0, 16 * sizeof(int), st));
    for (int c0 = 0; c0 < tn; c0 += p->chunk) {
        const int cn = std::min(p->chunk, tn - c0);
        const int n_tiles_m = (cn + kTileM - 1) / kTileM, T_pad = n_tiles_m * kTileM;
        const size_t part_bytes = (size_t)p->n_tiles_n * T_pad * 4 * sizeof(uint4);
        const size_t flag_bytes = (size_t)cn * g.n_mgau * sizeof(int2);
        if (p->part_cap < part_bytes) {
            cudaFree(p->dPart); p->dPart = nullptr; p->part_cap = 0;
            B200_CUDA_OK(cudaMalloc((void **)&p->dPart, part_bytes));
            p->part_cap = part_bytes;
        }
        if (p->cand_cap < (size_t)cn * g.n_mgau) {
            cudaFree(p->dCand); cudaFree(p->dBound); p->dCand = nullptr; p->dBound = nullptr; p->cand_cap = 0;
            B200_CUDA_OK(cudaMalloc((void **)&p->dCand, (size_t)cn * g.n_mgau * kCand * sizeof(int32_t)));
            B200_CUDA_OK(cudaMalloc((void **)&p->dBound, (size_t)cn * g.n_mgau * 3 * sizeof(float)));
            p->cand_cap = (size_t)cn * g.n_mgau;
        }
        if (p->flag_cap < flag_bytes) {
            cudaFree(p->dFlag); p->dFlag = nullptr; p->flag_cap = 0;
            B200_CUDA_OK(cudaMalloc((void **)&p->dFlag, flag_bytes));
            p->flag_cap = flag_bytes;
        }
        int2 *lc = lists + (size_t)c0 * g.n_mgau * g.n_feat * g.topn;
        for (int f = 0; f < g.n_feat; ++f) {
            const float *gx_half = nullptr;
            int rc = prep_features(d_feat + (size_t)(t0 + c0) * g.veclen, cn, g.veclen, g.featoff[f], g.featlen[f], p->ksteps[f],
                                   p->half[f], &p->dX, &p->x_cap, &p->dXh, &p->xh_cap, &gx_half, st);
            if (rc) return rc;
            TcParams prm;
            prm.gB = p->dB[f]; prm.gX = p->dX; prm.gMixw = nullptr; prm.raw = nullptr; prm.part = p->dPart;
            prm.T = cn; prm.T_pad = T_pad; prm.n_sen = 0; prm.n_tiles_m = n_tiles_m; prm.n_tiles_n = p->n_tiles_n;
            prm.ksteps = p->ksteps[f]; prm.aw = 1; prm.dbg = 0; prm.m31 = 63; prm.scaleA = nullptr; prm.fmt = nullptr;
            int m_chunks = 1;
            while ((long long)p->n_tiles_n * m_chunks < 16LL * p->n_sm && (n_tiles_m + m_chunks) / (m_chunks + 1) >= 8) ++m_chunks;
            prm.tiles_per_chunk = (n_tiles_m + m_chunks - 1) / m_chunks;
            prm.m_chunks = (n_tiles_m + prm.tiles_per_chunk - 1) / prm.tiles_per_chunk;
            prm.n_units = p->n_tiles_n * prm.m_chunks;
            memset(prm.logadd, 0, 256);
            const int grid = std::min(prm.n_units, p->n_sm);
            if (p->half[f].ksteps) {
                TcParams ph = prm;
                ph.gB = reinterpret_cast<const float *>(p->half[f].dB); ph.gX = gx_half; ph.ksteps = p->half[f].ksteps;
                ph.scaleA = p->half[f].dScale; ph.fmt = p->half[f].dFmt;
                if ((rc = launch_score_ks_half<32, 1>(ph, ph.ksteps, grid, st))) return rc;
                prm.fmt = p->half[f].dFmt;
            }
            rc = launch_score_ks<32, 1>(prm, p->ksteps[f], grid, st);
            if (rc) return rc;
            // flagged pairs of this launch: reset the per-launch counter, keep a running total in dCount[1]
            switch (g.topn) {
                case 1: rc = tied_select_launch<1>(p, g, f, d_feat, t0 + c0, cn, T_pad, lc, st); break;
                case 2: rc = tied_select_launch<2>(p, g, f, d_feat, t0 + c0, cn, T_pad, lc, st); break;
                case 3: rc = tied_select_launch<3>(p, g, f, d_feat, t0 + c0, cn, T_pad, lc, st); break;
                default: rc = tied_select_launch<4>(p, g, f, d_feat, t0 + c0, cn, T_pad, lc, st); break;
            }
            if (rc) return rc;
            tied_count_roll<<<1, 1, 0, st>>>(p->dCount, cn * g.n_mgau);
            B200_LAUNCH_CHECK();
            p->pairs += (long long)cn * g.n_mgau;
        }
    }
    return B200_OK;
}

}  // namespace b200
