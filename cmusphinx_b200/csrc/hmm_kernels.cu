// hmm_kernels.cu -- batched HMM Viterbi step, beam test + order-preserving
// compaction, active-senone gather (sm_100a).
//
// Reference: pocketsphinx/src/libpocketsphinx/hmm.c
//   hmm_vit_eval_5st_lr      :224-352      hmm_vit_eval_5st_lr_mpx :357-527
//   hmm_vit_eval_3st_lr      :531-609      hmm_vit_eval_3st_lr_mpx :611-709
// and the beam test of ngram_search_fwdtree.c:714-869, acmod_activate_hmm
// (acmod.c:1178-1217).  All arithmetic is int32 max-plus and is reproduced
// literally, including the quirks listed in SURVEY.md section 8(a):
// the stale exit score when s1 is dead (3st) / s3 dead (5st), the t2 leak from
// the exit step into state 2 when tp(0,2) is "zero", the nested tie-breaking
// order, and the BAD_SSID / "!= WORST_SCORE" guards of the mpx variants.
//
// Layout: structure of arrays, state-major ([state][hmm]) so that a warp's
// loads and stores of one field are contiguous 128-byte lines.  The frame's
// senone scores (int16, <= 64 K entries) and the transition table are staged
// in shared memory; the mpx senone-sequence table is read through L1/L2.
#include "hmm_dev.cuh"

namespace b200 {

#define BT(a, b) ((a) > (b))   // BETTER_THAN (hmm.h:85)
#define WT(a, b) ((a) < (b))   // WORSE_THAN

struct HmmRegs {
    int32_t sc[5], hi[5], out_sc, out_hi, best;
    uint16_t sid[5];
};

// tp row-major [from][to] with n+1 columns, stored negated (uint8).
template <int NE>
__device__ __forceinline__ int32_t tpv(const uint8_t *tp, int i, int j) { return -(int32_t)tp[i * (NE + 1) + j]; }
// The same matrix held in registers: hmm_run_kernel reads an HMM's whole matrix from shared memory with one or
// two 16-byte loads (rows padded to kTpStride) instead of a byte load per use -- with a different matrix per lane
// every byte load was a bank-conflicted wavefront of its own, ~18 per HMM and warp against 4 now, and the LSU
// data pipe is what bounds phase A.
template <int NE> struct TpRow { uint32_t w[(NE * (NE + 1) + 15) / 16 * 4]; };
template <int NE> constexpr int kTpStride = (NE * (NE + 1) + 15) / 16 * 16;
template <int NE>
__device__ __forceinline__ int32_t tpv(const TpRow<NE> &tp, int i, int j) {
    const int k = i * (NE + 1) + j;
    return -(int32_t)((tp.w[k >> 2] >> ((k & 3) * 8)) & 0xffu);
}
template <int NE>
__device__ __forceinline__ TpRow<NE> tp_row(const uint8_t *s_tp, int tm) {
    TpRow<NE> t;
    const uint4 *q = reinterpret_cast<const uint4 *>(s_tp + tm * kTpStride<NE>);
#pragma unroll
    for (int k = 0; k < kTpStride<NE> / 16; ++k) { const uint4 v = q[k]; t.w[4 * k] = v.x; t.w[4 * k + 1] = v.y; t.w[4 * k + 2] = v.z; t.w[4 * k + 3] = v.w; }
    return t;
}

template <class TP>
__device__ __forceinline__ void eval3(HmmRegs &h, const TP &tp, const int16_t *sen) {
    int32_t s3, s2, s1, s0, t2, t1, t0, best;
    s2 = h.sc[2] - sen[h.sid[2]];
    s1 = h.sc[1] - sen[h.sid[1]];
    s0 = h.sc[0] - sen[h.sid[0]];
    best = kWorstScore;
    t2 = (int32_t)0x80000000;
    if (BT(s1, kWorstScore)) {
        t1 = s2 + tpv<3>(tp, 2, 3);
        if (BT(tpv<3>(tp, 1, 3), B200_TMAT_WORST)) t2 = s1 + tpv<3>(tp, 1, 3);
        if (BT(t1, t2)) { s3 = t1; h.out_hi = h.hi[2]; }
        else { s3 = t2; h.out_hi = h.hi[1]; }
        if (WT(s3, kWorstScore)) s3 = kWorstScore;
        h.out_sc = s3;
        best = s3;
    }
    t0 = s2 + tpv<3>(tp, 2, 2);
    t1 = s1 + tpv<3>(tp, 1, 2);
    if (BT(tpv<3>(tp, 0, 2), B200_TMAT_WORST)) t2 = s0 + tpv<3>(tp, 0, 2);
    if (BT(t0, t1)) {
        if (BT(t2, t0)) { s2 = t2; h.hi[2] = h.hi[0]; } else s2 = t0;
    } else {
        if (BT(t2, t1)) { s2 = t2; h.hi[2] = h.hi[0]; } else { s2 = t1; h.hi[2] = h.hi[1]; }
    }
    if (WT(s2, kWorstScore)) s2 = kWorstScore;
    if (BT(s2, best)) best = s2;
    h.sc[2] = s2;
    t0 = s1 + tpv<3>(tp, 1, 1);
    t1 = s0 + tpv<3>(tp, 0, 1);
    if (BT(t0, t1)) s1 = t0; else { s1 = t1; h.hi[1] = h.hi[0]; }
    if (WT(s1, kWorstScore)) s1 = kWorstScore;
    if (BT(s1, best)) best = s1;
    h.sc[1] = s1;
    s0 = s0 + tpv<3>(tp, 0, 0);
    if (WT(s0, kWorstScore)) s0 = kWorstScore;
    if (BT(s0, best)) best = s0;
    h.sc[0] = s0;
    h.best = best;
}

#define MPX_SEN(st) sen[sseq[(size_t)h.sid[st] * NE + (st)]]

template <class TP>
__device__ __forceinline__ void eval3_mpx(HmmRegs &h, const TP &tp, const int16_t *sen,
                                          const uint16_t *__restrict__ sseq) {
    constexpr int NE = 3;
    int32_t s3, s2, s1, s0, t2, t1, t0, best;
    t2 = (int32_t)0x80000000;
    if (h.sid[2] == B200_BAD_SSID) s2 = t1 = kWorstScore;
    else { s2 = h.sc[2] - MPX_SEN(2); t1 = s2 + tpv<3>(tp, 2, 3); }
    if (h.sid[1] == B200_BAD_SSID) s1 = t2 = kWorstScore;
    else {
        s1 = h.sc[1] - MPX_SEN(1);
        if (BT(tpv<3>(tp, 1, 3), B200_TMAT_WORST)) t2 = s1 + tpv<3>(tp, 1, 3);
    }
    if (BT(t1, t2)) { s3 = t1; h.out_hi = h.hi[2]; }
    else { s3 = t2; h.out_hi = h.hi[1]; }
    if (WT(s3, kWorstScore)) s3 = kWorstScore;
    h.out_sc = s3;
    best = s3;
    s0 = h.sc[0] - MPX_SEN(0);
    t0 = t1 = kWorstScore;
    if (s2 != kWorstScore) t0 = s2 + tpv<3>(tp, 2, 2);
    if (s1 != kWorstScore) t1 = s1 + tpv<3>(tp, 1, 2);
    if (BT(tpv<3>(tp, 0, 2), B200_TMAT_WORST)) t2 = s0 + tpv<3>(tp, 0, 2);
    if (BT(t0, t1)) {
        if (BT(t2, t0)) { s2 = t2; h.hi[2] = h.hi[0]; h.sid[2] = h.sid[0]; } else s2 = t0;
    } else {
        if (BT(t2, t1)) { s2 = t2; h.hi[2] = h.hi[0]; h.sid[2] = h.sid[0]; }
        else { s2 = t1; h.hi[2] = h.hi[1]; h.sid[2] = h.sid[1]; }
    }
    if (WT(s2, kWorstScore)) s2 = kWorstScore;
    if (BT(s2, best)) best = s2;
    h.sc[2] = s2;
    t0 = kWorstScore;
    if (s1 != kWorstScore) t0 = s1 + tpv<3>(tp, 1, 1);
    t1 = s0 + tpv<3>(tp, 0, 1);
    if (BT(t0, t1)) s1 = t0; else { s1 = t1; h.hi[1] = h.hi[0]; h.sid[1] = h.sid[0]; }
    if (WT(s1, kWorstScore)) s1 = kWorstScore;
    if (BT(s1, best)) best = s1;
    h.sc[1] = s1;
    s0 += tpv<3>(tp, 0, 0);
    if (WT(s0, kWorstScore)) s0 = kWorstScore;
    if (BT(s0, best)) best = s0;
    h.sc[0] = s0;
    h.best = best;
}

// One "all transitions into state `to`" block of the 5-state non-mpx eval.
#define INTO3(S_TO, S_A, S_B, S_C, TO, A, B, C)                                      \
    t0 = S_A + tpv<5>(tp, A, TO); t1 = S_B + tpv<5>(tp, B, TO); t2 = S_C + tpv<5>(tp, C, TO); \
    if (BT(t0, t1)) { if (BT(t2, t0)) { S_TO = t2; h.hi[TO] = h.hi[C]; } else S_TO = t0; } \
    else { if (BT(t2, t1)) { S_TO = t2; h.hi[TO] = h.hi[C]; } else { S_TO = t1; h.hi[TO] = h.hi[B]; } } \
    if (WT(S_TO, kWorstScore)) S_TO = kWorstScore;                                   \
    if (BT(S_TO, best)) best = S_TO;                                                 \
    h.sc[TO] = S_TO;

template <class TP>
__device__ __forceinline__ void eval5(HmmRegs &h, const TP &tp, const int16_t *sen) {
    int32_t s5, s4, s3, s2, s1, s0, t2, t1, t0, best;
    best = kWorstScore;
    s4 = h.sc[4] - sen[h.sid[4]];
    s3 = h.sc[3] - sen[h.sid[3]];
    if (BT(s3, kWorstScore)) {
        t1 = s4 + tpv<5>(tp, 4, 5);
        t2 = s3 + tpv<5>(tp, 3, 5);
        if (BT(t1, t2)) { s5 = t1; h.out_hi = h.hi[4]; }
        else { s5 = t2; h.out_hi = h.hi[3]; }
        if (WT(s5, kWorstScore)) s5 = kWorstScore;
        h.out_sc = s5;
        best = s5;
    }
    s2 = h.sc[2] - sen[h.sid[2]];
    if (BT(s2, kWorstScore)) { INTO3(s4, s4, s3, s2, 4, 4, 3, 2) }
    s1 = h.sc[1] - sen[h.sid[1]];
    if (BT(s1, kWorstScore)) { INTO3(s3, s3, s2, s1, 3, 3, 2, 1) }
    s0 = h.sc[0] - sen[h.sid[0]];
    { INTO3(s2, s2, s1, s0, 2, 2, 1, 0) }
    t0 = s1 + tpv<5>(tp, 1, 1);
    t1 = s0 + tpv<5>(tp, 0, 1);
    if (BT(t0, t1)) s1 = t0; else { s1 = t1; h.hi[1] = h.hi[0]; }
    if (WT(s1, kWorstScore)) s1 = kWorstScore;
    if (BT(s1, best)) best = s1;
    h.sc[1] = s1;
    s0 = s0 + tpv<5>(tp, 0, 0);
    if (WT(s0, kWorstScore)) s0 = kWorstScore;
    if (BT(s0, best)) best = s0;
    h.sc[0] = s0;
    h.best = best;
}

// mpx flavour of the same block: guarded self/prev terms, ssid propagation.
#define INTO3_MPX(S_TO, S_A, S_B, TO, A, B, C)                                        \
    t0 = t1 = kWorstScore;                                                            \
    if (S_A != kWorstScore) t0 = S_A + tpv<5>(tp, A, TO);                             \
    if (S_B != kWorstScore) t1 = S_B + tpv<5>(tp, B, TO);                             \
    if (BT(t0, t1)) {                                                                 \
        if (BT(t2, t0)) { S_TO = t2; h.hi[TO] = h.hi[C]; h.sid[TO] = h.sid[C]; } else S_TO = t0; \
    } else {                                                                          \
        if (BT(t2, t1)) { S_TO = t2; h.hi[TO] = h.hi[C]; h.sid[TO] = h.sid[C]; }      \
        else { S_TO = t1; h.hi[TO] = h.hi[B]; h.sid[TO] = h.sid[B]; }                 \
    }                                                                                 \
    if (WT(S_TO, kWorstScore)) S_TO = kWorstScore;                                    \
    if (BT(S_TO, best)) best = S_TO;                                                  \
    h.sc[TO] = S_TO;

template <class TP>
__device__ __forceinline__ void eval5_mpx(HmmRegs &h, const TP &tp, const int16_t *sen,
                                          const uint16_t *__restrict__ sseq) {
    constexpr int NE = 5;
    int32_t s5, s4, s3, s2, s1, s0, t2, t1, t0, best;
    if (h.sid[4] == B200_BAD_SSID) s4 = t1 = kWorstScore;
    else { s4 = h.sc[4] - MPX_SEN(4); t1 = s4 + tpv<5>(tp, 4, 5); }
    if (h.sid[3] == B200_BAD_SSID) s3 = t2 = kWorstScore;
    else { s3 = h.sc[3] - MPX_SEN(3); t2 = s3 + tpv<5>(tp, 3, 5); }
    if (BT(t1, t2)) { s5 = t1; h.out_hi = h.hi[4]; }
    else { s5 = t2; h.out_hi = h.hi[3]; }
    if (WT(s5, kWorstScore)) s5 = kWorstScore;
    h.out_sc = s5;
    best = s5;
    if (h.sid[2] == B200_BAD_SSID) s2 = t2 = kWorstScore;
    else { s2 = h.sc[2] - MPX_SEN(2); t2 = s2 + tpv<5>(tp, 2, 4); }
    { INTO3_MPX(s4, s4, s3, 4, 4, 3, 2) }
    if (h.sid[1] == B200_BAD_SSID) s1 = t2 = kWorstScore;
    else { s1 = h.sc[1] - MPX_SEN(1); t2 = s1 + tpv<5>(tp, 1, 3); }
    { INTO3_MPX(s3, s3, s2, 3, 3, 2, 1) }
    s0 = h.sc[0] - MPX_SEN(0);
    t2 = s0 + tpv<5>(tp, 0, 2);
    {
        // state 2: same shape, but t2 is computed before the guards
        int32_t t2keep = t2;
        t0 = t1 = kWorstScore;
        if (s2 != kWorstScore) t0 = s2 + tpv<5>(tp, 2, 2);
        if (s1 != kWorstScore) t1 = s1 + tpv<5>(tp, 1, 2);
        t2 = t2keep;
        if (BT(t0, t1)) {
            if (BT(t2, t0)) { s2 = t2; h.hi[2] = h.hi[0]; h.sid[2] = h.sid[0]; } else s2 = t0;
        } else {
            if (BT(t2, t1)) { s2 = t2; h.hi[2] = h.hi[0]; h.sid[2] = h.sid[0]; }
            else { s2 = t1; h.hi[2] = h.hi[1]; h.sid[2] = h.sid[1]; }
        }
        if (WT(s2, kWorstScore)) s2 = kWorstScore;
        if (BT(s2, best)) best = s2;
        h.sc[2] = s2;
    }
    t0 = kWorstScore;
    if (s1 != kWorstScore) t0 = s1 + tpv<5>(tp, 1, 1);
    t1 = s0 + tpv<5>(tp, 0, 1);
    if (BT(t0, t1)) s1 = t0; else { s1 = t1; h.hi[1] = h.hi[0]; h.sid[1] = h.sid[0]; }
    if (WT(s1, kWorstScore)) s1 = kWorstScore;
    if (BT(s1, best)) best = s1;
    h.sc[1] = s1;
    s0 += tpv<5>(tp, 0, 0);
    if (WT(s0, kWorstScore)) s0 = kWorstScore;
    if (BT(s0, best)) best = s0;
    h.sc[0] = s0;
    h.best = best;
}

// ---------------------------------------------------------------- kernels
// All kernels take a 2-D grid: blockIdx.y = utterance.  An utterance owns the
// contiguous HMM range [utt_off[u], utt_off[u+1]) of the population, its own
// row of senone scores, its own frame-best / beam threshold and its own
// active-senone mask; the compacted survivor list is global and ordered by
// (utterance, HMM index).  n_utt == 1 is the reference's per-decoder case.
constexpr int kHmmBlock = 256;

// hmm_vit_eval_anytopo (hmm.c:711-786): what hmm_vit_eval dispatches to when
// n_emit_state is neither 3 nor 5 (1, 2 or 4 -- HMM_MAX_NSTATE is 5).  Generic
// upper-triangular topology: every transition that is not "zero" competes, ties
// keep the earlier candidate (self loop, then from = to-1, to-2, ...).  As in the
// reference, state 0's sum is not clamped, new scores are stored unclamped, and
// a missing senone scores WORST_SCORE (hmm_senscr, hmm.h:198-200).
template <int NE, class TP>
__device__ __forceinline__ void eval_any(HmmRegs &h, const TP &tp, const int16_t *sen, const uint16_t *sseq,
                                         bool mpx) {
    int32_t st[NE];
#pragma unroll
    for (int s = 0; s < NE; ++s) {
        uint32_t id = h.sid[s];
        if (mpx && id != B200_BAD_SSID) id = sseq[(size_t)id * NE + s];
        const int32_t ss = id == 0xffffu ? kWorstScore : -(int32_t)sen[id];
        int32_t v = h.sc[s] + ss;
        if (s > 0 && WT(v, kWorstScore)) v = kWorstScore;
        st[s] = v;
    }
    int32_t scr = kWorstScore, bh = 0;
    uint16_t bsid = 0;
    bool found = false;
#pragma unroll
    for (int f = NE - 1; f >= 0; --f) {
        const int32_t t = tpv<NE>(tp, f, NE);
        if (BT(t, B200_TMAT_WORST) && BT(st[f] + t, scr)) { scr = st[f] + t; bh = h.hi[f]; found = true; }
    }
    h.out_sc = scr;
    if (found) h.out_hi = bh;
    int32_t best = scr;
#pragma unroll
    for (int to = NE - 1; to >= 0; --to) {
        const int32_t tt = tpv<NE>(tp, to, to);
        scr = BT(tt, B200_TMAT_WORST) ? st[to] + tt : kWorstScore;
        found = false;
#pragma unroll
        for (int f = to - 1; f >= 0; --f) {
            const int32_t t = tpv<NE>(tp, f, to);
            if (BT(t, B200_TMAT_WORST) && BT(st[f] + t, scr)) { scr = st[f] + t; bh = h.hi[f]; bsid = h.sid[f]; found = true; }
        }
        h.sc[to] = scr;
        if (found) { h.hi[to] = bh; if (mpx) h.sid[to] = bsid; }
        if (WT(best, scr)) best = scr;
    }
    h.best = best;
}

// ------------------------------------------------------------------ the step
// One persistent cooperative kernel runs a whole RUN of search frames (round 2;
// it replaces five launches per frame: init, step, beam flag, scan, scatter).
// The grid is exactly one resident wave: gx CTAs stride the tiles (256 HMMs) of
// an utterance, gy CTA rows stride the utterances.  Per frame:
//   A  hmm_vit_eval of every HMM (state in registers, senone scores and the
//      transition table in shared memory), bestscore stored, the utterance's
//      frame best by atomicMax                              -- grid barrier --
//   B  beam test against best + beam, survivors counted per tile and per
//      utterance                                            -- grid barrier --
//   C  order-preserving scatter of the survivors' indices (offsets from the
//      tile counts: no keep-byte array, no separate scan kernel) and their
//      senones into the utterance's active mask (acmod_activate_hmm)
// Frame records rotate over three slots and the masks over two, so phase A of
// the next frame needs no third barrier.
__device__ __forceinline__ void grid_barrier(unsigned *ctr, unsigned n_cta, unsigned &epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        // release / acquire at gpu scope instead of two sequentially consistent fences around a relaxed atomic
        ++epoch;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        const unsigned target = epoch * n_cta;
        const long long t0 = clock64();
        unsigned v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= target) break;
            if (clock64() - t0 > 4000000000LL) __trap();        // a protocol bug must trap, not hang the box
        }
    }
    __syncthreads();
}

// all threads of all CTAs of the cluster; release / acquire at cluster scope orders the global-memory
// hand-overs between the phases (per-utterance best, tile counts, partial masks)
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// `ctr` / `n_cta`: the whole grid's counter, or -- in the row form -- the counter of this grid row and gridDim.x
template <bool CL>
__device__ __forceinline__ void sync_ctas(unsigned *ctr, unsigned n_cta, unsigned &epoch) {
    if (CL) cluster_barrier();
    else grid_barrier(ctr, n_cta, epoch);
}

// sum over the block of (a, b); every thread gets both totals
template <int BLK>
__device__ __forceinline__ int2 block_sum2(int a, int b, int32_t *s_red /* [2 * warps] */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    __syncthreads();
    if (lane == 0) { s_red[2 * w] = a; s_red[2 * w + 1] = b; }
    __syncthreads();
    int ta = 0, tb = 0;
    for (int k = 0; k < BLK / 32; ++k) { ta += s_red[2 * k]; tb += s_red[2 * k + 1]; }
    return make_int2(ta, tb);
}

// The active-senone mask of one utterance and frame: the OR of the partial masks its gx CTAs
// stored with plain stores (merging them with atomicOr made every CTA of an utterance hammer
// the same n_words addresses: 17 us per frame on 196 CTAs).  One thread per mask word.
struct MergePlan { int L, per_pass, sub, slot, n_mine; };
template <int BLK>
__device__ __forceinline__ MergePlan merge_plan(int gx, int n_words, int w0, int wstep) {
    // L lanes share a word (each ORs every L-th partial), BLK / L words per pass; this CTA owns the words
    // w0, w0 + wstep, ...; L shrinks until one pass covers them (then every load of the merge is issued at once)
    MergePlan m;
    m.n_mine = w0 < n_words ? (n_words - w0 + wstep - 1) / wstep : 0;
    int L = 1;
    while (L < 32 && L < gx) L <<= 1;
    while (L > 1 && BLK / L < m.n_mine) L >>= 1;
    m.L = L; m.per_pass = BLK / L; m.sub = threadIdx.x % L; m.slot = threadIdx.x / L;
    return m;
}
// first trip of pass j0: eight independent loads per lane (a plain loop waits for every partial in turn)
__device__ __forceinline__ void merge_load(const MergePlan &m, const uint32_t *part_u, int gx, int n_words, int w0, int wstep,
                                           int j0, int b0, uint32_t (&t)[8]) {
    const int j = j0 + m.slot, kk = w0 + j * wstep;
#pragma unroll
    for (int q = 0; q < 8; ++q) { const int b = b0 + q * m.L; t[q] = (j < m.n_mine && b < gx) ? __ldcg(part_u + (size_t)b * n_words + kk) : 0u; }
}
// the rest of the merge; t holds the loads of (pass 0, first trip) issued earlier by merge_load
template <int BLK>
__device__ __forceinline__ void merge_finish(const MergePlan &m, const uint32_t *part_u, int gx, int n_words, uint32_t *mask_u,
                                             int w0, int wstep, uint32_t (&t)[8]) {
    for (int j0 = 0; j0 < m.n_mine; j0 += m.per_pass) {
        const int j = j0 + m.slot, kk = w0 + j * wstep;
        uint32_t v = 0;
        for (int b0 = m.sub; b0 < gx; b0 += 8 * m.L) {
            if (j0 != 0 || b0 != m.sub) merge_load(m, part_u, gx, n_words, w0, wstep, j0, b0, t);
#pragma unroll
            for (int q = 0; q < 8; ++q) v |= t[q];
        }
        for (int o = m.L >> 1; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
        if (j < m.n_mine && m.sub == 0) mask_u[kk] = v;
    }
}
template <int BLK>
__device__ __forceinline__ void merge_mask(const uint32_t *part_u /* [gx][n_words] */, int gx, int n_words, uint32_t *mask_u,
                                           int w0, int wstep) {
    const MergePlan m = merge_plan<BLK>(gx, n_words, w0, wstep);
    uint32_t t[8];
    merge_load(m, part_u, gx, n_words, w0, wstep, 0, m.sub, t);
    merge_finish<BLK>(m, part_u, gx, n_words, mask_u, w0, wstep, t);
}

// BLK threads per CTA.  CL = false: cooperative launch, one resident wave of 256-thread CTAs, grid
// barriers, every frame over all utterances.  CL = true: every grid row is a thread-block CLUSTER
// (gridDim.x CTAs of 1024 threads) that takes ONE utterance at a time through ALL frames of the
// run with cluster barriers only -- utterances are independent, so no grid-wide barrier is left,
// and the utterance's state (76 B x 50 000 HMMs = 3.8 MB) stays in L2 from frame to frame instead
// of streaming from HBM.  Only the last frame's survivor list has to be ordered across
// utterances: the cluster form writes per-utterance lists (r.keep_tmp) and
// hmm_compact_last_kernel packs them.
#ifndef B200_RUN_CTAS
#define B200_RUN_CTAS 4
#endif
constexpr int kRunCtasPerSm = B200_RUN_CTAS;
constexpr int kPre = 2;             // HMMs per thread whose state loads phase A keeps in flight
constexpr int kBeamBatch = 8;       // tiles whose bestscore loads phase B issues before its first vote
template <int NE, int BLK, bool CL>
__global__ void __launch_bounds__(BLK, BLK == 256 ? kRunCtasPerSm : 1)
hmm_run_kernel(HmmDev c, HmmPop p, HmmRun r) {
    extern __shared__ uint8_t sm_raw[];
    int16_t *s_sen = reinterpret_cast<int16_t *>(sm_raw);
    const size_t sen_bytes = ((size_t)c.n_sen * 2 + 15) & ~(size_t)15;
    const size_t tp_bytes = (size_t)c.n_tmat * kTpStride<NE>;      // one padded row per matrix
    uint8_t *s_tp = sm_raw + sen_bytes;
    uint32_t *s_flag_w = reinterpret_cast<uint32_t *>(sm_raw + sen_bytes + tp_bytes);   // one flag byte per senone
    uint8_t *s_flag = reinterpret_cast<uint8_t *>(s_flag_w);
    // tile counts of one utterance | offsets of my tiles | keep ballots of my tiles (all my utterances)
    const int rows = (r.tpu + (int)gridDim.x - 1) / (int)gridDim.x + kBeamBatch - 1;   // my tiles per utterance (+ the batch overrun)
    int32_t *s_tc = reinterpret_cast<int32_t *>(s_flag + (size_t)((c.n_sen + 31) / 32) * 32);
    int32_t *s_off = s_tc + r.tpu;
    uint32_t *s_bal = reinterpret_cast<uint32_t *>(s_off + rows);
    constexpr int BS = BLK / 32 + BLK / 64;     // words per tile row: the warps' keep ballots, then their uint16 survivor prefixes
    int32_t *s_pre = reinterpret_cast<int32_t *>(sm_raw + r.pre_off);           // [kPre][2 NE + 2][BLK] phase A's prefetch slots
    __shared__ int32_t s_red[2 * (BLK / 32)];
    __shared__ int32_t s_wcnt[BLK / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int gx = gridDim.x, gy = gridDim.y, bx = blockIdx.x, by = blockIdx.y;
    // Row form (r.row_sync, cooperative launch): utterances are independent, so only the gx CTAs of a grid ROW -- the
    // ones that share utterances -- meet at a barrier; the rows drift apart and one row's latency-bound phases B / C
    // overlap the others' streaming phase A.  Like the cluster form it leaves per-utterance survivor lists
    // (hmm_compact_last_kernel packs the last frame's).
    const bool row = !CL && r.row_sync != 0;
    const bool own_list = CL || row;
    // Row form over a run of frames: ONE barrier per frame.  The frame's best needs one (every HMM evaluated before the beam
    // test) and the survivor list needs the tile counts of the whole utterance -- but nothing in frame f + 1 waits for the list
    // of frame f, and phase C reads shared memory and the tile counts only: it runs behind the barrier of frame f + 1, after
    // that frame's phase B, on the ballots / tile counts of the other parity.
    const bool defer = row && r.n_frames >= 2;
    const unsigned n_cta = row ? (unsigned)gx : (unsigned)gx * gy;
    unsigned *const bar = row ? r.bar + (size_t)by * 32 : r.bar;
    const int n = p.n_hmm, n_utt = p.n_utt;
    const int n_words = (c.n_sen + 31) / 32;
    unsigned epoch = 0;

    {   // transition table once; frame records and the first mask
        const int ntp = c.n_tmat * NE * (NE + 1);
        for (int i = tid; i < ntp; i += BLK) s_tp[(i / (NE * (NE + 1))) * kTpStride<NE> + i % (NE * (NE + 1))] = c.tp[i];
        if (bx == 0)
            for (int u = by; u < n_utt; u += gy) {
                if (tid < 3) { HmmFrame f; f.best = kWorstScore; f.n_keep = 0; f.thresh = kWorstScore; f.pad = 0; r.fr3[(size_t)tid * n_utt + u] = f; }
            }
    }
    sync_ctas<CL>(bar, n_cta, epoch);

    for (int u0 = by; u0 < (CL ? n_utt : by + 1); u0 += gy) {       // CL: one utterance at a time, all its frames
    const int u_lo = CL ? u0 : by, u_hi = CL ? u0 + 1 : n_utt;
    for (int f = 0; f < r.n_frames + (defer ? 1 : 0); ++f) {    // (deferred: one more turn for the last frame's phase C)
        const bool live = f < r.n_frames;
        const bool probe = r.probe && f == r.n_frames - 1 && bx == 0 && by == 0 && tid == 0;
        if (probe) r.probe[0] = clock64();
        HmmFrame *fr = r.fr3 + (size_t)((r.slot0 + f) % 3) * n_utt;
        const int16_t *sen_frame = r.sen_base + (size_t)((r.frame0 + f) % r.n_cycle) * r.frame_stride;
        // ------------------------------------------------ A: hmm_vit_eval
        for (int u = u_lo; live && u < u_hi; u += gy) {
            const int lo = p.utt_off[u], hi = p.utt_off[u + 1];
            if (lo + bx * BLK >= hi) continue;            // uniform per block
            const int16_t *senscr = sen_frame + (size_t)u * c.n_sen;
            __syncthreads();                                    // the previous row's readers are done
            {
                const int n16 = ((reinterpret_cast<size_t>(senscr) & 15) == 0) ? (c.n_sen * 2) / 16 : 0;
                const int4 *src = reinterpret_cast<const int4 *>(senscr);
                int4 *dst = reinterpret_cast<int4 *>(s_sen);
                for (int i = tid; i < n16; i += BLK) dst[i] = src[i];
                for (int i = n16 * 8 + tid; i < c.n_sen; i += BLK) s_sen[i] = senscr[i];
            }
            __syncthreads();
            int32_t blockbest = kWorstScore;
            // Software pipeline through shared memory: a thread's 2 NE + 2 int32 values of its next kPre HMMs are in flight
            // as 4-byte cp.async copies into slots only this thread touches (no barrier: cp.async.wait_group orders a thread's
            // own copies), the five small fields of the next HMM as plain loads.  Against holding the whole next HMM in
            // registers (13 loads, one deep) this keeps 2.5x the bytes in flight per thread with 6 fewer registers.
            constexpr int NA = 2 * NE + 2;
            uint32_t pre_base = (uint32_t)__cvta_generic_to_shared(s_pre) + tid * 4;
            auto pre_issue = [&](int i, int slot) {
                if (i < hi) {
                    const uint32_t d = pre_base + slot * (NA * BLK * 4);
#pragma unroll
                    for (int s = 0; s < NE; ++s) {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + s * BLK * 4), "l"(p.score + (size_t)s * n + i) : "memory");
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + (NE + s) * BLK * 4), "l"(p.history + (size_t)s * n + i) : "memory");
                    }
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 2 * NE * BLK * 4), "l"(p.out_score + i) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + (2 * NE + 1) * BLK * 4), "l"(p.out_history + i) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            auto load_small = [&](int i, HmmRegs &h, int &tm, bool &mpx) {
#pragma unroll
                for (int s = 0; s < NE; ++s) h.sid[s] = p.senid[(size_t)s * n + i];
                tm = (int)p.tmatid[i];
                mpx = p.mpx[i] != 0;
            };
            const int step = gx * BLK;
            // Pair form (NE <= 3, 256-thread CTAs, even n_hmm and utterance start): two of the CTA's FULL tiles per trip, half
            // of the threads each, a thread taking two ADJACENT HMMs with 8-byte copies, loads and stores -- the same 13 + 9
            // memory instructions (and their address arithmetic) now serve two HMMs, and the two evaluations are independent
            // chains for the scheduler.  Left-over and ragged tiles go through the single loop below.
            int k_done = 0;                                     // my tiles already evaluated
            if constexpr (NE <= 3 && BLK == 256 && !CL) {
                const int n_tiles_a = (hi - lo + BLK - 1) / BLK;
                const int k_mine = (n_tiles_a - bx + gx - 1) / gx;
                int k_full = k_mine;                            // my tiles that are complete
                if (k_mine > 0 && bx + (k_mine - 1) * gx == n_tiles_a - 1 && ((hi - lo) % BLK) != 0) --k_full;
                const int n_trips = (r.pair && ((n | lo) & 1) == 0) ? k_full / 2 : 0;
                if (n_trips > 0) {
                    constexpr int H = BLK / 2;
                    const int q = tid / H, pi = tid % H;
                    const uint32_t pbase = (uint32_t)__cvta_generic_to_shared(s_pre) + tid * 8;
                    auto i_of = [&](int trip) { return lo + (bx + (2 * trip + q) * gx) * BLK + 2 * pi; };
                    auto issue2 = [&](int trip) {                  // ONE stage: the copies of trip t + 1 fly during trip t's evaluation
                        if (trip < n_trips) {
                            const int i2 = i_of(trip);
                            const uint32_t d = pbase;
#pragma unroll
                            for (int s = 0; s < NE; ++s) {
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d + s * BLK * 8), "l"(p.score + (size_t)s * n + i2) : "memory");
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d + (NE + s) * BLK * 8), "l"(p.history + (size_t)s * n + i2) : "memory");
                            }
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d + 2 * NE * BLK * 8), "l"(p.out_score + i2) : "memory");
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d + (2 * NE + 1) * BLK * 8), "l"(p.out_history + i2) : "memory");
                        }
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    };
                    uint32_t sidw[NE], tmw; uint32_t mpw;       // the small fields of both HMMs, packed as loaded
                    auto small2 = [&](int trip) {
                        const int i2 = i_of(trip);
#pragma unroll
                        for (int s = 0; s < NE; ++s) sidw[s] = *reinterpret_cast<const uint32_t *>(p.senid + (size_t)s * n + i2);
                        tmw = *reinterpret_cast<const uint32_t *>(p.tmatid + i2);
                        mpw = *reinterpret_cast<const uint16_t *>(p.mpx + i2);
                    };
                    issue2(0);
                    small2(0);
                    for (int trip = 0; trip < n_trips; ++trip) {
                        const int i2 = i_of(trip);
                        HmmRegs ha, hb;
#pragma unroll
                        for (int s = 0; s < NE; ++s) { ha.sid[s] = (uint16_t)(sidw[s] & 0xffffu); hb.sid[s] = (uint16_t)(sidw[s] >> 16); }
                        const int tma = (int)(int16_t)(tmw & 0xffffu), tmb = (int)(int16_t)(tmw >> 16);
                        const bool mpa = (mpw & 0xffu) != 0, mpb = (mpw >> 8) != 0;
                        if (trip + 1 < n_trips) small2(trip + 1);
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                        {
                            const int2 *qq = reinterpret_cast<const int2 *>(s_pre) + tid;
#pragma unroll
                            for (int s = 0; s < NE; ++s) {
                                const int2 a = qq[s * BLK], b = qq[(NE + s) * BLK];
                                ha.sc[s] = a.x; hb.sc[s] = a.y; ha.hi[s] = b.x; hb.hi[s] = b.y;
                            }
                            const int2 a = qq[2 * NE * BLK], b = qq[(2 * NE + 1) * BLK];
                            ha.out_sc = a.x; hb.out_sc = a.y; ha.out_hi = b.x; hb.out_hi = b.y;
                        }
                        issue2(trip + 1);                       // (the slot's values are in registers)
                        {
                            const TpRow<NE> tp = tp_row<NE>(s_tp, tma);
                            if constexpr (NE == 3) { if (mpa) eval3_mpx(ha, tp, s_sen, c.sseq); else eval3(ha, tp, s_sen); }
                            else eval_any<NE>(ha, tp, s_sen, c.sseq, mpa);
                        }
                        {
                            const TpRow<NE> tp = tp_row<NE>(s_tp, tmb);
                            if constexpr (NE == 3) { if (mpb) eval3_mpx(hb, tp, s_sen, c.sseq); else eval3(hb, tp, s_sen); }
                            else eval_any<NE>(hb, tp, s_sen, c.sseq, mpb);
                        }
#pragma unroll
                        for (int s = 0; s < NE; ++s) {
                            *reinterpret_cast<int2 *>(p.score + (size_t)s * n + i2) = make_int2(ha.sc[s], hb.sc[s]);
                            *reinterpret_cast<int2 *>(p.history + (size_t)s * n + i2) = make_int2(ha.hi[s], hb.hi[s]);
                        }
                        if (mpa | mpb) {                        // (a plain HMM's ids are unchanged: writing them back is harmless)
#pragma unroll
                            for (int s = 1; s < NE; ++s)
                                *reinterpret_cast<uint32_t *>(p.senid + (size_t)s * n + i2) = (uint32_t)ha.sid[s] | ((uint32_t)hb.sid[s] << 16);
                        }
                        *reinterpret_cast<int2 *>(p.out_score + i2) = make_int2(ha.out_sc, hb.out_sc);
                        *reinterpret_cast<int2 *>(p.out_history + i2) = make_int2(ha.out_hi, hb.out_hi);
                        *reinterpret_cast<int2 *>(p.bestscore + i2) = make_int2(ha.best, hb.best);
                        blockbest = max(blockbest, max(ha.best, hb.best));
                    }
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncthreads();                            // the single loop lays its 4-byte slots over other threads' 8-byte ones
                    k_done = 2 * n_trips;
                }
            }
            int i = lo + (bx + k_done * gx) * BLK + tid;
#pragma unroll
            for (int d = 0; d < kPre; ++d) pre_issue(i + d * step, d);
            HmmRegs h; int tm = 0; bool mpx = false;
            if (i < hi) load_small(i, h, tm, mpx);
            int slot = 0;
            while (i < hi) {
                HmmRegs hn; int tm_n = 0; bool mpx_n = false;
                if (i + step < hi) load_small(i + step, hn, tm_n, mpx_n);
                asm volatile("cp.async.wait_group %0;" ::"n"(kPre - 1) : "memory");
                {
                    const int32_t *q = s_pre + (size_t)slot * NA * BLK + tid;
#pragma unroll
                    for (int s = 0; s < NE; ++s) { h.sc[s] = q[s * BLK]; h.hi[s] = q[(NE + s) * BLK]; }
                    h.out_sc = q[2 * NE * BLK]; h.out_hi = q[(2 * NE + 1) * BLK];
                }
                pre_issue(i + kPre * step, slot);               // (the slot's values are in registers)
                slot = slot + 1 == kPre ? 0 : slot + 1;
                const TpRow<NE> tp = tp_row<NE>(s_tp, tm);
                if constexpr (NE == 3) { if (mpx) eval3_mpx(h, tp, s_sen, c.sseq); else eval3(h, tp, s_sen); }
                else if constexpr (NE == 5) { if (mpx) eval5_mpx(h, tp, s_sen, c.sseq); else eval5(h, tp, s_sen); }
                else eval_any<NE>(h, tp, s_sen, c.sseq, mpx);
#pragma unroll
                for (int s = 0; s < NE; ++s) {
                    p.score[(size_t)s * n + i] = h.sc[s];
                    p.history[(size_t)s * n + i] = h.hi[s];
                }
                if (mpx) {
#pragma unroll
                    for (int s = 1; s < NE; ++s) p.senid[(size_t)s * n + i] = h.sid[s];
                }
                p.out_score[i] = h.out_sc;
                p.out_history[i] = h.out_hi;
                p.bestscore[i] = h.best;
                blockbest = max(blockbest, h.best);
#pragma unroll
                for (int s = 0; s < NE; ++s) h.sid[s] = hn.sid[s];
                tm = tm_n; mpx = mpx_n; i += step;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            for (int o = 16; o > 0; o >>= 1) blockbest = max(blockbest, __shfl_xor_sync(0xffffffffu, blockbest, o));
            if (lane == 0) s_wcnt[w] = blockbest;
            __syncthreads();
            if (tid == 0) {
                int32_t b = s_wcnt[0];
                for (int k = 1; k < BLK / 32; ++k) b = max(b, s_wcnt[k]);
                atomicMax(&fr[u].best, b);
            }
        }
        if (!r.do_beam) continue;                               // (eval only: one frame per launch)
        if (probe) r.probe[1] = clock64();
        sync_ctas<CL>(bar, n_cta, epoch);
        if (probe) r.probe[2] = clock64();

        // ------------------------------------------------ B: beam test, counts
        // All loads of the CTA's tiles of an utterance are issued before the first vote; the keep
        // ballots stay in shared memory for phase C (no keep-byte array, no second read).
        for (int u = u_lo; live && u < u_hi; u += gy) {
            const int lo = p.utt_off[u], hi = p.utt_off[u + 1];
            // the previous frame's partial masks are complete (barrier 1): their loads are issued here and merged after the
            // beam test -- nothing in this frame waits for that mask
            const MergePlan mp_ = merge_plan<BLK>(gx, n_words, bx, gx);
            const uint32_t *mpart = r.mask_part + ((size_t)((r.mask0 + f - 1) & 1) * n_utt + u) * gx * n_words;
            uint32_t mt[8];
            if (f > 0) merge_load(mp_, mpart, gx, n_words, bx, gx, 0, mp_.sub, mt);
            const int32_t thresh = fr[u].best + r.beam;
            if (probe) r.probe[6] = clock64();
            const int n_tiles = (hi - lo + BLK - 1) / BLK;
            uint32_t *bal_u = s_bal + ((size_t)(defer ? (f & 1) * ((n_utt + gy - 1) / gy) : 0) + (CL ? 0 : (u - by) / gy)) * rows * BS;
            uint32_t *part_u = r.mask_part + (((size_t)((r.mask0 + f) & 1) * n_utt + u) * gx + bx) * n_words;
            for (int k = tid; k < n_words * 8; k += BLK) s_flag_w[k] = 0u;
            __syncthreads();                                    // (the flags are clear before the first survivor sets one)
            // Four of my tiles at a time: bestscore, mpx and the senone ids of every HMM are loaded together (unconditionally,
            // on a clamped index), then the beam vote, and the survivors flag their senones -- acmod_activate_hmm needs only
            // the keep bits, not the survivors' positions in the list, so it runs here and not behind the second barrier
            // with the scatter: its loads share the round trip of the beam test's.
            for (int row0 = 0; bx + row0 * gx < n_tiles; row0 += 4) {
                int32_t bs[4]; uint32_t sid[4][NE]; bool mp[4], act[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int t = bx + (row0 + j) * gx;
                    const int i = lo + t * BLK + tid;
                    const int ii = min(i, hi - 1);
                    bs[j] = (t < n_tiles && i < hi) ? p.bestscore[ii] : (int32_t)0x80000000;
                    mp[j] = p.mpx[ii] != 0;
#pragma unroll
                    for (int s = 0; s < NE; ++s) sid[j][s] = p.senid[(size_t)s * n + ii];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    act[j] = BT(bs[j], thresh);
                    const unsigned bal = __ballot_sync(0xffffffffu, act[j]);
                    if (lane == 0 && bx + (row0 + j) * gx < n_tiles) bal_u[(row0 + j) * BS + w] = bal;
                }
                if (__any_sync(0xffffffffu, mp[0] | mp[1] | mp[2] | mp[3])) {      // (most warps hold no multiplex HMM)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
#pragma unroll
                        for (int s = 0; s < NE; ++s)
                            if (mp[j]) sid[j][s] = (act[j] && sid[j][s] != B200_BAD_SSID) ? c.sseq[(size_t)sid[j][s] * NE + s] : 0xffffffffu;
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (!act[j]) continue;
#pragma unroll
                    for (int s = 0; s < NE; ++s)
                        if (sid[j][s] != 0xffffffffu) s_flag[sid[j][s]] = 1;
                }
            }
            __syncthreads();
            if (probe) r.probe[7] = clock64();
            int cta_cnt = 0;
            for (int rr = tid; bx + rr * gx < n_tiles; rr += BLK) {
                int cnt = 0;
                uint16_t *pre = reinterpret_cast<uint16_t *>(bal_u + rr * BS + BLK / 32);     // survivors in the warps before warp k
#pragma unroll
                for (int k = 0; k < BLK / 32; ++k) { pre[k] = (uint16_t)cnt; cnt += __popc(bal_u[rr * BS + k]); }
                r.tile_count[((size_t)(defer ? (f & 1) : 0) * n_utt + u) * r.tpu + bx + rr * gx] = cnt;
                cta_cnt += cnt;
            }
            cta_cnt = block_sum2<BLK>(cta_cnt, 0, s_red).x;
            if (tid == 0) {
                if (cta_cnt) atomicAdd(&fr[u].n_keep, cta_cnt);
                if (bx == 0) fr[u].thresh = thresh;
                if (defer && bx == 0) {                         // the record of the frame after next (its phase A follows the next barrier)
                    HmmFrame z; z.best = kWorstScore; z.n_keep = 0; z.thresh = kWorstScore; z.pad = 0;
                    (r.fr3 + (size_t)((r.slot0 + f + 2) % 3) * n_utt)[u] = z;
                }
            }
            // this CTA's partial mask of the frame (the block_sum2 barriers above publish the flags): 32 flag bytes (0 / 1) -> one
            // mask word per thread, two 16-byte reads and a multiply that packs four bytes into a nibble
            for (int kk = tid; kk < n_words; kk += BLK) {
                const uint4 a = reinterpret_cast<const uint4 *>(s_flag)[2 * kk], b = reinterpret_cast<const uint4 *>(s_flag)[2 * kk + 1];
                auto nib = [](uint32_t x) { return (x * 0x01020408u) >> 24 & 0xfu; };
                part_u[kk] = nib(a.x) | nib(a.y) << 4 | nib(a.z) << 8 | nib(a.w) << 12 | nib(b.x) << 16 | nib(b.y) << 20 | nib(b.z) << 24 | nib(b.w) << 28;
            }
            if (f > 0) merge_finish<BLK>(mp_, mpart, gx, n_words, r.mask2 + ((size_t)((r.mask0 + f - 1) & 1) * n_utt + u) * n_words, bx, gx, mt);
            __syncthreads();                                    // (the next utterance of this CTA clears the flags)
        }
        if (probe) r.probe[3] = clock64();
        // ------------------------------------------------ C: scatter (of frame g)
        auto phase_c = [&](int g) {
        HmmFrame *fr_n2 = r.fr3 + (size_t)((r.slot0 + g + 2) % 3) * n_utt;
        HmmFrame *fr = r.fr3 + (size_t)((r.slot0 + g) % 3) * n_utt;
        for (int u = u_lo; u < u_hi; u += gy) {
            const int lo = p.utt_off[u], hi = p.utt_off[u + 1];
            const int n_tiles = (hi - lo + BLK - 1) / BLK;
            const uint32_t *bal_u = s_bal + ((size_t)(defer ? (g & 1) * ((n_utt + gy - 1) / gy) : 0) + (CL ? 0 : (u - by) / gy)) * rows * BS;
            // survivors of the utterances before this one; the utterance's tile counts
            int part = 0;
            if (!own_list) for (int k = tid; k < u; k += BLK) part += fr[k].n_keep;
            for (int k = tid; k < n_tiles; k += BLK) s_tc[k] = __ldcg(r.tile_count + ((size_t)(defer ? (g & 1) : 0) * n_utt + u) * r.tpu + k);
            const int base_all = block_sum2<BLK>(part, 0, s_red).x;  // (its barriers also publish s_tc)
            if (probe) r.probe[8] = clock64();
            const int base = own_list ? lo : base_all;               // a list per utterance, packed after the run
            int32_t *keep_dst = own_list ? r.keep_tmp : r.keep_idx;
            if (bx == 0 && tid == 0) {
                if (!own_list && u == n_utt - 1) *r.total = base + fr[u].n_keep;
                if (!defer) {
                    HmmFrame z; z.best = kWorstScore; z.n_keep = 0; z.thresh = kWorstScore; z.pad = 0;
                    fr_n2[u] = z;                               // the record of the frame after next
                }
            }
            if (bx >= n_tiles) continue;                        // (uniform) no tile of this utterance
            // exclusive scan of the utterance's tile counts, in place (one pass: a chunk per thread,
            // then a scan of the 256 chunk sums)
            {
                const int per = (n_tiles + BLK - 1) / BLK;
                const int k0 = min(n_tiles, tid * per), k1 = min(n_tiles, k0 + per);
                int sum = 0;
                for (int k = k0; k < k1; ++k) sum += s_tc[k];
                int x = sum;
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                if (lane == 31) s_wcnt[w] = x;
                __syncthreads();
                int wbase = 0;
                for (int k = 0; k < w; ++k) wbase += s_wcnt[k];
                int run = base + wbase + x - sum;
                for (int k = k0; k < k1; ++k) { const int v = s_tc[k]; s_tc[k] = run; run += v; }
            }
            __syncthreads();
            if (probe) { r.probe[9] = clock64(); r.probe[10] = r.probe[9]; }
            // the survivors' indices, in order: tile offset + the warp's prefix (phase B) + the lane's rank in its ballot
            for (int row = 0; bx + row * gx < n_tiles; ++row) {
                const unsigned bal = bal_u[row * BS + w];
                if ((bal >> lane) & 1u) {
                    const int woff = reinterpret_cast<const uint16_t *>(bal_u + row * BS + BLK / 32)[w];
                    keep_dst[s_tc[bx + row * gx] + woff + __popc(bal & ((1u << lane) - 1u))] = lo + (bx + row * gx) * BLK + tid;
                }
            }
            __syncthreads();                                    // (s_tc is rewritten for the next utterance of this CTA)
        }
        };
        if (!defer) {
            sync_ctas<CL>(bar, n_cta, epoch);
            if (probe) r.probe[4] = clock64();
            phase_c(f);
        } else {
            if (probe) r.probe[4] = clock64();
            if (f > 0) phase_c(f - 1);
        }
        if (probe) r.probe[5] = clock64();
    }
    if (r.do_beam && r.n_frames > 0) {                          // the last frame's mask
        sync_ctas<CL>(bar, n_cta, epoch);
        const int f = r.n_frames - 1;
        for (int u = u_lo; u < u_hi; u += gy)
            merge_mask<BLK>(r.mask_part + ((size_t)((r.mask0 + f) & 1) * n_utt + u) * gx * n_words, gx, n_words,
                       r.mask2 + ((size_t)((r.mask0 + f) & 1) * n_utt + u) * n_words, bx, gx);
    }
    }                                                           // next utterance of this cluster
}

// The cluster form leaves the last frame's survivors as one list per utterance at keep_tmp + utt_off[u];
// pack them in (utterance, HMM index) order and publish the total.  One block per utterance.
__global__ void __launch_bounds__(256)
hmm_compact_last_kernel(HmmPop p, const HmmFrame *__restrict__ fr, const int32_t *__restrict__ keep_tmp,
                        int32_t *__restrict__ keep_idx, int32_t *__restrict__ total) {
    __shared__ int s_base;
    const int u = blockIdx.x;
    if (threadIdx.x == 0) {
        int b = 0;
        for (int k = 0; k < u; ++k) b += fr[k].n_keep;
        s_base = b;
        if (u == p.n_utt - 1) *total = b + fr[u].n_keep;
    }
    __syncthreads();
    const int n = fr[u].n_keep, lo = p.utt_off[u], base = s_base;
    for (int i = threadIdx.x; i < n; i += 256) keep_idx[base + i] = keep_tmp[lo + i];
}

// ------------------------------------------------------------ the resident form
// When the whole population fits one HMM per thread of one resident wave (tiles of all utterances <= CTAs the device
// keeps resident: the literal config 4, one utterance x 50 000 HMMs = 196 CTAs), a run of frames never re-reads the state:
// every thread loads its HMM once, keeps it in registers through all frames of the run and stores it after the last one.
// Per frame only the senone row (prefetched one frame ahead with cp.async into a second buffer), the frame records, the
// tile counts, the survivor list and the partial masks move; phases, barriers (per utterance), records, masks and lists
// are those of hmm_run_kernel's row form, which the same tests compare it with.
template <int NE>
__global__ void __launch_bounds__(kHmmBlock, 4)
hmm_resident_kernel(HmmDev c, HmmPop p, HmmRun r) {
    constexpr int BLK = kHmmBlock;
    extern __shared__ uint8_t sm_raw[];
    const size_t sen_bytes = ((size_t)c.n_sen * 2 + 15) & ~(size_t)15;
    int16_t *s_sen2 = reinterpret_cast<int16_t *>(sm_raw);                       // two rows
    uint8_t *s_tp = sm_raw + 2 * sen_bytes;
    uint32_t *s_flag_w = reinterpret_cast<uint32_t *>(s_tp + (size_t)c.n_tmat * kTpStride<NE>);
    uint8_t *s_flag = reinterpret_cast<uint8_t *>(s_flag_w);
    __shared__ int32_t s_red[2 * (BLK / 32)];
    __shared__ int32_t s_wcnt[BLK / 32];
    __shared__ uint32_t s_bal[BLK / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int gx = gridDim.x, bx = blockIdx.x, u = blockIdx.y;
    const int n = p.n_hmm, n_utt = p.n_utt, n_words = (c.n_sen + 31) / 32;
    unsigned *const bar = r.bar + (size_t)u * 32;
    unsigned epoch = 0;
    const int lo = p.utt_off[u], hi = p.utt_off[u + 1];
    const int n_tiles = (hi - lo + BLK - 1) / BLK;
    const bool has = bx < n_tiles;                          // (a shorter utterance leaves some CTAs of its row without a tile)
    const int i = lo + bx * BLK + tid;
    const bool on = has && i < hi;
    {
        const int ntp = c.n_tmat * NE * (NE + 1);
        for (int k = tid; k < ntp; k += BLK) s_tp[(k / (NE * (NE + 1))) * kTpStride<NE> + k % (NE * (NE + 1))] = c.tp[k];
        if (bx == 0 && tid < 3) { HmmFrame z; z.best = kWorstScore; z.n_keep = 0; z.thresh = kWorstScore; z.pad = 0; r.fr3[(size_t)tid * n_utt + u] = z; }
    }
    HmmRegs h; int tm = 0; bool mpx = false;
    h.best = kWorstScore;
    if (on) {
#pragma unroll
        for (int s = 0; s < NE; ++s) { h.sc[s] = p.score[(size_t)s * n + i]; h.hi[s] = p.history[(size_t)s * n + i]; h.sid[s] = p.senid[(size_t)s * n + i]; }
        h.out_sc = p.out_score[i]; h.out_hi = p.out_history[i];
        tm = (int)p.tmatid[i]; mpx = p.mpx[i] != 0;
    }
    // the senone row of a frame -> buffer (frame & 1); 16-byte cp.async when the row allows it
    auto row_of = [&](int f) { return r.sen_base + (size_t)((r.frame0 + f) % r.n_cycle) * r.frame_stride + (size_t)u * c.n_sen; };
    auto stage_row = [&](int f) {
        const int16_t *src = row_of(f);
        int16_t *dst = s_sen2 + (size_t)(f & 1) * (sen_bytes / 2);
        const int n16 = ((reinterpret_cast<size_t>(src) & 15) == 0) ? (c.n_sen * 2) / 16 : 0;
        const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
        for (int k = tid; k < n16; k += BLK)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + k * 16), "l"(reinterpret_cast<const int4 *>(src) + k) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (int k = n16 * 8 + tid; k < c.n_sen; k += BLK) dst[k] = src[k];
    };
    if (has) stage_row(0);
    grid_barrier(bar, gx, epoch);

    // ONE barrier per frame: the frame's best needs one (every HMM evaluated before the beam test), and the survivor
    // list needs the tile counts of the whole utterance -- but nothing in frame f + 1 waits for the list or the mask of
    // frame f, so phase C of frame f runs behind the barrier of frame f + 1, next to that frame's beam test, on what
    // the thread kept of frame f in registers (its ballot, prefix and senone ids).  Iteration f: A(f) | barrier |
    // B(f), C(f - 1), merge of the partial masks of f - 2.  Tile counts alternate between two buffers.
    const int F = r.n_frames;
    unsigned bal_p = 0; int woff_p = 0; uint32_t rid_p[NE];
#pragma unroll
    for (int s = 0; s < NE; ++s) rid_p[s] = 0xffffffffu;
    for (int f = 0; f <= F; ++f) {
        HmmFrame *fr = r.fr3 + (size_t)((r.slot0 + f) % 3) * n_utt;
        uint32_t rid[NE];
#pragma unroll
        for (int s = 0; s < NE; ++s) rid[s] = 0xffffffffu;
        // ------------------------------------------------ A(f)
        if (has && f < F) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                                // the row of frame f is complete; the other buffer's readers are done
            if (f + 1 < F) stage_row(f + 1);
            const int16_t *sen = s_sen2 + (size_t)(f & 1) * (sen_bytes / 2);
            int32_t blockbest = kWorstScore;
            if (on) {
                const TpRow<NE> tp = tp_row<NE>(s_tp, tm);
                if constexpr (NE == 3) { if (mpx) eval3_mpx(h, tp, sen, c.sseq); else eval3(h, tp, sen); }
                else if constexpr (NE == 5) { if (mpx) eval5_mpx(h, tp, sen, c.sseq); else eval5(h, tp, sen); }
                else eval_any<NE>(h, tp, sen, c.sseq, mpx);
                blockbest = h.best;
#pragma unroll
                for (int s = 0; s < NE; ++s) {              // the senones this HMM activates if it survives the frame
                    uint32_t id = h.sid[s];
                    if (mpx) id = id != B200_BAD_SSID ? (uint32_t)c.sseq[(size_t)id * NE + s] : 0xffffffffu;
                    rid[s] = id;
                }
            }
            for (int o = 16; o > 0; o >>= 1) blockbest = max(blockbest, __shfl_xor_sync(0xffffffffu, blockbest, o));
            if (lane == 0) s_wcnt[w] = blockbest;
            __syncthreads();
            if (tid == 0) {
                int32_t b = s_wcnt[0];
                for (int k = 1; k < BLK / 32; ++k) b = max(b, s_wcnt[k]);
                atomicMax(&fr[u].best, b);
            }
        }
        grid_barrier(bar, gx, epoch);

        // the partial masks of frame f - 2 are complete: loads now, merge at the end of the iteration
        const MergePlan mp_ = merge_plan<BLK>(gx, n_words, bx, gx);
        const uint32_t *mpart = r.mask_part + ((size_t)((r.mask0 + f - 2) & 1) * n_utt + u) * gx * n_words;
        uint32_t mt[8];
        if (f >= 2) merge_load(mp_, mpart, gx, n_words, bx, gx, 0, mp_.sub, mt);
        int before = 0;                                     // survivors of frame f - 1 in the tiles ahead of mine (loads issued with the others)
        if (f > 0 && has)
            for (int k = tid; k < bx; k += BLK) before += __ldcg(r.tile_count + ((size_t)((f - 1) & 1) * n_utt + u) * r.tpu + k);
        // ------------------------------------------------ B(f)
        unsigned bal = 0; int woff = 0;
        if (f < F) {
            if (bx == 0 && tid == 0) {
                HmmFrame z; z.best = kWorstScore; z.n_keep = 0; z.thresh = kWorstScore; z.pad = 0;
                (r.fr3 + (size_t)((r.slot0 + f + 2) % 3) * n_utt)[u] = z;      // the record of the frame after next (its A follows the next barrier)
            }
            const int32_t thresh = __ldcg(&fr[u].best) + r.beam;
            const bool keep = on && BT(h.best, thresh);
            bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_bal[w] = bal;
            __syncthreads();
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < BLK / 32; ++k) { const int pc = __popc(s_bal[k]); if (k < w) woff += pc; cnt += pc; }
            if (tid == 0) {
                if (has) r.tile_count[((size_t)(f & 1) * n_utt + u) * r.tpu + bx] = cnt;
                if (cnt) atomicAdd(&fr[u].n_keep, cnt);
                if (bx == 0) fr[u].thresh = thresh;
            }
        }
        // ------------------------------------------------ C(f - 1)
        if (f > 0) {
            uint32_t *part_u = r.mask_part + (((size_t)((r.mask0 + f - 1) & 1) * n_utt + u) * gx + bx) * n_words;
            if (!has) {
                for (int k = tid; k < n_words; k += BLK) part_u[k] = 0u;
            } else {
                for (int k = tid; k < n_words * 8; k += BLK) s_flag_w[k] = 0u;
                const int base = lo + block_sum2<BLK>(before, 0, s_red).x;      // (its barriers also publish the cleared flags and free s_bal)
                if ((bal_p >> lane) & 1u) {
                    r.keep_tmp[base + woff_p + __popc(bal_p & ((1u << lane) - 1u))] = i;
#pragma unroll
                    for (int s = 0; s < NE; ++s)
                        if (rid_p[s] != 0xffffffffu) s_flag[rid_p[s]] = 1;
                }
                __syncthreads();
                for (int kk = tid; kk < n_words; kk += BLK) {
                    const uint4 a = reinterpret_cast<const uint4 *>(s_flag)[2 * kk], b = reinterpret_cast<const uint4 *>(s_flag)[2 * kk + 1];
                    auto nib = [](uint32_t x) { return (x * 0x01020408u) >> 24 & 0xfu; };
                    part_u[kk] = nib(a.x) | nib(a.y) << 4 | nib(a.z) << 8 | nib(a.w) << 12 | nib(b.x) << 16 | nib(b.y) << 20 | nib(b.z) << 24 | nib(b.w) << 28;
                }
            }
        }
        if (f >= 2) merge_finish<BLK>(mp_, mpart, gx, n_words, r.mask2 + ((size_t)((r.mask0 + f - 2) & 1) * n_utt + u) * n_words, bx, gx, mt);
        bal_p = bal;                                        // the warp's keep ballot of frame f
        woff_p = woff;
#pragma unroll
        for (int s = 0; s < NE; ++s) rid_p[s] = rid[s];
    }
    if (on) {                                               // the state goes back once
#pragma unroll
        for (int s = 0; s < NE; ++s) { p.score[(size_t)s * n + i] = h.sc[s]; p.history[(size_t)s * n + i] = h.hi[s]; }
        if (mpx) {
#pragma unroll
            for (int s = 1; s < NE; ++s) p.senid[(size_t)s * n + i] = h.sid[s];
        }
        p.out_score[i] = h.out_sc; p.out_history[i] = h.out_hi; p.bestscore[i] = h.best;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    grid_barrier(bar, gx, epoch);
    {
        const int f = r.n_frames - 1;
        merge_mask<BLK>(r.mask_part + ((size_t)((r.mask0 + f) & 1) * n_utt + u) * gx * n_words, gx, n_words,
                        r.mask2 + ((size_t)((r.mask0 + f) & 1) * n_utt + u) * n_words, bx, gx);
    }
}

// ------------------------------------------------------------ host launcher
static size_t run_smem_base(const HmmDev &c, int tpu, int gx, int utts_per_cta, int blk);
static size_t run_smem(const HmmDev &c, int tpu, int gx, int utts_per_cta, int blk, int *pre_off = nullptr) {
    const size_t base = (run_smem_base(c, tpu, gx, utts_per_cta, blk) + 15) & ~(size_t)15;
    if (pre_off) *pre_off = (int)base;
    // the single loop's kPre stages of 4-byte slots, or (NE <= 3) the pair loop's one stage of 8-byte slots (16 kB either way for NE = 3:
    // four CTAs then fit the 164 kB shared-memory carve-out and the L1 keeps 92 kB -- with 32 kB of slots it drops to 28 kB and phase A runs 50 % slower)
    return base + std::max((size_t)kPre * (2 * c.n_emit + 2) * blk * 4, c.n_emit <= 3 ? (size_t)(2 * c.n_emit + 2) * blk * 8 : (size_t)0);
}
static size_t run_smem_base(const HmmDev &c, int tpu, int gx, int utts_per_cta, int blk) {
    const int rows = (tpu + gx - 1) / gx + kBeamBatch - 1;
    return (((size_t)c.n_sen * 2 + 15) & ~(size_t)15) + (size_t)c.n_tmat * ((c.n_emit * (c.n_emit + 1) + 15) / 16 * 16) +
           (size_t)((c.n_sen + 31) / 32) * 32 + ((size_t)tpu + rows + (size_t)2 * utts_per_cta * rows * (blk / 32 + blk / 64)) * 4 + 16;      // (ballots of two frames: see `defer`)
}

#define B200_HMM_NE(...)                                   \
    switch (c.n_emit) {                                    \
    case 1: { constexpr int NE = 1; __VA_ARGS__; } break;  \
    case 2: { constexpr int NE = 2; __VA_ARGS__; } break;  \
    case 3: { constexpr int NE = 3; __VA_ARGS__; } break;  \
    case 4: { constexpr int NE = 4; __VA_ARGS__; } break;  \
    default: { constexpr int NE = 5; __VA_ARGS__; } break; \
    }

constexpr int kClusterBlock = 1024;

// The cluster form (see hmm_run_kernel): grid rows = clusters of `cs` CTAs x 1024 threads, as many
// rows as the device can keep resident, each taking utterances row, row + rows, ... through the
// whole run.  Returns 1 when the form does not apply (then the cooperative form runs).
static int hmm_launch_run_cluster(const HmmDev &c, const HmmPop &p, const HmmRun &run_in, cudaStream_t st) {
    static int enabled = -1;
    // Opt-in (B200_HMM_CLUSTER=1): measured SLOWER than the cooperative form on B200 -- 17.4 vs 14.5 us per
    // frame for one utterance x 50 000 HMMs, 175 vs 95 us for 64 utterances.  The cluster barriers are
    // cheaper (1.2-1.5 k cycles against 2.8-3.0 k for the grid barrier), but a frame's time is the chain of
    // dependent global round trips INSIDE the phases, and 16 CTAs x 1024 threads walk it three times per
    // phase where 196 CTAs x 256 threads walk it once (phase A 13.2 k cycles against 4.3 k; profiles/README.md).
    if (enabled < 0) { const char *e = getenv("B200_HMM_CLUSTER"); enabled = (e && atoi(e) != 0) ? 1 : 0; }
    if (!enabled || !run_in.do_beam) return 1;
    HmmRun r = run_in;
    r.tpu = (p.max_per_utt + kClusterBlock - 1) / kClusterBlock;
    static AttrOnce attr;
    if (attr.need()) {
        cudaError_t e = cudaSuccess;
#define B200_SET(NE_)                                                                                                          \
        e = cudaFuncSetAttribute(hmm_run_kernel<NE_, kClusterBlock, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(hmm_run_kernel<NE_, kClusterBlock, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); \
        if (e != cudaSuccess) { cudaGetLastError(); enabled = 0; return 1; }
        B200_SET(1) B200_SET(2) B200_SET(3) B200_SET(4) B200_SET(5)
#undef B200_SET
    }
    for (int cs = 16; cs >= 8; cs >>= 1) {
        const size_t sh = run_smem(c, r.tpu, cs, 1, kClusterBlock, &r.pre_off);
        if (sh > 200 * 1024) continue;
        if ((size_t)2 * p.n_utt * cs * ((c.n_sen + 31) / 32) > run_in.mask_part_words) continue;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(cs, 1, 1); cfg.blockDim = dim3(kClusterBlock, 1, 1); cfg.dynamicSmemBytes = sh; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n_clusters = 0;
        cudaError_t e = cudaSuccess;
        B200_HMM_NE(e = cudaOccupancyMaxActiveClusters(&n_clusters, hmm_run_kernel<NE, kClusterBlock, true>, &cfg));
        if (e != cudaSuccess || n_clusters < 1) { cudaGetLastError(); continue; }
        cfg.gridDim = dim3(cs, std::min(p.n_utt, n_clusters), 1);
        HmmDev cc = c; HmmPop pp = p;
        B200_HMM_NE(e = cudaLaunchKernelEx(&cfg, hmm_run_kernel<NE, kClusterBlock, true>, cc, pp, r));
        if (e != cudaSuccess) { cudaGetLastError(); continue; }
        g_launches.fetch_add(1, std::memory_order_relaxed);
        const HmmFrame *fr_last = r.fr3 + (size_t)((r.slot0 + r.n_frames - 1) % 3) * p.n_utt;
        hmm_compact_last_kernel<<<p.n_utt, 256, 0, st>>>(pp, fr_last, r.keep_tmp, r.keep_idx, r.total);
        B200_LAUNCH_CHECK();
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return B200_OK;
    }
    return 1;
}

// The resident form (see hmm_resident_kernel).  Returns 1 when it does not apply.
static int hmm_launch_run_resident(const HmmDev &c, const HmmPop &p, const HmmRun &run_in, cudaStream_t st) {
    static int enabled = -1;
    if (enabled < 0) { const char *e = getenv("B200_HMM_RESIDENT"); enabled = (e && atoi(e) == 0) ? 0 : 1; }
    if (!enabled || !run_in.do_beam || run_in.n_frames < 2 || p.n_utt > kHmmBarRows) return 1;
    const int gx = (p.max_per_utt + kHmmBlock - 1) / kHmmBlock, gy = p.n_utt;
    int n_sm = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if ((long)gx * gy > 8L * n_sm) return 1;
    const size_t sh = 2 * (((size_t)c.n_sen * 2 + 15) & ~(size_t)15) + (size_t)c.n_tmat * ((c.n_emit * (c.n_emit + 1) + 15) / 16 * 16) +
                      (size_t)((c.n_sen + 31) / 32) * 32 + 16;
    if (sh > 200 * 1024) return 1;
    if ((size_t)2 * p.n_utt * gx * ((c.n_sen + 31) / 32) > run_in.mask_part_words) return 1;
    static AttrOnce attr;
    if (attr.need()) {
        cudaError_t e = cudaSuccess;
        // (every instantiation: the once-flag is per device, not per NE)
        e = cudaFuncSetAttribute(hmm_resident_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); B200_CUDA_OK(e);
        e = cudaFuncSetAttribute(hmm_resident_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); B200_CUDA_OK(e);
        e = cudaFuncSetAttribute(hmm_resident_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); B200_CUDA_OK(e);
        e = cudaFuncSetAttribute(hmm_resident_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); B200_CUDA_OK(e);
        e = cudaFuncSetAttribute(hmm_resident_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); B200_CUDA_OK(e);
    }
    int per_sm = 0;
    B200_HMM_NE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hmm_resident_kernel<NE>, kHmmBlock, sh));
    if ((long)per_sm * n_sm < (long)gx * gy) return 1;
    HmmRun r = run_in;
    HmmDev cc = c; HmmPop pp = p;
    B200_CUDA_OK(cudaMemsetAsync(r.bar, 0, (size_t)gy * 32 * sizeof(unsigned), st));
    void *args[] = {(void *)&cc, (void *)&pp, (void *)&r};
    cudaError_t e = cudaSuccess;
    B200_HMM_NE(e = cudaLaunchCooperativeKernel((const void *)hmm_resident_kernel<NE>, dim3(gx, gy), dim3(kHmmBlock), args, sh, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    B200_CUDA_OK(e);
    const HmmFrame *fr_last = r.fr3 + (size_t)((r.slot0 + r.n_frames - 1) % 3) * p.n_utt;
    hmm_compact_last_kernel<<<p.n_utt, 256, 0, st>>>(pp, fr_last, r.keep_tmp, r.keep_idx, r.total);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int hmm_launch_run(const HmmDev &c, const HmmPop &p, const HmmRun &run_in, cudaStream_t st) {
    if (p.n_hmm <= 0 || p.n_utt <= 0 || run_in.n_frames <= 0) return B200_OK;
    if (c.n_emit < 1 || c.n_emit > 5) { set_error("n_emit_state %d outside 1..5 (HMM_MAX_NSTATE)", c.n_emit); return B200_ERR_UNSUP; }
    {
        const int rc = hmm_launch_run_cluster(c, p, run_in, st);
        if (rc != 1) return rc;
    }
    {
        const int rc = hmm_launch_run_resident(c, p, run_in, st);
        if (rc != 1) return rc;
    }
    static AttrOnce attr;
    if (attr.need()) {
        cudaError_t e = cudaSuccess;
        e = cudaFuncSetAttribute(hmm_run_kernel<1, kHmmBlock, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); B200_CUDA_OK(e);
        e = cudaFuncSetAttribute(hmm_run_kernel<2, kHmmBlock, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); B200_CUDA_OK(e);
        e = cudaFuncSetAttribute(hmm_run_kernel<3, kHmmBlock, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); B200_CUDA_OK(e);
        e = cudaFuncSetAttribute(hmm_run_kernel<4, kHmmBlock, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); B200_CUDA_OK(e);
        e = cudaFuncSetAttribute(hmm_run_kernel<5, kHmmBlock, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); B200_CUDA_OK(e);
    }
    // exactly one resident wave of CTAs (occupancy x SM count); the shared-memory need depends on
    // the grid shape (tiles per CTA), so: shape from the occupancy at the base need, then re-check
    const int bpu = (p.max_per_utt + kHmmBlock - 1) / kHmmBlock;
    int n_sm = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    int gx = 1, gy = 1, pre_off = 0;
    size_t sh = 0;
    for (int per_sm_try = (c.n_emit == 3 ? 5 : 4); per_sm_try >= 1; --per_sm_try) {
        const int wave = per_sm_try * n_sm;
        gy = std::max(1, std::min(p.n_utt, wave));
        gx = std::max(1, std::min(bpu, wave / gy));
        sh = run_smem(c, bpu, gx, (p.n_utt + gy - 1) / gy, kHmmBlock, &pre_off);
        if (sh > 200 * 1024) continue;
        int per_sm = 0;
        B200_HMM_NE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hmm_run_kernel<NE, kHmmBlock, false>, kHmmBlock, sh));
        if (per_sm >= per_sm_try) break;
        if (per_sm_try == 1) { set_error("hmm step: population does not fit one resident wave (%zu B shared memory per CTA)", sh); return B200_ERR_UNSUP; }
    }
    if (sh > 200 * 1024) { set_error("hmm step needs %zu B shared memory", sh); return B200_ERR_UNSUP; }
    HmmRun r = run_in;
    r.pre_off = pre_off;
    {
        static int pair_on = -1;
        if (pair_on < 0) { const char *e = getenv("B200_HMM_PAIR"); pair_on = (e && atoi(e) == 0) ? 0 : 1; }
        r.pair = pair_on;
    }
    if ((size_t)2 * p.n_utt * gx * ((c.n_sen + 31) / 32) > run_in.mask_part_words) {
        set_error("hmm step: partial-mask buffer too small (%zu words)", run_in.mask_part_words);
        return B200_ERR_ARG;
    }
    HmmDev cc = c; HmmPop pp = p;
    // row barriers when there is more than one row and a run of frames to drift over (B200_HMM_ROWSYNC=0: grid barriers)
    static int row_ok = -1;
    if (row_ok < 0) { const char *e = getenv("B200_HMM_ROWSYNC"); row_ok = (e && atoi(e) == 0) ? 0 : 1; }
    r.row_sync = (row_ok && r.do_beam && gy > 1 && r.n_frames > 1 && gy <= kHmmBarRows) ? 1 : 0;
    B200_CUDA_OK(cudaMemsetAsync(r.bar, 0, r.row_sync ? (size_t)gy * 32 * sizeof(unsigned) : sizeof(unsigned), st));
    void *args[] = {(void *)&cc, (void *)&pp, (void *)&r};
    cudaError_t e = cudaSuccess;
    B200_HMM_NE(e = cudaLaunchCooperativeKernel((const void *)hmm_run_kernel<NE, kHmmBlock, false>, dim3(gx, gy), dim3(kHmmBlock), args, sh, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    B200_CUDA_OK(e);
    if (r.row_sync) {
        const HmmFrame *fr_last = r.fr3 + (size_t)((r.slot0 + r.n_frames - 1) % 3) * p.n_utt;
        hmm_compact_last_kernel<<<p.n_utt, 256, 0, st>>>(pp, fr_last, r.keep_tmp, r.keep_idx, r.total);
        B200_LAUNCH_CHECK();
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return B200_OK;
}
#undef B200_HMM_NE


// ------------------------------------------------------------ evaluation of a LIST of channels
// eval_root_chan + eval_nonroot_chan (PS/ngram_search_fwdtree.c:598-634): hmm_vit_eval for the root
// channels whose frame stamp is the current frame and for every entry of the frame's active list, the
// best of their scores per utterance.  The population is n_utt copies of one lexical tree's channels
// (channel c of utterance u = HMM u * n_chan + c), the lists are those fwdtree_prune_kernel writes
// (csrc/fwdtree_prune.cu): evaluate and prune alternate on the device without a host copy of any
// state.  Gathers by index: the senone scores and the transition table are read through L1 / L2.
template <int NE>
__global__ void __launch_bounds__(kHmmBlock)
hmm_eval_list_kernel(HmmDev c, HmmPop p, HmmList l) {
    __shared__ int32_t s_best[kHmmBlock / 32];
    const int u = blockIdx.y, tid = threadIdx.x;
    const int n_act = l.n_act[u];
    const int e = blockIdx.x * kHmmBlock + tid;
    if (blockIdx.x * kHmmBlock >= l.n_root + n_act) return;          // uniform
    const int32_t fi = l.par[(size_t)u * 8];
    const size_t base = (size_t)u * l.n_chan;
    int32_t best = kWorstScore;
    int ch = -1;
    if (e < l.n_root) { if (l.frame[base + e] == fi) ch = e; }       // :604
    else if (e < l.n_root + n_act) ch = l.acl[(size_t)u * l.list_cap + (e - l.n_root)];
    if (ch >= 0) {
        const size_t i = base + ch;
        const int n = p.n_hmm;
        const int16_t *sen = l.senscr + (size_t)u * c.n_sen;
        HmmRegs h;
#pragma unroll
        for (int s = 0; s < NE; ++s) {
            h.sc[s] = p.score[(size_t)s * n + i];
            h.hi[s] = p.history[(size_t)s * n + i];
            h.sid[s] = p.senid[(size_t)s * n + i];
        }
        h.out_sc = p.out_score[i];
        h.out_hi = p.out_history[i];
        const uint8_t *tp = c.tp + (int)p.tmatid[i] * NE * (NE + 1);
        const bool mpx = p.mpx[i] != 0;
        if (NE == 3) { if (mpx) eval3_mpx(h, tp, sen, c.sseq); else eval3(h, tp, sen); }
        else if (NE == 5) { if (mpx) eval5_mpx(h, tp, sen, c.sseq); else eval5(h, tp, sen); }
        else eval_any<NE>(h, tp, sen, c.sseq, mpx);
#pragma unroll
        for (int s = 0; s < NE; ++s) {
            p.score[(size_t)s * n + i] = h.sc[s];
            p.history[(size_t)s * n + i] = h.hi[s];
        }
        if (mpx) {
#pragma unroll
            for (int s = 1; s < NE; ++s) p.senid[(size_t)s * n + i] = h.sid[s];
        }
        p.out_score[i] = h.out_sc;
        p.out_history[i] = h.out_hi;
        p.bestscore[i] = h.best;
        best = h.best;
    }
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((tid & 31) == 0) s_best[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < kHmmBlock / 32; ++k) best = max(best, s_best[k]);
        if (best > kWorstScore) atomicMax(&l.best[u], best);
    }
}

__global__ void hmm_fill_kernel(int32_t *dst, int n, int32_t v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

int hmm_launch_eval_list(const HmmDev &c, const HmmPop &p, const HmmList &l, int n_utt, cudaStream_t st) {
    hmm_fill_kernel<<<(n_utt + 255) / 256, 256, 0, st>>>(l.best, n_utt, kWorstScore);
    B200_LAUNCH_CHECK();
    const dim3 grid((l.n_root + l.list_cap + kHmmBlock - 1) / kHmmBlock, n_utt);
    switch (c.n_emit) {
        case 1: hmm_eval_list_kernel<1><<<grid, kHmmBlock, 0, st>>>(c, p, l); break;
        case 2: hmm_eval_list_kernel<2><<<grid, kHmmBlock, 0, st>>>(c, p, l); break;
        case 3: hmm_eval_list_kernel<3><<<grid, kHmmBlock, 0, st>>>(c, p, l); break;
        case 4: hmm_eval_list_kernel<4><<<grid, kHmmBlock, 0, st>>>(c, p, l); break;
        default: hmm_eval_list_kernel<5><<<grid, kHmmBlock, 0, st>>>(c, p, l); break;
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// ------------------------------------------------------------ maintenance
// hmm_normalize (PS/hmm.c:207-218) for the whole population: every state score
// and exit score that is BETTER_THAN WORST_SCORE loses its utterance's value.
__global__ void __launch_bounds__(kHmmBlock)
hmm_normalize_kernel(HmmPop p, int n_emit, const int32_t *__restrict__ best_per_utt, const HmmFrame *__restrict__ fr) {
    const int u = blockIdx.y;
    const int i = p.utt_off[u] + blockIdx.x * kHmmBlock + threadIdx.x;
    if (i >= p.utt_off[u + 1]) return;
    const int32_t b = best_per_utt ? best_per_utt[u] : fr[u].best;
    for (int s = 0; s < n_emit; ++s) {
        const int32_t v = p.score[(size_t)s * p.n_hmm + i];
        if (BT(v, kWorstScore)) p.score[(size_t)s * p.n_hmm + i] = v - b;
    }
    const int32_t o = p.out_score[i];
    if (BT(o, kWorstScore)) p.out_score[i] = o - b;
}

// hmm_clear_scores (PS/hmm.c:169-180) for the HMMs the beam step dropped
// (keep == 0): the `else` arm of prune_nonroot_chan, PS/ngram_search_fwdtree.c:864-866.
__global__ void __launch_bounds__(kHmmBlock)
hmm_clear_pruned_kernel(HmmPop p, int n_emit, const HmmFrame *__restrict__ fr) {
    const int u = blockIdx.y;
    const int i = p.utt_off[u] + blockIdx.x * kHmmBlock + threadIdx.x;
    if (i >= p.utt_off[u + 1]) return;
    if (BT(p.bestscore[i], fr[u].thresh)) return;         // kept by the last beam step
    for (int s = 0; s < n_emit; ++s) p.score[(size_t)s * p.n_hmm + i] = kWorstScore;
    p.out_score[i] = kWorstScore;
    p.bestscore[i] = kWorstScore;
}

// Batched hmm_enter (PS/hmm.c:199-205) with the test its callers make first
// (`score BETTER_THAN hmm_in_score`, PS/ngram_search_fwdtree.c:757,846): entry k
// wants to put (score[k], hist[k]) into state 0 of HMM idx[k].  Sequential
// semantics over the list: the best score wins, the FIRST entry among equal
// scores keeps its history, an entry that does not beat the resident score
// changes nothing.  Three passes: atomicMax on the score, atomicMin of the list
// position among the winners, history write.
__global__ void hmm_enter_max_kernel(HmmPop p, const int32_t *__restrict__ idx, const int32_t *__restrict__ score,
                                     int n, int32_t *__restrict__ winner) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = idx[k];
    winner[i] = 0x7fffffff;
    if (BT(score[k], p.score[i])) atomicMax(&p.score[i], score[k]);   // state 0 row
}
__global__ void hmm_enter_pick_kernel(HmmPop p, const int32_t *__restrict__ idx, const int32_t *__restrict__ score,
                                      const int32_t *__restrict__ old0, int n, int32_t *__restrict__ winner) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = idx[k];
    if (score[k] == p.score[i] && BT(score[k], old0[k])) atomicMin(&winner[i], k);
}
__global__ void hmm_enter_hist_kernel(HmmPop p, const int32_t *__restrict__ idx, const int32_t *__restrict__ hist,
                                      int n, const int32_t *__restrict__ winner, uint8_t *__restrict__ entered) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = idx[k];
    if (winner[i] == k) { p.history[i] = hist[k]; if (entered) entered[i] = 1; }
}
__global__ void hmm_enter_snapshot_kernel(HmmPop p, const int32_t *__restrict__ idx, int n, int32_t *__restrict__ old0) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) old0[k] = p.score[idx[k]];
}

int hmm_launch_normalize(const HmmPop &p, int n_emit, const int32_t *d_best_per_utt, const HmmFrame *fr, cudaStream_t st) {
    if (p.n_hmm <= 0) return B200_OK;
    const int bpu = (p.max_per_utt + kHmmBlock - 1) / kHmmBlock;
    hmm_normalize_kernel<<<dim3(bpu, p.n_utt), kHmmBlock, 0, st>>>(p, n_emit, d_best_per_utt, fr);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int hmm_launch_clear_pruned(const HmmPop &p, int n_emit, const HmmFrame *fr, cudaStream_t st) {
    if (p.n_hmm <= 0) return B200_OK;
    const int bpu = (p.max_per_utt + kHmmBlock - 1) / kHmmBlock;
    hmm_clear_pruned_kernel<<<dim3(bpu, p.n_utt), kHmmBlock, 0, st>>>(p, n_emit, fr);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

int hmm_launch_enter(const HmmPop &p, const int32_t *d_idx, const int32_t *d_score, const int32_t *d_hist, int n,
                     int32_t *d_winner /* [n_hmm] scratch */, int32_t *d_old0 /* [n] scratch */, uint8_t *d_entered,
                     cudaStream_t st) {
    if (n <= 0) return B200_OK;
    const int g = (n + 255) / 256;
    hmm_enter_snapshot_kernel<<<g, 256, 0, st>>>(p, d_idx, n, d_old0);
    B200_LAUNCH_CHECK();
    hmm_enter_max_kernel<<<g, 256, 0, st>>>(p, d_idx, d_score, n, d_winner);
    B200_LAUNCH_CHECK();
    hmm_enter_pick_kernel<<<g, 256, 0, st>>>(p, d_idx, d_score, d_old0, n, d_winner);
    B200_LAUNCH_CHECK();
    hmm_enter_hist_kernel<<<g, 256, 0, st>>>(p, d_idx, d_hist, n, d_winner, d_entered);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

// ------------------------------------------------------------ phone-loop look-ahead search
// phone_loop_search_step (PS/phone_loop_search.c:253-291) minus its acmod calls, for a batch of utterances
// in lock step: renormalize_hmms (:171-184) when best + 2 beam underflows, evaluate_hmms (:186-210),
// prune_hmms (:212-233), phone_transition (:235-268), and the look-ahead penalties the forward tree
// search reads through phone_loop_search_score (PS/phone_loop_search.h:104).  One CTA per utterance, one
// thread per CI phone.  phone_transition is a double loop whose inner test reads what earlier sources
// wrote: it is kept sequential over the SOURCE phone (a block-uniform loop of n_phones steps) and parallel
// over the target, which reproduces the reference for any input.
struct PhoneLoop {
    int n_phones, n_emit, n_sen, n_utt;
    const uint8_t *tp; const uint16_t *senid;     // [n_emit][n_phones], shared by the utterances
    const int16_t *tmatid;
    int32_t *score, *history;                     // [n_emit][n_utt * n_phones]
    int32_t *out_score, *out_history, *bestscore, *frame;
    int32_t *best;                                // [n_utt] pls->best_score
    int32_t *renorm;                              // [n_utt] norm applied this frame, or 0
    int32_t beam, pbeam, pip;
};

template <int NE>
__global__ void __launch_bounds__(1024)
phone_loop_step_kernel(PhoneLoop q, const int16_t *__restrict__ senscr_all, int frame_idx, int32_t *__restrict__ pls_pen) {
    extern __shared__ int32_t sm_pl[];
    int32_t *s_frame = sm_pl, *s_out = sm_pl + q.n_phones, *s_outh = sm_pl + 2 * q.n_phones;
    __shared__ int32_t s_red[32];
    __shared__ int32_t s_best;
    const int u = blockIdx.x, i = threadIdx.x, n = q.n_phones;
    const size_t N = (size_t)q.n_utt * n, at = (size_t)u * n + i;
    const bool on = i < n;
    const int16_t *sen = senscr_all + (size_t)u * q.n_sen;
    const int32_t nf = frame_idx + 1;
    HmmRegs h; int32_t fr = -1;
    if (on) {
#pragma unroll
        for (int s = 0; s < NE; ++s) { h.sc[s] = q.score[(size_t)s * N + at]; h.hi[s] = q.history[(size_t)s * N + at]; h.sid[s] = q.senid[(size_t)s * n + i]; }
        h.out_sc = q.out_score[at]; h.out_hi = q.out_history[at]; h.best = q.bestscore[at];
        fr = q.frame[at];
    }
    // renormalize_hmms
    const int32_t prev_best = q.best[u];
    int32_t norm = 0;
    if (WT(prev_best + 2 * q.beam, kWorstScore)) {                       // :273
        norm = prev_best;
        if (on) {
#pragma unroll
            for (int s = 0; s < NE; ++s) if (BT(h.sc[s], kWorstScore)) h.sc[s] -= norm;
            if (BT(h.out_sc, kWorstScore)) h.out_sc -= norm;
        }
    }
    // evaluate_hmms
    int32_t b = kWorstScore;
    if (on && fr >= frame_idx) {                                         // :199
        const uint8_t *tp = q.tp + (int)q.tmatid[i] * NE * (NE + 1);
        if (NE == 3) eval3(h, tp, sen);
        else if (NE == 5) eval5(h, tp, sen);
        else eval_any<NE>(h, tp, sen, nullptr, false);
        b = h.best;
    }
    for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    if ((i & 31) == 0) s_red[i >> 5] = b;
    __syncthreads();
    if (i == 0) {
        int32_t m = s_red[0];
        for (int k = 1; k < (int)((blockDim.x + 31) >> 5); ++k) m = max(m, s_red[k]);
        s_best = m;
    }
    __syncthreads();
    const int32_t bs = s_best;
    // prune_hmms
    if (on && fr >= frame_idx) {
        if (BT(h.best, bs + q.beam)) fr = nf;
        else {
#pragma unroll
            for (int s = 0; s < NE; ++s) h.sc[s] = kWorstScore;
            h.out_sc = kWorstScore; h.best = kWorstScore;
        }
    }
    if (on) { s_frame[i] = fr; s_out[i] = h.out_sc; s_outh[i] = h.out_hi; }
    __syncthreads();
    // phone_transition: sequential over the source, parallel over the target
    const int32_t thresh = bs + q.pbeam;
    for (int k = 0; k < n; ++k) {
        const bool src = s_frame[k] == nf;                               // :248 (as the walk finds it)
        const int32_t nps = s_out[k] + q.pip;
        __syncthreads();
        if (src && BT(nps, thresh) && on && (fr < frame_idx || BT(nps, h.sc[0]))) {   // :258-259
            h.sc[0] = nps; h.hi[0] = s_outh[k]; fr = nf;
            s_frame[i] = nf;
        }
        __syncthreads();
    }
    if (on) {
#pragma unroll
        for (int s = 0; s < NE; ++s) { q.score[(size_t)s * N + at] = h.sc[s]; q.history[(size_t)s * N + at] = h.hi[s]; }
        q.out_score[at] = h.out_sc; q.out_history[at] = h.out_hi; q.bestscore[at] = h.best; q.frame[at] = fr;
        if (pls_pen) pls_pen[at] = h.best - bs;                          // phone_loop_search_score
    }
    if (i == 0) { q.best[u] = bs; q.renorm[u] = norm; }
}

}  // namespace b200

using namespace b200;

struct b200_phoneloop {
    PhoneLoop q{};
    int device = 0;
    std::vector<void *> owned;
    int16_t *d_senscr = nullptr;      // host-form staging
    int32_t *d_pen = nullptr;
    cudaStream_t st = nullptr;
};

namespace {
template <typename T>
int pl_alloc(b200_phoneloop *h, T **dst, size_t n, const T *src) {
    void *p = nullptr;
    B200_CUDA_OK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    h->owned.push_back(p);
    if (src && n) B200_CUDA_OK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (T *)p;
    return B200_OK;
}
}  // namespace

extern "C" b200_phoneloop_t *b200_phone_loop_create(int n_phones, int n_emit, const uint8_t *tp, int n_tmat, const uint16_t *senid,
                                                    const int16_t *tmatid, int n_sen, int32_t beam, int32_t pbeam, int32_t pip,
                                                    int n_utt, int device) {
    if (n_phones < 1 || n_phones > 1024 || n_emit < 1 || n_emit > 5 || !tp || n_tmat < 1 || !senid || !tmatid || n_sen < 1 || n_utt < 1) {
        set_error("b200_phone_loop_create: bad argument"); return nullptr;
    }
    for (int i = 0; i < n_phones; ++i) {
        if (tmatid[i] < 0 || tmatid[i] >= n_tmat) { set_error("b200_phone_loop_create: tmatid[%d] = %d", i, tmatid[i]); return nullptr; }
        for (int s = 0; s < n_emit; ++s)
            if (senid[(size_t)s * n_phones + i] >= n_sen) { set_error("b200_phone_loop_create: senid[%d][%d] out of range", s, i); return nullptr; }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: libb200sphinx has no CPU fallback"); return nullptr; }
    if (device < 0 || device >= ndev || cudaSetDevice(device) != cudaSuccess) { set_error("bad device %d", device); return nullptr; }
    auto *h = new b200_phoneloop;
    h->device = device;
    PhoneLoop &q = h->q;
    q.n_phones = n_phones; q.n_emit = n_emit; q.n_sen = n_sen; q.n_utt = n_utt; q.beam = beam; q.pbeam = pbeam; q.pip = pip;
    const size_t N = (size_t)n_utt * n_phones;
    uint8_t *d_tp = nullptr; uint16_t *d_sid = nullptr; int16_t *d_tm = nullptr;
    if (pl_alloc(h, &d_tp, (size_t)n_tmat * n_emit * (n_emit + 1), tp) || pl_alloc(h, &d_sid, (size_t)n_emit * n_phones, senid) ||
        pl_alloc(h, &d_tm, (size_t)n_phones, tmatid) || pl_alloc<int32_t>(h, &q.score, N * n_emit, nullptr) ||
        pl_alloc<int32_t>(h, &q.history, N * n_emit, nullptr) || pl_alloc<int32_t>(h, &q.out_score, N, nullptr) ||
        pl_alloc<int32_t>(h, &q.out_history, N, nullptr) || pl_alloc<int32_t>(h, &q.bestscore, N, nullptr) ||
        pl_alloc<int32_t>(h, &q.frame, N, nullptr) || pl_alloc<int32_t>(h, &q.best, (size_t)n_utt, nullptr) ||
        pl_alloc<int32_t>(h, &q.renorm, (size_t)n_utt, nullptr) || pl_alloc<int16_t>(h, &h->d_senscr, (size_t)n_utt * n_sen, nullptr) ||
        pl_alloc<int32_t>(h, &h->d_pen, N, nullptr) || cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess) {
        b200_phone_loop_free(h);
        return nullptr;
    }
    q.tp = d_tp; q.senid = d_sid; q.tmatid = d_tm;
    if (b200_phone_loop_start(h) != B200_OK) { b200_phone_loop_free(h); return nullptr; }
    return h;
}

extern "C" void b200_phone_loop_free(b200_phoneloop_t *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (void *p : h->owned) cudaFree(p);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

// phone_loop_search_start (:153-169): hmm_clear + hmm_enter(hmm, 0, -1, 0) for every phone, best_score = 0
extern "C" int b200_phone_loop_start(b200_phoneloop_t *h) {
    if (!h) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(h->device));
    const PhoneLoop &q = h->q;
    const size_t N = (size_t)q.n_utt * q.n_phones;
    std::vector<int32_t> sc(N * q.n_emit, kWorstScore), hi(N * q.n_emit, -1), w(N, kWorstScore), m1(N, -1), z(N, 0), zb(q.n_utt, 0);
    std::fill(sc.begin(), sc.begin() + N, 0);                         // state 0 of every phone entered with score 0
    return b200_phone_loop_set_state(h, sc.data(), hi.data(), w.data(), m1.data(), w.data(), z.data(), zb.data());
}

extern "C" int b200_phone_loop_set_state(b200_phoneloop_t *h, const int32_t *score, const int32_t *history, const int32_t *out_score,
                                         const int32_t *out_history, const int32_t *bestscore, const int32_t *frame, const int32_t *best) {
    if (!h || !score || !history || !out_score || !out_history || !bestscore || !frame || !best) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(h->device));
    const PhoneLoop &q = h->q;
    const size_t N = (size_t)q.n_utt * q.n_phones;
    B200_CUDA_OK(cudaStreamSynchronize(h->st));
    B200_CUDA_OK(cudaMemcpy(q.score, score, N * q.n_emit * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(q.history, history, N * q.n_emit * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(q.out_score, out_score, N * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(q.out_history, out_history, N * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(q.bestscore, bestscore, N * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(q.frame, frame, N * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(q.best, best, (size_t)q.n_utt * 4, cudaMemcpyHostToDevice));
    return B200_OK;
}

extern "C" int b200_phone_loop_get_state(b200_phoneloop_t *h, int32_t *score, int32_t *history, int32_t *out_score, int32_t *out_history,
                                         int32_t *bestscore, int32_t *frame, int32_t *best, int32_t *renorm) {
    if (!h) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(h->device));
    const PhoneLoop &q = h->q;
    const size_t N = (size_t)q.n_utt * q.n_phones;
    B200_CUDA_OK(cudaStreamSynchronize(h->st));
    if (score) B200_CUDA_OK(cudaMemcpy(score, q.score, N * q.n_emit * 4, cudaMemcpyDeviceToHost));
    if (history) B200_CUDA_OK(cudaMemcpy(history, q.history, N * q.n_emit * 4, cudaMemcpyDeviceToHost));
    if (out_score) B200_CUDA_OK(cudaMemcpy(out_score, q.out_score, N * 4, cudaMemcpyDeviceToHost));
    if (out_history) B200_CUDA_OK(cudaMemcpy(out_history, q.out_history, N * 4, cudaMemcpyDeviceToHost));
    if (bestscore) B200_CUDA_OK(cudaMemcpy(bestscore, q.bestscore, N * 4, cudaMemcpyDeviceToHost));
    if (frame) B200_CUDA_OK(cudaMemcpy(frame, q.frame, N * 4, cudaMemcpyDeviceToHost));
    if (best) B200_CUDA_OK(cudaMemcpy(best, q.best, (size_t)q.n_utt * 4, cudaMemcpyDeviceToHost));
    if (renorm) B200_CUDA_OK(cudaMemcpy(renorm, q.renorm, (size_t)q.n_utt * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}

extern "C" int b200_phone_loop_step_dev(b200_phoneloop_t *h, const int16_t *d_senscr, int frame_idx, int32_t *d_pls_pen, void *stream) {
    if (!h || !d_senscr || frame_idx < 0) { set_error("b200_phone_loop_step_dev: bad argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(h->device));
    const PhoneLoop &q = h->q;
    const int threads = ((q.n_phones + 31) / 32) * 32;
    const size_t sh = (size_t)3 * q.n_phones * 4;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->st;
    switch (q.n_emit) {
        case 1: phone_loop_step_kernel<1><<<q.n_utt, threads, sh, st>>>(q, d_senscr, frame_idx, d_pls_pen); break;
        case 2: phone_loop_step_kernel<2><<<q.n_utt, threads, sh, st>>>(q, d_senscr, frame_idx, d_pls_pen); break;
        case 3: phone_loop_step_kernel<3><<<q.n_utt, threads, sh, st>>>(q, d_senscr, frame_idx, d_pls_pen); break;
        case 4: phone_loop_step_kernel<4><<<q.n_utt, threads, sh, st>>>(q, d_senscr, frame_idx, d_pls_pen); break;
        default: phone_loop_step_kernel<5><<<q.n_utt, threads, sh, st>>>(q, d_senscr, frame_idx, d_pls_pen); break;
    }
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_phone_loop_step_host(b200_phoneloop_t *h, const int16_t *senscr, int frame_idx, int32_t *pls_pen, int32_t *best) {
    if (!h || !senscr) { set_error("b200_phone_loop_step_host: bad argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(h->device));
    const PhoneLoop &q = h->q;
    B200_CUDA_OK(cudaMemcpyAsync(h->d_senscr, senscr, (size_t)q.n_utt * q.n_sen * 2, cudaMemcpyHostToDevice, h->st));
    if (int rc = b200_phone_loop_step_dev(h, h->d_senscr, frame_idx, h->d_pen, h->st)) return rc;
    B200_CUDA_OK(cudaStreamSynchronize(h->st));
    if (pls_pen) B200_CUDA_OK(cudaMemcpy(pls_pen, h->d_pen, (size_t)q.n_utt * q.n_phones * 4, cudaMemcpyDeviceToHost));
    if (best) B200_CUDA_OK(cudaMemcpy(best, q.best, (size_t)q.n_utt * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}
