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

__device__ __forceinline__ void eval3(HmmRegs &h, const uint8_t *tp, const int16_t *sen) {
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

__device__ __forceinline__ void eval3_mpx(HmmRegs &h, const uint8_t *tp, const int16_t *sen,
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

__device__ __forceinline__ void eval5(HmmRegs &h, const uint8_t *tp, const int16_t *sen) {
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

__device__ __forceinline__ void eval5_mpx(HmmRegs &h, const uint8_t *tp, const int16_t *sen,
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
template <int NE>
__device__ __forceinline__ void eval_any(HmmRegs &h, const uint8_t *tp, const int16_t *sen, const uint16_t *sseq,
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

template <int NE>
__global__ void __launch_bounds__(kHmmBlock, NE == 3 ? 6 : 4)
hmm_step_kernel(HmmDev c, HmmPop p, const int16_t *__restrict__ senscr_all, HmmFrame *fr) {
    extern __shared__ uint8_t sm_raw[];
    int16_t *s_sen = reinterpret_cast<int16_t *>(sm_raw);
    uint8_t *s_tp = sm_raw + (((size_t)c.n_sen * 2 + 15) & ~(size_t)15);
    const int tid = threadIdx.x;
    const int u = blockIdx.y;
    const int lo = p.utt_off[u], hi = p.utt_off[u + 1];
    if (lo + (int)blockIdx.x * kHmmBlock >= hi) return;   // uniform per block
    const int16_t *senscr = senscr_all + (size_t)u * c.n_sen;
    // stage the frame's senone scores (vectorised) and the transition table
    {
        const int n16 = ((reinterpret_cast<size_t>(senscr) & 15) == 0) ? (c.n_sen * 2) / 16 : 0;
        const int4 *src = reinterpret_cast<const int4 *>(senscr);
        int4 *dst = reinterpret_cast<int4 *>(s_sen);
        for (int i = tid; i < n16; i += kHmmBlock) dst[i] = src[i];
        for (int i = n16 * 8 + tid; i < c.n_sen; i += kHmmBlock) s_sen[i] = senscr[i];
        const int ntp = c.n_tmat * NE * (NE + 1);
        for (int i = tid; i < ntp; i += kHmmBlock) s_tp[i] = c.tp[i];
    }
    __syncthreads();
    int32_t blockbest = kWorstScore;
    const int n = p.n_hmm;
    for (int i = lo + blockIdx.x * kHmmBlock + tid; i < hi; i += gridDim.x * kHmmBlock) {
        HmmRegs h;
#pragma unroll
        for (int s = 0; s < NE; ++s) {
            h.sc[s] = p.score[(size_t)s * n + i];
            h.hi[s] = p.history[(size_t)s * n + i];
            h.sid[s] = p.senid[(size_t)s * n + i];
        }
        h.out_sc = p.out_score[i];
        h.out_hi = p.out_history[i];
        const uint8_t *tp = s_tp + (int)p.tmatid[i] * NE * (NE + 1);
        const bool mpx = p.mpx[i] != 0;
        if (NE == 3) { if (mpx) eval3_mpx(h, tp, s_sen, c.sseq); else eval3(h, tp, s_sen); }
        else if (NE == 5) { if (mpx) eval5_mpx(h, tp, s_sen, c.sseq); else eval5(h, tp, s_sen); }
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
    }
    for (int o = 16; o > 0; o >>= 1) blockbest = max(blockbest, __shfl_xor_sync(0xffffffffu, blockbest, o));
    __shared__ int32_t s_best[kHmmBlock / 32];
    if ((tid & 31) == 0) s_best[tid >> 5] = blockbest;
    __syncthreads();
    if (tid == 0) {
        int32_t b = s_best[0];
        for (int w = 1; w < kHmmBlock / 32; ++w) b = max(b, s_best[w]);
        atomicMax(&fr[u].best, b);
    }
}

__global__ void hmm_frame_init_kernel(HmmFrame *fr, int n_utt, uint32_t *mask, int n_mask_words) {
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int i = i0; i < n_mask_words; i += stride) mask[i] = 0u;
    for (int i = i0; i < n_utt; i += stride) { fr[i].best = kWorstScore; fr[i].n_keep = 0; }
}

// Pass 1 of the order-preserving compaction: keep flag + per-tile count.  A
// tile is kTileIters * kHmmBlock consecutive HMMs of one utterance (few, fat
// CTAs: the scan below and the mask merge of pass 3 scale with the tile count).
// kTileIters = 8 for big populations, 1 when that would leave most SMs idle.
template <int kTileIters>
__global__ void __launch_bounds__(kHmmBlock)
hmm_beam_flag_kernel(HmmPop p, const HmmFrame *fr, int32_t beam, uint8_t *keep, int32_t *block_count) {
    __shared__ int32_t s_cnt[kHmmBlock / 32];
    constexpr int kTile = kTileIters * kHmmBlock;
    const int u = blockIdx.y, tid = threadIdx.x;
    const int hi = p.utt_off[u + 1];
    const int base = p.utt_off[u] + blockIdx.x * kTile + tid;
    const int32_t thresh = fr[u].best + beam;
    int32_t bs[kTileIters];
#pragma unroll
    for (int j = 0; j < kTileIters; ++j) {
        const int i = base + j * kHmmBlock;
        bs[j] = i < hi ? p.bestscore[i] : (int32_t)0x80000000;
    }
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < kTileIters; ++j) {
        const int i = base + j * kHmmBlock;
        const bool k = i < hi && BT(bs[j], thresh);
        if (i < hi) keep[i] = k ? 1 : 0;
        cnt += k ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((tid & 31) == 0) s_cnt[tid >> 5] = cnt;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < kHmmBlock / 32; ++w) t += s_cnt[w];
        block_count[u * gridDim.x + blockIdx.x] = t;
    }
}

// Pass 2: exclusive scan of the block counts by one block (every thread owns a
// contiguous run of counts: one pass, three barriers); per-utterance survivor
// counts land in fr[u].n_keep, the total in *total.
__global__ void __launch_bounds__(1024)
hmm_scan_kernel(int32_t *block_count, int n_blocks, int blocks_per_utt, HmmFrame *fr, int n_utt, int32_t *total) {
    __shared__ int32_t s_warp[32];
    __shared__ int32_t s_total;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int per = (n_blocks + 1023) / 1024;
    const int b0 = min(n_blocks, tid * per), b1 = min(n_blocks, b0 + per);
    int32_t sum = 0;
    for (int i = b0; i < b1; ++i) sum += block_count[i];
    int32_t x = sum;
    for (int o = 1; o < 32; o <<= 1) { int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
        int32_t ws = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1) { int32_t y = __shfl_up_sync(0xffffffffu, ws, o); if (lane >= o) ws += y; }
        s_warp[lane] = ws;
    }
    __syncthreads();
    const int32_t incl = x + (w ? s_warp[w - 1] : 0);
    int32_t run = incl - sum;
    for (int i = b0; i < b1; ++i) { const int32_t v = block_count[i]; block_count[i] = run; run += v; }
    if (tid == 1023) { s_total = incl; *total = incl; }
    __syncthreads();
    for (int u = tid; u < n_utt; u += 1024) {
        const int32_t b = block_count[u * blocks_per_utt];
        const int32_t e = (u + 1 < n_utt) ? block_count[(u + 1) * blocks_per_utt] : s_total;
        fr[u].n_keep = e - b;
    }
}

// Pass 3: scatter survivors (order preserved) and OR their senones into the
// utterance's active mask (acmod_activate_hmm).  One tile per CTA: the mask is
// accumulated in shared memory and merged into the utterance's mask once.
template <int NE, int kTileIters>
__global__ void __launch_bounds__(kHmmBlock)
hmm_scatter_kernel(HmmDev c, HmmPop p, const uint8_t *keep, const int32_t *block_off,
                   int32_t *keep_idx, uint32_t *mask_all) {
    // one flag BYTE per senone, set with plain stores (every writer stores the same 1, so no
    // atomics and no serialisation of the many survivors that share a mask word), packed
    // into words by ballots at the end
    extern __shared__ uint32_t s_flag_w[];
    uint8_t *s_flag = reinterpret_cast<uint8_t *>(s_flag_w);
    constexpr int kTile = kTileIters * kHmmBlock;
    constexpr int kWarps = kHmmBlock / 32, kCnt = kTileIters * kWarps;   // (row, warp) survivor counts
    static_assert(kCnt <= 64, "offset scan handles two entries per lane");
    __shared__ int32_t s_off[64];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int u = blockIdx.y;
    const int lo = p.utt_off[u], hi = p.utt_off[u + 1];
    if (lo + (int)blockIdx.x * kTile >= hi) return;
    const int n_words = (c.n_sen + 31) / 32;
    uint32_t *mask = mask_all + (size_t)u * n_words;
    for (int k = tid; k < n_words * 8; k += kHmmBlock) s_flag_w[k] = 0u;
    const int off = block_off[u * gridDim.x + blockIdx.x];
    const int n = p.n_hmm;
    const int base = lo + blockIdx.x * kTile + tid;
    // all keep flags of the tile first, then ONE exclusive scan over the (row, warp)
    // counts: the rows below need no barrier between them and their loads overlap
    uint8_t kp[kTileIters];
    unsigned bal[kTileIters];
#pragma unroll
    for (int j = 0; j < kTileIters; ++j) {
        const int i = base + j * kHmmBlock;
        kp[j] = i < hi ? keep[i] : 0;
    }
#pragma unroll
    for (int j = 0; j < kTileIters; ++j) {
        bal[j] = __ballot_sync(0xffffffffu, kp[j] != 0);
        if (lane == 0) s_off[j * kWarps + w] = __popc(bal[j]);
    }
    __syncthreads();
    if (w == 0) {
        const int32_t v0 = lane < kCnt ? s_off[lane] : 0, v1 = lane + 32 < kCnt ? s_off[lane + 32] : 0;
        int32_t x0 = v0, x1 = v1;
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t y0 = __shfl_up_sync(0xffffffffu, x0, o), y1 = __shfl_up_sync(0xffffffffu, x1, o);
            if (lane >= o) { x0 += y0; x1 += y1; }
        }
        const int32_t tot0 = __shfl_sync(0xffffffffu, x0, 31);
        if (lane < kCnt) s_off[lane] = x0 - v0;
        if (lane + 32 < kCnt) s_off[lane + 32] = tot0 + x1 - v1;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kTileIters; ++j) {
        if (!kp[j]) continue;
        const int i = base + j * kHmmBlock;
        keep_idx[off + s_off[j * kWarps + w] + __popc(bal[j] & ((1u << lane) - 1u))] = i;
        const bool mpx = p.mpx[i] != 0;
#pragma unroll
        for (int s = 0; s < NE; ++s) {
            uint32_t id = p.senid[(size_t)s * n + i];
            if (mpx) {
                if (id == B200_BAD_SSID) continue;
                id = c.sseq[(size_t)id * NE + s];
            }
            s_flag[id] = 1;
        }
    }
    __syncthreads();
    for (int kk = w; kk < n_words; kk += kWarps) {      // warp-uniform loop
        const unsigned word = __ballot_sync(0xffffffffu, s_flag[kk * 32 + lane] != 0);
        if (lane == 0 && word) atomicOr(&mask[kk], word);
    }
}

// ------------------------------------------------------------ host launchers
static size_t step_smem(const HmmDev &c) {
    return (((size_t)c.n_sen * 2 + 15) & ~(size_t)15) + (size_t)c.n_tmat * c.n_emit * (c.n_emit + 1) + 16;
}

int hmm_launch_step(const HmmDev &c, const HmmPop &p, const int16_t *d_senscr, int32_t beam,
                    HmmFrame *fr, uint8_t *keep, int32_t *block_count, int32_t *keep_idx,
                    uint32_t *mask, int32_t *total, int do_beam, cudaStream_t st) {
    const int n_words = (c.n_sen + 31) / 32;
    if (p.n_hmm <= 0 || p.n_utt <= 0) return B200_OK;
    const int bpu = (p.max_per_utt + kHmmBlock - 1) / kHmmBlock;   // blocks per utterance
    const int n_mask = n_words * p.n_utt;
    hmm_frame_init_kernel<<<std::max(1, std::min(148, (n_mask + 255) / 256)), 256, 0, st>>>(fr, p.n_utt, mask, n_mask);
    B200_LAUNCH_CHECK();
    const size_t sh = step_smem(c);
    if (sh > 200 * 1024) { set_error("hmm step needs %zu B shared memory", sh); return B200_ERR_UNSUP; }
    static AttrOnce attr;
    if (c.n_emit < 1 || c.n_emit > 5) { set_error("n_emit_state %d outside 1..5 (HMM_MAX_NSTATE)", c.n_emit); return B200_ERR_UNSUP; }
#define B200_HMM_NE(...)                                   \
    switch (c.n_emit) {                                    \
    case 1: { constexpr int NE = 1; __VA_ARGS__; } break;  \
    case 2: { constexpr int NE = 2; __VA_ARGS__; } break;  \
    case 3: { constexpr int NE = 3; __VA_ARGS__; } break;  \
    case 4: { constexpr int NE = 4; __VA_ARGS__; } break;  \
    default: { constexpr int NE = 5; __VA_ARGS__; } break; \
    }
    if (attr.need()) {
        for (int ne = 1; ne <= 5; ++ne) {
            cudaError_t e = cudaSuccess;
            switch (ne) {
            case 1: e = cudaFuncSetAttribute(hmm_step_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); break;
            case 2: e = cudaFuncSetAttribute(hmm_step_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); break;
            case 3: e = cudaFuncSetAttribute(hmm_step_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); break;
            case 4: e = cudaFuncSetAttribute(hmm_step_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); break;
            default: e = cudaFuncSetAttribute(hmm_step_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); break;
            }
            B200_CUDA_OK(e);
        }
    }
    // exactly one resident wave of CTAs (occupancy x SM count), split evenly over the
    // utterances; each CTA strides its utterance's range
    static size_t occ_sh[6] = {0, 0, 0, 0, 0, 0};
    static int occ_wave[6] = {0, 0, 0, 0, 0, 0};   // CTAs in one resident wave, per kernel flavour
    const int fl = c.n_emit;
    if (occ_sh[fl] != sh || occ_wave[fl] == 0) {
        int per_sm = 0, n_sm = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        B200_HMM_NE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hmm_step_kernel<NE>, kHmmBlock, sh));
        occ_wave[fl] = std::max(1, per_sm) * n_sm;
        occ_sh[fl] = sh;
    }
    int gx = std::max(1, std::min(bpu, occ_wave[fl] / p.n_utt));
    dim3 grid(gx, p.n_utt);
    B200_HMM_NE(hmm_step_kernel<NE><<<grid, kHmmBlock, sh, st>>>(c, p, d_senscr, fr));
    B200_LAUNCH_CHECK();
    if (!do_beam) return B200_OK;
    // fat tiles (8 x 256 HMMs per CTA) when that still fills the machine twice over, else one HMM per thread
    const bool fat = (long long)((p.max_per_utt + 8 * kHmmBlock - 1) / (8 * kHmmBlock)) * p.n_utt >= 2 * 148;
    const int tile = (fat ? 8 : 1) * kHmmBlock;
    const int tpu = (p.max_per_utt + tile - 1) / tile;            // tiles per utterance (<= bpu)
    dim3 g2(tpu, p.n_utt);
    if (fat) hmm_beam_flag_kernel<8><<<g2, kHmmBlock, 0, st>>>(p, fr, beam, keep, block_count);
    else hmm_beam_flag_kernel<1><<<g2, kHmmBlock, 0, st>>>(p, fr, beam, keep, block_count);
    B200_LAUNCH_CHECK();
    hmm_scan_kernel<<<1, 1024, 0, st>>>(block_count, tpu * p.n_utt, tpu, fr, p.n_utt, total);
    B200_LAUNCH_CHECK();
    const size_t msh = (size_t)n_words * 32;     // one flag byte per senone (<= 64 KB)
    if (msh > 48 * 1024) {   // more than 12 288 senones: the flag bytes need the opt-in shared-memory size
        cudaError_t e = cudaSuccess;
        B200_HMM_NE(e = fat ? cudaFuncSetAttribute(hmm_scatter_kernel<NE, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024)
                            : cudaFuncSetAttribute(hmm_scatter_kernel<NE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024));
        B200_CUDA_OK(e);
    }
    if (fat) { B200_HMM_NE(hmm_scatter_kernel<NE, 8><<<g2, kHmmBlock, msh, st>>>(c, p, keep, block_count, keep_idx, mask)); }
    else { B200_HMM_NE(hmm_scatter_kernel<NE, 1><<<g2, kHmmBlock, msh, st>>>(c, p, keep, block_count, keep_idx, mask)); }
#undef B200_HMM_NE
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
hmm_clear_pruned_kernel(HmmPop p, int n_emit, const uint8_t *__restrict__ keep) {
    const int i = blockIdx.x * kHmmBlock + threadIdx.x;
    if (i >= p.n_hmm || keep[i]) return;
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

int hmm_launch_clear_pruned(const HmmPop &p, int n_emit, const uint8_t *keep, cudaStream_t st) {
    if (p.n_hmm <= 0) return B200_OK;
    hmm_clear_pruned_kernel<<<(p.n_hmm + kHmmBlock - 1) / kHmmBlock, kHmmBlock, 0, st>>>(p, n_emit, keep);
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

}  // namespace b200
