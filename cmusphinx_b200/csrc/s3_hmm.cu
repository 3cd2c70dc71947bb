// s3_hmm.cu -- sphinx3's flavour of hmm_vit_eval (sphinx3/src/libs3decoder/libam/hmm.c:285-873)
// for a batch of HMMs: int32 transition log-probabilities and int32 senone scores that are ADDED
// (sphinx3 scores are plain logs, higher is better), WORST_SCORE = S3_LOGPROB_ZERO = 0xc8000000
// (s3types.h:192), int32 senone-sequence ids with -1 = none (hmm.h:220-226), senone ids always
// through ctx->sseq[ssid][state] (non-mpx HMMs too).  The five evaluators differ from each other
// -- and from the pocketsphinx copy (hmm_kernels.cu) -- in which states they guard, clamp and fold
// into the best score; each is restated with its own guards:
//   s3_eval5      hmm.c:285-414   non-mpx 5-state: exit / state 4 / state 3 updated only when the
//                                 state two below is alive (s3, s2, s1 > WORST_SCORE)
//   s3_eval5_mpx  hmm.c:416-586   missing ssid == -1, `!= WORST_SCORE` guards, state 0 unguarded
//   s3_eval3      hmm.c:592-671   t2 starts at INT_MIN and only exists when tp(1,3) / tp(0,2) do
//   s3_eval3_mpx  hmm.c:673-774   scores clamped before use, t2 (= s1 + tp(1,3)) can leak into state 2
//   s3_eval_any   hmm.c:776-850   generic topology, unclamped stores
// One thread per HMM; state-major SoA like the pocketsphinx population.
#include "dev_common.cuh"

#include <climits>
#include <vector>

namespace b200 {

namespace {

constexpr int32_t kS3Worst = (int32_t)0xc8000000;

struct S3Ctx {
    int n_emit, n_tmat, n_sseq, n_sen;
    const int32_t *tp;        // [n_tmat][n_emit][n_emit + 1]
    const int16_t *sseq;      // [n_sseq][n_emit]  (s3senid_t)
};

struct S3Regs {
    int32_t sc[5], hi[5], ssid[5];
    int32_t out_sc, out_hi, best;
};

#define TPV(i, j) tp[(i) * (NE + 1) + (j)]

template <int NE>
__device__ __forceinline__ int32_t sen_of(const S3Ctx &c, const int32_t *sen, int32_t ssid, int st) {
    return sen[c.sseq[(size_t)ssid * NE + st]];
}

__device__ __forceinline__ void s3_eval5(S3Regs &h, const S3Ctx &c, const int32_t *tp, const int32_t *sen) {
    constexpr int NE = 5;
    const int32_t ss = h.ssid[0];
    int32_t s5, s4, s3, s2, s1, s0, t2, t1, t0, best = kS3Worst;
    s4 = h.sc[4] + sen_of<NE>(c, sen, ss, 4);
    s3 = h.sc[3] + sen_of<NE>(c, sen, ss, 3);
    if (s3 > kS3Worst) {
        t1 = s4 + TPV(4, 5);
        t2 = s3 + TPV(3, 5);
        if (t1 > t2) { s5 = t1; h.out_hi = h.hi[4]; } else { s5 = t2; h.out_hi = h.hi[3]; }
        if (s5 < kS3Worst) s5 = kS3Worst;
        h.out_sc = s5;
        best = s5;
    }
    s2 = h.sc[2] + sen_of<NE>(c, sen, ss, 2);
    if (s2 > kS3Worst) {
        t0 = s4 + TPV(4, 4); t1 = s3 + TPV(3, 4); t2 = s2 + TPV(2, 4);
        if (t0 > t1) { if (t2 > t0) { s4 = t2; h.hi[4] = h.hi[2]; } else s4 = t0; }
        else { if (t2 > t1) { s4 = t2; h.hi[4] = h.hi[2]; } else { s4 = t1; h.hi[4] = h.hi[3]; } }
        if (s4 < kS3Worst) s4 = kS3Worst;
        if (s4 > best) best = s4;
        h.sc[4] = s4;
    }
    s1 = h.sc[1] + sen_of<NE>(c, sen, ss, 1);
    if (s1 > kS3Worst) {
        t0 = s3 + TPV(3, 3); t1 = s2 + TPV(2, 3); t2 = s1 + TPV(1, 3);
        if (t0 > t1) { if (t2 > t0) { s3 = t2; h.hi[3] = h.hi[1]; } else s3 = t0; }
        else { if (t2 > t1) { s3 = t2; h.hi[3] = h.hi[1]; } else { s3 = t1; h.hi[3] = h.hi[2]; } }
        if (s3 < kS3Worst) s3 = kS3Worst;
        if (s3 > best) best = s3;
        h.sc[3] = s3;
    }
    s0 = h.sc[0] + sen_of<NE>(c, sen, ss, 0);
    t0 = s2 + TPV(2, 2); t1 = s1 + TPV(1, 2); t2 = s0 + TPV(0, 2);
    if (t0 > t1) { if (t2 > t0) { s2 = t2; h.hi[2] = h.hi[0]; } else s2 = t0; }
    else { if (t2 > t1) { s2 = t2; h.hi[2] = h.hi[0]; } else { s2 = t1; h.hi[2] = h.hi[1]; } }
    if (s2 < kS3Worst) s2 = kS3Worst;
    if (s2 > best) best = s2;
    h.sc[2] = s2;
    t0 = s1 + TPV(1, 1); t1 = s0 + TPV(0, 1);
    if (t0 > t1) s1 = t0; else { s1 = t1; h.hi[1] = h.hi[0]; }
    if (s1 < kS3Worst) s1 = kS3Worst;
    if (s1 > best) best = s1;
    h.sc[1] = s1;
    s0 = s0 + TPV(0, 0);
    if (s0 < kS3Worst) s0 = kS3Worst;
    if (s0 > best) best = s0;
    h.sc[0] = s0;
    h.best = best;
}

__device__ __forceinline__ void s3_eval5_mpx(S3Regs &h, const S3Ctx &c, const int32_t *tp, const int32_t *sen) {
    constexpr int NE = 5;
    int32_t *ssid = h.ssid;
    int32_t s5, s4, s3, s2, s1, s0, t2, t1, t0, best;
    if (ssid[4] == -1) s4 = t1 = kS3Worst;
    else { s4 = h.sc[4] + sen_of<NE>(c, sen, ssid[4], 4); t1 = s4 + TPV(4, 5); }
    if (ssid[3] == -1) s3 = t2 = kS3Worst;
    else { s3 = h.sc[3] + sen_of<NE>(c, sen, ssid[3], 3); t2 = s3 + TPV(3, 5); }
    if (t1 > t2) { s5 = t1; h.out_hi = h.hi[4]; } else { s5 = t2; h.out_hi = h.hi[3]; }
    if (s5 < kS3Worst) s5 = kS3Worst;
    h.out_sc = s5;
    best = s5;
    if (ssid[2] == -1) s2 = t2 = kS3Worst;
    else { s2 = h.sc[2] + sen_of<NE>(c, sen, ssid[2], 2); t2 = s2 + TPV(2, 4); }
    t0 = t1 = kS3Worst;
    if (s4 != kS3Worst) t0 = s4 + TPV(4, 4);
    if (s3 != kS3Worst) t1 = s3 + TPV(3, 4);
    if (t0 > t1) { if (t2 > t0) { s4 = t2; h.hi[4] = h.hi[2]; ssid[4] = ssid[2]; } else s4 = t0; }
    else { if (t2 > t1) { s4 = t2; h.hi[4] = h.hi[2]; ssid[4] = ssid[2]; } else { s4 = t1; h.hi[4] = h.hi[3]; ssid[4] = ssid[3]; } }
    if (s4 < kS3Worst) s4 = kS3Worst;
    if (s4 > best) best = s4;
    h.sc[4] = s4;
    if (ssid[1] == -1) s1 = t2 = kS3Worst;
    else { s1 = h.sc[1] + sen_of<NE>(c, sen, ssid[1], 1); t2 = s1 + TPV(1, 3); }
    t0 = t1 = kS3Worst;
    if (s3 != kS3Worst) t0 = s3 + TPV(3, 3);
    if (s2 != kS3Worst) t1 = s2 + TPV(2, 3);
    if (t0 > t1) { if (t2 > t0) { s3 = t2; h.hi[3] = h.hi[1]; ssid[3] = ssid[1]; } else s3 = t0; }
    else { if (t2 > t1) { s3 = t2; h.hi[3] = h.hi[1]; ssid[3] = ssid[1]; } else { s3 = t1; h.hi[3] = h.hi[2]; ssid[3] = ssid[2]; } }
    if (s3 < kS3Worst) s3 = kS3Worst;
    if (s3 > best) best = s3;
    h.sc[3] = s3;
    s0 = h.sc[0] + sen_of<NE>(c, sen, ssid[0], 0);
    t0 = t1 = kS3Worst;
    if (s2 != kS3Worst) t0 = s2 + TPV(2, 2);
    if (s1 != kS3Worst) t1 = s1 + TPV(1, 2);
    t2 = s0 + TPV(0, 2);
    if (t0 > t1) { if (t2 > t0) { s2 = t2; h.hi[2] = h.hi[0]; ssid[2] = ssid[0]; } else s2 = t0; }
    else { if (t2 > t1) { s2 = t2; h.hi[2] = h.hi[0]; ssid[2] = ssid[0]; } else { s2 = t1; h.hi[2] = h.hi[1]; ssid[2] = ssid[1]; } }
    if (s2 < kS3Worst) s2 = kS3Worst;
    if (s2 > best) best = s2;
    h.sc[2] = s2;
    t0 = kS3Worst;
    if (s1 != kS3Worst) t0 = s1 + TPV(1, 1);
    t1 = s0 + TPV(0, 1);
    if (t0 > t1) s1 = t0; else { s1 = t1; h.hi[1] = h.hi[0]; ssid[1] = ssid[0]; }
    if (s1 < kS3Worst) s1 = kS3Worst;
    if (s1 > best) best = s1;
    h.sc[1] = s1;
    s0 += TPV(0, 0);
    if (s0 < kS3Worst) s0 = kS3Worst;
    if (s0 > best) best = s0;
    h.sc[0] = s0;
    h.best = best;
}

__device__ __forceinline__ void s3_eval3(S3Regs &h, const S3Ctx &c, const int32_t *tp, const int32_t *sen) {
    constexpr int NE = 3;
    const int32_t ss = h.ssid[0];
    int32_t s3, s2, s1, s0, t2, t1, t0, best;
    s2 = h.sc[2] + sen_of<NE>(c, sen, ss, 2);
    s1 = h.sc[1] + sen_of<NE>(c, sen, ss, 1);
    s0 = h.sc[0] + sen_of<NE>(c, sen, ss, 0);
    t0 = t1 = best = kS3Worst;
    t2 = INT_MIN;
    if (s2 > kS3Worst) { t1 = s2 + TPV(2, 3); t0 = s2 + TPV(2, 2); }
    if (s1 > kS3Worst && TPV(1, 3) > kS3Worst) t2 = s1 + TPV(1, 3);
    if (t1 > t2) { s3 = t1; h.out_hi = h.hi[2]; } else { s3 = t2; h.out_hi = h.hi[1]; }
    if (s3 < kS3Worst) s3 = kS3Worst;
    h.out_sc = s3;
    best = s3;
    t1 = t2 = kS3Worst;
    if (s1 > kS3Worst) t1 = s1 + TPV(1, 2);
    if (TPV(0, 2) > kS3Worst) t2 = s0 + TPV(0, 2);
    if (t0 > t1) { if (t2 > t0) { s2 = t2; h.hi[2] = h.hi[0]; } else s2 = t0; }
    else { if (t2 > t1) { s2 = t2; h.hi[2] = h.hi[0]; } else { s2 = t1; h.hi[2] = h.hi[1]; } }
    if (s2 < kS3Worst) s2 = kS3Worst;
    if (s2 > best) best = s2;
    h.sc[2] = s2;
    t0 = t1 = kS3Worst;
    if (s1 > kS3Worst) t0 = s1 + TPV(1, 1);
    if (s0 > kS3Worst) t1 = s0 + TPV(0, 1);
    if (t0 > t1) s1 = t0; else { s1 = t1; h.hi[1] = h.hi[0]; }
    if (s1 < kS3Worst) s1 = kS3Worst;
    if (s1 > best) best = s1;
    h.sc[1] = s1;
    s0 = s0 + TPV(0, 0);
    if (s0 < kS3Worst) s0 = kS3Worst;
    if (s0 > best) best = s0;
    h.sc[0] = s0;
    h.best = best;
}

__device__ __forceinline__ void s3_eval3_mpx(S3Regs &h, const S3Ctx &c, const int32_t *tp, const int32_t *sen) {
    constexpr int NE = 3;
    int32_t *ssid = h.ssid;
    int32_t s3, s2, s1, s0, t2, t1, t0, best;
    t2 = INT_MIN;
    if (ssid[2] == -1) s2 = t1 = kS3Worst;
    else {
        s2 = h.sc[2] + sen_of<NE>(c, sen, ssid[2], 2);
        if (s2 < kS3Worst) s2 = kS3Worst;
        t1 = s2 + TPV(2, 3);
    }
    if (ssid[1] == -1) s1 = kS3Worst;
    else {
        s1 = h.sc[1] + sen_of<NE>(c, sen, ssid[1], 1);
        if (s1 < kS3Worst) s1 = kS3Worst;
        t2 = s1 + TPV(1, 3);
    }
    if (t1 > t2) { s3 = t1; h.out_hi = h.hi[2]; } else { s3 = t2; h.out_hi = h.hi[1]; }
    if (s3 < kS3Worst) s3 = kS3Worst;
    h.out_sc = s3;
    best = s3;
    s0 = h.sc[0] + sen_of<NE>(c, sen, ssid[0], 0);
    if (s0 < kS3Worst) s0 = kS3Worst;
    t0 = t1 = kS3Worst;
    if (s2 != kS3Worst) t0 = s2 + TPV(2, 2);
    if (s1 != kS3Worst) t1 = s1 + TPV(1, 2);
    if (TPV(0, 2) > kS3Worst) t2 = s0 + TPV(0, 2);
    if (t0 > t1) { if (t2 > t0) { s2 = t2; h.hi[2] = h.hi[0]; ssid[2] = ssid[0]; } else s2 = t0; }
    else { if (t2 > t1) { s2 = t2; h.hi[2] = h.hi[0]; ssid[2] = ssid[0]; } else { s2 = t1; h.hi[2] = h.hi[1]; ssid[2] = ssid[1]; } }
    if (s2 < kS3Worst) s2 = kS3Worst;
    if (s2 > best) best = s2;
    h.sc[2] = s2;
    t0 = kS3Worst;
    if (s1 != kS3Worst) t0 = s1 + TPV(1, 1);
    t1 = s0 + TPV(0, 1);
    if (t0 > t1) s1 = t0; else { s1 = t1; h.hi[1] = h.hi[0]; ssid[1] = ssid[0]; }
    if (s1 < kS3Worst) s1 = kS3Worst;
    if (s1 > best) best = s1;
    h.sc[1] = s1;
    s0 += TPV(0, 0);
    if (s0 < kS3Worst) s0 = kS3Worst;
    if (s0 > best) best = s0;
    h.sc[0] = s0;
    h.best = best;
}

template <int NE>
__device__ __forceinline__ void s3_eval_any(S3Regs &h, const S3Ctx &c, const int32_t *tp, const int32_t *sen, bool mpx) {
    int32_t st[NE];
#pragma unroll
    for (int s = 0; s < NE; ++s) {
        const int32_t id = mpx ? h.ssid[s] : h.ssid[0];
        const int32_t ss = id == -1 ? kS3Worst : sen_of<NE>(c, sen, id, s);
        int32_t v = h.sc[s] + ss;
        if (s > 0 && v < kS3Worst) v = kS3Worst;
        st[s] = v;
    }
    int32_t scr = kS3Worst, bestfrom = -1;
#pragma unroll
    for (int f = NE - 1; f >= 0; --f) {
        const int32_t t = TPV(f, NE);
        if (t > kS3Worst && st[f] + t > scr) { scr = st[f] + t; bestfrom = f; }
    }
    h.out_sc = scr;
    int32_t hi_new[NE], ss_new[NE], sc_new[NE];
    if (bestfrom >= 0) h.out_hi = h.hi[bestfrom];
    int32_t best = scr;
#pragma unroll
    for (int s = 0; s < NE; ++s) { hi_new[s] = h.hi[s]; ss_new[s] = h.ssid[s]; }
    // states are updated from the last to the first and only read lower-numbered ones:
    // in-place updates never feed a later read, so the copies are just for clarity
#pragma unroll
    for (int to = NE - 1; to >= 0; --to) {
        const int32_t tt = TPV(to, to);
        scr = tt > kS3Worst ? st[to] + tt : kS3Worst;
        bestfrom = -1;
#pragma unroll
        for (int f = to - 1; f >= 0; --f) {
            const int32_t t = TPV(f, to);
            if (t > kS3Worst && st[f] + t > scr) { scr = st[f] + t; bestfrom = f; }
        }
        sc_new[to] = scr;
        if (bestfrom >= 0) { hi_new[to] = h.hi[bestfrom]; if (mpx) ss_new[to] = h.ssid[bestfrom]; }
        if (best < scr) best = scr;
    }
#pragma unroll
    for (int s = 0; s < NE; ++s) { h.sc[s] = sc_new[s]; h.hi[s] = hi_new[s]; h.ssid[s] = ss_new[s]; }
    h.best = best;
}
#undef TPV

struct S3Pop {
    int n_hmm;
    int32_t *score, *history, *ssid;       // [n_emit][n_hmm]
    int32_t *out_score, *out_history, *bestscore;
    const int32_t *tmatid;
    const uint8_t *mpx;
};

template <int NE>
__global__ void __launch_bounds__(256)
s3_hmm_eval_kernel(S3Ctx c, S3Pop p, const int32_t *__restrict__ sen, int32_t *frame_best) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    int32_t best = kS3Worst;
    if (i < p.n_hmm) {
        const int n = p.n_hmm;
        S3Regs h;
#pragma unroll
        for (int s = 0; s < NE; ++s) {
            h.sc[s] = p.score[(size_t)s * n + i];
            h.hi[s] = p.history[(size_t)s * n + i];
            h.ssid[s] = p.ssid[(size_t)s * n + i];
        }
        h.out_sc = p.out_score[i];
        h.out_hi = p.out_history[i];
        const int32_t *tp = c.tp + (size_t)p.tmatid[i] * NE * (NE + 1);
        const bool mpx = p.mpx[i] != 0;
        if (NE == 3) { if (mpx) s3_eval3_mpx(h, c, tp, sen); else s3_eval3(h, c, tp, sen); }
        else if (NE == 5) { if (mpx) s3_eval5_mpx(h, c, tp, sen); else s3_eval5(h, c, tp, sen); }
        else s3_eval_any<NE>(h, c, tp, sen, mpx);
#pragma unroll
        for (int s = 0; s < NE; ++s) {
            p.score[(size_t)s * n + i] = h.sc[s];
            p.history[(size_t)s * n + i] = h.hi[s];
            if (mpx) p.ssid[(size_t)s * n + i] = h.ssid[s];
        }
        p.out_score[i] = h.out_sc;
        p.out_history[i] = h.out_hi;
        p.bestscore[i] = h.best;
        best = h.best;
    }
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) atomicMax(frame_best, best);
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200_s3hmm_eval_host(int n_emit, const int32_t *tp, int n_tmat, const int16_t *sseq, int n_sseq, int n_sen,
                                    b200_s3hmm_soa_t *h, const int32_t *senscr, int n_frames, int32_t *best_out,
                                    int device) {
    if (n_emit < 1 || n_emit > 5 || !tp || n_tmat <= 0 || !sseq || n_sseq <= 0 || n_sen <= 0 || !h || n_frames < 0 ||
        (n_frames > 0 && !senscr)) { set_error("b200_s3hmm_eval_host: bad argument"); return B200_ERR_ARG; }
    const int n = h->n_hmm;
    if (n < 0 || (n > 0 && (!h->score || !h->history || !h->ssid || !h->out_score || !h->out_history || !h->bestscore ||
                            !h->tmatid || !h->mpx))) { set_error("b200_s3hmm_eval_host: null SoA field"); return B200_ERR_ARG; }
    const bool fixed_topo = n_emit == 3 || n_emit == 5;     // the specialised evaluators read state 0's senone unguarded
    for (int i = 0; i < n; ++i) {
        if (h->tmatid[i] < 0 || h->tmatid[i] >= n_tmat) { set_error("tmatid[%d]=%d out of range", i, h->tmatid[i]); return B200_ERR_ARG; }
        for (int s = 0; s < (h->mpx[i] ? n_emit : 1); ++s) {
            const int32_t id = h->ssid[(size_t)s * n + i];
            const int32_t lowest = (fixed_topo && (s == 0 || !h->mpx[i])) ? 0 : -1;
            if (id < lowest || id >= n_sseq) { set_error("ssid[%d][%d]=%d out of range", s, i, id); return B200_ERR_ARG; }
        }
    }
    for (size_t k = 0; k < (size_t)n_sseq * n_emit; ++k)
        if (sseq[k] < 0 || sseq[k] >= n_sen) { set_error("sseq entry %d out of range", (int)sseq[k]); return B200_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: libb200sphinx has no CPU fallback"); return B200_ERR_CUDA; }
    if (device < 0 || device >= ndev) { set_error("bad device %d", device); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(device));
    if (n == 0) { for (int f = 0; f < n_frames; ++f) if (best_out) best_out[f] = kS3Worst; return B200_OK; }
    struct Buf { void *p = nullptr; ~Buf() { cudaFree(p); } };
    Buf bt, bs, bsen, b1, b2, b3, b4, b5, b6, b7, b8, bb;
    auto up = [](Buf &b, const void *src, size_t bytes) {
        if (cudaMalloc(&b.p, bytes) != cudaSuccess) return false;
        return !src || cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    const size_t N = (size_t)n, ne = (size_t)n_emit;
    if (!up(bt, tp, (size_t)n_tmat * ne * (ne + 1) * 4) || !up(bs, sseq, (size_t)n_sseq * ne * 2) ||
        !up(bsen, senscr, (size_t)std::max(n_frames, 1) * n_sen * 4) || !up(b1, h->score, N * ne * 4) ||
        !up(b2, h->history, N * ne * 4) || !up(b3, h->ssid, N * ne * 4) || !up(b4, h->out_score, N * 4) ||
        !up(b5, h->out_history, N * 4) || !up(b6, h->bestscore, N * 4) || !up(b7, h->tmatid, N * 4) || !up(b8, h->mpx, N) ||
        !up(bb, nullptr, (size_t)std::max(n_frames, 1) * 4)) {
        set_error("b200_s3hmm_eval_host: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        return B200_ERR_CUDA;
    }
    S3Ctx c{n_emit, n_tmat, n_sseq, n_sen, (const int32_t *)bt.p, (const int16_t *)bs.p};
    S3Pop p{n, (int32_t *)b1.p, (int32_t *)b2.p, (int32_t *)b3.p, (int32_t *)b4.p, (int32_t *)b5.p, (int32_t *)b6.p,
            (const int32_t *)b7.p, (const uint8_t *)b8.p};
    std::vector<int32_t> init(std::max(n_frames, 1), kS3Worst);
    B200_CUDA_OK(cudaMemcpy(bb.p, init.data(), init.size() * 4, cudaMemcpyHostToDevice));
    const int grid = (n + 255) / 256;
    for (int f = 0; f < n_frames; ++f) {
        const int32_t *sen = (const int32_t *)bsen.p + (size_t)f * n_sen;
        int32_t *fb = (int32_t *)bb.p + f;
        switch (n_emit) {
            case 1: s3_hmm_eval_kernel<1><<<grid, 256>>>(c, p, sen, fb); break;
            case 2: s3_hmm_eval_kernel<2><<<grid, 256>>>(c, p, sen, fb); break;
            case 3: s3_hmm_eval_kernel<3><<<grid, 256>>>(c, p, sen, fb); break;
            case 4: s3_hmm_eval_kernel<4><<<grid, 256>>>(c, p, sen, fb); break;
            default: s3_hmm_eval_kernel<5><<<grid, 256>>>(c, p, sen, fb); break;
        }
        B200_LAUNCH_CHECK();
    }
    B200_CUDA_OK(cudaDeviceSynchronize());
    if (best_out && n_frames > 0) B200_CUDA_OK(cudaMemcpy(best_out, bb.p, (size_t)n_frames * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(h->score, b1.p, N * ne * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(h->history, b2.p, N * ne * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(h->ssid, b3.p, N * ne * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(h->out_score, b4.p, N * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(h->out_history, b5.p, N * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(h->bestscore, b6.p, N * 4, cudaMemcpyDeviceToHost));
    return B200_OK;
}
