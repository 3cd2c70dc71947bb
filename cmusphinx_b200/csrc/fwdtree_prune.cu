// fwdtree_prune.cu -- the prune / phone-transition stage of the forward tree search on the GPU
// (SURVEY.md section 8(f)-1): prune_root_chan followed by prune_nonroot_chan,
// pocketsphinx/src/libpocketsphinx/ngram_search_fwdtree.c:714-790 and :792-869, with
// hmm_enter (hmm.c:197-203) and hmm_clear_scores (hmm.c:169-181), for a batch of utterances
// that share one lexical tree.
//
// The reference walks the root channels in index order and then the frame's active list in
// LIST order, and its result depends on that order in three ways:
//   (1) the next frame's active list is the sequence of its appends (`*(nacl++) = hmm`): a
//       channel is appended by whichever of "it survives the beam" (at its own list position) and
//       "its parent enters it" (at the parent's position, in `alt` order behind the parent's own
//       append) comes first;
//   (2) a channel that fails the beam is cleared (hmm_clear_scores) unless its parent entered it
//       EARLIER in the walk; if the parent comes later, the entry test
//       `pl_newphone_score BETTER_THAN hmm_in_score` sees the cleared WORST_SCORE, always passes,
//       and the channel restarts with only its state 0 alive instead of keeping states 1.. ;
//   (3) the last-phone candidates (lastphn_cand) are appended in walk order.
// Every non-root channel has exactly one parent (the lexical tree IS a tree; checked at
// b200_chantree_create), so each channel sees at most two events per frame -- its own beam test at
// time tau(c) and its parent's entry attempt at time tau(parent) -- and the outcome of both is a
// closed-form function of the PRE-prune state and of which of the two times is smaller.  "Time" is
// the position in the walk: root r -> r, active-list position i -> n_root + i.  The kernel
// therefore never walks anything in sequence:
//   phase 0  tau[] for the frame's roots and active-list members;
//   phase 1  one thread per walk element decides, from unmodified state only: keep / clear, whether
//            it appends itself, and for each child whether it enters it and whether that entry
//            appends the child (decision byte per child) -- and counts its appends and candidates;
//   scan     exclusive prefix sums of the two counts in walk order = the reference's append order;
//   phase 2  writes: hmm_enter of the entered children, frame stamps, hmm_clear_scores (minus
//            state 0 when the parent's later entry owns it), the next list and the candidates at
//            their scanned offsets.
// One CTA per utterance (the walk of one utterance is a few thousand elements; a batch of
// utterances fills the GPU), HBM traffic = the touched rows only.  Bit-exact, list order included,
// against the reference on every frame of real decodes (tests/test_fwdtree_prune.py).
#include "dev_common.cuh"

#include <algorithm>
#include <vector>

namespace b200 {

namespace {

constexpr int32_t kWorst = (int32_t)0xE0000000;   // PS/hmm.h:74
constexpr int kPruneThreads = 1024;

struct PruneTree {
    int n_root, n_chan, n_ci, n_edge, n_pw;
    const int32_t *child_off, *child, *ciphone, *pw_off, *pw_wid, *pw_lp, *parent;
};

struct PruneArgs {
    int ne; long stride;                 // state s of channel c of utterance u: score[s * stride + u * n_chan + c]
    int32_t *score, *history, *out_score, *out_history, *bestscore, *frame;
    const int32_t *par;                  // [n_utt][8]
    const int32_t *pls_pen;              // [n_utt][n_ci] or null
    const int32_t *acl, *n_act; int list_cap;
    int32_t *nacl, *n_nacl;
    int32_t *cand, *n_cand; int cand_cap;
    // scratch
    int32_t *tau;                        // [n_utt][n_chan], -1 outside the call
    uint8_t *dec;                        // [n_utt][n_chan]: bit 0 the parent enters the channel, bit 1 and that appends it
    unsigned long long *ecount;          // [n_utt][n_root + list_cap]: appends | candidates << 32, then their exclusive scan
    uint8_t *eflag;                      // [n_utt][n_root + list_cap]
};

enum : uint8_t { kKeep = 1, kSelfAppend = 2, kClear = 4, kClearSkip0 = 8 };

struct Thr { int32_t fi, thresh, newphone, lastphn, pip, nwpen; bool pls; };

// The parent's attempt to enter child c (fwdtree.c:746-757 / 824-840).  All reads are of the
// pre-prune state.  p_tau = the parent's walk position; returns bit 0 entered, bit 1 appended by it.
__device__ __forceinline__ unsigned edge_eval(const Thr &t, int32_t nps, int32_t pen, int32_t p_tau, int32_t c_tau,
                                              int32_t c_frame, int32_t c_in, int32_t c_best, int32_t &pl) {
    pl = nps + pen;
    if (!(t.pls || nps > t.newphone) || !(pl > t.newphone)) return 0u;
    if (c_tau < 0) {                                   // not on this frame's list
        const bool e = c_frame < t.fi || pl > c_in;
        return e ? 3u : 0u;
    }
    if (p_tau < c_tau) {                               // the parent comes first: the child is as hmm_vit_eval left it
        const bool e = c_frame < t.fi || pl > c_in;
        return e ? 3u : 0u;
    }
    if (c_best > t.thresh) {                           // the child kept itself (and appended itself) earlier
        return pl > c_in ? 1u : 0u;
    }
    return pl > kWorst ? 3u : 0u;                      // the child was cleared earlier: in-score is WORST_SCORE
}

__global__ void __launch_bounds__(kPruneThreads) fwdtree_prune_kernel(PruneTree tr, PruneArgs a) {
    const int u = blockIdx.x, tid = threadIdx.x;
    const size_t base = (size_t)u * tr.n_chan;
    int32_t *score0 = a.score + base, *hist0 = a.history + base;
    int32_t *out_score = a.out_score + base, *out_hist = a.out_history + base, *best = a.bestscore + base, *frame = a.frame + base;
    const int32_t *par = a.par + (size_t)u * 8;
    const int32_t *pen = a.pls_pen ? a.pls_pen + (size_t)u * tr.n_ci : nullptr;
    const int32_t *acl = a.acl + (size_t)u * a.list_cap;
    const int n_act = a.n_act[u];
    const int E = tr.n_root + n_act;
    int32_t *tau = a.tau + base;
    uint8_t *dec = a.dec + base;
    unsigned long long *ecount = a.ecount + (size_t)u * (tr.n_root + a.list_cap);
    uint8_t *eflag = a.eflag + (size_t)u * (tr.n_root + a.list_cap);
    int32_t *nacl = a.nacl + (size_t)u * a.list_cap;
    int32_t *cand = a.cand + (size_t)u * a.cand_cap * 3;

    Thr t;
    t.fi = par[0]; t.thresh = par[1] + par[2]; t.newphone = par[1] + par[3]; t.lastphn = par[1] + par[4];
    t.pip = par[5]; t.nwpen = par[6]; t.pls = par[7] != 0 && pen != nullptr;
    const int32_t nf = t.fi + 1;

    // ---- phase 0: walk positions
    for (int e = tid; e < E; e += kPruneThreads) {
        if (e < tr.n_root) tau[e] = frame[e] >= t.fi ? e : -1;      // :737
        else tau[acl[e - tr.n_root]] = e;
    }
    __syncthreads();

    // ---- phase 1: decisions and counts, from unmodified state
    for (int e = tid; e < E; e += kPruneThreads) {
        const int c = e < tr.n_root ? e : acl[e - tr.n_root];
        unsigned n_app = 0, n_cand = 0; uint8_t fl = 0;
        if (tau[c] == e) {                                          // (an inactive root has tau -1)
            const bool keep = best[c] > t.thresh;                   // :740, :815
            if (e >= tr.n_root) {
                // what did / will the parent do to this channel?
                const int p = tr.parent[c];
                const int32_t pt = tau[p];
                unsigned d = 0;
                if (pt >= 0 && best[p] > t.thresh) {
                    int32_t pl;
                    d = edge_eval(t, out_score[p] + t.pip, t.pls ? pen[tr.ciphone[c]] : 0, pt, e, frame[c], score0[c], best[c], pl);
                }
                const bool entered_before = (d & 1u) && pt < e;
                if (keep) fl = kKeep | (entered_before ? 0 : kSelfAppend);     // :817-820
                else if (!entered_before) fl = kClear | ((d & 1u) ? kClearSkip0 : 0);   // :863-865
                n_app = (fl & kSelfAppend) ? 1u : 0u;
            } else if (keep) fl = kKeep;
            if (keep) {
                const int32_t nps = out_score[c] + t.pip;
                for (int k = tr.child_off[c]; k < tr.child_off[c + 1]; ++k) {
                    const int c2 = tr.child[k];
                    int32_t pl;
                    const unsigned d = edge_eval(t, nps, t.pls ? pen[tr.ciphone[c2]] : 0, e, tau[c2], frame[c2], score0[c2], best[c2], pl);
                    dec[c2] = (uint8_t)d;
                    n_app += d >> 1;
                }
                if (t.pls || nps > t.lastphn)                                   // :763, :845
                    for (int k = tr.pw_off[c]; k < tr.pw_off[c + 1]; ++k)
                        n_cand += (nps + (t.pls ? pen[tr.pw_lp[k]] : 0)) > t.lastphn ? 1u : 0u;
            }
        }
        ecount[e] = (unsigned long long)n_app | ((unsigned long long)n_cand << 32);
        eflag[e] = fl;
    }
    __syncthreads();

    // ---- exclusive scan of (appends, candidates) in walk order
    __shared__ unsigned long long s_warp[kPruneThreads / 32];
    __shared__ unsigned long long s_carry;
    if (tid == 0) s_carry = 0ull;
    __syncthreads();
    for (int b0 = 0; b0 < E; b0 += kPruneThreads) {
        const int e = b0 + tid;
        const unsigned long long v = e < E ? ecount[e] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if ((tid & 31) >= o) x += y;
        }
        if ((tid & 31) == 31) s_warp[tid >> 5] = x;
        __syncthreads();
        if (tid < 32) {
            unsigned long long w = s_warp[tid], xs = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, xs, o);
                if (tid >= o) xs += y;
            }
            s_warp[tid] = xs - w;                                  // exclusive over warps
        }
        __syncthreads();
        const unsigned long long carry = s_carry;
        if (e < E) ecount[e] = carry + s_warp[tid >> 5] + (x - v);
        __syncthreads();
        if (tid == kPruneThreads - 1) s_carry = carry + s_warp[tid >> 5] + x;
        __syncthreads();
    }
    if (tid == 0) {
        a.n_nacl[u] = (int32_t)(s_carry & 0xffffffffull);
        a.n_cand[u] = (int32_t)(s_carry >> 32);
    }

    // ---- phase 2: writes
    for (int e = tid; e < E; e += kPruneThreads) {
        const uint8_t fl = eflag[e];
        const int c = e < tr.n_root ? e : acl[e - tr.n_root];
        unsigned off = (unsigned)(ecount[e] & 0xffffffffull), coff = (unsigned)(ecount[e] >> 32);
        if (e >= tr.n_root) tau[c] = -1;
        if (fl & kSelfAppend) { if (off < (unsigned)a.list_cap) nacl[off] = c; ++off; }
        if (fl & kKeep) {
            frame[c] = nf;
            const int32_t nps = out_score[c] + t.pip, oh = out_hist[c];
            for (int k = tr.child_off[c]; k < tr.child_off[c + 1]; ++k) {
                const int c2 = tr.child[k];
                const unsigned d = dec[c2];
                if (d & 1u) {                                                   // hmm_enter
                    score0[c2] = nps + (t.pls ? pen[tr.ciphone[c2]] : 0);
                    hist0[c2] = oh;
                    frame[c2] = nf;
                    if (d & 2u) { if (off < (unsigned)a.list_cap) nacl[off] = c2; ++off; }
                }
            }
            if (t.pls || nps > t.lastphn)
                for (int k = tr.pw_off[c]; k < tr.pw_off[c + 1]; ++k) {
                    const int32_t pl = nps + (t.pls ? pen[tr.pw_lp[k]] : 0);
                    if (pl > t.lastphn) {
                        if (coff < (unsigned)a.cand_cap) {
                            cand[3 * coff] = tr.pw_wid[k]; cand[3 * coff + 1] = pl - t.nwpen; cand[3 * coff + 2] = oh;   // :772-777
                        }
                        ++coff;
                    }
                }
        } else if (fl & kClear) {                                               // hmm_clear_scores
            for (int s = (fl & kClearSkip0) ? 1 : 0; s < a.ne; ++s) a.score[(size_t)s * a.stride + base + c] = kWorst;
            out_score[c] = kWorst;
            best[c] = kWorst;
        }
    }
}

// The two other places where ngram_fwdtree_search touches the tree's channels:
//   mode 0  renormalize_scores (ngram_search_fwdtree.c:557-576, the tree part): hmm_normalize (hmm.c:205-216) of
//           the roots stamped with the current frame and of every active-list entry, norm[u] = the
//           utterance's best score;
//   mode 1  deactivate_channels (:1418-1431): hmm_clear_scores of the roots still stamped with the
//           current frame after word_transition (the ones prune_root_chan did not keep).
__global__ void __launch_bounds__(256) fwdtree_maint_kernel(PruneTree tr, PruneArgs a, const int32_t *norm, int mode) {
    const int u = blockIdx.y;
    const int e = blockIdx.x * 256 + threadIdx.x;
    const int n_act = mode == 0 ? a.n_act[u] : 0;
    if (e >= tr.n_root + n_act) return;
    const size_t base = (size_t)u * tr.n_chan;
    const int32_t fi = a.par[(size_t)u * 8];
    int c;
    if (e < tr.n_root) { if (a.frame[base + e] != fi) return; c = e; }
    else c = a.acl[(size_t)u * a.list_cap + (e - tr.n_root)];
    if (mode == 0) {
        const int32_t b = norm[u];
        for (int s = 0; s < a.ne; ++s) {
            const int32_t v = a.score[(size_t)s * a.stride + base + c];
            if (v > kWorst) a.score[(size_t)s * a.stride + base + c] = v - b;
        }
        const int32_t o = a.out_score[base + c];
        if (o > kWorst) a.out_score[base + c] = o - b;
    } else {
        for (int s = 0; s < a.ne; ++s) a.score[(size_t)s * a.stride + base + c] = kWorst;
        a.out_score[base + c] = kWorst;
        a.bestscore[base + c] = kWorst;
    }
}

}  // namespace

}  // namespace b200

using namespace b200;

struct b200_chantree {
    PruneTree t{};
    int device = 0, n_emit = 3;
    std::vector<void *> owned;
    // scratch + staging, sized for (n_utt, list_cap)
    int cap_utt = 0, cap_list = 0;
    int32_t *d_tau = nullptr; uint8_t *d_dec = nullptr; unsigned long long *d_ecount = nullptr; uint8_t *d_eflag = nullptr;
    // host-form staging
    int32_t *d_state = nullptr; size_t state_cap = 0;
    int32_t *d_lists = nullptr; size_t lists_cap = 0;
};

namespace {

template <typename T>
int dev_copy(b200_chantree *h, const T *src, size_t n, const T **dst) {
    void *p = nullptr;
    B200_CUDA_OK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    h->owned.push_back(p);
    if (n) B200_CUDA_OK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T *)p;
    return B200_OK;
}

int scratch_reserve(b200_chantree *h, int n_utt, int list_cap) {
    if (n_utt <= h->cap_utt && list_cap <= h->cap_list) return B200_OK;
    cudaFree(h->d_tau); cudaFree(h->d_dec); cudaFree(h->d_ecount); cudaFree(h->d_eflag);
    h->d_tau = nullptr; h->d_dec = nullptr; h->d_ecount = nullptr; h->d_eflag = nullptr; h->cap_utt = h->cap_list = 0;
    const int U = std::max(n_utt, h->cap_utt), Lc = std::max(list_cap, h->cap_list);
    const size_t nc = (size_t)U * h->t.n_chan, ne = (size_t)U * (h->t.n_root + Lc);
    B200_CUDA_OK(cudaMalloc((void **)&h->d_tau, std::max<size_t>(nc, 1) * 4));
    B200_CUDA_OK(cudaMalloc((void **)&h->d_dec, std::max<size_t>(nc, 1)));
    B200_CUDA_OK(cudaMalloc((void **)&h->d_ecount, std::max<size_t>(ne, 1) * 8));
    B200_CUDA_OK(cudaMalloc((void **)&h->d_eflag, std::max<size_t>(ne, 1)));
    B200_CUDA_OK(cudaMemset(h->d_tau, 0xff, std::max<size_t>(nc, 1) * 4));
    B200_CUDA_OK(cudaMemset(h->d_dec, 0, std::max<size_t>(nc, 1)));
    h->cap_utt = U; h->cap_list = Lc;
    return B200_OK;
}

}  // namespace

extern "C" b200_chantree_t *b200_chantree_create(int n_root, int n_chan, const int32_t *child_off, const int32_t *child,
                                                  const int32_t *ciphone, const int32_t *pw_off, const int32_t *pw_wid,
                                                  const int32_t *pw_lastphone, int n_ci, int n_emit, int device) {
    if (n_root < 0 || n_chan < n_root || !child_off || !pw_off || !ciphone || n_ci < 1 || n_emit < 1 || n_emit > 5) {
        set_error("b200_chantree_create: bad argument"); return nullptr;
    }
    const int n_edge = child_off[n_chan], n_pw = pw_off[n_chan];
    if (child_off[0] != 0 || pw_off[0] != 0 || n_edge < 0 || n_pw < 0 || (n_edge && !child) || (n_pw && (!pw_wid || !pw_lastphone))) {
        set_error("b200_chantree_create: bad CSR offsets"); return nullptr;
    }
    std::vector<int32_t> parent((size_t)std::max(n_chan, 1), -1);
    for (int c = 0; c < n_chan; ++c) {
        if (child_off[c + 1] < child_off[c] || pw_off[c + 1] < pw_off[c]) { set_error("b200_chantree_create: CSR offsets decrease at %d", c); return nullptr; }
        if (ciphone[c] < 0 || ciphone[c] >= n_ci) { set_error("b200_chantree_create: ciphone[%d] = %d", c, ciphone[c]); return nullptr; }
        for (int k = child_off[c]; k < child_off[c + 1]; ++k) {
            const int d = child[k];
            if (d < n_root || d >= n_chan) { set_error("b200_chantree_create: child %d of %d is not a non-root channel", d, c); return nullptr; }
            if (parent[d] != -1) { set_error("b200_chantree_create: channel %d has two parents (%d, %d): not a tree", d, parent[d], c); return nullptr; }
            parent[d] = c;
        }
    }
    for (int k = 0; k < n_pw; ++k)
        if (pw_lastphone[k] < 0 || pw_lastphone[k] >= n_ci) { set_error("b200_chantree_create: last phone %d out of range", pw_lastphone[k]); return nullptr; }
    for (int c = n_root; c < n_chan; ++c)
        if (parent[c] < 0) { set_error("b200_chantree_create: non-root channel %d has no parent", c); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: libb200sphinx has no CPU fallback"); return nullptr; }
    if (device < 0 || device >= ndev) { set_error("bad device %d", device); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
    auto *h = new b200_chantree;
    h->device = device; h->n_emit = n_emit;
    h->t.n_root = n_root; h->t.n_chan = n_chan; h->t.n_ci = n_ci; h->t.n_edge = n_edge; h->t.n_pw = n_pw;
    if (dev_copy(h, child_off, (size_t)n_chan + 1, &h->t.child_off) || dev_copy(h, child, (size_t)n_edge, &h->t.child) ||
        dev_copy(h, ciphone, (size_t)n_chan, &h->t.ciphone) || dev_copy(h, pw_off, (size_t)n_chan + 1, &h->t.pw_off) ||
        dev_copy(h, pw_wid, (size_t)n_pw, &h->t.pw_wid) || dev_copy(h, pw_lastphone, (size_t)n_pw, &h->t.pw_lp) ||
        dev_copy(h, parent.data(), (size_t)n_chan, &h->t.parent)) {
        b200_chantree_free(h);
        return nullptr;
    }
    return h;
}

extern "C" void b200_chantree_free(b200_chantree_t *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (void *p : h->owned) cudaFree(p);
    cudaFree(h->d_tau); cudaFree(h->d_dec); cudaFree(h->d_ecount); cudaFree(h->d_eflag); cudaFree(h->d_state); cudaFree(h->d_lists);
    delete h;
}

extern "C" int b200_chantree_cand_cap(const b200_chantree_t *h) { return h ? h->t.n_pw : 0; }

extern "C" int b200_fwdtree_prune_dev(b200_chantree_t *h, int n_utt, const b200_prune_dev_t *d, void *stream) {
    if (!h || !d || n_utt < 1 || d->list_cap < 0 || d->cand_cap < 0 || !d->score || !d->history || !d->out_score ||
        !d->out_history || !d->bestscore || !d->frame || !d->par || !d->n_act || !d->n_nacl || !d->n_cand ||
        (d->list_cap && (!d->acl || !d->nacl)) || (d->cand_cap && !d->cand) || d->state_stride < (long)n_utt * h->t.n_chan) {
        set_error("b200_fwdtree_prune_dev: bad argument"); return B200_ERR_ARG;
    }
    B200_CUDA_OK(cudaSetDevice(h->device));
    if (int rc = scratch_reserve(h, n_utt, d->list_cap)) return rc;
    PruneArgs a{};
    a.ne = h->n_emit; a.stride = d->state_stride;
    a.score = d->score; a.history = d->history; a.out_score = d->out_score; a.out_history = d->out_history;
    a.bestscore = d->bestscore; a.frame = d->frame;
    a.par = d->par; a.pls_pen = d->pls_pen; a.acl = d->acl; a.n_act = d->n_act; a.list_cap = d->list_cap;
    a.nacl = d->nacl; a.n_nacl = d->n_nacl; a.cand = d->cand; a.n_cand = d->n_cand; a.cand_cap = d->cand_cap;
    a.tau = h->d_tau; a.dec = h->d_dec; a.ecount = h->d_ecount; a.eflag = h->d_eflag;
    fwdtree_prune_kernel<<<n_utt, kPruneThreads, 0, (cudaStream_t)stream>>>(h->t, a);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_fwdtree_prune_host(b200_chantree_t *h, int n_utt, const int32_t *par, const int32_t *pls_pen,
                                       const int32_t *acl, const int32_t *n_act, int list_cap, int32_t *score,
                                       int32_t *history, int32_t *out_score, int32_t *out_history, int32_t *bestscore,
                                       int32_t *frame, int32_t *nacl, int32_t *n_nacl, int32_t *cand, int32_t *n_cand,
                                       int cand_cap) {
    if (!h || n_utt < 1 || !par || !n_act || list_cap < 0 || cand_cap < 0 || !score || !history || !out_score || !out_history ||
        !bestscore || !frame || !n_nacl || !n_cand || (list_cap && (!acl || !nacl)) || (cand_cap && !cand)) {
        set_error("b200_fwdtree_prune_host: bad argument"); return B200_ERR_ARG;
    }
    const int nc = h->t.n_chan, ne = h->n_emit;
    for (int u = 0; u < n_utt; ++u) {
        if (n_act[u] < 0 || n_act[u] > list_cap) { set_error("b200_fwdtree_prune_host: n_act[%d] = %d exceeds the list capacity", u, n_act[u]); return B200_ERR_ARG; }
        for (int i = 0; i < n_act[u]; ++i) {
            const int c = acl[(size_t)u * list_cap + i];
            if (c < h->t.n_root || c >= nc) { set_error("b200_fwdtree_prune_host: active-list entry %d is not a non-root channel", c); return B200_ERR_ARG; }
        }
    }
    B200_CUDA_OK(cudaSetDevice(h->device));
    const size_t N = (size_t)n_utt * nc;
    const size_t state_words = N * (2 * (size_t)ne + 4);
    if (state_words > h->state_cap) {
        cudaFree(h->d_state); h->d_state = nullptr; h->state_cap = 0;
        B200_CUDA_OK(cudaMalloc((void **)&h->d_state, std::max<size_t>(state_words, 1) * 4));
        h->state_cap = state_words;
    }
    const size_t L = (size_t)n_utt * list_cap, Cc = (size_t)n_utt * cand_cap * 3;
    const size_t list_words = (size_t)n_utt * (8 + h->t.n_ci + 3) + 2 * L + Cc;
    if (list_words > h->lists_cap) {
        cudaFree(h->d_lists); h->d_lists = nullptr; h->lists_cap = 0;
        B200_CUDA_OK(cudaMalloc((void **)&h->d_lists, std::max<size_t>(list_words, 1) * 4));
        h->lists_cap = list_words;
    }
    b200_prune_dev_t d{};
    int32_t *s = h->d_state;
    d.score = s; s += N * ne; d.history = s; s += N * ne; d.out_score = s; s += N; d.out_history = s; s += N;
    d.bestscore = s; s += N; d.frame = s;
    d.state_stride = (long)N;
    int32_t *l = h->d_lists;
    int32_t *d_par = l; l += (size_t)n_utt * 8;
    int32_t *d_pen = l; l += (size_t)n_utt * h->t.n_ci;
    int32_t *d_nact = l; l += n_utt; d.n_nacl = l; l += n_utt; d.n_cand = l; l += n_utt;
    int32_t *d_acl = l; l += L; d.nacl = l; l += L; d.cand = l;
    d.par = d_par; d.pls_pen = pls_pen ? d_pen : nullptr; d.acl = d_acl; d.n_act = d_nact; d.list_cap = list_cap; d.cand_cap = cand_cap;
    B200_CUDA_OK(cudaMemcpy(d.score, score, N * ne * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(d.history, history, N * ne * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(d.out_score, out_score, N * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(d.out_history, out_history, N * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(d.bestscore, bestscore, N * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(d.frame, frame, N * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(d_par, par, (size_t)n_utt * 8 * 4, cudaMemcpyHostToDevice));
    if (pls_pen) B200_CUDA_OK(cudaMemcpy(d_pen, pls_pen, (size_t)n_utt * h->t.n_ci * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(d_nact, n_act, (size_t)n_utt * 4, cudaMemcpyHostToDevice));
    if (L) B200_CUDA_OK(cudaMemcpy(d_acl, acl, L * 4, cudaMemcpyHostToDevice));
    if (int rc = b200_fwdtree_prune_dev(h, n_utt, &d, nullptr)) return rc;
    B200_CUDA_OK(cudaDeviceSynchronize());
    B200_CUDA_OK(cudaMemcpy(score, d.score, N * ne * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(history, d.history, N * ne * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(out_score, d.out_score, N * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(bestscore, d.bestscore, N * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(frame, d.frame, N * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(n_nacl, d.n_nacl, (size_t)n_utt * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(n_cand, d.n_cand, (size_t)n_utt * 4, cudaMemcpyDeviceToHost));
    if (L) B200_CUDA_OK(cudaMemcpy(nacl, d.nacl, L * 4, cudaMemcpyDeviceToHost));
    if (Cc) B200_CUDA_OK(cudaMemcpy(cand, d.cand, Cc * 4, cudaMemcpyDeviceToHost));
    for (int u = 0; u < n_utt; ++u)
        if (n_nacl[u] > list_cap || n_cand[u] > cand_cap) {
            set_error("b200_fwdtree_prune_host: utterance %d needs %d list / %d candidate slots", u, n_nacl[u], n_cand[u]);
            return B200_ERR_ARG;
        }
    return B200_OK;
}

static int fwdtree_maint(b200_chantree_t *h, int n_utt, const b200_prune_dev_t *d, const int32_t *d_norm, int mode, void *stream) {
    if (!h || !d || n_utt < 1 || !d->score || !d->out_score || !d->bestscore || !d->frame || !d->par ||
        (mode == 0 && (!d_norm || !d->n_act || d->list_cap < 0 || (d->list_cap && !d->acl))) ||
        d->state_stride < (long)n_utt * h->t.n_chan) { set_error("b200_fwdtree_renorm_dev / _deactivate_dev: bad argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(h->device));
    PruneArgs a{};
    a.ne = h->n_emit; a.stride = d->state_stride;
    a.score = d->score; a.history = d->history; a.out_score = d->out_score; a.out_history = d->out_history;
    a.bestscore = d->bestscore; a.frame = d->frame; a.par = d->par; a.acl = d->acl; a.n_act = d->n_act; a.list_cap = d->list_cap;
    const int elems = h->t.n_root + (mode == 0 ? d->list_cap : 0);
    fwdtree_maint_kernel<<<dim3((elems + 255) / 256, n_utt), 256, 0, (cudaStream_t)stream>>>(h->t, a, d_norm, mode);
    B200_LAUNCH_CHECK();
    return B200_OK;
}

extern "C" int b200_fwdtree_renorm_dev(b200_chantree_t *h, int n_utt, const b200_prune_dev_t *d, const int32_t *d_norm, void *stream) {
    return fwdtree_maint(h, n_utt, d, d_norm, 0, stream);
}

extern "C" int b200_fwdtree_deactivate_dev(b200_chantree_t *h, int n_utt, const b200_prune_dev_t *d, void *stream) {
    return fwdtree_maint(h, n_utt, d, nullptr, 1, stream);
}
