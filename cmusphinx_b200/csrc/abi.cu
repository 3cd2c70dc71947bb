// abi.cu -- the extern "C" surface of libb200sphinx.so (device side).
//
// Owns the device-resident models, scratch and streams; every scoring entry
// point ends in sm_100a kernel launches (gmm_exact.cu, mahal_tc.cu,
// hmm_kernels.cu).  There is no host fallback: without a CUDA device the
// constructors fail and say so.
#include "gmm_dev.cuh"
#include "hmm_dev.cuh"

#include <algorithm>
#include <cstring>
#include <new>

namespace b200 {
std::atomic<long long> g_launches{0};
}
using namespace b200;

namespace {

template <typename T>
int dev_alloc_copy(T **dst, const T *src, size_t n) {
    *dst = nullptr;
    if (n == 0) return B200_OK;
    B200_CUDA_OK(cudaMalloc((void **)dst, n * sizeof(T)));
    if (src) B200_CUDA_OK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return B200_OK;
}

constexpr size_t kListScratchBytes = (size_t)256 << 20;
constexpr int kHostChunkFrames = 8192;

}  // namespace

struct b200_mgau {
    int kind = 0;  // 0 ms, 1 ptm, 2 s2_semi
    b200_mgau_cfg_t cfg{};
    GmmDev g{};
    int device = 0;
    bool cont = false;  // ms with identity senone->codebook map
    int path = 0;
    TcPlan *tc = nullptr;
    TcTied *tct = nullptr;   // tensor-core codebook stage of the tied back-ends
    float *d_mean = nullptr, *d_var = nullptr, *d_det = nullptr;
    uint8_t *d_mixw = nullptr, *d_sen2cb = nullptr;
    uint32_t *d_sen2mgau = nullptr;
    size_t n_param = 0, n_det = 0;
    // scratch
    int2 *d_lists = nullptr; size_t lists_cap = 0;  // bytes
    float *d_feat[2] = {nullptr, nullptr}; int16_t *d_out[2] = {nullptr, nullptr};
    size_t feat_cap[2] = {0, 0}, out_cap[2] = {0, 0};
    cudaStream_t st[2] = {nullptr, nullptr};
    // ring of event quadruples: one per timed scoring call, so a caller can run
    // K asynchronous steps and read every step's kernel times afterwards
    static constexpr int kRing = 64;
    cudaEvent_t ring[kRing][5] = {};   // ev0 start, ev1 after prep, ev2 after scoring (+ fix-ups), ev3 end, ev4 between the score kernel and its fix-up kernels
    cudaEvent_t *ev = ring[0];
    long long n_timed = 0;
    float last_ms[4] = {0, 0, 0, 0};
    // per-frame / utterance cache
    float *d_ufeat = nullptr; size_t ufeat_cap = 0;
    int16_t *d_uraw = nullptr; size_t uraw_cap = 0;
    int2 *d_ulists = nullptr; size_t ulists_cap = 0;
    int2 *d_carry = nullptr; int carry_frame = -2;   // s2_semi -ds: finished list of the last frame scored
    int utt_T = 0;
    uint8_t *d_active = nullptr; size_t active_cap = 0;
    int16_t *d_row = nullptr;
    int16_t *h_row = nullptr;   // pinned; kernels write it directly (UVA), no per-frame D2H call
    uint8_t *h_active = nullptr; size_t h_active_cap = 0;   // pinned copy of the caller's active list, read by kernels directly
    float *h_frame = nullptr;   // pinned, one frame of features
    // ms / s2_semi: the utterance's score rows on the host (pinned), copied once at utt_begin --
    // frame_eval calls are then served without touching the GPU
    int16_t *h_uraw = nullptr; size_t h_uraw_cap = 0;
};

namespace {

int ensure(void **p, size_t *cap, size_t bytes) {
    if (*cap >= bytes) return B200_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    B200_CUDA_OK(cudaMalloc(p, bytes));
    *cap = bytes;
    return B200_OK;
}

size_t list_bytes_per_frame(const b200_mgau *m) {
    return (size_t)m->g.n_mgau * m->g.n_feat * m->g.topn * sizeof(int2);
}

int frames_per_list_chunk(const b200_mgau *m) {
    size_t per = list_bytes_per_frame(m);
    size_t n = kListScratchBytes / (per ? per : 1);
    if (n < 128) n = 128;
    if (n > 32768) n = 32768;
    return (int)n;
}

b200_mgau *mgau_common(int kind, const b200_mgau_cfg_t *cfg, const float *mean, const float *var,
                       const float *det) {
    if (!cfg || !mean || !var || !det) { set_error("null argument"); return nullptr; }
    if (cfg->n_feat < 1 || cfg->n_feat > B200_MAX_STREAMS) { set_error("n_feat %d unsupported (1..%d)", cfg->n_feat, B200_MAX_STREAMS); return nullptr; }
    if (cfg->topn < 1 || cfg->topn > B200_MAX_TOPN) { set_error("topn %d unsupported (1..%d)", cfg->topn, B200_MAX_TOPN); return nullptr; }
    if (cfg->topn > cfg->n_density) { set_error("topn %d > n_density %d", cfg->topn, cfg->n_density); return nullptr; }
    // -ds: ms_mgau never reads it (ignored, as in the reference); s2_semi re-scores the previous
    // frame's codewords on the skipped frames (tied_ds_kernel); the reference's ptm path leaves
    // RAW un-normalised scores in its lists on skipped frames (ptm_mgau.c:247-248 returns before
    // the normalisation) and then indexes its 256-byte log-add table with them: undefined, refused.
    if (cfg->ds_ratio > 1 && kind == 1) { set_error("-ds %d: the reference's ptm back-end is undefined for -ds > 1 (un-normalised list scores, ptm_mgau.c:247-248)", cfg->ds_ratio); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: libb200sphinx has no CPU fallback");
        return nullptr;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { set_error("bad device %d", cfg->device); return nullptr; }
    if (cudaSetDevice(cfg->device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
    b200_mgau *m = new (std::nothrow) b200_mgau();
    if (!m) return nullptr;
    m->kind = kind; m->cfg = *cfg; m->device = cfg->device;
    GmmDev &g = m->g;
    g.n_mgau = cfg->n_mgau; g.n_feat = cfg->n_feat; g.n_density = cfg->n_density; g.n_sen = cfg->n_sen;
    g.topn = cfg->topn; g.aw = cfg->aw > 0 ? cfg->aw : 1;
    g.veclen = 0; g.maxlen = 0;
    for (int f = 0; f < cfg->n_feat; ++f) {
        g.featlen[f] = cfg->featlen[f]; g.featoff[f] = g.veclen;
        g.veclen += cfg->featlen[f]; g.maxlen = std::max(g.maxlen, cfg->featlen[f]);
        g.topn_beam[f] = kind == 2 ? cfg->topn_beam[f] : 0;   // only s2_semi_mgau_init reads -topn_beam
    }
    m->n_param = (size_t)g.n_mgau * g.n_density * g.veclen;
    m->n_det = (size_t)g.n_mgau * g.n_feat * g.n_density;
    uint32_t tab[256];
    LogMath lm(cfg->logbase, kShift, true);
    if (lm.width != 1) { set_error("log base %f too small for an 8-bit add table", cfg->logbase); delete m; return nullptr; }
    for (int i = 0; i < 256; ++i) g.logadd[i] = (uint8_t)(i < (int)lm.table.size() ? lm.table[i] : 0);
    (void)tab;
    if (dev_alloc_copy(&m->d_mean, mean, m->n_param) || dev_alloc_copy(&m->d_var, var, m->n_param) ||
        dev_alloc_copy(&m->d_det, det, m->n_det)) { b200_mgau_free(m); return nullptr; }
    g.mean = m->d_mean; g.var = m->d_var; g.det = m->d_det;
    for (int i = 0; i < 2; ++i)
        if (cudaStreamCreateWithFlags(&m->st[i], cudaStreamNonBlocking) != cudaSuccess) { set_error("stream create failed"); b200_mgau_free(m); return nullptr; }
    for (int r = 0; r < b200_mgau::kRing; ++r)
        for (int i = 0; i < 5; ++i)
            if (cudaEventCreate(&m->ring[r][i]) != cudaSuccess) { set_error("event create failed"); b200_mgau_free(m); return nullptr; }
    if (cudaMalloc((void **)&m->d_row, (size_t)g.n_sen * 2 + 16) != cudaSuccess ||
        cudaMallocHost((void **)&m->h_row, (size_t)g.n_sen * 2 + 16) != cudaSuccess ||
        cudaMallocHost((void **)&m->h_frame, (size_t)g.veclen * 4 + 16) != cudaSuccess) {
        set_error("row buffers alloc failed"); b200_mgau_free(m); return nullptr;
    }
    return m;
}

// Dense un-normalised (ms, normalize=false) or final scores for frames [0,T)
// of d_feat.  Events: ev0 start, ev1 after operand prep (tensor-core path
// only), ev2 after the scoring kernels, ev3 after normalisation.
int score_dense_dev(b200_mgau *m, const float *d_feat, int T, int16_t *d_out, cudaStream_t st,
                    bool normalize, bool timed) {
    const GmmDev &g = m->g;
    if (T <= 0) return B200_OK;
    int rc;
    if (timed) {
        m->ev = m->ring[m->n_timed % b200_mgau::kRing];
        ++m->n_timed;
        cudaEventRecord(m->ev[0], st);
    }
    if (m->kind == 0 && m->path == 1 && m->tc) {
        int T_pad = 0;
        if (timed) cudaEventRecord(m->ev[4], st);       // (re-recorded inside when the fix-up kernels run)
        if ((rc = tc_score_raw(m->tc, d_feat, T, st, timed ? &m->ev[1] : nullptr, &T_pad, timed ? &m->ev[4] : nullptr))) return rc;
        if (timed) cudaEventRecord(m->ev[2], st);
        if ((rc = tc_finish(m->tc, T, T_pad, normalize ? 1 : 0, d_out, st))) return rc;
        if (timed) cudaEventRecord(m->ev[3], st);
        return B200_OK;
    }
    if (timed) cudaEventRecord(m->ev[1], st);
    if (m->kind == 0 && m->cont) {
        for (int t0 = 0; t0 < T; t0 += 65535 * 128) {   // grid.y limit
            int tn = std::min(T - t0, 65535 * 128);
            if ((rc = gmm_launch_topn(g, 0, d_feat, T, t0, tn, nullptr, d_out, 1, st))) return rc;
        }
    } else {
        int chunk = frames_per_list_chunk(m);
        const int ds = m->kind == 2 ? std::max(1, m->cfg.ds_ratio) : 1;
        if (ds > 1) chunk = std::max(ds, chunk - chunk % ds);   // every chunk starts on a fully evaluated frame
        if ((rc = ensure((void **)&m->d_lists, &m->lists_cap, (size_t)chunk * list_bytes_per_frame(m)))) return rc;
        for (int t0 = 0; t0 < T; t0 += chunk) {
            int tn = std::min(T - t0, chunk);
            if (m->kind != 0 && m->path == 1 && m->tct) rc = tc_tied_lists(m->tct, g, d_feat, t0, tn, m->d_lists, st);
            else rc = gmm_launch_topn(g, m->kind, d_feat, T, t0, tn, m->d_lists, nullptr, 0, st);
            if (rc) return rc;
            if (ds > 1 && (rc = gmm_launch_tied_ds(g, d_feat, t0, tn, 0, ds, m->d_lists, nullptr, st))) return rc;
            if (m->kind == 0) rc = gmm_launch_ms_senone(g, m->d_lists, T, t0, tn, d_out, st);
            else rc = gmm_launch_tied_senone(g, m->d_lists, T, t0, tn, m->kind == 2, nullptr, 0, d_out, st);
            if (rc) return rc;
        }
    }
    if (timed) cudaEventRecord(m->ev[2], st);
    if (m->kind == 0 && normalize && (rc = gmm_launch_normalize(d_out, T, g.n_sen, st))) return rc;
    if (timed) cudaEventRecord(m->ev[3], st);
    return B200_OK;
}

void fetch_times(b200_mgau *m) {
    // ev0 start, ev1 after prep (or start), ev2 after main, ev3 end
    float a = 0, b = 0, c = 0, tot = 0;
    cudaEventElapsedTime(&tot, m->ev[0], m->ev[3]);
    cudaEventElapsedTime(&a, m->ev[0], m->ev[1]);
    cudaEventElapsedTime(&b, m->ev[1], m->ev[2]);
    cudaEventElapsedTime(&c, m->ev[2], m->ev[3]);
    m->last_ms[0] = tot; m->last_ms[1] = a; m->last_ms[2] = b; m->last_ms[3] = c;
}

}  // namespace

extern "C" {

int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

long long b200_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------- create
b200_mgau_t *b200_ms_create(const b200_mgau_cfg_t *cfg, const float *mean, const float *var,
                            const float *det, const uint8_t *mixw, const uint32_t *sen2mgau) {
    if (!mixw || !sen2mgau) { set_error("null argument"); return nullptr; }
    b200_mgau *m = mgau_common(0, cfg, mean, var, det);
    if (!m) return nullptr;
    GmmDev &g = m->g;
    // transpose mixw [sen][feat][cw] -> [feat][cw][sen]
    std::vector<uint8_t> t((size_t)g.n_sen * g.n_feat * g.n_density);
    for (int s = 0; s < g.n_sen; ++s)
        for (int f = 0; f < g.n_feat; ++f)
            for (int c = 0; c < g.n_density; ++c)
                t[((size_t)f * g.n_density + c) * g.n_sen + s] = mixw[((size_t)s * g.n_feat + f) * g.n_density + c];
    m->cont = (g.n_mgau == g.n_sen);
    for (int s = 0; s < g.n_sen; ++s) {
        if ((int)sen2mgau[s] >= g.n_mgau) { set_error("sen2mgau[%d]=%u out of range", s, sen2mgau[s]); b200_mgau_free(m); return nullptr; }
        if ((int)sen2mgau[s] != s) m->cont = false;
    }
    if (dev_alloc_copy(&m->d_mixw, t.data(), t.size()) || dev_alloc_copy(&m->d_sen2mgau, sen2mgau, (size_t)g.n_sen)) {
        b200_mgau_free(m); return nullptr;
    }
    g.mixw_t = m->d_mixw; g.sen2mgau = m->d_sen2mgau; g.sen2cb = nullptr; g.n_clust = 0; g.row_bytes = g.n_sen;
    m->path = 0;
    if (m->cont && tc_shape_supported(g)) {
        m->tc = tc_plan_create(g, mean, var, det, mixw, m->device);
        if (m->tc) m->path = 1;
    }
    return m;
}

static b200_mgau_t *tied_create(int kind, const b200_mgau_cfg_t *cfg, const float *mean, const float *var,
                                const float *det, const uint8_t *mixw, int n_clust, const uint8_t *mixw_cb,
                                const uint8_t *sen2cb) {
    if (!mixw) { set_error("null argument"); return nullptr; }
    if (n_clust && !mixw_cb) { set_error("cluster codebook missing"); return nullptr; }
    if (kind == 2 && cfg && cfg->n_mgau != 1) { set_error("s2_semi needs exactly one codebook"); return nullptr; }
    if (kind == 1 && cfg && cfg->n_mgau > 256) { set_error("number of codebooks exceeds 256: %d", cfg->n_mgau); return nullptr; }
    b200_mgau *m = mgau_common(kind, cfg, mean, var, det);
    if (!m) return nullptr;
    GmmDev &g = m->g;
    g.n_clust = n_clust ? 16 : 0;
    g.row_bytes = n_clust ? (g.n_sen + 1) / 2 : g.n_sen;
    memset(g.mixw_cb, 0, 16);
    if (n_clust) memcpy(g.mixw_cb, mixw_cb, 16);
    std::vector<uint8_t> s2c((size_t)g.n_sen, 0);
    if (kind == 1) {
        if (!sen2cb) { set_error("ptm needs sen2cb"); b200_mgau_free(m); return nullptr; }
        for (int s = 0; s < g.n_sen; ++s) {
            if (sen2cb[s] >= g.n_mgau) { set_error("sen2cb[%d]=%d out of range", s, sen2cb[s]); b200_mgau_free(m); return nullptr; }
            s2c[s] = sen2cb[s];
        }
    }
    if (dev_alloc_copy(&m->d_mixw, mixw, (size_t)g.n_feat * g.n_density * g.row_bytes) ||
        dev_alloc_copy(&m->d_sen2cb, s2c.data(), s2c.size())) { b200_mgau_free(m); return nullptr; }
    g.mixw_t = m->d_mixw; g.sen2cb = m->d_sen2cb; g.sen2mgau = nullptr;
    if (gmm_tied_smem(g, g.n_sen + g.n_sen / 255 + 1) > 200 * 1024) {
        set_error("model too large for the tied senone kernel's shared memory"); b200_mgau_free(m); return nullptr;
    }
    // tensor-core codebook stage (bit-identical lists); the exact kernel stays as path 0
    m->path = 0;
    m->tct = tc_tied_create(g, kind, mean, var, det, m->device);
    if (m->tct) m->path = 1;
    return m;
}

b200_mgau_t *b200_ptm_create(const b200_mgau_cfg_t *cfg, const float *mean, const float *var, const float *det,
                             const uint8_t *mixw, int n_clust, const uint8_t *mixw_cb, const uint8_t *sen2cb) {
    return tied_create(1, cfg, mean, var, det, mixw, n_clust, mixw_cb, sen2cb);
}

b200_mgau_t *b200_semi_create(const b200_mgau_cfg_t *cfg, const float *mean, const float *var, const float *det,
                              const uint8_t *mixw, int n_clust, const uint8_t *mixw_cb) {
    return tied_create(2, cfg, mean, var, det, mixw, n_clust, mixw_cb, nullptr);
}

b200_mgau_t *b200_ms_load(const char *meanfile, const char *varfile, const char *mixwfile, const char *senmgau,
                          const uint8_t *sen2cb, double varfloor, double mixwfloor, int topn, int aw,
                          double logbase, int device) {
    int32_t dm[4], dv[4], dw[4], vl[64], vl2[64];
    if (b200_s3_read_gauden(meanfile, dm, vl, nullptr) || b200_s3_read_gauden(varfile, dv, vl2, nullptr)) return nullptr;
    if (dm[0] != dv[0] || dm[1] != dv[1] || dm[2] != dv[2]) { set_error("mixture-gaussians dimensions for means and variances differ"); return nullptr; }
    if (dm[1] > B200_MAX_STREAMS) { set_error("n_feat %d unsupported", dm[1]); return nullptr; }
    for (int f = 0; f < dm[1]; ++f) if (vl[f] != vl2[f]) { set_error("feature lengths for means and variances differ"); return nullptr; }
    std::vector<float> mean((size_t)dm[3]), var((size_t)dm[3]);
    if (b200_s3_read_gauden(meanfile, dm, vl, mean.data()) || b200_s3_read_gauden(varfile, dv, vl2, var.data())) return nullptr;
    if (b200_s3_read_mixw(mixwfile, dw, nullptr)) return nullptr;
    std::vector<float> mixw((size_t)dw[3]);
    if (b200_s3_read_mixw(mixwfile, dw, mixw.data())) return nullptr;
    if (dw[1] != dm[1]) { set_error("#feature mismatch: gauden=%d senone=%d", dm[1], dw[1]); return nullptr; }
    if (dw[2] != dm[2]) { set_error("#densities mismatch: gauden=%d senone=%d", dm[2], dw[2]); return nullptr; }
    const int n_sen = dw[0], n_feat = dm[1], n_density = dm[2], n_mgau = dm[0];
    // precompute per stream block (layout [mgau][feat][density][len f])
    std::vector<float> det((size_t)n_mgau * n_feat * n_density);
    int veclen = 0;
    for (int f = 0; f < n_feat; ++f) veclen += vl[f];
    for (int mg = 0; mg < n_mgau; ++mg) {
        int off = 0;
        for (int f = 0; f < n_feat; ++f) {
            float *vp = var.data() + (size_t)mg * n_density * veclen + (size_t)n_density * off;
            float *dp = det.data() + ((size_t)mg * n_feat + f) * n_density;
            if (b200_gauden_precompute(vp, dp, n_density, vl[f], (float)varfloor, logbase)) return nullptr;
            off += vl[f];
        }
    }
    std::vector<uint8_t> q((size_t)dw[3]);
    if (b200_mixw_quantize_ms(mixw.data(), q.data(), n_sen, n_feat, n_density, (float)mixwfloor, logbase)) return nullptr;
    std::vector<uint32_t> map((size_t)n_sen, 0u);
    std::string sm = senmgau ? senmgau : "";
    if (sm.empty()) sm = (n_mgau == 1) ? ".semi." : ".cont.";
    if (sm == ".semi.") { /* all zero */ }
    else if (sm == ".ptm.") {
        if (!sen2cb) { set_error(".ptm. mapping needs sen2cb"); return nullptr; }
        for (int s = 0; s < n_sen; ++s) map[s] = sen2cb[s];
    } else if (sm == ".cont." || sm == ".s3cont.") {
        if (n_sen <= 1) { set_error("#senone=%d; must be >1", n_sen); return nullptr; }
        for (int s = 0; s < n_sen; ++s) map[s] = s;
    } else { set_error("senone-codebook map files are not supported (%s)", sm.c_str()); return nullptr; }
    b200_mgau_cfg_t cfg{};
    cfg.n_mgau = n_mgau; cfg.n_feat = n_feat; cfg.n_density = n_density; cfg.n_sen = n_sen;
    for (int f = 0; f < n_feat; ++f) cfg.featlen[f] = vl[f];
    cfg.topn = (topn == 0 || topn > n_density) ? n_density : topn;   // ms_mgau.c:121-127
    cfg.aw = aw; cfg.ds_ratio = 1; cfg.logbase = logbase; cfg.device = device;
    return b200_ms_create(&cfg, mean.data(), var.data(), det.data(), q.data(), map.data());
}

void b200_mgau_free(b200_mgau_t *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->h_uraw) cudaFreeHost(m->h_uraw);
    if (m->tc) tc_plan_free(m->tc);
    if (m->tct) tc_tied_free(m->tct);
    cudaFree(m->d_mean); cudaFree(m->d_var); cudaFree(m->d_det); cudaFree(m->d_mixw);
    cudaFree(m->d_sen2cb); cudaFree(m->d_sen2mgau); cudaFree(m->d_lists);
    for (int i = 0; i < 2; ++i) { cudaFree(m->d_feat[i]); cudaFree(m->d_out[i]); if (m->st[i]) cudaStreamDestroy(m->st[i]); }
    for (int r = 0; r < b200_mgau::kRing; ++r)
        for (int i = 0; i < 5; ++i) if (m->ring[r][i]) cudaEventDestroy(m->ring[r][i]);
    cudaFree(m->d_ufeat); cudaFree(m->d_uraw); cudaFree(m->d_ulists); cudaFree(m->d_carry); cudaFree(m->d_active); cudaFree(m->d_row);
    if (m->h_row) cudaFreeHost(m->h_row);
    if (m->h_active) cudaFreeHost(m->h_active);
    if (m->h_frame) cudaFreeHost(m->h_frame);
    delete m;
}

const char *b200_mgau_name(const b200_mgau_t *m) {
    if (!m) return "";
    return m->kind == 0 ? "b200_ms" : (m->kind == 1 ? "b200_ptm" : "b200_semi");
}
int b200_mgau_n_sen(const b200_mgau_t *m) { return m ? m->g.n_sen : 0; }
int b200_mgau_featdim(const b200_mgau_t *m) { return m ? m->g.veclen : 0; }

int b200_mgau_update_params(b200_mgau_t *m, const float *mean, const float *var, const float *det) {
    if (!m || !mean || !var || !det) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(m->device));
    B200_CUDA_OK(cudaMemcpy(m->d_mean, mean, m->n_param * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(m->d_var, var, m->n_param * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(m->d_det, det, m->n_det * 4, cudaMemcpyHostToDevice));
    if (m->tc) {
        // the tensor-core operand is derived from (mean, var, det): rebuild it
        std::vector<uint8_t> mixw_sfc((size_t)m->g.n_sen * m->g.n_feat * m->g.n_density);
        std::vector<uint8_t> t(mixw_sfc.size());
        B200_CUDA_OK(cudaMemcpy(t.data(), m->d_mixw, t.size(), cudaMemcpyDeviceToHost));
        const GmmDev &g = m->g;
        for (int s = 0; s < g.n_sen; ++s)
            for (int f = 0; f < g.n_feat; ++f)
                for (int c = 0; c < g.n_density; ++c)
                    mixw_sfc[((size_t)s * g.n_feat + f) * g.n_density + c] = t[((size_t)f * g.n_density + c) * g.n_sen + s];
        tc_plan_free(m->tc);
        m->tc = tc_plan_create(m->g, mean, var, det, mixw_sfc.data(), m->device);
        if (!m->tc) m->path = 0;
    }
    if (m->tct) {
        const int was = m->path;
        tc_tied_free(m->tct);
        m->tct = tc_tied_create(m->g, m->kind, mean, var, det, m->device);
        m->path = m->tct ? was : 0;
    }
    return B200_OK;
}

int b200_mgau_tied_stats(b200_mgau_t *m, long long out[3]) {
    if (!m || !out) { set_error("null argument"); return B200_ERR_ARG; }
    out[0] = out[1] = out[2] = 0;
    if (m->tct) tc_tied_stats(m->tct, out);
    return B200_OK;
}

int b200_mgau_cont_stats(b200_mgau_t *m, long long out[7]) {
    if (!m || !out) { set_error("null argument"); return B200_ERR_ARG; }
    for (int i = 0; i < 7; ++i) out[i] = 0;
    if (!m->tc) return B200_OK;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    if (tc_last_stats(m->tc, out)) { set_error("tc_last_stats: %s", cudaGetErrorString(cudaGetLastError())); return B200_ERR_CUDA; }
    return B200_OK;
}

int b200_mgau_set_path(b200_mgau_t *m, int path) {
    if (!m) return B200_ERR_ARG;
    if (path == 0) { m->path = 0; return B200_OK; }
    if (path == 1 && m->kind != 0) {
        if (!m->tct) { set_error("tensor-core path unavailable for this model shape"); return B200_ERR_UNSUP; }
        m->path = 1; return B200_OK;
    }
    if (path == 1) {
        if (!m->tc) { set_error("tensor-core path unavailable for this model shape"); return B200_ERR_UNSUP; }
        m->path = 1; return B200_OK;
    }
    set_error("unknown path %d", path);
    return B200_ERR_ARG;
}
int b200_mgau_get_path(const b200_mgau_t *m) { return m ? m->path : -1; }

int b200_mgau_tc_last_format(b200_mgau_t *m) {
    if (!m || !m->tc) return -1;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    return tc_last_format(m->tc);
}

// ------------------------------------------------------------ dense scoring
int b200_mgau_score_dev(b200_mgau_t *m, const float *d_feat, int T, int16_t *d_out, void *stream) {
    if (!m || (T > 0 && (!d_feat || !d_out))) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(m->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : m->st[0];
    int rc = score_dense_dev(m, d_feat, T, d_out, st, true, true);
    if (rc) return rc;
    if (!stream) {
        B200_CUDA_OK(cudaStreamSynchronize(st));
        fetch_times(m);
    }
    return B200_OK;
}

int b200_mgau_score_host(b200_mgau_t *m, const float *feat, int T, int16_t *out) {
    if (!m || (T > 0 && (!feat || !out))) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(m->device));
    const GmmDev &g = m->g;
    int chunk = std::min(T, kHostChunkFrames);
    if (chunk <= 0) return B200_OK;
    if (m->kind == 2 && m->cfg.ds_ratio > 1 && chunk < T)   // chunks are scored as frames 0.. : keep them aligned to -ds
        chunk = std::max(m->cfg.ds_ratio, chunk - chunk % m->cfg.ds_ratio);
    const size_t fb = (size_t)chunk * g.veclen * 4, ob = (size_t)chunk * g.n_sen * 2;
    for (int i = 0; i < 2; ++i) {
        int rc = ensure((void **)&m->d_feat[i], &m->feat_cap[i], fb); if (rc) return rc;
        rc = ensure((void **)&m->d_out[i], &m->out_cap[i], ob); if (rc) return rc;
    }
    // two streams ping-pong: H2D(k+1) and D2H(k-1) overlap compute(k).  The list
    // scratch is shared, so compute itself is serialised through events.
    cudaEvent_t done[2];
    for (int i = 0; i < 2; ++i) B200_CUDA_OK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    int k = 0, rc = B200_OK;
    for (int t0 = 0; t0 < T; t0 += chunk, ++k) {
        const int b = k & 1, tn = std::min(chunk, T - t0);
        cudaStream_t st = m->st[b];
        cudaMemcpyAsync(m->d_feat[b], feat + (size_t)t0 * g.veclen, (size_t)tn * g.veclen * 4, cudaMemcpyHostToDevice, st);
        if (k > 0) cudaStreamWaitEvent(st, done[b ^ 1], 0);
        rc = score_dense_dev(m, m->d_feat[b], tn, m->d_out[b], st, true, false);
        if (rc) break;
        cudaEventRecord(done[b], st);
        cudaMemcpyAsync(out + (size_t)t0 * g.n_sen, m->d_out[b], (size_t)tn * g.n_sen * 2, cudaMemcpyDeviceToHost, st);
    }
    cudaError_t e0 = cudaStreamSynchronize(m->st[0]), e1 = cudaStreamSynchronize(m->st[1]);
    for (int i = 0; i < 2; ++i) cudaEventDestroy(done[i]);
    if (rc) return rc;
    B200_CUDA_OK(e0);
    B200_CUDA_OK(e1);
    return B200_OK;
}

float b200_mgau_last_ms(const b200_mgau_t *m, int which) {
    if (!m || which < 0 || which > 3) return -1.f;
    return m->last_ms[which];
}

float b200_mgau_timing_avg(b200_mgau_t *m, int n_calls, int which) {
    if (!m || which < 0 || which > 4 || n_calls < 1) return -1.f;
    if (n_calls > b200_mgau::kRing) n_calls = b200_mgau::kRing;
    if ((long long)n_calls > m->n_timed) n_calls = (int)m->n_timed;
    if (n_calls < 1) return -1.f;
    if (cudaSetDevice(m->device) != cudaSuccess) return -1.f;
    double sum = 0;
    for (int k = 0; k < n_calls; ++k) {
        cudaEvent_t *e = m->ring[(m->n_timed - 1 - k) % b200_mgau::kRing];
        float ms = 0;
        cudaError_t rc = which == 0 ? cudaEventElapsedTime(&ms, e[0], e[3])
                       : which == 4 ? cudaEventElapsedTime(&ms, e[4], e[2])
                                    : cudaEventElapsedTime(&ms, e[which - 1], e[which]);
        if (rc != cudaSuccess) { set_error("timing not available: %s", cudaGetErrorString(rc)); cudaGetLastError(); return -1.f; }
        sum += ms;
    }
    return (float)(sum / n_calls);
}

// ------------------------------------------------- utterance / frame serving
int b200_mgau_utt_begin(b200_mgau_t *m, const float *feat, int T) { return b200_mgau_utt_begin_at(m, feat, T, 0); }

int b200_mgau_utt_begin_at(b200_mgau_t *m, const float *feat, int T, int frame0) {
    if (!m || T < 0 || frame0 < 0 || (T > 0 && !feat)) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(m->device));
    const GmmDev &g = m->g;
    m->utt_T = 0;
    if (T == 0) return B200_OK;
    cudaStream_t st = m->st[0];
    int rc = ensure((void **)&m->d_ufeat, &m->ufeat_cap, (size_t)T * g.veclen * 4);
    if (rc) return rc;
    B200_CUDA_OK(cudaMemcpyAsync(m->d_ufeat, feat, (size_t)T * g.veclen * 4, cudaMemcpyHostToDevice, st));
    if (m->kind == 0) {
        if ((rc = ensure((void **)&m->d_uraw, &m->uraw_cap, (size_t)T * g.n_sen * 2))) return rc;
        if ((rc = score_dense_dev(m, m->d_ufeat, T, m->d_uraw, st, false, false))) return rc;
    } else {
        if ((rc = ensure((void **)&m->d_ulists, &m->ulists_cap, (size_t)T * list_bytes_per_frame(m)))) return rc;
        for (int t0 = 0; t0 < T; t0 += 65535 * 128) {
            int tn = std::min(T - t0, 65535 * 128);
            // lists for frame t land at index (t - 0): pass t0 = 0 base by offsetting the pointer
            int2 *dst = m->d_ulists + (size_t)t0 * g.n_mgau * g.n_feat * g.topn;
            if (m->path == 1 && m->tct) rc = tc_tied_lists(m->tct, g, m->d_ufeat, t0, tn, dst, st);
            else rc = gmm_launch_topn(g, m->kind, m->d_ufeat, T, t0, tn, dst, nullptr, 0, st);
            if (rc) return rc;
        }
        const int ds = m->kind == 2 ? std::max(1, m->cfg.ds_ratio) : 1;
        if (ds > 1) {
            const size_t lb = list_bytes_per_frame(m);
            const bool need = frame0 % ds != 0;
            if (need && m->carry_frame != frame0 - 1) {
                set_error("-ds %d: frame %d needs the list of frame %d, which was not the last frame scored", ds, frame0, frame0 - 1);
                return B200_ERR_ARG;
            }
            if (!m->d_carry) B200_CUDA_OK(cudaMalloc((void **)&m->d_carry, lb));
            if ((rc = gmm_launch_tied_ds(g, m->d_ufeat, 0, T, frame0, ds, m->d_ulists, need ? m->d_carry : nullptr, st))) return rc;
            B200_CUDA_OK(cudaMemcpyAsync(m->d_carry, (const char *)m->d_ulists + (size_t)(T - 1) * lb, lb, cudaMemcpyDeviceToDevice, st));
            m->carry_frame = frame0 + T - 1;
        }
    }
    if (m->kind != 1) {
        // ms: un-normalised rows (the normalisation depends on the caller's active set,
        // PS/ms_mgau.c:226-248, and is applied by serve_frame on the host); s2_semi: final
        // scores (normalised by each stream's own best codeword, independent of the active set)
        if (m->kind == 2) {
            if ((rc = ensure((void **)&m->d_uraw, &m->uraw_cap, (size_t)T * g.n_sen * 2))) return rc;
            if ((rc = gmm_launch_tied_senone(g, m->d_ulists, T, 0, T, 1, nullptr, 0, m->d_uraw, st))) return rc;
        }
        const size_t bytes = (size_t)T * g.n_sen * 2;
        if (m->h_uraw_cap < bytes) {
            if (m->h_uraw) cudaFreeHost(m->h_uraw);
            m->h_uraw = nullptr; m->h_uraw_cap = 0;
            B200_CUDA_OK(cudaMallocHost((void **)&m->h_uraw, bytes));
            m->h_uraw_cap = bytes;
        }
        B200_CUDA_OK(cudaMemcpyAsync(m->h_uraw, m->d_uraw, bytes, cudaMemcpyDeviceToHost, st));
    }
    B200_CUDA_OK(cudaStreamSynchronize(st));
    m->utt_T = T;
    return B200_OK;
}

// ms / s2_semi: one frame_eval from the host copy of the utterance's rows (no GPU work).
static int serve_frame_host(b200_mgau_t *m, int frame, int16_t *senscr, const uint8_t *senone_active,
                            int32_t n_active, int32_t compallsen) {
    const int n_sen = m->g.n_sen;
    const int16_t *row = m->h_uraw + (size_t)frame * n_sen;
    if (!compallsen && (n_active < 0 || (n_active > 0 && !senone_active))) { set_error("bad active list"); return B200_ERR_ARG; }
    if (m->kind == 2) {
        // s2_semi_mgau_frame_eval: memset 0, then the active senones (s2_semi_mgau.c:840-886)
        if (compallsen) { memcpy(senscr, row, (size_t)n_sen * 2); return B200_OK; }
        memset(senscr, 0, (size_t)n_sen * 2);
        int s = 0;
        for (int i = 0; i < n_active; ++i) { s += senone_active[i]; if (s >= n_sen) break; senscr[s] = row[s]; }
        return B200_OK;
    }
    // ms_cont_mgau_frame_eval: subtract the best of the scored (all / active) senones (ms_mgau.c:188-248)
    int32_t best = 0x7fffffff;
    if (compallsen) {
        for (int i = 0; i < n_sen; ++i) best = std::min(best, (int32_t)row[i]);
        for (int i = 0; i < n_sen; ++i) {
            const int32_t v = (int32_t)row[i] - best;
            senscr[i] = (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
        }
        return B200_OK;
    }
    int s = 0;
    for (int i = 0; i < n_active; ++i) { s += senone_active[i]; if (s >= n_sen) break; best = std::min(best, (int32_t)row[s]); }
    s = 0;
    for (int i = 0; i < n_active; ++i) {
        s += senone_active[i];
        if (s >= n_sen) break;
        const int32_t v = (int32_t)row[s] - best;
        senscr[s] = (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
    }
    return B200_OK;
}

static int serve_frame(b200_mgau_t *m, int frame, int16_t *senscr, const uint8_t *senone_active,
                       int32_t n_active, int32_t compallsen) {
    const GmmDev &g = m->g;
    if (m->kind != 1) return serve_frame_host(m, frame, senscr, senone_active, n_active, compallsen);
    cudaStream_t st = m->st[0];
    const bool use_active = !compallsen;
    // ptm: the normalisation depends on which CODEBOOKS the active senones touch
    // (PS/ptm_mgau.c:267-288), so the mixing stage runs on the device per frame.
    // Per-frame serving is latency bound (one call per frame of the decoder's
    // search): the active list and the result row live in pinned host memory
    // that the kernels read / write directly, so a frame costs one or two
    // launches and one stream synchronisation, no copy calls.
    const uint8_t *act = nullptr;
    if (use_active) {
        if (n_active < 0 || (n_active > 0 && !senone_active)) { set_error("bad active list"); return B200_ERR_ARG; }
        if (m->h_active_cap < (size_t)std::max(n_active, 1)) {
            if (m->h_active) cudaFreeHost(m->h_active);
            m->h_active = nullptr; m->h_active_cap = 0;
            const size_t cap = (size_t)std::max(n_active, g.n_sen + g.n_sen / 255 + 64);
            B200_CUDA_OK(cudaMallocHost((void **)&m->h_active, cap));
            m->h_active_cap = cap;
        }
        if (n_active > 0) memcpy(m->h_active, senone_active, (size_t)n_active);
        act = m->h_active;
    }
    int rc;
    if (m->kind == 0) {
        const int16_t *raw = m->d_uraw + (size_t)frame * g.n_sen;
        if (use_active) {
            if (n_active == 0) return B200_OK;
            if ((rc = gmm_launch_ms_active_normalize(raw, act, n_active, m->h_row, st))) return rc;   // writes active entries only
        } else {
            B200_CUDA_OK(cudaMemcpyAsync(m->d_row, raw, (size_t)g.n_sen * 2, cudaMemcpyDeviceToDevice, st));
            if ((rc = gmm_launch_normalize(m->d_row, 1, g.n_sen, st))) return rc;
            B200_CUDA_OK(cudaMemcpyAsync(m->h_row, m->d_row, (size_t)g.n_sen * 2, cudaMemcpyDeviceToHost, st));
        }
    } else {
        const int2 *l = m->d_ulists + (size_t)frame * g.n_mgau * g.n_feat * g.topn;
        // out row index is t - t0 with T=1,t0=0 -> write straight into the pinned row
        if ((rc = gmm_launch_tied_senone(g, l, 1, 0, 1, m->kind == 2, act, use_active ? n_active : 0, m->h_row, st))) return rc;
    }
    B200_CUDA_OK(cudaStreamSynchronize(st));
    if (m->kind == 0 && use_active) {
        // the reference writes only the active entries of the caller's array
        int s = 0;
        for (int i = 0; i < n_active; ++i) { s += senone_active[i]; senscr[s] = m->h_row[s]; }
    } else {
        memcpy(senscr, m->h_row, (size_t)g.n_sen * 2);
    }
    return B200_OK;
}

int b200_mgau_utt_frame(b200_mgau_t *m, int16_t *senscr, const uint8_t *senone_active, int32_t n_senone_active,
                        int32_t frame, int32_t compallsen) {
    if (!m || !senscr) { set_error("null argument"); return B200_ERR_ARG; }
    if (frame < 0 || frame >= m->utt_T) { set_error("frame %d outside the scored utterance (0..%d)", frame, m->utt_T); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(m->device));
    return serve_frame(m, frame, senscr, senone_active, n_senone_active, compallsen);
}

int b200_mgau_frame_eval(b200_mgau_t *m, int16_t *senscr, const uint8_t *senone_active, int32_t n_senone_active,
                         const float *const *feat, int32_t frame, int32_t compallsen) {
    if (!m || !senscr || !feat) { set_error("null argument"); return B200_ERR_ARG; }
    const GmmDev &g = m->g;
    for (int f = 0; f < g.n_feat; ++f) {
        if (!feat[f]) { set_error("null stream pointer"); return B200_ERR_ARG; }
        memcpy(m->h_frame + g.featoff[f], feat[f], (size_t)g.featlen[f] * 4);
    }
    int rc = b200_mgau_utt_begin_at(m, m->h_frame, 1, frame < 0 ? 0 : frame);   // the frame number matters to -ds only
    if (rc) return rc;
    return serve_frame(m, 0, senscr, senone_active, n_senone_active, compallsen);
}

// ---------------------------------------------------------------------- HMM
}  // extern "C"

struct b200_hmmctx {
    HmmDev c{};
    int device = 0;
    uint8_t *d_tp = nullptr; uint16_t *d_sseq = nullptr;
    HmmPop p{};
    size_t pop_cap = 0;  // in HMMs
    HmmFrame *d_fr = nullptr; int fr_cap = 0;          // [3][n_utt] rotating frame records
    int fr_slot = 0;                                   // slot of the last frame that ran
    int32_t *d_block_count = nullptr, *d_keep_idx = nullptr;   // survivors per tile; survivor list
    int32_t *d_keep_tmp = nullptr;                             // per-utterance lists of the cluster kernel
    size_t bc_cap = 0;
    uint32_t *d_mask = nullptr; size_t mask_cap = 0;   // [2][n_utt][n_words]
    uint32_t *d_mask_part = nullptr; size_t mask_part_cap = 0;   // words: [2][n_utt][CTAs per utterance][n_words]
    int mask_par = 0;                                  // parity of the last frame's mask
    unsigned *d_bar = nullptr;                         // grid barrier of the persistent step kernel
    cudaStream_t last_st = nullptr;                    // stream of the last step (a caller may pass its own)
    bool stepped = false;
    int32_t *d_utt_off = nullptr; int utt_cap = 0;
    int32_t *d_total = nullptr;
    std::vector<int32_t> h_utt_off;
    int16_t *d_senscr = nullptr; size_t senscr_cap = 0;
    int32_t *d_winner = nullptr; size_t winner_cap = 0;      // hmm_enter scratch [n_hmm]
    int32_t *d_enter = nullptr; size_t enter_cap = 0;        // hmm_enter scratch: idx | score | hist | old0, n each
    cudaStream_t st = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    float last_ms = 0;
};

namespace {

void pop_free(b200_hmmctx *c) {
    cudaFree(c->p.score); cudaFree(c->p.history); cudaFree(c->p.out_score); cudaFree(c->p.out_history);
    cudaFree(c->p.bestscore); cudaFree(c->p.senid); cudaFree(c->p.tmatid); cudaFree(c->p.mpx);
    cudaFree(c->d_keep_idx); cudaFree(c->d_keep_tmp);
    c->p = HmmPop{}; c->d_keep_idx = nullptr; c->d_keep_tmp = nullptr; c->pop_cap = 0;
}

int pop_reserve(b200_hmmctx *c, int n) {
    if ((size_t)n <= c->pop_cap) { c->p.n_hmm = n; return B200_OK; }
    pop_free(c);
    const int ne = c->c.n_emit;
    const size_t N = (size_t)n;
    B200_CUDA_OK(cudaMalloc((void **)&c->p.score, N * ne * 4));
    B200_CUDA_OK(cudaMalloc((void **)&c->p.history, N * ne * 4));
    B200_CUDA_OK(cudaMalloc((void **)&c->p.out_score, N * 4));
    B200_CUDA_OK(cudaMalloc((void **)&c->p.out_history, N * 4));
    B200_CUDA_OK(cudaMalloc((void **)&c->p.bestscore, N * 4));
    B200_CUDA_OK(cudaMalloc((void **)&c->p.senid, N * ne * 2));
    B200_CUDA_OK(cudaMalloc((void **)&c->p.tmatid, N * 2));
    B200_CUDA_OK(cudaMalloc((void **)&c->p.mpx, N));
    B200_CUDA_OK(cudaMalloc((void **)&c->d_keep_idx, N * 4));
    B200_CUDA_OK(cudaMalloc((void **)&c->d_keep_tmp, N * 4));
    c->pop_cap = N; c->p.n_hmm = n;
    return B200_OK;
}

// Partition the resident population into utterances (device copies of the
// offsets, per-utterance frame state, masks and block counters).
int set_utts(b200_hmmctx *c, int n_utt, const int32_t *off) {
    if (n_utt < 1 || !off || off[0] != 0 || off[n_utt] != c->p.n_hmm) { set_error("bad utterance offsets"); return B200_ERR_ARG; }
    int mx = 0;
    for (int u = 0; u < n_utt; ++u) {
        if (off[u + 1] < off[u]) { set_error("utterance offsets not monotone"); return B200_ERR_ARG; }
        mx = std::max(mx, off[u + 1] - off[u]);
    }
    const int n_words = (c->c.n_sen + 31) / 32;
    if (n_utt + 1 > c->utt_cap) {
        cudaFree(c->d_utt_off); c->d_utt_off = nullptr; c->utt_cap = 0;
        B200_CUDA_OK(cudaMalloc((void **)&c->d_utt_off, (size_t)(n_utt + 1) * 4));
        c->utt_cap = n_utt + 1;
    }
    if (n_utt > c->fr_cap) {
        cudaFree(c->d_fr); c->d_fr = nullptr; c->fr_cap = 0;
        B200_CUDA_OK(cudaMalloc((void **)&c->d_fr, (size_t)3 * n_utt * sizeof(HmmFrame)));
        c->fr_cap = n_utt;
    }
    if ((size_t)n_utt * n_words > c->mask_cap) {
        cudaFree(c->d_mask); c->d_mask = nullptr; c->mask_cap = 0;
        B200_CUDA_OK(cudaMalloc((void **)&c->d_mask, (size_t)2 * n_utt * n_words * 4));
        c->mask_cap = (size_t)n_utt * n_words;
    }
    {   // partial masks: at most one resident wave of CTAs (<= 8 per SM) is split over the utterances
        const int wave = 8 * 148;
        const int gy = std::max(1, std::min(n_utt, wave));
        const size_t gx_max = (size_t)std::max(16, std::min((mx + 255) / 256, wave / gy));   // (>= the cluster kernel's 16 CTAs per utterance)
        const size_t need = (size_t)2 * n_utt * gx_max * n_words;
        if (need > c->mask_part_cap) {
            cudaFree(c->d_mask_part); c->d_mask_part = nullptr; c->mask_part_cap = 0;
            B200_CUDA_OK(cudaMalloc((void **)&c->d_mask_part, need * 4));
            c->mask_part_cap = need;
        }
    }
    const size_t nb = (size_t)2 * ((mx + 255) / 256) * n_utt + 1;     // (the resident kernel alternates between two sets of tile counts)
    if (nb > c->bc_cap) {
        cudaFree(c->d_block_count); c->d_block_count = nullptr; c->bc_cap = 0;
        B200_CUDA_OK(cudaMalloc((void **)&c->d_block_count, nb * 4));
        c->bc_cap = nb;
    }
    B200_CUDA_OK(cudaMemcpy(c->d_utt_off, off, (size_t)(n_utt + 1) * 4, cudaMemcpyHostToDevice));
    c->h_utt_off.assign(off, off + n_utt + 1);
    c->p.n_utt = n_utt; c->p.max_per_utt = mx; c->p.utt_off = c->d_utt_off;
    c->fr_slot = 0; c->mask_par = 0; c->stepped = false;
    return B200_OK;
}

int check_soa(const b200_hmmctx *c, const b200_hmm_soa_t *h) {
    if (!c || !h) { set_error("null argument"); return B200_ERR_ARG; }
    if (h->n_hmm < 0) { set_error("negative n_hmm"); return B200_ERR_ARG; }
    if (h->n_hmm > 0 && (!h->score || !h->history || !h->out_score || !h->out_history || !h->senid ||
                         !h->tmatid || !h->mpx || !h->bestscore)) { set_error("null SoA field"); return B200_ERR_ARG; }
    const size_t N = (size_t)h->n_hmm;
    for (int i = 0; i < h->n_hmm; ++i) {
        if (h->tmatid[i] < 0 || h->tmatid[i] >= c->c.n_tmat) { set_error("tmatid[%d]=%d out of range", i, h->tmatid[i]); return B200_ERR_ARG; }
        // the kernels index shared-memory tables with these ids: senone ids (or, for mpx HMMs,
        // senone-sequence ids with BAD_SSID = 0xffff meaning "no state")
        for (int s = 0; s < c->c.n_emit; ++s) {
            const unsigned id = h->senid[(size_t)s * N + i];
            if (h->mpx[i]) {
                if (id != B200_BAD_SSID && (int)id >= c->c.n_sseq) { set_error("ssid[%d][%d]=%u out of range (%d senone sequences)", s, i, id, c->c.n_sseq); return B200_ERR_ARG; }
            } else if ((int)id >= c->c.n_sen) { set_error("senid[%d][%d]=%u out of range (%d senones)", s, i, id, c->c.n_sen); return B200_ERR_ARG; }
        }
    }
    return B200_OK;
}

}  // namespace

extern "C" {

b200_hmmctx_t *b200_hmm_ctx_create(int n_emit, const uint8_t *tp, int n_tmat, const uint16_t *sseq, int n_sseq,
                                   int n_sen, int device) {
    if (n_emit < 1 || n_emit > 5) { set_error("n_emit_state %d unsupported (1..5 = HMM_MAX_NSTATE, PS/hmm.h:90)", n_emit); return nullptr; }
    if (!tp || n_tmat <= 0 || n_sen <= 0 || n_sen > 65535 || (n_sseq > 0 && !sseq)) { set_error("bad hmm context arguments"); return nullptr; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: libb200sphinx has no CPU fallback"); return nullptr; }
    if (device < 0 || device >= ndev || cudaSetDevice(device) != cudaSuccess) { set_error("bad device %d", device); return nullptr; }
    b200_hmmctx *c = new (std::nothrow) b200_hmmctx();
    if (!c) return nullptr;
    c->device = device;
    c->c.n_emit = n_emit; c->c.n_tmat = n_tmat; c->c.n_sseq = n_sseq; c->c.n_sen = n_sen;
    const int n_words = (n_sen + 31) / 32;
    if (dev_alloc_copy(&c->d_tp, tp, (size_t)n_tmat * n_emit * (n_emit + 1)) ||
        dev_alloc_copy(&c->d_sseq, sseq, (size_t)std::max(n_sseq, 1) * n_emit * (n_sseq > 0 ? 1 : 0)) ||
        cudaMalloc((void **)&c->d_total, 4) != cudaSuccess || cudaMalloc((void **)&c->d_bar, (size_t)kHmmBarRows * 32 * sizeof(unsigned)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev[0]) != cudaSuccess || cudaEventCreate(&c->ev[1]) != cudaSuccess) {
        set_error("hmm context allocation failed"); b200_hmm_ctx_free(c); return nullptr;
    }
    c->c.tp = c->d_tp; c->c.sseq = c->d_sseq;
    return c;
}

void b200_hmm_ctx_free(b200_hmmctx_t *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    pop_free(c);
    cudaFree(c->d_tp); cudaFree(c->d_sseq); cudaFree(c->d_fr); cudaFree(c->d_mask); cudaFree(c->d_mask_part); cudaFree(c->d_senscr); cudaFree(c->d_winner); cudaFree(c->d_enter);
    cudaFree(c->d_block_count); cudaFree(c->d_utt_off); cudaFree(c->d_total); cudaFree(c->d_bar);
    if (c->st) cudaStreamDestroy(c->st);
    for (int i = 0; i < 2; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    delete c;
}

int b200_hmm_pop_upload(b200_hmmctx_t *c, const b200_hmm_soa_t *h) {
    int rc = check_soa(c, h);
    if (rc) return rc;
    B200_CUDA_OK(cudaSetDevice(c->device));
    if ((rc = pop_reserve(c, h->n_hmm))) return rc;
    const size_t N = (size_t)h->n_hmm;
    const int ne = c->c.n_emit;
    {
        const int32_t one[2] = {0, h->n_hmm};
        if ((rc = set_utts(c, 1, one))) return rc;
    }
    if (N == 0) return B200_OK;
    B200_CUDA_OK(cudaMemcpy(c->p.score, h->score, N * ne * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(c->p.history, h->history, N * ne * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(c->p.out_score, h->out_score, N * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(c->p.out_history, h->out_history, N * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(c->p.bestscore, h->bestscore, N * 4, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(c->p.senid, h->senid, N * ne * 2, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(c->p.tmatid, h->tmatid, N * 2, cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(c->p.mpx, h->mpx, N, cudaMemcpyHostToDevice));
    return B200_OK;
}

int b200_hmm_pop_download(b200_hmmctx_t *c, b200_hmm_soa_t *h) {
    if (!c || !h) { set_error("null argument"); return B200_ERR_ARG; }
    if (h->n_hmm != c->p.n_hmm) { set_error("population size mismatch: %d vs %d", h->n_hmm, c->p.n_hmm); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(c->device));
    const size_t N = (size_t)h->n_hmm;
    const int ne = c->c.n_emit;
    if (N == 0) return B200_OK;
    B200_CUDA_OK(cudaStreamSynchronize(c->st));
    B200_CUDA_OK(cudaMemcpy(h->score, c->p.score, N * ne * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(h->history, c->p.history, N * ne * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(h->out_score, c->p.out_score, N * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(h->out_history, c->p.out_history, N * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(h->bestscore, c->p.bestscore, N * 4, cudaMemcpyDeviceToHost));
    B200_CUDA_OK(cudaMemcpy(h->senid, c->p.senid, N * ne * 2, cudaMemcpyDeviceToHost));
    return B200_OK;
}

// Device addresses of the resident population (for callers that chain their own device-side
// stages behind the step kernels, e.g. b200_fwdtree_prune_dev) and the stream its kernels run on.
int b200_hmm_pop_device(b200_hmmctx_t *c, b200_hmm_soa_t *dev, void **stream) {
    if (!c || !dev) { set_error("null argument"); return B200_ERR_ARG; }
    if (c->p.n_hmm <= 0) { set_error("b200_hmm_pop_device: no resident population"); return B200_ERR_ARG; }
    dev->n_hmm = c->p.n_hmm;
    dev->score = c->p.score; dev->history = c->p.history; dev->out_score = c->p.out_score; dev->out_history = c->p.out_history;
    dev->senid = c->p.senid; dev->tmatid = c->p.tmatid; dev->mpx = c->p.mpx; dev->bestscore = c->p.bestscore;
    if (stream) *stream = (void *)c->st;
    return B200_OK;
}

// eval_root_chan + eval_nonroot_chan (PS/ngram_search_fwdtree.c:598-634) on the resident population.
int b200_hmm_eval_list_dev(b200_hmmctx_t *c, int n_root, int n_chan, const int32_t *d_frame, const int32_t *d_par,
                           const int32_t *d_acl, const int32_t *d_n_act, int list_cap, const int16_t *d_senscr,
                           int32_t *d_best, void *stream) {
    if (!c || n_root < 0 || n_chan < 1 || n_root > n_chan || !d_frame || !d_par || !d_n_act || list_cap < 0 ||
        (list_cap && !d_acl) || !d_senscr || !d_best) { set_error("b200_hmm_eval_list_dev: bad argument"); return B200_ERR_ARG; }
    if (c->p.n_hmm <= 0 || c->p.n_hmm % n_chan) { set_error("b200_hmm_eval_list_dev: the resident population (%d HMMs) is not a whole number of %d-channel trees", c->p.n_hmm, n_chan); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(c->device));
    HmmList l{};
    l.n_root = n_root; l.n_chan = n_chan; l.list_cap = list_cap; l.frame = d_frame; l.par = d_par; l.acl = d_acl; l.n_act = d_n_act;
    l.senscr = d_senscr; l.best = d_best;
    return hmm_launch_eval_list(c->c, c->p, l, c->p.n_hmm / n_chan, stream ? (cudaStream_t)stream : c->st);
}

// One launch of the persistent step kernel: n_frames frames, frame f on the scores at
// d_senscr + ((frame0 + f) % n_cycle) * frame_stride.
static int hmm_run(b200_hmmctx *c, const int16_t *d_senscr, long frame_stride, int n_cycle, int n_frames, int32_t beam,
                   int do_beam, cudaStream_t st) {
    if (c->p.n_hmm <= 0 || n_frames <= 0) return B200_OK;
    HmmRun r{};
    r.sen_base = d_senscr; r.frame_stride = frame_stride; r.n_cycle = n_cycle; r.frame0 = 0; r.n_frames = n_frames;
    r.beam = beam; r.do_beam = do_beam;
    r.fr3 = c->d_fr; r.slot0 = (c->fr_slot + 1) % 3;
    r.tile_count = c->d_block_count; r.tpu = (c->p.max_per_utt + 255) / 256;
    r.keep_idx = c->d_keep_idx; r.keep_tmp = c->d_keep_tmp;
    r.mask2 = c->d_mask; r.mask0 = (c->mask_par + 1) & 1;
    r.mask_part = c->d_mask_part; r.mask_part_words = c->mask_part_cap;
    r.total = c->d_total; r.bar = c->d_bar;
    static long long *d_probe = nullptr;
    if (getenv("B200_HMM_PROBE") && !d_probe) cudaMalloc((void **)&d_probe, 128);
    r.probe = d_probe;
    const int rc = hmm_launch_run(c->c, c->p, r, st);
    if (rc) return rc;
    c->fr_slot = (r.slot0 + n_frames - 1) % 3;
    c->last_st = st;
    c->mask_par = (r.mask0 + n_frames - 1) & 1;
    c->stepped = do_beam != 0;
    if (d_probe) {
        long long h[11];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, d_probe, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "hmm probe (cycles): A %lld  bar1 %lld  B %lld (mask merge %lld, beam votes %lld, counts %lld)  bar2 %lld  C %lld (counts in %lld, scan %lld, scatter %lld, mask out %lld)\n",
                h[1] - h[0], h[2] - h[1], h[3] - h[2], h[6] - h[2], h[7] - h[6], h[3] - h[7], h[4] - h[3], h[5] - h[4], h[8] - h[4], h[9] - h[8], h[10] - h[9], h[5] - h[10]);
    }
    return B200_OK;
}

int b200_hmm_step_dev(b200_hmmctx_t *c, const int16_t *d_senscr, int32_t beam, void *stream) {
    if (!c || !d_senscr) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->st;
    cudaEventRecord(c->ev[0], st);
    const int rc = hmm_run(c, d_senscr, 0, 1, 1, beam, 1, st);
    cudaEventRecord(c->ev[1], st);
    return rc;
}

// A run of frames is ONE launch of the persistent kernel (two grid barriers per
// frame instead of five kernel launches).
int b200_hmm_run_dev(b200_hmmctx_t *c, const int16_t *d_senscr, long frame_stride, int n_cycle, int n_frames,
                     int32_t beam, void *stream) {
    if (!c || !d_senscr || n_cycle < 1 || n_frames < 0 || frame_stride < 0) { set_error("bad argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->st;
    cudaEventRecord(c->ev[0], st);
    const int rc = hmm_run(c, d_senscr, frame_stride, n_cycle, n_frames, beam, 1, st);
    cudaEventRecord(c->ev[1], st);
    return rc;
}

int b200_hmm_pop_set_utts(b200_hmmctx_t *c, int n_utt, const int32_t *utt_off) {
    if (!c) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(c->device));
    return set_utts(c, n_utt, utt_off);
}

int b200_hmm_step_results(b200_hmmctx_t *c, int32_t *best, int32_t *n_keep, int32_t *keep_idx, uint32_t *sen_mask) {
    if (!c) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(c->device));
    B200_CUDA_OK(cudaStreamSynchronize(c->st));
    if (c->last_st && c->last_st != c->st) B200_CUDA_OK(cudaStreamSynchronize(c->last_st));   // the step ran on the caller's stream
    const int nu = std::max(c->p.n_utt, 1);
    std::vector<HmmFrame> fr(nu);
    int32_t total = 0;
    if (c->p.n_hmm > 0) {
        B200_CUDA_OK(cudaMemcpy(fr.data(), c->d_fr + (size_t)c->fr_slot * nu, sizeof(HmmFrame) * nu, cudaMemcpyDeviceToHost));
        B200_CUDA_OK(cudaMemcpy(&total, c->d_total, 4, cudaMemcpyDeviceToHost));
    } else {
        for (auto &f : fr) { f.best = B200_WORST_SCORE; f.n_keep = 0; }
    }
    for (int u = 0; u < nu; ++u) {
        if (best) best[u] = fr[u].best;
        if (n_keep) n_keep[u] = fr[u].n_keep;
    }
    if (keep_idx && total > 0) B200_CUDA_OK(cudaMemcpy(keep_idx, c->d_keep_idx, (size_t)total * 4, cudaMemcpyDeviceToHost));
    if (sen_mask && c->p.n_hmm > 0)
        B200_CUDA_OK(cudaMemcpy(sen_mask, c->d_mask + (size_t)c->mask_par * nu * ((c->c.n_sen + 31) / 32),
                                (size_t)nu * ((c->c.n_sen + 31) / 32) * 4, cudaMemcpyDeviceToHost));
    cudaEventElapsedTime(&c->last_ms, c->ev[0], c->ev[1]);
    return B200_OK;
}

int b200_hmm_step_host(b200_hmmctx_t *c, const int16_t *senscr, int32_t beam) {
    if (!c || !senscr) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(c->device));
    const size_t nb = (size_t)c->c.n_sen * 2 * std::max(c->p.n_utt, 1);
    int rc = ensure((void **)&c->d_senscr, &c->senscr_cap, nb);
    if (rc) return rc;
    B200_CUDA_OK(cudaMemcpyAsync(c->d_senscr, senscr, nb, cudaMemcpyHostToDevice, c->st));
    return b200_hmm_step_dev(c, c->d_senscr, beam, nullptr);
}

int b200_hmm_eval_host(b200_hmmctx_t *c, b200_hmm_soa_t *h, const int16_t *senscr, int n_frames, int32_t *best_out) {
    if (!senscr || n_frames < 0) { set_error("bad arguments"); return B200_ERR_ARG; }
    int rc = b200_hmm_pop_upload(c, h);
    if (rc) return rc;
    if (h->n_hmm == 0) { for (int f = 0; f < n_frames; ++f) if (best_out) best_out[f] = B200_WORST_SCORE; return B200_OK; }
    if ((rc = ensure((void **)&c->d_senscr, &c->senscr_cap, (size_t)c->c.n_sen * 2 * std::max(n_frames, 1)))) return rc;
    B200_CUDA_OK(cudaMemcpyAsync(c->d_senscr, senscr, (size_t)c->c.n_sen * 2 * n_frames, cudaMemcpyHostToDevice, c->st));
    for (int f = 0; f < n_frames; ++f) {
        rc = hmm_run(c, c->d_senscr + (size_t)f * c->c.n_sen, 0, 1, 1, 0, 0, c->st);
        if (rc) return rc;
        if (best_out) B200_CUDA_OK(cudaMemcpyAsync(&best_out[f], (const int32_t *)(c->d_fr + (size_t)c->fr_slot * std::max(c->p.n_utt, 1)), 4,
                                                   cudaMemcpyDeviceToHost, c->st));
    }
    B200_CUDA_OK(cudaStreamSynchronize(c->st));
    return b200_hmm_pop_download(c, h);
}

float b200_hmm_last_ms(const b200_hmmctx_t *c) { return c ? c->last_ms : -1.f; }

int b200_hmm_normalize_dev(b200_hmmctx_t *c, const int32_t *d_best_per_utt, void *stream) {
    if (!c) { set_error("null argument"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(c->device));
    return hmm_launch_normalize(c->p, c->c.n_emit, d_best_per_utt, c->d_fr + (size_t)c->fr_slot * std::max(c->p.n_utt, 1),
                                stream ? (cudaStream_t)stream : c->st);
}

int b200_hmm_clear_pruned_dev(b200_hmmctx_t *c, void *stream) {
    if (!c || !c->stepped) { set_error("no beam step has run on this population"); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(c->device));
    return hmm_launch_clear_pruned(c->p, c->c.n_emit, c->d_fr + (size_t)c->fr_slot * std::max(c->p.n_utt, 1),
                                   stream ? (cudaStream_t)stream : c->st);
}

int b200_hmm_enter_dev(b200_hmmctx_t *c, const int32_t *d_idx, const int32_t *d_score, const int32_t *d_hist, int n,
                       void *stream) {
    if (!c || n < 0 || (n > 0 && (!d_idx || !d_score || !d_hist))) { set_error("b200_hmm_enter_dev: bad argument"); return B200_ERR_ARG; }
    if (n == 0 || c->p.n_hmm == 0) return B200_OK;
    B200_CUDA_OK(cudaSetDevice(c->device));
    int rc = ensure((void **)&c->d_winner, &c->winner_cap, (size_t)c->p.n_hmm * 4);
    if (rc) return rc;
    if ((rc = ensure((void **)&c->d_enter, &c->enter_cap, (size_t)n * 16))) return rc;
    return hmm_launch_enter(c->p, d_idx, d_score, d_hist, n, c->d_winner, c->d_enter + (size_t)3 * n, nullptr,
                            stream ? (cudaStream_t)stream : c->st);
}

int b200_hmm_enter_host(b200_hmmctx_t *c, const int32_t *idx, const int32_t *score, const int32_t *hist, int n) {
    if (!c || n < 0 || (n > 0 && (!idx || !score || !hist))) { set_error("b200_hmm_enter_host: bad argument"); return B200_ERR_ARG; }
    if (n == 0) return B200_OK;
    for (int k = 0; k < n; ++k)
        if (idx[k] < 0 || idx[k] >= c->p.n_hmm) { set_error("hmm_enter: HMM index %d out of range", idx[k]); return B200_ERR_ARG; }
    B200_CUDA_OK(cudaSetDevice(c->device));
    int rc = ensure((void **)&c->d_enter, &c->enter_cap, (size_t)n * 16);
    if (rc) return rc;
    B200_CUDA_OK(cudaMemcpyAsync(c->d_enter, idx, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
    B200_CUDA_OK(cudaMemcpyAsync(c->d_enter + n, score, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
    B200_CUDA_OK(cudaMemcpyAsync(c->d_enter + 2 * (size_t)n, hist, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
    if ((rc = b200_hmm_enter_dev(c, c->d_enter, c->d_enter + n, c->d_enter + 2 * (size_t)n, n, c->st))) return rc;
    B200_CUDA_OK(cudaStreamSynchronize(c->st));
    return B200_OK;
}

// ------------------------------------------------------------ memory helpers
void *b200_dev_alloc(size_t bytes, int device) {
    void *p = nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p, bytes) != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
void b200_dev_free(void *p) { if (p) cudaFree(p); }
int b200_dev_upload(void *dst, const void *src, size_t bytes) {
    B200_CUDA_OK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return B200_OK;
}
int b200_dev_download(void *dst, const void *src, size_t bytes) {
    B200_CUDA_OK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return B200_OK;
}
void *b200_host_alloc_pinned(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { set_error("cudaMallocHost(%zu) failed", bytes); cudaGetLastError(); return nullptr; }
    return p;
}
void b200_host_free_pinned(void *p) { if (p) cudaFreeHost(p); }
int b200_dev_sync(int device) {
    B200_CUDA_OK(cudaSetDevice(device));
    B200_CUDA_OK(cudaDeviceSynchronize());
    return B200_OK;
}

}  // extern "C"
