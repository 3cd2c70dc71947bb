// host_model.cc -- host-side (load-time) half of libb200sphinx.so.
//
// Everything here runs once at model load and must be BIT-IDENTICAL to what
// the reference computes on the host, because the device kernels consume the
// resulting integer-valued tables.  It therefore uses the same libm calls
// (log, sqrt in double) and the same truncation points as the reference:
//
//   log-add table        sphinxbase/src/libsphinxbase/util/logmath.c:61-161
//   logmath_log/add      .../logmath.c:391-452
//   det / scaled 1/2var  pocketsphinx/src/libpocketsphinx/ms_gauden.c:314-359
//   ms mixw quantiser    .../ms_senone.c:236-258
//   tied mixw quantiser  .../ptm_mgau.c:720-742, s2_semi_mgau.c:1155-1177
//   tmat quantiser       .../tmat.c:275-296
//   S3 binary container  sphinxbase/src/libsphinxbase/util/bio.c:137-300
//   sendump container    .../s2_semi_mgau.c:888-1089
//   active list deltas   .../acmod.c:1219-1271
#include "b200_internal.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace b200 {

thread_local std::string g_last_error;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

// ------------------------------------------------------------------ logmath
LogMath::LogMath(double b, int sh, bool use_table) : base(b), shift(sh) {
    log_of_base = std::log(base);
    inv_log_of_base = 1.0 / log_of_base;
    zero = (int32_t)0x80000000 >> (shift + 2);
    if (!use_table) return;
    // Pass 1: find the index at which the rounded entry reaches zero.
    double byx = 1.0;
    uint32_t i = 0;
    for (;; ++i) {
        double lobyx = std::log(1.0 + byx) * inv_log_of_base;
        int32_t k = (int32_t)(lobyx + 0.5 * (1 << shift)) >> shift;
        if (k <= 0) break;
        byx /= base;
    }
    i >>= shift;
    if (i < 255) i = 255;
    table.assign(i + 1, 0u);
    uint32_t maxyx = (uint32_t)(std::log(2.0) / std::log(base) + 0.5) >> shift;
    width = maxyx < 256 ? 1 : (maxyx < 65536 ? 2 : 4);
    // Pass 2: fill; within one shifted bucket the first (largest) value wins.
    byx = 1.0;
    for (i = 0;; ++i) {
        double lobyx = std::log(1.0 + byx) * inv_log_of_base;
        int32_t k = (int32_t)(lobyx + 0.5 * (1 << shift)) >> shift;
        uint32_t &slot = table[i >> shift];
        if (slot == 0) {
            uint32_t kk = (uint32_t)k;
            if (width == 1) kk &= 0xffu;
            else if (width == 2) kk &= 0xffffu;
            slot = kk;
        }
        if (k <= 0) break;
        byx /= base;
    }
}

int32_t LogMath::log(double p) const {
    if (p <= 0) return zero;
    return (int32_t)(std::log(p) * inv_log_of_base) >> shift;
}

int32_t LogMath::ln_to_log(double ln_p) const {
    return (int32_t)(ln_p * inv_log_of_base) >> shift;
}

int32_t LogMath::add(int32_t x, int32_t y) const {
    if (x <= zero) return y;
    if (y <= zero) return x;
    int32_t d, r;
    if (x > y) { d = x - y; r = x; } else { d = y - x; r = y; }
    if (d < 0) return r;
    if ((size_t)d >= table.size()) return r;
    return r + (int32_t)table[d];
}

// ------------------------------------------------------- vector utilities
// pocketsphinx/src/libpocketsphinx/vector.c:90-128
static double sum_norm(float *v, int n) {
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += v[i];
    if (sum != 0.0) {
        double f = 1.0 / sum;
        for (int i = 0; i < n; ++i) v[i] = (float)(v[i] * f);
    }
    return sum;
}
static void floor_all(float *v, int n, double flr) {
    for (int i = 0; i < n; ++i)
        if (v[i] < flr) v[i] = (float)flr;
}
static void floor_nz(float *v, int n, double flr) {
    for (int i = 0; i < n; ++i)
        if (v[i] != 0.0 && v[i] < flr) v[i] = (float)flr;
}

// ------------------------------------------------------------- S3 container
struct S3File {
    FILE *fp = nullptr;
    bool swap = false;
    bool chksum_present = false;
    uint32_t chksum = 0;
    ~S3File() { if (fp) fclose(fp); }
};

static inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }

static bool s3_open(S3File &f, const char *path) {
    f.fp = fopen(path, "rb");
    if (!f.fp) { set_error("cannot open '%s'", path); return false; }
    char line[16384], key[4096], val[4096];
    if (!fgets(line, sizeof line, f.fp)) { set_error("%s: empty", path); return false; }
    if (strcmp(line, "s3\n") != 0) {
        set_error("%s: not an s3 binary file (old headerless format unsupported)", path);
        return false;
    }
    for (;;) {
        if (!fgets(line, sizeof line, f.fp)) { set_error("%s: premature EOF in header", path); return false; }
        int n = 0;
        if (sscanf(line, "%4095s%n", key, &n) != 1) { set_error("%s: header format error", path); return false; }
        if (strcmp(key, "endhdr") == 0) break;
        if (key[0] == '#') continue;
        if (sscanf(line + n, "%4095s", val) != 1) { set_error("%s: header value missing", path); return false; }
        if (strcmp(key, "chksum0") == 0) f.chksum_present = true;
    }
    uint32_t magic;
    if (fread(&magic, 4, 1, f.fp) != 1) { set_error("%s: no byte-order magic", path); return false; }
    if (magic == 0x11223344u) f.swap = false;
    else if (bswap32(magic) == 0x11223344u) f.swap = true;
    else { set_error("%s: bad byte-order magic %08x", path, magic); return false; }
    return true;
}

// 4-byte element read with byte swap + rotating checksum (bio.c:266-345).
static bool s3_read32(S3File &f, void *dst, size_t n) {
    if (fread(dst, 4, n, f.fp) != n) return false;
    uint32_t *p = (uint32_t *)dst;
    for (size_t i = 0; i < n; ++i) {
        if (f.swap) p[i] = bswap32(p[i]);
        f.chksum = ((f.chksum << 20) | (f.chksum >> 12)) + p[i];
    }
    return true;
}

static bool s3_finish(S3File &f, const char *path) {
    if (f.chksum_present) {
        uint32_t file_sum;
        if (fread(&file_sum, 4, 1, f.fp) != 1) { set_error("%s: checksum missing", path); return false; }
        if (f.swap) file_sum = bswap32(file_sum);
        if (file_sum != f.chksum) { set_error("%s: checksum mismatch", path); return false; }
    }
    char c;
    if (fread(&c, 1, 1, f.fp) == 1) { set_error("%s: more data than expected", path); return false; }
    return true;
}

}  // namespace b200

using namespace b200;

extern "C" {

const char *b200_last_error(void) { return g_last_error.c_str(); }
int b200_abi_version(void) { return 1; }

int b200_logadd_table(double base, int shift, uint32_t *out, int max_out) {
    if (base <= 1.0) { set_error("base must be > 1"); return B200_ERR_ARG; }
    LogMath lm(base, shift, true);
    int n = (int)lm.table.size();
    for (int i = 0; i < n && i < max_out; ++i) out[i] = lm.table[i];
    return n;
}

int32_t b200_logmath_log(double base, int shift, double p) {
    return LogMath(base, shift, false).log(p);
}

int32_t b200_logmath_add(double base, int shift, int32_t x, int32_t y) {
    return LogMath(base, shift, true).add(x, y);
}

int b200_gauden_precompute(float *var, float *det, long n_vec, int len,
                           float varfloor, double logbase) {
    if (!var || !det || len <= 0 || varfloor <= 0) { set_error("bad precompute args"); return B200_ERR_ARG; }
    LogMath lm(logbase, 0, false);
    for (long v = 0; v < n_vec; ++v) {
        float *vp = var + (size_t)v * len;
        float d = 0.0f;
        for (int i = 0; i < len; ++i) {
            float fv = vp[i];
            if (fv < varfloor) fv = varfloor;
            // float += (float)int, exactly as the reference accumulates det.
            d += (float)lm.log(1.0 / std::sqrt(fv * 2.0 * M_PI));
            vp[i] = (float)lm.ln_to_log(1.0 / (fv * 2.0));
        }
        det[v] = d;
    }
    return B200_OK;
}

int b200_mixw_quantize_ms(float *mixw, uint8_t *out, int n_sen, int n_feat,
                          int n_cw, float mixwfloor, double logbase) {
    if (mixwfloor <= 0.0 || mixwfloor >= 1.0) { set_error("mixwfloor not in (0,1)"); return B200_ERR_ARG; }
    LogMath lm(logbase, 0, false);
    for (size_t r = 0; r < (size_t)n_sen * n_feat; ++r) {
        float *pdf = mixw + r * n_cw;
        sum_norm(pdf, n_cw);
        floor_all(pdf, n_cw, mixwfloor);
        sum_norm(pdf, n_cw);
        for (int c = 0; c < n_cw; ++c) {
            int32_t p = -lm.log(pdf[c]);
            p += (1 << (B200_SENSCR_SHIFT - 1)) - 1;
            out[r * n_cw + c] = (p < (255 << B200_SENSCR_SHIFT)) ? (uint8_t)(p >> B200_SENSCR_SHIFT) : 255;
        }
    }
    return B200_OK;
}

int b200_mixw_quantize_tied(float *mixw, uint8_t *out, int n_sen, int n_feat,
                            int n_cw, float mixwfloor, double logbase) {
    LogMath lm8(logbase, B200_SENSCR_SHIFT, false);
    for (int s = 0; s < n_sen; ++s)
        for (int f = 0; f < n_feat; ++f) {
            float *pdf = mixw + ((size_t)s * n_feat + f) * n_cw;
            sum_norm(pdf, n_cw);
            floor_all(pdf, n_cw, mixwfloor);
            sum_norm(pdf, n_cw);
            for (int c = 0; c < n_cw; ++c) {
                int32_t q = -lm8.log(pdf[c]);
                if (q > 159 || q < 0) q = 159;
                out[((size_t)f * n_cw + c) * n_sen + s] = (uint8_t)q;
            }
        }
    return B200_OK;
}

int b200_tmat_quantize(float *tp, uint8_t *out, int n_tmat, int n_src,
                       double tmatfloor, double logbase) {
    LogMath lm(logbase, 0, false);
    int n_dst = n_src + 1;
    for (int t = 0; t < n_tmat; ++t)
        for (int j = 0; j < n_src; ++j) {
            float *row = tp + ((size_t)t * n_src + j) * n_dst;
            sum_norm(row, n_dst);
            floor_nz(row, n_dst, tmatfloor);
            sum_norm(row, n_dst);
            for (int k = 0; k < n_dst; ++k) {
                int32_t l = (-lm.log(row[k])) >> B200_SENSCR_SHIFT;
                if (l > 255) l = 255;
                out[((size_t)t * n_src + j) * n_dst + k] = (uint8_t)l;
            }
        }
    return B200_OK;
}

int b200_s3_read_gauden(const char *path, int32_t dims[4], int32_t *veclen, float *data) {
    S3File f;
    if (!s3_open(f, path)) return B200_ERR_IO;
    int32_t hdr[3];
    if (!s3_read32(f, hdr, 3)) { set_error("%s: short header", path); return B200_ERR_IO; }
    if (hdr[1] < 1 || hdr[1] > 64) { set_error("%s: bad n_feat %d", path, hdr[1]); return B200_ERR_IO; }
    std::vector<int32_t> vl(hdr[1]);
    if (!s3_read32(f, vl.data(), hdr[1])) { set_error("%s: short veclen", path); return B200_ERR_IO; }
    int32_t n;
    if (!s3_read32(f, &n, 1)) { set_error("%s: short count", path); return B200_ERR_IO; }
    long blk = 0;
    for (int i = 0; i < hdr[1]; ++i) blk += vl[i];
    if ((long)n != (long)hdr[0] * hdr[2] * blk) {
        set_error("%s: #floats(%d) doesn't match dimensions %d x %d x %ld", path, n, hdr[0], hdr[2], blk);
        return B200_ERR_IO;
    }
    dims[0] = hdr[0]; dims[1] = hdr[1]; dims[2] = hdr[2]; dims[3] = n;
    if (veclen) for (int i = 0; i < hdr[1]; ++i) veclen[i] = vl[i];
    if (!data) return B200_OK;
    if (!s3_read32(f, data, (size_t)n)) { set_error("%s: short data", path); return B200_ERR_IO; }
    return s3_finish(f, path) ? B200_OK : B200_ERR_IO;
}

static int read_4dim(const char *path, int32_t dims[4], float *data) {
    S3File f;
    if (!s3_open(f, path)) return B200_ERR_IO;
    if (!s3_read32(f, dims, 4)) { set_error("%s: short header", path); return B200_ERR_IO; }
    if ((long)dims[3] != (long)dims[0] * dims[1] * dims[2]) {
        set_error("%s: #floats(%d) doesn't match dimensions %d x %d x %d", path, dims[3], dims[0], dims[1], dims[2]);
        return B200_ERR_IO;
    }
    if (!data) return B200_OK;
    if (!s3_read32(f, data, (size_t)dims[3])) { set_error("%s: short data", path); return B200_ERR_IO; }
    return s3_finish(f, path) ? B200_OK : B200_ERR_IO;
}

int b200_s3_read_mixw(const char *path, int32_t dims[4], float *data) { return read_4dim(path, dims, data); }

int b200_s3_read_tmat(const char *path, int32_t dims[4], float *data) {
    int rc = read_4dim(path, dims, data);
    if (rc == B200_OK && dims[2] != dims[1] + 1) {
        set_error("%s: #from-states(%d) != #to-states(%d)-1", path, dims[1], dims[2]);
        return B200_ERR_IO;
    }
    return rc;
}

// sendump: length-prefixed strings, then optional cluster codebook, then rows.
int b200_s3_read_sendump(const char *path, int32_t dims[5], uint8_t *mixw, uint8_t cb[16]) {
    FILE *fp = fopen(path, "rb");
    if (!fp) { set_error("cannot open '%s'", path); return B200_ERR_IO; }
    struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{fp};
    bool swap = false;
    auto rd32 = [&](int32_t *v) -> bool {
        if (fread(v, 4, 1, fp) != 1) return false;
        if (swap) *v = (int32_t)bswap32((uint32_t)*v);
        return true;
    };
    int32_t n;
    if (fread(&n, 4, 1, fp) != 1) { set_error("%s: empty", path); return B200_ERR_IO; }
    if (n < 1 || n > 999) {
        n = (int32_t)bswap32((uint32_t)n);
        if (n < 1 || n > 999) { set_error("%s: title length %x out of range", path, n); return B200_ERR_IO; }
        swap = true;
    }
    std::vector<char> line(1000);
    if (fread(line.data(), 1, n, fp) != (size_t)n || line[n - 1] != '\0') { set_error("%s: bad title", path); return B200_ERR_IO; }
    if (!rd32(&n) || n < 1 || n > 999) { set_error("%s: bad header length", path); return B200_ERR_IO; }
    if (fread(line.data(), 1, n, fp) != (size_t)n || line[n - 1] != '\0') { set_error("%s: bad header", path); return B200_ERR_IO; }
    // Defaults come from the caller's model (dims[0..2] on input), as the
    // reference seeds them from s->n_feat / s->n_density / mdef n_sen.
    int n_feat = dims[0], n_density = dims[1], n_sen = dims[2], n_clust = 0, n_bits = 8;
    for (;;) {
        if (!rd32(&n)) { set_error("%s: truncated strings", path); return B200_ERR_IO; }
        if (n == 0) break;
        if (n < 0 || n > 999 || fread(line.data(), 1, n, fp) != (size_t)n) { set_error("%s: bad string", path); return B200_ERR_IO; }
        line[n < 999 ? n : 999] = '\0';
        sscanf(line.data(), "feature_count %d", &n_feat);
        sscanf(line.data(), "mixture_count %d", &n_density);
        sscanf(line.data(), "model_count %d", &n_sen);
        sscanf(line.data(), "cluster_count %d", &n_clust);
        sscanf(line.data(), "cluster_bits %d", &n_bits);
    }
    int32_t r = n_density, c = n_sen;
    if (n_clust == 0) {
        // Older files carry (possibly padded) #rows, #columns here.
        if (!rd32(&r) || !rd32(&c)) { set_error("%s: no rows/cols", path); return B200_ERR_IO; }
    }
    if (!(n_clust == 0 || n_clust == 15 || n_clust == 16)) { set_error("%s: cluster count must be 0, 15 or 16", path); return B200_ERR_IO; }
    if (n_clust == 15) n_clust = 16;
    if (!(n_bits == 8 || n_bits == 4)) { set_error("%s: cluster bits must be 4 or 8", path); return B200_ERR_IO; }
    if (r != n_density) { set_error("%s: padded row count %d != %d unsupported", path, r, n_density); return B200_ERR_UNSUP; }
    // (the mixing kernels index rows with a stride of n_sen: a padded column count would misalign every row)
    if (c != n_sen) { set_error("%s: padded column count %d != %d senones unsupported", path, c, n_sen); return B200_ERR_UNSUP; }
    int row_bytes = (n_bits == 4) ? (c + 1) / 2 : c;
    dims[0] = n_feat; dims[1] = n_density; dims[2] = n_sen; dims[3] = n_clust; dims[4] = row_bytes;
    if (!mixw) return B200_OK;
    if (n_clust) {
        if (fread(cb, 1, n_clust, fp) != (size_t)n_clust) { set_error("%s: short cluster codebook", path); return B200_ERR_IO; }
    }
    size_t total = (size_t)n_feat * n_density * row_bytes;
    if (fread(mixw, 1, total, fp) != total) { set_error("%s: short mixw data", path); return B200_ERR_IO; }
    return B200_OK;
}

int b200_flags2list(const uint32_t *mask, int n_sen, uint8_t *deltas, int max_out) {
    int n = 0, last = 0;
    for (int s = 0; s < n_sen; ++s) {
        if (!(mask[s >> 5] & (1u << (s & 31)))) continue;
        int delta = s - last;
        while (delta > 255) {
            if (n >= max_out) return B200_ERR_ARG;
            deltas[n++] = 255;
            delta -= 255;
        }
        if (n >= max_out) return B200_ERR_ARG;
        deltas[n++] = (uint8_t)delta;
        last = s;
    }
    return n;
}

}  // extern "C"

// ------------------------------------------------------------ .sen files
// Senone-score dumps (PS/acmod.c:349-361 header, :885-923 acmod_write_scores,
// :928-982 acmod_read_scores_internal): S3 header {version 0.1, mdef_file,
// n_sen, logbase}, then per frame int16 n_active and either n_sen int16 scores
// (all active) or n_active uint8 deltas + n_active int16 scores.
extern "C" int b200_sen_write(const char *path, const char *mdef_file, int n_sen, double logbase,
                              const int16_t *scores, int n_frames, const uint8_t *const *active,
                              const int32_t *n_active) {
    using namespace b200;
    if (!path || !scores || n_sen <= 0 || n_sen > 32767 || n_frames < 0) { set_error("b200_sen_write: bad argument"); return B200_ERR_ARG; }
    FILE *fp = fopen(path, "wb");
    if (!fp) { set_error("cannot create '%s'", path); return B200_ERR_IO; }
    fprintf(fp, "s3\nversion 0.1\nmdef_file %s\nn_sen %d\nlogbase %f\nendhdr\n", mdef_file ? mdef_file : "(null)", n_sen, logbase);
    const uint32_t magic = 0x11223344u;
    bool ok = fwrite(&magic, 4, 1, fp) == 1;
    for (int t = 0; t < n_frames && ok; ++t) {
        const int16_t *row = scores + (size_t)t * n_sen;
        const int na = (active && n_active && active[t]) ? n_active[t] : n_sen;
        const int16_t na16 = (int16_t)na;
        ok = fwrite(&na16, 2, 1, fp) == 1;
        if (na == n_sen) {
            ok = ok && fwrite(row, 2, (size_t)n_sen, fp) == (size_t)n_sen;
        } else {
            ok = ok && fwrite(active[t], 1, (size_t)na, fp) == (size_t)na;
            for (int i = 0, n = 0; i < na && ok; ++i) {
                n += active[t][i];
                if (n >= n_sen) { set_error("b200_sen_write: active list of frame %d leaves the senone range", t); fclose(fp); return B200_ERR_ARG; }
                ok = fwrite(row + n, 2, 1, fp) == 1;
            }
        }
    }
    if (fclose(fp) != 0) ok = false;
    if (!ok) { set_error("write to '%s' failed", path); return B200_ERR_IO; }
    return B200_OK;
}

extern "C" int b200_sen_read(const char *path, int32_t dims[2], double *logbase, int16_t *scores, int32_t *n_active) {
    using namespace b200;
    if (!path || !dims) { set_error("b200_sen_read: bad argument"); return B200_ERR_ARG; }
    FILE *fp = fopen(path, "rb");
    if (!fp) { set_error("cannot open '%s'", path); return B200_ERR_IO; }
    char line[16384], key[4096], val[4096];
    int n_sen = 0; double lb = 0.0; bool ok = true;
    if (!fgets(line, sizeof line, fp) || strcmp(line, "s3\n") != 0) { set_error("%s: not an s3 file", path); fclose(fp); return B200_ERR_IO; }
    for (;;) {
        if (!fgets(line, sizeof line, fp)) { ok = false; break; }
        if (sscanf(line, "%4095s", key) != 1) { ok = false; break; }
        if (strcmp(key, "endhdr") == 0) break;
        if (sscanf(line + strlen(key), "%4095s", val) != 1) continue;
        if (strcmp(key, "n_sen") == 0) n_sen = atoi(val);
        if (strcmp(key, "logbase") == 0) lb = atof(val);
    }
    uint32_t magic = 0;
    if (!ok || fread(&magic, 4, 1, fp) != 1 || n_sen <= 0) { set_error("%s: bad senone-dump header", path); fclose(fp); return B200_ERR_IO; }
    const bool swap = magic != 0x11223344u;
    if (swap && bswap32(magic) != 0x11223344u) { set_error("%s: bad byte-order magic", path); fclose(fp); return B200_ERR_IO; }
    auto sw16 = [&](int16_t v) { return swap ? (int16_t)__builtin_bswap16((uint16_t)v) : v; };
    const int max_frames = scores ? dims[1] : 0x7fffffff;
    int t = 0;
    std::vector<uint8_t> deltas;
    std::vector<int16_t> row((size_t)n_sen);
    for (; t < max_frames; ++t) {
        int16_t na16;
        if (fread(&na16, 2, 1, fp) != 1) break;
        const int na = sw16(na16);
        if (na < 0 || na > n_sen) { set_error("%s: frame %d has %d active senones", path, t, na); fclose(fp); return B200_ERR_IO; }
        std::fill(row.begin(), row.end(), (int16_t)0x7fff);     // SENSCR_DUMMY, PS/hmm.h
        if (na == n_sen) {
            if (fread(row.data(), 2, (size_t)n_sen, fp) != (size_t)n_sen) break;
            for (auto &v : row) v = sw16(v);
        } else {
            deltas.resize((size_t)na);
            if (na && fread(deltas.data(), 1, (size_t)na, fp) != (size_t)na) break;
            bool short_read = false;
            for (int i = 0, n = 0; i < na; ++i) {
                n += deltas[i];
                int16_t v;
                if (n >= n_sen || fread(&v, 2, 1, fp) != 1) { short_read = true; break; }
                row[n] = sw16(v);
            }
            if (short_read) break;
        }
        if (scores) std::copy(row.begin(), row.end(), scores + (size_t)t * n_sen);
        if (n_active) n_active[t] = na;
    }
    fclose(fp);
    dims[0] = n_sen; dims[1] = t;
    if (logbase) *logbase = lb;
    return B200_OK;
}

// ------------------------------------------------------------ mdef maps
// The two maps the scorers need from the model definition:
//   sen2cimap[s]  CI phone owning senone s   (ptm: senone -> codebook, PS/ptm_mgau.c:836-848)
//   cd2cisen[s]   CI senone at the same state position (sphinx3 CI-GMM selection,
//                 sphinx3/include/mdef.h:199-201; PS/bin_mdef.c:462-497)
// built exactly like bin_mdef_read (PS/bin_mdef.c:330-507): walk the phones in
// order, state position j of phone i -> senone sseq[ssid][j]; the FIRST phone
// that uses a senone defines its CI phone.  Reads binary "BMDF" files and the
// text format 0.3 (PS/mdef.c:505-602).
namespace {
struct MdefMaps { int n_ciphone = 0, n_emit = 0, n_ci_sen = 0, n_sen = 0; std::vector<int16_t> sen2ci, cd2ci; };

bool mdef_finish(MdefMaps &m, const std::vector<int32_t> &phone_ci, const std::vector<std::vector<int32_t>> &phone_sen) {
    m.sen2ci.assign((size_t)m.n_sen, -1);
    m.cd2ci.assign((size_t)m.n_sen, -1);
    for (int i = 0; i < m.n_ci_sen && i < m.n_sen; ++i) m.cd2ci[i] = (int16_t)i;
    for (size_t i = 0; i < phone_sen.size(); ++i) {
        const int ci = phone_ci[i];
        for (size_t j = 0; j < phone_sen[i].size(); ++j) {
            const int sn = phone_sen[i][j];
            if (sn < 0 || sn >= m.n_sen || ci < 0 || ci >= m.n_ciphone) { b200::set_error("mdef: senone/phone id out of range"); return false; }
            if (m.sen2ci[sn] == -1) m.sen2ci[sn] = (int16_t)ci;
            if (j < phone_sen[ci].size()) m.cd2ci[sn] = (int16_t)phone_sen[ci][j];
        }
    }
    return true;
}

bool mdef_read_bin(FILE *fh, bool swap, MdefMaps &m) {
    auto rd32 = [&](int32_t &v) { if (fread(&v, 4, 1, fh) != 1) return false; if (swap) v = (int32_t)b200::bswap32((uint32_t)v); return true; };
    int32_t ver, hdr, h[10];
    if (!rd32(ver) || !rd32(hdr) || fseek(fh, hdr, SEEK_CUR) != 0) return false;
    for (auto &v : h) if (!rd32(v)) return false;
    const int n_ciphone = h[0], n_phone = h[1], n_emit = h[2], n_sseq = h[6], n_cd_tree = h[8];
    m.n_ciphone = n_ciphone; m.n_emit = n_emit; m.n_ci_sen = h[3]; m.n_sen = h[4];
    if (n_emit <= 0) { b200::set_error("mdef: heterogeneous topologies are not supported"); return false; }
    const long pos = ftell(fh);
    fseek(fh, 0, SEEK_END);
    const long end = ftell(fh);
    fseek(fh, pos, SEEK_SET);
    std::vector<uint8_t> buf((size_t)(end - pos));
    if (fread(buf.data(), 1, buf.size(), fh) != buf.size()) return false;
    size_t o = 0;
    for (int i = 0; i < n_ciphone; ++i) { while (o < buf.size() && buf[o]) ++o; ++o; }
    o = (o + 3) & ~(size_t)3;
    o += (size_t)n_cd_tree * 8;                       // cd_tree_t: int16 ctx, int16 n_down, int32 down
    const size_t need = o + (size_t)n_phone * 12 + 4;
    if (need > buf.size()) return false;
    std::vector<int32_t> ssid((size_t)n_phone), ci((size_t)n_phone);
    for (int i = 0; i < n_phone; ++i) {
        int32_t v; memcpy(&v, &buf[o + (size_t)i * 12], 4);
        ssid[i] = swap ? (int32_t)b200::bswap32((uint32_t)v) : v;
        ci[i] = i < n_ciphone ? i : (int8_t)buf[o + (size_t)i * 12 + 9];   // info.cd.ctx[0]
    }
    o += (size_t)n_phone * 12;
    int32_t sseq_size; memcpy(&sseq_size, &buf[o], 4);
    if (swap) sseq_size = (int32_t)b200::bswap32((uint32_t)sseq_size);
    o += 4;
    if (o + (size_t)sseq_size * 2 > buf.size() || sseq_size < n_sseq * n_emit) return false;
    std::vector<std::vector<int32_t>> ps((size_t)n_phone);
    for (int i = 0; i < n_phone; ++i) {
        if (ssid[i] < 0 || ssid[i] >= n_sseq) return false;
        for (int j = 0; j < n_emit; ++j) {
            uint16_t v; memcpy(&v, &buf[o + ((size_t)ssid[i] * n_emit + j) * 2], 2);
            if (swap) v = __builtin_bswap16(v);
            ps[i].push_back(v);
        }
    }
    return mdef_finish(m, ci, ps);
}

bool mdef_read_text(FILE *fh, MdefMaps &m) {
    char line[16384];
    int n_base = -1, n_tri = -1, n_state_map = -1;
    bool versioned = false;
    std::vector<std::string> ciname;
    std::vector<int32_t> ci;
    std::vector<std::vector<int32_t>> ps;
    while (fgets(line, sizeof line, fh)) {
        char *p = line;
        while (*p == ' ' || *p == '\t') ++p;
        if (*p == '#' || *p == '\n' || *p == 0) continue;
        if (!versioned) { versioned = true; if (strncmp(p, "0.3", 3) != 0) { b200::set_error("mdef: text version is not 0.3"); return false; } continue; }
        int v; char key[64];
        if (n_base < 0 || n_tri < 0 || n_state_map < 0 || m.n_sen == 0 || m.n_ci_sen == 0 || m.n_ciphone == -7) {
            if (sscanf(p, "%d %63s", &v, key) == 2) {
                if (!strcmp(key, "n_base")) { n_base = v; continue; }
                if (!strcmp(key, "n_tri")) { n_tri = v; continue; }
                if (!strcmp(key, "n_state_map")) { n_state_map = v; continue; }
                if (!strcmp(key, "n_tied_state")) { m.n_sen = v; continue; }
                if (!strcmp(key, "n_tied_ci_state")) { m.n_ci_sen = v; continue; }
                if (!strcmp(key, "n_tied_tmat")) continue;
            }
        }
        // phone line: base lft rt p attrib tmat s0 s1 ... N
        char base[256], lft[256], rt[256], wpos[64], attrib[256];
        int tmat, used = 0;
        if (sscanf(p, "%255s %255s %255s %63s %255s %d%n", base, lft, rt, wpos, attrib, &tmat, &used) != 6) continue;
        std::vector<int32_t> st;
        char *q = p + used;
        for (;;) {
            while (*q == ' ' || *q == '\t') ++q;
            if (*q == 'N' || *q == 0 || *q == '\n') break;
            st.push_back((int32_t)strtol(q, &q, 10));
        }
        int cid;
        if ((int)ciname.size() < n_base) { cid = (int)ciname.size(); ciname.push_back(base); }
        else {
            cid = -1;
            for (size_t k = 0; k < ciname.size(); ++k) if (ciname[k] == base) { cid = (int)k; break; }
            if (cid < 0) { b200::set_error("mdef: triphone of unknown base phone %s", base); return false; }
        }
        ci.push_back(cid);
        ps.push_back(st);
    }
    if (n_base <= 0 || ps.empty()) { b200::set_error("mdef: no phones found"); return false; }
    m.n_ciphone = n_base; m.n_emit = (int)ps[0].size();
    return mdef_finish(m, ci, ps);
}
}  // namespace

extern "C" int b200_mdef_read_maps(const char *path, int32_t dims[4], int16_t *sen2cimap, int16_t *cd2cisen) {
    using namespace b200;
    if (!path || !dims) { set_error("b200_mdef_read_maps: bad argument"); return B200_ERR_ARG; }
    FILE *fh = fopen(path, "rb");
    if (!fh) { set_error("cannot open '%s'", path); return B200_ERR_IO; }
    uint32_t magic = 0;
    MdefMaps m;
    bool ok;
    if (fread(&magic, 4, 1, fh) == 1 && (magic == 0x46444d42u || magic == 0x424d4446u)) {
        ok = mdef_read_bin(fh, magic == 0x424d4446u, m);
    } else {
        rewind(fh);
        ok = mdef_read_text(fh, m);
    }
    fclose(fh);
    if (!ok) { if (g_last_error.empty() || g_last_error.find("mdef") == std::string::npos) set_error("%s: malformed model definition", path); return B200_ERR_IO; }
    dims[0] = m.n_sen; dims[1] = m.n_ci_sen; dims[2] = m.n_ciphone; dims[3] = m.n_emit;
    if (sen2cimap) std::copy(m.sen2ci.begin(), m.sen2ci.end(), sen2cimap);
    if (cd2cisen) std::copy(m.cd2ci.begin(), m.cd2ci.end(), cd2cisen);
    return B200_OK;
}


// ---------------------------------------------------------------- hmm_t AoS <-> SoA
extern "C" int b200_hmm_pack(const void *const *hmms, int n, int n_emit, b200_hmm_soa_t *soa) {
    if (n < 0 || n_emit < 1 || n_emit > 5 || !soa || (n > 0 && (!hmms || !soa->score || !soa->history || !soa->out_score ||
        !soa->out_history || !soa->senid || !soa->tmatid || !soa->mpx || !soa->bestscore))) {
        b200::set_error("b200_hmm_pack: bad argument");
        return B200_ERR_ARG;
    }
    const size_t N = (size_t)n;
    for (int i = 0; i < n; ++i) {
        const b200_ps_hmm_t *h = static_cast<const b200_ps_hmm_t *>(hmms[i]);
        if (!h || h->n_emit_state != n_emit) { b200::set_error("b200_hmm_pack: hmm %d has %d emitting states, context has %d", i, h ? h->n_emit_state : -1, n_emit); return B200_ERR_ARG; }
        for (int s = 0; s < n_emit; ++s) {
            soa->score[s * N + i] = h->score[s];
            soa->history[s * N + i] = h->history[s];
            soa->senid[s * N + i] = h->senid[s];
        }
        soa->out_score[i] = h->out_score;
        soa->out_history[i] = h->out_history;
        soa->tmatid[i] = h->tmatid;
        soa->mpx[i] = h->mpx;
        soa->bestscore[i] = h->bestscore;
    }
    soa->n_hmm = n;
    return B200_OK;
}

extern "C" void b200_hmm_unpack_one(const b200_hmm_soa_t *soa, int n_emit, int i, void *hmm) {
    b200_ps_hmm_t *h = static_cast<b200_ps_hmm_t *>(hmm);
    const size_t N = (size_t)soa->n_hmm;
    for (int s = 0; s < n_emit; ++s) {
        h->score[s] = soa->score[s * N + i];
        h->history[s] = soa->history[s * N + i];
        if (h->mpx) h->senid[s] = soa->senid[s * N + i];
    }
    h->out_score = soa->out_score[i];
    h->out_history = soa->out_history[i];
    h->bestscore = soa->bestscore[i];
}

extern "C" int b200_hmm_unpack(const b200_hmm_soa_t *soa, int n_emit, void *const *hmms) {
    if (!soa || n_emit < 1 || n_emit > 5 || (soa->n_hmm > 0 && !hmms)) { b200::set_error("b200_hmm_unpack: bad argument"); return B200_ERR_ARG; }
    for (int i = 0; i < soa->n_hmm; ++i) b200_hmm_unpack_one(soa, n_emit, i, hmms[i]);
    return B200_OK;
}
