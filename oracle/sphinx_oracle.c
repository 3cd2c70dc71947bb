/* oracle/sphinx_oracle.c -- TEST INFRASTRUCTURE ONLY (see sphinx_oracle.h).
 *
 * Plain-C restatement of the pocketsphinx hot path, written from the
 * reference's behaviour (file:line cited per function), never linked into the
 * product.  Pinned against the reference's own compiled code (oracle/_ref,
 * tests/test_oracle_vs_ref.py) and the golden vectors in tests/golden/.
 */
#include "sphinx_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SHIFT ORC_SENSCR_SHIFT

/* ------------------------------------------------------------------ logmath
 * sphinxbase/src/libsphinxbase/util/logmath.c:61-161 (init), :391-436 (add),
 * :446-465 (log, ln_to_log). */
orc_logmath_t *
orc_logmath_init(double base, int shift, int use_table)
{
    orc_logmath_t *lm;
    uint32_t i, n;
    double byx;

    if (base <= 1.0)
        return NULL;
    lm = calloc(1, sizeof(*lm));
    lm->base = base;
    lm->inv_log_of_base = 1.0 / log(base);
    lm->shift = shift;
    lm->zero = INT_MIN >> (shift + 2);
    if (!use_table)
        return lm;
    for (i = 0, byx = 1.0;; ++i, byx /= base) {
        int32_t k = (int32_t)(log(1.0 + byx) * lm->inv_log_of_base + 0.5 * (1 << shift)) >> shift;
        if (k <= 0)
            break;
    }
    n = i >> shift;
    if (n < 255)
        n = 255;
    lm->table_size = n + 1;
    lm->table = calloc(lm->table_size, sizeof(uint32_t));
    {
        uint32_t maxyx = (uint32_t)(log(2.0) / log(base) + 0.5) >> shift;
        uint32_t mask = maxyx < 256 ? 0xffu : (maxyx < 65536 ? 0xffffu : 0xffffffffu);
        for (i = 0, byx = 1.0;; ++i, byx /= base) {
            int32_t k = (int32_t)(log(1.0 + byx) * lm->inv_log_of_base + 0.5 * (1 << shift)) >> shift;
            if (lm->table[i >> shift] == 0)
                lm->table[i >> shift] = (uint32_t)k & mask;
            if (k <= 0)
                break;
        }
    }
    return lm;
}

void
orc_logmath_free(orc_logmath_t *lm)
{
    if (lm) {
        free(lm->table);
        free(lm);
    }
}

int
orc_logmath_table(const orc_logmath_t *lm, int32_t *out, int max_out)
{
    uint32_t i;
    for (i = 0; i < lm->table_size && (int)i < max_out; ++i)
        out[i] = (int32_t)lm->table[i];
    return (int)lm->table_size;
}

int32_t
orc_logmath_log(const orc_logmath_t *lm, double p)
{
    if (p <= 0)
        return lm->zero;
    return (int32_t)(log(p) * lm->inv_log_of_base) >> lm->shift;
}

int32_t
orc_logmath_ln_to_log(const orc_logmath_t *lm, double ln_p)
{
    return (int32_t)(ln_p * lm->inv_log_of_base) >> lm->shift;
}

int32_t
orc_logmath_add(const orc_logmath_t *lm, int32_t x, int32_t y)
{
    int32_t d, r;
    if (x <= lm->zero)
        return y;
    if (y <= lm->zero)
        return x;
    if (x > y) { d = x - y; r = x; }
    else { d = y - x; r = y; }
    if (d < 0 || (uint32_t)d >= lm->table_size)
        return r;
    return r + (int32_t)lm->table[d];
}

/* tied_mgau_common.h:104-121.  NB the reference does NOT range-check d: with
 * ptm's min-normalisation (ptm_mgau.c:267-288) d can exceed the 256-entry
 * table and the reference reads past it (heap-dependent garbage).  The oracle
 * returns the intended value (correction 0 beyond the table), so oracle ==
 * reference only on frames where the reference stays inside its table. */
static int
fast_add(const orc_logmath_t *lm, int mlx, int mly)
{
    int d, r;
    if (mlx > mly) { d = mlx - mly; r = mly; }
    else { d = mly - mlx; r = mlx; }
    return r - ((unsigned)d < lm->table_size ? (int)lm->table[d] : 0);
}

/* ------------------------------------------------------ load-time precompute */
/* pocketsphinx/src/libpocketsphinx/vector.c:90-128 */
static double
sum_norm(float *v, int n)
{
    double s = 0.0;
    int i;
    for (i = 0; i < n; ++i)
        s += v[i];
    if (s != 0.0) {
        double f = 1.0 / s;
        for (i = 0; i < n; ++i)
            v[i] *= f;
    }
    return s;
}

/* ms_gauden.c:314-359 */
int
orc_gauden_precompute(float *var, float *det, long n_vec, int len, float varfloor, double logbase)
{
    orc_logmath_t *lm = orc_logmath_init(logbase, 0, 0);
    long v;
    int i;
    for (v = 0; v < n_vec; ++v) {
        float *vp = var + v * len;
        det[v] = 0;
        for (i = 0; i < len; ++i) {
            if (vp[i] < varfloor)
                vp[i] = varfloor;
            det[v] += (float)orc_logmath_log(lm, 1.0 / sqrt(vp[i] * 2.0 * M_PI));
            vp[i] = (float)orc_logmath_ln_to_log(lm, 1.0 / (vp[i] * 2.0));
        }
    }
    orc_logmath_free(lm);
    return 0;
}

/* ms_senone.c:236-258 */
int
orc_mixw_quantize(float *mixw, uint8_t *out, int n_sen, int n_feat, int n_cw, float mixwfloor, double logbase)
{
    orc_logmath_t *lm = orc_logmath_init(logbase, 0, 0);
    long r, nr = (long)n_sen * n_feat;
    int c;
    for (r = 0; r < nr; ++r) {
        float *pdf = mixw + r * n_cw;
        sum_norm(pdf, n_cw);
        for (c = 0; c < n_cw; ++c)
            if (pdf[c] < mixwfloor)
                pdf[c] = mixwfloor;
        sum_norm(pdf, n_cw);
        for (c = 0; c < n_cw; ++c) {
            int32_t p = -orc_logmath_log(lm, pdf[c]) + (1 << (SHIFT - 1)) - 1;
            out[r * n_cw + c] = (p < (255 << SHIFT)) ? (uint8_t)(p >> SHIFT) : 255;
        }
    }
    orc_logmath_free(lm);
    return 0;
}

/* ptm_mgau.c:720-742 / s2_semi_mgau.c:1155-1177 */
int
orc_mixw_quantize_tied(float *mixw, uint8_t *out, int n_sen, int n_feat, int n_cw, float mixwfloor, double logbase)
{
    orc_logmath_t *lm8 = orc_logmath_init(logbase, SHIFT, 0);
    int s, f, c;
    for (s = 0; s < n_sen; ++s)
        for (f = 0; f < n_feat; ++f) {
            float *pdf = mixw + ((long)s * n_feat + f) * n_cw;
            sum_norm(pdf, n_cw);
            for (c = 0; c < n_cw; ++c)
                if (pdf[c] < mixwfloor)
                    pdf[c] = mixwfloor;
            sum_norm(pdf, n_cw);
            for (c = 0; c < n_cw; ++c) {
                int32_t q = -orc_logmath_log(lm8, pdf[c]);
                if (q > 159 || q < 0)
                    q = 159;
                out[((long)f * n_cw + c) * n_sen + s] = (uint8_t)q;
            }
        }
    orc_logmath_free(lm8);
    return 0;
}

/* tmat.c:275-296 */
int
orc_tmat_quantize(float *tp, uint8_t *out, int n_tmat, int n_src, double tpfloor, double logbase)
{
    orc_logmath_t *lm = orc_logmath_init(logbase, 0, 0);
    int t, j, k, n_dst = n_src + 1;
    for (t = 0; t < n_tmat; ++t)
        for (j = 0; j < n_src; ++j) {
            float *row = tp + ((long)t * n_src + j) * n_dst;
            sum_norm(row, n_dst);
            for (k = 0; k < n_dst; ++k)
                if (row[k] != 0.0 && row[k] < tpfloor)
                    row[k] = (float)tpfloor;
            sum_norm(row, n_dst);
            for (k = 0; k < n_dst; ++k) {
                int ltp = -orc_logmath_log(lm, row[k]) >> SHIFT;
                if (ltp > 255)
                    ltp = 255;
                out[((long)t * n_src + j) * n_dst + k] = (uint8_t)ltp;
            }
        }
    orc_logmath_free(lm);
    return 0;
}

/* acmod.c:1219-1271 */
int
orc_flags2list(const uint32_t *mask, int n_sen, uint8_t *deltas)
{
    int s, n = 0, l = 0;
    for (s = 0; s < n_sen; ++s) {
        int delta;
        if (!(mask[s / 32] & (1u << (s % 32))))
            continue;
        delta = s - l;
        while (delta > 255) {
            deltas[n++] = 255;
            delta -= 255;
        }
        deltas[n++] = (uint8_t)delta;
        l = s;
    }
    return n;
}

/* ------------------------------------------------------------ ms back-end */
orc_ms_model_t *
orc_ms_model_new(int n_mgau, int n_feat, const int *featlen, int n_density, int n_sen, int topn, int aw,
                 const float *mean, const float *var, const float *det, const uint8_t *mixw,
                 const uint32_t *sen2mgau, double logbase)
{
    orc_ms_model_t *m = calloc(1, sizeof(*m));
    int f;
    m->n_mgau = n_mgau; m->n_feat = n_feat; m->n_density = n_density; m->n_sen = n_sen;
    m->topn = (topn == 0 || topn > n_density) ? n_density : topn;   /* ms_mgau.c:121-127 */
    m->aw = aw > 0 ? aw : 1;
    for (f = 0; f < n_feat; ++f) {
        m->featlen[f] = featlen[f];
        m->featoff[f] = m->veclen;
        m->veclen += featlen[f];
    }
    m->mean = mean; m->var = var; m->det = det; m->mixw = mixw; m->sen2mgau = sen2mgau;
    m->lmath10 = orc_logmath_init(logbase, SHIFT, 1);
    return m;
}

void
orc_ms_model_free(orc_ms_model_t *m)
{
    if (m) {
        orc_logmath_free(m->lmath10);
        free(m);
    }
}

/* ms_gauden.c:417-523: one codebook, one stream. */
void
orc_ms_compute_dist(const orc_ms_model_t *m, int mgau, int feat, const float *obs, int32_t *ids, float *dists)
{
    int len = m->featlen[feat], nd = m->n_density, n = m->topn, d, i, j;
    const float *mp = m->mean + (long)mgau * nd * m->veclen + (long)nd * m->featoff[feat];
    const float *vp = m->var + (long)mgau * nd * m->veclen + (long)nd * m->featoff[feat];
    const float *dp = m->det + ((long)mgau * m->n_feat + feat) * nd;

    if (n >= nd) {   /* compute_dist_all: everything, index order */
        for (d = 0; d < nd; ++d) {
            float dval = dp[d];
            for (i = 0; i < len; ++i) {
                float diff = obs[i] - mp[d * len + i];
                dval -= diff * diff * vp[d * len + i];
            }
            dists[d] = dval;
            ids[d] = d;
        }
        return;
    }
    for (i = 0; i < n; ++i) {
        dists[i] = (float)ORC_WORST_DIST;
        ids[i] = 0;
    }
    for (d = 0; d < nd; ++d) {
        float dval = dp[d];
        for (i = 0; i < len && dval >= dists[n - 1]; ++i) {
            float diff = obs[i] - mp[d * len + i];
            dval -= diff * diff * vp[d * len + i];
        }
        if (i < len || dval < dists[n - 1])
            continue;
        for (i = 0; i < n && dval < dists[i]; ++i)
            ;
        for (j = n - 1; j > i; --j) {
            dists[j] = dists[j - 1];
            ids[j] = ids[j - 1];
        }
        dists[i] = dval;
        ids[i] = d;
    }
}

/* ms_senone.c:372-421 */
static int32_t
ms_senone_eval(const orc_ms_model_t *m, int s, const int32_t *ids, const float *dists)
{
    int n = m->topn, f, t;
    int32_t scr = 0;
    for (f = 0; f < m->n_feat; ++f) {
        const int32_t *fi = ids + f * n;
        const float *fd = dists + f * n;
        const uint8_t *pdf = m->mixw + ((long)s * m->n_feat + f) * m->n_density;
        int32_t fden = ((int32_t)fd[0] + ((1 << SHIFT) - 1)) >> SHIFT;
        int32_t fscr = fden - pdf[fi[0]];
        for (t = 1; t < n; ++t) {
            fden = ((int32_t)fd[t] + ((1 << SHIFT) - 1)) >> SHIFT;
            fscr = orc_logmath_add(m->lmath10, fscr, fden - pdf[fi[t]]);
        }
        scr -= fscr;
    }
    scr /= m->aw;
    if (scr > 32767) scr = 32767;
    if (scr < -32768) scr = -32768;
    return scr;
}

/* ms_mgau.c:162-252 */
int
orc_ms_frame_eval(const orc_ms_model_t *m, const float *feat, const uint8_t *senone_active,
                  int n_senone_active, int compallsen, int16_t *senscr)
{
    int n = m->topn, nf = m->n_feat, g, f, s, i, l;
    int32_t *ids = malloc(sizeof(int32_t) * (size_t)m->n_mgau * nf * n);
    float *dists = malloc(sizeof(float) * (size_t)m->n_mgau * nf * n);
    char *act = calloc(m->n_mgau, 1);
    int32_t best = 0x7fffffff;

    if (compallsen)
        memset(act, 1, m->n_mgau);
    else
        for (i = 0, l = 0; i < n_senone_active; ++i) {
            l += senone_active[i];
            act[m->sen2mgau[l]] = 1;
        }
    for (g = 0; g < m->n_mgau; ++g)
        if (act[g])
            for (f = 0; f < nf; ++f)
                orc_ms_compute_dist(m, g, f, feat + m->featoff[f], ids + ((long)g * nf + f) * n,
                                    dists + ((long)g * nf + f) * n);
    if (compallsen) {
        for (s = 0; s < m->n_sen; ++s) {
            long o = (long)m->sen2mgau[s] * nf * n;
            senscr[s] = (int16_t)ms_senone_eval(m, s, ids + o, dists + o);
            if (best > senscr[s]) best = senscr[s];
        }
        for (s = 0; s < m->n_sen; ++s) {
            int32_t bs = senscr[s] - best;
            senscr[s] = (int16_t)(bs > 32767 ? 32767 : (bs < -32768 ? -32768 : bs));
        }
    }
    else {
        for (i = 0, s = 0; i < n_senone_active; ++i) {
            long o;
            s += senone_active[i];
            o = (long)m->sen2mgau[s] * nf * n;
            senscr[s] = (int16_t)ms_senone_eval(m, s, ids + o, dists + o);
            if (best > senscr[s]) best = senscr[s];
        }
        for (i = 0, s = 0; i < n_senone_active; ++i) {
            int32_t bs;
            s += senone_active[i];
            bs = senscr[s] - best;
            senscr[s] = (int16_t)(bs > 32767 ? 32767 : (bs < -32768 ? -32768 : bs));
        }
    }
    free(ids); free(dists); free(act);
    return 0;
}

int
orc_ms_eval_all(const orc_ms_model_t *m, const float *feat, int T, int16_t *out)
{
    int t;
    for (t = 0; t < T; ++t)
        orc_ms_frame_eval(m, feat + (long)t * m->veclen, NULL, 0, 1, out + (long)t * m->n_sen);
    return 0;
}

/* ------------------------------------------------------- ptm / s2_semi */
typedef struct { int32_t cw, score; } topn_t;

struct orc_tied_model {
    int kind, n_mgau, n_feat, n_density, n_sen, topn, veclen;
    int featlen[8], featoff[8];
    const float *mean, *var, *det;
    const uint8_t *mixw, *mixw_cb, *sen2cb;
    int row_bytes, n_clust;
    orc_logmath_t *lm8;
    topn_t *hist[2];          /* [n_mgau][n_feat][topn] x 2 slots */
    topn_t *f;                /* current slot */
    char *cb_active;
    int frame_idx;
    int topn_beam[8];         /* s2_semi -topn_beam per stream, 0 = off */
    int hist_n[2][8];         /* mgau_norm's return value per slot and stream (topn_hist_n) */
    int ds_ratio;             /* s2_semi -ds (0/1 = every frame) */
};

orc_tied_model_t *
orc_tied_new(int kind, int n_mgau, int n_feat, const int *featlen, int n_density, int n_sen, int topn,
             const float *mean, const float *var, const float *det, const uint8_t *mixw, int row_bytes,
             int n_clust, const uint8_t *mixw_cb, const uint8_t *sen2cb, double logbase)
{
    orc_tied_model_t *m = calloc(1, sizeof(*m));
    int f;
    m->kind = kind; m->n_mgau = n_mgau; m->n_feat = n_feat; m->n_density = n_density;
    m->n_sen = n_sen; m->topn = topn;
    for (f = 0; f < n_feat; ++f) {
        m->featlen[f] = featlen[f];
        m->featoff[f] = m->veclen;
        m->veclen += featlen[f];
    }
    m->mean = mean; m->var = var; m->det = det; m->mixw = mixw; m->row_bytes = row_bytes;
    m->n_clust = n_clust; m->mixw_cb = mixw_cb; m->sen2cb = sen2cb;
    m->lm8 = orc_logmath_init(logbase, SHIFT, 1);
    m->hist[0] = malloc(sizeof(topn_t) * (size_t)n_mgau * n_feat * topn);
    m->hist[1] = malloc(sizeof(topn_t) * (size_t)n_mgau * n_feat * topn);
    m->cb_active = malloc(n_mgau);
    orc_tied_reset(m);
    return m;
}

void
orc_tied_free(orc_tied_model_t *m)
{
    if (!m) return;
    orc_logmath_free(m->lm8);
    free(m->hist[0]); free(m->hist[1]); free(m->cb_active); free(m);
}

/* s2_semi_mgau.c:1300-1303: -topn_beam per stream (ignored by ptm, which never reads it) */
void
orc_tied_set_topn_beam(orc_tied_model_t *m, const int *beam)
{
    int f;
    for (f = 0; f < m->n_feat; ++f) m->topn_beam[f] = m->kind == 1 ? 0 : beam[f];
}

/* s2_semi_mgau.c:1298: -ds (the ptm flavour of -ds is undefined in the reference and not restated) */
void
orc_tied_set_ds(orc_tied_model_t *m, int ds_ratio)
{
    m->ds_ratio = m->kind == 1 ? 1 : ds_ratio;
}

/* ptm_mgau.c:846-865, s2_semi_mgau.c:1313-1324 */
void
orc_tied_reset(orc_tied_model_t *m)
{
    int h, i, k, n = m->n_mgau * m->n_feat;
    for (h = 0; h < 2; ++h)
        for (i = 0; i < n; ++i)
            for (k = 0; k < m->topn; ++k) {
                m->hist[h][i * m->topn + k].cw = k;
                m->hist[h][i * m->topn + k].score = ORC_WORST_DIST;
            }
    m->f = m->hist[0];
    m->frame_idx = 0;
    memset(m->cb_active, 1, m->n_mgau);
}

static float
tied_dist(const orc_tied_model_t *m, int cb, int feat, int cw, const float *z)
{
    int len = m->featlen[feat], i;
    long base = (long)cb * m->n_density * m->veclen + (long)m->n_density * m->featoff[feat] + (long)cw * len;
    float d = m->det[((long)cb * m->n_feat + feat) * m->n_density + cw];
    for (i = 0; i < len; ++i) {
        float diff = z[i] - m->mean[base + i];
        float sq = diff * diff;
        float c = sq * m->var[base + i];
        d = d - c;
    }
    return d;
}

/* ptm_mgau.c:98-144, s2_semi_mgau.c:80-117: re-score last frame's codewords */
static void
tied_eval_topn(orc_tied_model_t *m, int cb, int feat, const float *z)
{
    topn_t *topn = m->f + ((long)cb * m->n_feat + feat) * m->topn;
    int i, j;
    for (i = 0; i < m->topn; ++i) {
        topn_t v;
        int32_t d = (int32_t)tied_dist(m, cb, feat, topn[i].cw, z);
        topn[i].score = d;
        if (i == 0)
            continue;
        v = topn[i];
        for (j = i - 1; j >= 0 && d > topn[j].score; --j)
            topn[j + 1] = topn[j];
        topn[j + 1] = v;
    }
}

/* ptm_mgau.c:159-231, s2_semi_mgau.c:119-173: scan the codebook.  The
 * dimension loop's early exit only skips work (partial sums decrease
 * monotonically) except that it is evaluated on the value BEFORE each
 * dimension (ptm: before each group of four after the first len%4), which
 * is what the literal loop below reproduces. */
static void
tied_eval_cb(orc_tied_model_t *m, int cb, int feat, const float *z)
{
    topn_t *topn = m->f + ((long)cb * m->n_feat + feat) * m->topn;
    topn_t *worst = topn + m->topn - 1;
    int len = m->featlen[feat], cw, i, j;
    long pbase = (long)cb * m->n_density * m->veclen + (long)m->n_density * m->featoff[feat];
    const float *det = m->det + ((long)cb * m->n_feat + feat) * m->n_density;

    for (cw = 0; cw < m->n_density; ++cw) {
        const float *mean = m->mean + pbase + (long)cw * len;
        const float *var = m->var + pbase + (long)cw * len;
        float d = det[cw];
        topn_t *cur;
        if (m->kind == 1) {
            float thresh = (float)worst->score;
            for (j = 0; j < len % 4 && d >= thresh; ++j) {
                float diff = z[j] - mean[j];
                d = d - (diff * diff) * var[j];
            }
            for (; j < len && d >= thresh; j += 4) {
                float c0, c1, c2, c3, df;
                df = z[j] - mean[j];         c0 = (df * df) * var[j];
                df = z[j + 1] - mean[j + 1]; c1 = (df * df) * var[j + 1];
                df = z[j + 2] - mean[j + 2]; c2 = (df * df) * var[j + 2];
                df = z[j + 3] - mean[j + 3]; c3 = (df * df) * var[j + 3];
                d = d - c0; d = d - c1; d = d - c2; d = d - c3;
            }
            if (j < len)
                continue;
            if (d < thresh)
                continue;
        }
        else {
            for (j = 0; j < len && d >= worst->score; ++j) {
                float diff = z[j] - mean[j];
                d = d - (diff * diff) * var[j];
            }
            if (j < len)
                continue;
            if ((int32_t)d < worst->score)
                continue;
        }
        for (i = 0; i < m->topn; ++i)
            if (topn[i].cw == cw)
                break;
        if (i < m->topn)
            continue;
        for (cur = worst - 1; cur >= topn && (int32_t)d >= cur->score; --cur)
            cur[1] = cur[0];
        ++cur;
        cur->cw = cw;
        cur->score = (int32_t)d;
    }
}

static int
tied_mixw(const orc_tied_model_t *m, int f, int cw, int sen)
{
    const uint8_t *row = m->mixw + ((long)f * m->n_density + cw) * m->row_bytes;
    if (m->n_clust) {
        int b = row[sen / 2];
        return m->mixw_cb[(sen & 1) ? (b >> 4) : (b & 0x0f)];
    }
    return row[sen];
}

/* ptm_mgau.c:405-450 + :236-400; s2_semi_mgau.c:840-886 + :189-207 + get_scores_* */
int
orc_tied_frame_eval(orc_tied_model_t *m, const float *feat, const uint8_t *senone_active,
                    int n_senone_active, int compallsen, int frame, int16_t *senscr)
{
    int C = m->n_mgau, F = m->n_feat, N = m->topn, i, j, k, f, l;
    size_t lbytes = sizeof(topn_t) * (size_t)C * F * N;
    topn_t *lastf;

    m->f = m->hist[frame % 2];
    lastf = m->hist[(frame + 1) % 2];
    if (frame >= m->frame_idx) {
        memcpy(m->f, lastf, lbytes);
        if (m->kind == 1) {
            if (compallsen)
                memset(m->cb_active, 1, C);
            else {
                memset(m->cb_active, 0, C);
                for (i = 0, l = 0; i < n_senone_active; ++i) {
                    l += senone_active[i];
                    m->cb_active[m->sen2cb[l]] = 1;
                }
            }
            for (i = 0; i < C; ++i)
                for (j = 0; j < F; ++j)
                    tied_eval_topn(m, i, j, feat + m->featoff[j]);
            for (i = 0; i < C; ++i)
                if (m->cb_active[i])
                    for (j = 0; j < F; ++j)
                        tied_eval_cb(m, i, j, feat + m->featoff[j]);
            for (j = 0; j < F; ++j) {
                int32_t norm = 0x7fffffff;
                for (i = 0; i < C; ++i)
                    if (m->cb_active[i] && norm > (m->f[((long)i * F + j) * N].score >> SHIFT))
                        norm = m->f[((long)i * F + j) * N].score >> SHIFT;
                for (i = 0; i < C; ++i) {
                    if (!m->cb_active[i])
                        continue;
                    for (k = 0; k < N; ++k) {
                        topn_t *e = &m->f[((long)i * F + j) * N + k];
                        e->score = -((e->score >> SHIFT) - norm);
                        if (e->score > 96)
                            e->score = 96;
                    }
                }
            }
        }
        else {
            for (j = 0; j < F; ++j) {
                topn_t *t = m->f + (long)j * N;
                int32_t norm;
                tied_eval_topn(m, 0, j, feat + m->featoff[j]);
                if (m->ds_ratio <= 1 || frame % m->ds_ratio == 0)   /* s2_semi_mgau.c:182-183 */
                    tied_eval_cb(m, 0, j, feat + m->featoff[j]);
                norm = t[0].score >> SHIFT;
                for (k = 0; k < N; ++k) {
                    t[k].score = -((t[k].score >> SHIFT) - norm);
                    if (t[k].score > 96)
                        t[k].score = 96;
                    if (m->topn_beam[j] && t[k].score > m->topn_beam[j])   /* s2_semi_mgau.c:203-204 */
                        break;
                }
                m->hist_n[frame % 2][j] = k;
            }
        }
        m->frame_idx = frame + 1;   /* what acmod_advance does (acmod.c:880) */
    }

    memset(senscr, 0, sizeof(int16_t) * m->n_sen);
    {
        int n = compallsen ? m->n_sen : n_senone_active;
        int32_t best = 0x7fffffff;
        for (i = 0, l = 0; i < n; ++i) {
            int sen, cb;
            int32_t ascore = 0;
            if (compallsen) sen = i;
            else { l += senone_active[i]; sen = l; }
            cb = m->kind == 1 ? m->sen2cb[sen] : 0;
            if (m->kind == 1 && !m->cb_active[cb])
                for (f = 0; f < F; ++f)
                    for (k = 0; k < N; ++k)
                        m->f[((long)cb * F + f) * N + k].score = 96;
            for (f = 0; f < F; ++f) {
                const topn_t *t = m->f + ((long)cb * F + f) * N;
                int fden = 0;
                const int nf = m->kind == 1 ? N : m->hist_n[frame % 2][f];
                for (k = 0; k < nf; ++k) {
                    int w = tied_mixw(m, f, t[k].cw, sen) + t[k].score;
                    fden = (k == 0) ? w : fast_add(m->lm8, fden, w);
                }
                ascore += fden;
            }
            if (m->kind == 1) {
                if (ascore < best) best = ascore;
                senscr[sen] = (int16_t)ascore;
            }
            else
                senscr[sen] = (int16_t)(senscr[sen] + ascore);
        }
        if (m->kind == 1)
            for (i = 0; i < m->n_sen; ++i)
                senscr[i] = (int16_t)(senscr[i] - best);
    }
    return 0;
}

int
orc_tied_eval_all(orc_tied_model_t *m, const float *feat, int T, int16_t *out)
{
    int t;
    for (t = 0; t < T; ++t)
        orc_tied_frame_eval(m, feat + (long)t * m->veclen, NULL, 0, 1, t, out + (long)t * m->n_sen);
    return 0;
}

void
orc_tied_lists(const orc_tied_model_t *m, int32_t *cw, int32_t *score)
{
    long i, n = (long)m->n_mgau * m->n_feat * m->topn;
    for (i = 0; i < n; ++i) {
        cw[i] = m->f[i].cw;
        score[i] = m->f[i].score;
    }
}

/* ------------------------------------------------------------------- HMM
 * hmm.c:224-807, restated over HMM-major arrays. */
#define W ORC_WORST_SCORE

static int32_t
clampw(int32_t s)
{
    return s < W ? W : s;
}

/* Picks among self (t0), previous (t1) and skip (t2) exactly like the nested
 * if/else of hmm.c:556-571: returns 0, 1 or 2. */
static int
pick3(int32_t t0, int32_t t1, int32_t t2)
{
    if (t0 > t1)
        return t2 > t0 ? 2 : 0;
    return t2 > t1 ? 2 : 1;
}

static int32_t
hmm_eval_one(int ne, const uint8_t *tp, const uint16_t *sseq, const int16_t *senscr, int32_t *sc, int32_t *hi,
             int32_t *out_sc, int32_t *out_hi, uint16_t *sid, int mpx)
{
    int32_t s[5], best, t0, t1, t2, x;
    int nc = ne + 1, st, w;
#define TP(i, j) (-(int32_t)tp[(i) * nc + (j)])
#define SEN(st) (mpx ? senscr[sseq[(long)sid[st] * ne + (st)]] : senscr[sid[st]])

    if (ne != 3 && ne != 5) {   /* hmm_vit_eval_anytopo, hmm.c:711-786 (n_emit_state 1, 2 or 4) */
        int to, from, bestfrom;
        int32_t newscr, scr;
        for (st = 0; st < ne; ++st) {
            /* hmm_senscr (hmm.h:198-200): WORST_SCORE for a missing senone */
            int32_t ss;
            if (mpx) ss = (sid[st] == ORC_BAD_SSID || sseq[(long)sid[st] * ne + st] == 0xffff) ? W : -(int32_t)SEN(st);
            else ss = sid[st] == 0xffff ? W : -(int32_t)SEN(st);
            s[st] = sc[st] + ss;
            if (st > 0 && s[st] < W) s[st] = W;      /* state 0 is not clamped (:720) */
        }
        scr = W; bestfrom = -1;
        for (from = ne - 1; from >= 0; --from)
            if (TP(from, ne) > ORC_TMAT_WORST && (newscr = s[from] + TP(from, ne)) > scr) { scr = newscr; bestfrom = from; }
        *out_sc = scr;
        if (bestfrom >= 0) *out_hi = hi[bestfrom];
        best = scr;
        for (to = ne - 1; to >= 0; --to) {
            scr = TP(to, to) > ORC_TMAT_WORST ? s[to] + TP(to, to) : W;
            bestfrom = -1;
            for (from = to - 1; from >= 0; --from)
                if (TP(from, to) > ORC_TMAT_WORST && (newscr = s[from] + TP(from, to)) > scr) { scr = newscr; bestfrom = from; }
            sc[to] = scr;                              /* stored unclamped (:766-774) */
            if (bestfrom >= 0) { hi[to] = hi[bestfrom]; if (mpx) sid[to] = sid[bestfrom]; }
            if (best < scr) best = scr;
        }
        return best;
    }
    if (ne == 3 && !mpx) {   /* hmm.c:531-609 */
        s[2] = sc[2] - SEN(2); s[1] = sc[1] - SEN(1); s[0] = sc[0] - SEN(0);
        best = W;
        t2 = INT_MIN;
        if (s[1] > W) {
            t1 = s[2] + TP(2, 3);
            if (TP(1, 3) > ORC_TMAT_WORST) t2 = s[1] + TP(1, 3);
            if (t1 > t2) { x = t1; *out_hi = hi[2]; } else { x = t2; *out_hi = hi[1]; }
            *out_sc = best = clampw(x);
        }
        t0 = s[2] + TP(2, 2); t1 = s[1] + TP(1, 2);
        if (TP(0, 2) > ORC_TMAT_WORST) t2 = s[0] + TP(0, 2);
        w = pick3(t0, t1, t2);
        x = w == 0 ? t0 : (w == 1 ? t1 : t2);
        if (w == 2) hi[2] = hi[0]; else if (w == 1) hi[2] = hi[1];
        sc[2] = clampw(x); if (sc[2] > best) best = sc[2];
        t0 = s[1] + TP(1, 1); t1 = s[0] + TP(0, 1);
        if (t0 > t1) x = t0; else { x = t1; hi[1] = hi[0]; }
        sc[1] = clampw(x); if (sc[1] > best) best = sc[1];
        sc[0] = clampw(s[0] + TP(0, 0)); if (sc[0] > best) best = sc[0];
        return best;
    }
    if (ne == 3 && mpx) {    /* hmm.c:611-709 */
        t2 = INT_MIN;
        if (sid[2] == ORC_BAD_SSID) s[2] = t1 = W;
        else { s[2] = sc[2] - SEN(2); t1 = s[2] + TP(2, 3); }
        if (sid[1] == ORC_BAD_SSID) s[1] = t2 = W;
        else { s[1] = sc[1] - SEN(1); if (TP(1, 3) > ORC_TMAT_WORST) t2 = s[1] + TP(1, 3); }
        if (t1 > t2) { x = t1; *out_hi = hi[2]; } else { x = t2; *out_hi = hi[1]; }
        *out_sc = best = clampw(x);
        s[0] = sc[0] - SEN(0);
        t0 = t1 = W;
        if (s[2] != W) t0 = s[2] + TP(2, 2);
        if (s[1] != W) t1 = s[1] + TP(1, 2);
        if (TP(0, 2) > ORC_TMAT_WORST) t2 = s[0] + TP(0, 2);
        w = pick3(t0, t1, t2);
        x = w == 0 ? t0 : (w == 1 ? t1 : t2);
        if (w == 2) { hi[2] = hi[0]; sid[2] = sid[0]; } else if (w == 1) { hi[2] = hi[1]; sid[2] = sid[1]; }
        sc[2] = clampw(x); if (sc[2] > best) best = sc[2];
        t0 = W;
        if (s[1] != W) t0 = s[1] + TP(1, 1);
        t1 = s[0] + TP(0, 1);
        if (t0 > t1) x = t0; else { x = t1; hi[1] = hi[0]; sid[1] = sid[0]; }
        sc[1] = clampw(x); if (sc[1] > best) best = sc[1];
        sc[0] = clampw(s[0] + TP(0, 0)); if (sc[0] > best) best = sc[0];
        return best;
    }
    if (ne == 5 && !mpx) {   /* hmm.c:224-352 */
        best = W;
        s[4] = sc[4] - SEN(4); s[3] = sc[3] - SEN(3);
        if (s[3] > W) {
            t1 = s[4] + TP(4, 5); t2 = s[3] + TP(3, 5);
            if (t1 > t2) { x = t1; *out_hi = hi[4]; } else { x = t2; *out_hi = hi[3]; }
            *out_sc = best = clampw(x);
        }
        for (st = 4; st >= 2; --st) {
            /* state st's block runs only if the state two below is alive
             * (states 4 and 3); state 2's block always runs. */
            s[st - 2] = sc[st - 2] - SEN(st - 2);
            if (st > 2 && !(s[st - 2] > W))
                continue;
            t0 = s[st] + TP(st, st); t1 = s[st - 1] + TP(st - 1, st); t2 = s[st - 2] + TP(st - 2, st);
            w = pick3(t0, t1, t2);
            x = w == 0 ? t0 : (w == 1 ? t1 : t2);
            if (w == 2) hi[st] = hi[st - 2]; else if (w == 1) hi[st] = hi[st - 1];
            s[st] = clampw(x); if (s[st] > best) best = s[st];
            sc[st] = s[st];
        }
        t0 = s[1] + TP(1, 1); t1 = s[0] + TP(0, 1);
        if (t0 > t1) x = t0; else { x = t1; hi[1] = hi[0]; }
        sc[1] = clampw(x); if (sc[1] > best) best = sc[1];
        sc[0] = clampw(s[0] + TP(0, 0)); if (sc[0] > best) best = sc[0];
        return best;
    }
    /* ne == 5 && mpx: hmm.c:357-527 */
    if (sid[4] == ORC_BAD_SSID) s[4] = t1 = W;
    else { s[4] = sc[4] - SEN(4); t1 = s[4] + TP(4, 5); }
    if (sid[3] == ORC_BAD_SSID) s[3] = t2 = W;
    else { s[3] = sc[3] - SEN(3); t2 = s[3] + TP(3, 5); }
    if (t1 > t2) { x = t1; *out_hi = hi[4]; } else { x = t2; *out_hi = hi[3]; }
    *out_sc = best = clampw(x);
    for (st = 4; st >= 2; --st) {
        int lo = st - 2;
        if (lo == 0) { s[0] = sc[0] - SEN(0); t2 = s[0] + TP(0, st); }
        else if (sid[lo] == ORC_BAD_SSID) s[lo] = t2 = W;
        else { s[lo] = sc[lo] - SEN(lo); t2 = s[lo] + TP(lo, st); }
        t0 = t1 = W;
        if (s[st] != W) t0 = s[st] + TP(st, st);
        if (s[st - 1] != W) t1 = s[st - 1] + TP(st - 1, st);
        w = pick3(t0, t1, t2);
        x = w == 0 ? t0 : (w == 1 ? t1 : t2);
        if (w == 2) { hi[st] = hi[lo]; sid[st] = sid[lo]; }
        else if (w == 1) { hi[st] = hi[st - 1]; sid[st] = sid[st - 1]; }
        /* NB: s[st] keeps the pre-update sum for the next block's "t1" only
         * through s[st-1]; s[st] itself is not read again. */
        x = clampw(x); if (x > best) best = x;
        sc[st] = x;
    }
    t0 = W;
    if (s[1] != W) t0 = s[1] + TP(1, 1);
    t1 = s[0] + TP(0, 1);
    if (t0 > t1) x = t0; else { x = t1; hi[1] = hi[0]; sid[1] = sid[0]; }
    sc[1] = clampw(x); if (sc[1] > best) best = sc[1];
    sc[0] = clampw(s[0] + TP(0, 0)); if (sc[0] > best) best = sc[0];
    return best;
#undef TP
#undef SEN
}

int32_t
orc_hmm_eval_batch(int n_emit, int n_hmm, const uint8_t *tp, int n_tmat, const uint16_t *sseq, int n_sseq,
                   const int16_t *senscr, int32_t *score, int32_t *history, int32_t *out_score,
                   int32_t *out_history, uint16_t *senid, const uint16_t *ssid, const int16_t *tmatid,
                   const uint8_t *mpx, int32_t *bestscore, int n_frames_repeat)
{
    int32_t best = W;
    int i, r;
    (void)n_tmat; (void)n_sseq; (void)ssid;
    if (n_emit < 1 || n_emit > 5)   /* HMM_MAX_NSTATE = 5 (hmm.h:90) */
        return W;
    for (r = 0; r < (n_frames_repeat > 0 ? n_frames_repeat : 1); ++r) {
        best = W;
        for (i = 0; i < n_hmm; ++i) {
            int32_t b = hmm_eval_one(n_emit, tp + (long)tmatid[i] * n_emit * (n_emit + 1), sseq, senscr,
                                     score + (long)i * n_emit, history + (long)i * n_emit, &out_score[i],
                                     &out_history[i], senid + (long)i * n_emit, mpx[i]);
            bestscore[i] = b;
            if (b > best) best = b;
        }
    }
    return best;
}

/* ======================================================================
 * sphinx3 flavour of the path: approx_cont_mgau_frame_eval and friends.
 * S3 = sphinx3/src/libs3decoder.  Scores are int32 logs (base -logbase,
 * default 1.0003, shift 0), higher = better; the Mahalanobis sum runs in
 * float64 over float32 differences (S3/libam/cont_mgau.c:1033-1205).
 * ====================================================================== */
#define S3_ZERO ((int32_t)0xc8000000)   /* sphinx3/include/s3types.h:192 */
#define S3_NO_BSTIDX (-1)               /* sphinx3/include/cont_mgau.h:134 */
#define S3_NOT_UPDATED (-100)           /* :135 */

struct orc_s3_model {
    int n_sen, n_ci_sen, max_comp, veclen;
    int *n_comp;                 /* [n_sen] after mgau_uninit_compact */
    float *mean, *var, *lrd;     /* [s][max_comp][veclen], [s][max_comp] */
    int32_t *mixw;               /* [s][max_comp] */
    int32_t *cd2cisen;
    double distfloor, f;
    orc_logmath_t *lmath;
    int32_t ci_pbeam; int max_cd; int ds_ratio; float tighten_factor;
    int32_t *bstidx, *bstscr, *updatetime;
    int32_t *ci_occ, *idx;
    int64_t n_sen_eval, n_gau_eval;
    /* sub-vector quantised shortlists (S3/libam/subvq.c), optional */
    int svq_n_sv, svq_size, svq_eval;        /* #sub-vectors, codewords per sub-vector, VQ_EVAL */
    int *svq_veclen, **svq_featdim;          /* [n_sv], [n_sv][veclen] */
    float **svq_mean, **svq_var, **svq_lrd;  /* [n_sv] -> [size][veclen], [size][veclen] (1/(2 var)), [size] */
    double svq_distfloor;
    int32_t *svq_map;                        /* [n_sen][max_comp][n_sv] compacted + linearised, -1 = none */
    int32_t *svq_dist;                       /* [n_sv * size] scores of the current frame */
    int32_t svq_beam;
    int32_t *svq_sl;                         /* [max_comp + 1] */
    /* Gaussian selector (S3/libam/gs.c), optional */
    int gs_n_code, gs_featlen, gs_best;      /* gs_best: closest codeword of the current frame */
    float *gs_codeword;                      /* [n_code][featlen] */
    uint32_t *gs_map;                        /* [n_sen][n_code] bit c = component c is in the shortlist */
};

/* S3/libcommon/vector.c:181-204 */
static int s3_vec_is_zero(const float *v, int n) { int i; for (i = 0; i < n && v[i] == 0.0; ++i); return i == n; }
static int s3_vec_is_nan(const float *v, int n) { int i; for (i = 0; i < n; ++i) if (isnan(v[i])) return 1; return 0; }

/* mgau_init (S3/libam/cont_mgau.c:900-958) on arrays instead of files:
 * mixw read/normalise/log (:480-680), mgau_uninit_compact (:700-790),
 * mgau_var_floor (:798-825), mgau_precomp (:852-894), distfloor (:951). */
orc_s3_model_t *
orc_s3_new(int n_sen, int n_comp, int veclen, const float *mean, const float *var, const float *mixw,
           double varfloor, double mixwfloor, double logbase, const int32_t *cd2cisen, int n_ci_sen)
{
    orc_s3_model_t *m = calloc(1, sizeof *m);
    size_t nv = (size_t)n_sen * n_comp * veclen, nc = (size_t)n_sen * n_comp;
    int s, c, i, c2;
    float *pdf = malloc(sizeof(float) * n_comp);
    m->n_sen = n_sen; m->n_ci_sen = n_ci_sen; m->max_comp = n_comp; m->veclen = veclen;
    m->lmath = orc_logmath_init(logbase, 0, 1);
    m->n_comp = malloc(sizeof(int) * n_sen);
    m->mean = malloc(sizeof(float) * nv); memcpy(m->mean, mean, sizeof(float) * nv);
    m->var = malloc(sizeof(float) * nv); memcpy(m->var, var, sizeof(float) * nv);
    m->lrd = calloc(nc, sizeof(float));
    m->mixw = calloc(nc, sizeof(int32_t));
    m->cd2cisen = malloc(sizeof(int32_t) * n_sen); memcpy(m->cd2cisen, cd2cisen, sizeof(int32_t) * n_sen);
    m->bstidx = malloc(sizeof(int32_t) * n_sen); m->bstscr = malloc(sizeof(int32_t) * n_sen);
    m->updatetime = malloc(sizeof(int32_t) * n_sen);
    m->ci_occ = calloc(n_ci_sen > 0 ? n_ci_sen : 1, sizeof(int32_t)); m->idx = calloc(n_ci_sen > 0 ? n_ci_sen : 1, sizeof(int32_t));
    /* mixture weights */
    for (s = 0; s < n_sen; ++s) {
        memcpy(pdf, mixw + (size_t)s * n_comp, sizeof(float) * n_comp);
        if (s3_vec_is_zero(pdf, n_comp)) {
            for (c = 0; c < n_comp; ++c) m->mixw[(size_t)s * n_comp + c] = S3_ZERO;
        } else {
            double sum = 0.0, f;
            for (c = 0; c < n_comp; ++c) if (pdf[c] != 0.0 && pdf[c] < mixwfloor) pdf[c] = (float)mixwfloor;
            for (c = 0; c < n_comp; ++c) sum += pdf[c];
            if (sum != 0.0) { f = 1.0 / sum; for (c = 0; c < n_comp; ++c) pdf[c] = (float)((double)pdf[c] * f); }
            for (c = 0; c < n_comp; ++c)
                m->mixw[(size_t)s * n_comp + c] = (pdf[c] != 0.0) ? (pdf[c] <= 0.0 ? S3_ZERO : orc_logmath_log(m->lmath, pdf[c])) : S3_ZERO;
        }
    }
    free(pdf);
    /* compaction of uninitialised components */
    for (s = 0; s < n_sen; ++s) {
        for (c = 0, c2 = 0; c < n_comp; ++c) {
            float *mu = m->mean + ((size_t)s * n_comp + c) * veclen, *va = m->var + ((size_t)s * n_comp + c) * veclen;
            int keep = !(s3_vec_is_nan(mu, veclen) || s3_vec_is_nan(va, veclen) || s3_vec_is_zero(va, veclen));
            if (keep) {
                if (c2 != c) {
                    memcpy(m->mean + ((size_t)s * n_comp + c2) * veclen, mu, sizeof(float) * veclen);
                    memcpy(m->var + ((size_t)s * n_comp + c2) * veclen, va, sizeof(float) * veclen);
                    m->mixw[(size_t)s * n_comp + c2] = m->mixw[(size_t)s * n_comp + c];
                }
                ++c2;
            }
        }
        m->n_comp[s] = c2;
    }
    /* variance floor, then precompute */
    for (s = 0; s < n_sen; ++s)
        for (c = 0; c < m->n_comp[s]; ++c) {
            float *va = m->var + ((size_t)s * n_comp + c) * veclen;
            double lrd = 0.0;
            if (varfloor > 0.0)
                for (i = 0; i < veclen; ++i) if (va[i] < varfloor) va[i] = (float)varfloor;
            for (i = 0; i < veclen; ++i) {
                lrd += log(va[i]);
                va[i] = (float)(1.0 / (va[i] * 2.0));
            }
            lrd += veclen * log(2.0 * M_PI);
            m->lrd[(size_t)s * n_comp + c] = (float)(-0.5 * lrd);
        }
    m->distfloor = (double)S3_ZERO * log(logbase);   /* logmath_log_to_ln, logmath.c:467-471 */
    m->f = 1.0 / log(logbase);
    m->ci_pbeam = orc_logmath_log(m->lmath, 1e-80); m->max_cd = 100000; m->ds_ratio = 1; m->tighten_factor = 0.5f;
    orc_s3_utt_reset(m);
    for (s = 0; s < n_sen; ++s) m->bstscr[s] = S3_ZERO;
    return m;
}

void
orc_s3_free(orc_s3_model_t *m)
{
    if (m && m->gs_n_code) { free(m->gs_codeword); free(m->gs_map); if (!m->svq_n_sv) free(m->svq_sl); }
    if (m && m->svq_n_sv) {
        int k;
        for (k = 0; k < m->svq_n_sv; ++k) { free(m->svq_featdim[k]); free(m->svq_mean[k]); free(m->svq_var[k]); free(m->svq_lrd[k]); }
        free(m->svq_veclen); free(m->svq_featdim); free(m->svq_mean); free(m->svq_var); free(m->svq_lrd);
        free(m->svq_map); free(m->svq_dist); free(m->svq_sl);
    }
    if (!m) return;
    free(m->n_comp); free(m->mean); free(m->var); free(m->lrd); free(m->mixw); free(m->cd2cisen);
    free(m->bstidx); free(m->bstscr); free(m->updatetime); free(m->ci_occ); free(m->idx);
    orc_logmath_free(m->lmath); free(m);
}

/* fast_gmm_init (S3/libam/fast_algo_struct.c:420-467): beam given as a
 * probability, converted with logs3. */
void
orc_s3_set_fast(orc_s3_model_t *m, double ci_pbeam, int max_cd, int ds_ratio, float tighten_factor)
{
    m->ci_pbeam = ci_pbeam <= 0.0 ? S3_ZERO : orc_logmath_log(m->lmath, ci_pbeam);
    m->max_cd = max_cd; m->ds_ratio = ds_ratio; m->tighten_factor = tighten_factor;
}

int32_t orc_s3_ci_pbeam(const orc_s3_model_t *m) { return m->ci_pbeam; }

/* per-utterance re-initialisation, S3/libsearch/srch_time_switch_tree.c:484-490 */
void
orc_s3_utt_reset(orc_s3_model_t *m)
{
    int s;
    for (s = 0; s < m->n_sen; ++s) { m->bstidx[s] = S3_NO_BSTIDX; m->updatetime[s] = S3_NOT_UPDATED; }
}

void
orc_s3_params(const orc_s3_model_t *m, int32_t *n_comp, float *mean, float *var, float *lrd, int32_t *mixw, double *scal)
{
    size_t nv = (size_t)m->n_sen * m->max_comp * m->veclen, nc = (size_t)m->n_sen * m->max_comp;
    int s;
    for (s = 0; s < m->n_sen; ++s) n_comp[s] = m->n_comp[s];
    memcpy(mean, m->mean, sizeof(float) * nv); memcpy(var, m->var, sizeof(float) * nv);
    memcpy(lrd, m->lrd, sizeof(float) * nc); memcpy(mixw, m->mixw, sizeof(int32_t) * nc);
    scal[0] = m->distfloor; scal[1] = m->f;
}

void
orc_s3_state(const orc_s3_model_t *m, int32_t *bstidx, int32_t *updatetime)
{
    memcpy(bstidx, m->bstidx, sizeof(int32_t) * m->n_sen); memcpy(updatetime, m->updatetime, sizeof(int32_t) * m->n_sen);
}

/* one density: the float64 accumulation of cont_mgau.c:1062-1068 */
static double
s3_dval(const orc_s3_model_t *m, int s, int c, const float *x)
{
    const float *mu = m->mean + ((size_t)s * m->max_comp + c) * m->veclen;
    const float *va = m->var + ((size_t)s * m->max_comp + c) * m->veclen;
    double dval = m->lrd[(size_t)s * m->max_comp + c], diff;
    int i;
    for (i = 0; i < m->veclen; ++i) {
        diff = x[i] - mu[i];          /* float32 subtraction, then widened */
        dval -= diff * diff * va[i];
    }
    return dval;
}

/* mgau_eval (cont_mgau.c:1171-1205) with mgau_eval_all (:1033-1123) and
 * mgau_eval_active (:1125-1165); diagonal covariances only. */
int32_t
orc_s3_mgau_eval(orc_s3_model_t *m, int s, const int32_t *active, const float *x, int fr, int update_best_id)
{
    const int32_t *mw = m->mixw + (size_t)s * m->max_comp;
    int32_t score = S3_ZERO, gauscr;
    int c, j, nc = m->n_comp[s];
    double d1, d2;
    if (update_best_id) { m->bstidx[s] = S3_NO_BSTIDX; m->bstscr[s] = S3_ZERO; m->updatetime[s] = fr; }
    if (!active) {
        for (c = 0; c < nc - 1; c += 2) {
            d1 = s3_dval(m, s, c, x); d2 = s3_dval(m, s, c + 1, x);
            if (d1 < m->distfloor) d1 = m->distfloor;
            if (d2 < m->distfloor) d2 = m->distfloor;
            gauscr = (int32_t)(m->f * d1) + mw[c];
            score = orc_logmath_add(m->lmath, score, gauscr);
            if (gauscr > m->bstscr[s]) { m->bstidx[s] = c; m->bstscr[s] = gauscr; }   /* :1080, no update_best_id test */
            gauscr = (int32_t)(m->f * d2) + mw[c + 1];
            score = orc_logmath_add(m->lmath, score, gauscr);
            if (update_best_id && gauscr > m->bstscr[s]) { m->bstidx[s] = c + 1; m->bstscr[s] = gauscr; }
        }
        if (c < nc) {
            d1 = s3_dval(m, s, c, x);
            if (d1 < m->distfloor) d1 = m->distfloor;
            gauscr = (int32_t)(m->f * d1) + mw[c];
            score = orc_logmath_add(m->lmath, score, gauscr);
            if (update_best_id && gauscr > m->bstscr[s]) { m->bstidx[s] = c; m->bstscr[s] = gauscr; }
        }
    } else {
        for (j = 0; active[j] >= 0; ++j) {
            c = active[j];
            d1 = s3_dval(m, s, c, x);
            if (d1 < m->distfloor) d1 = m->distfloor;
            gauscr = (int32_t)(m->f * d1) + mw[c];
            score = orc_logmath_add(m->lmath, score, gauscr);
            if (update_best_id && gauscr > m->bstscr[s]) { m->bstidx[s] = c; m->bstscr[s] = gauscr; }
        }
    }
    if (score <= S3_ZERO) score = S3_ZERO;
    return score;
}

/* ---- sub-vector quantised Gaussian selection (S3/libam/subvq.c) ----
 * orc_s3_set_svq = what subvq_init (:206-373) does after reading the file: variance floor +
 * vector_maha_precomp per codeword (subvq_maha_precomp :106-123, S3/libcommon/vector.c:127-134,
 * 327-339), map compaction (:127-181: entries < 0 mark unused components, the remaining ones move
 * up) and linearisation (:191-203: index = sub-vector * vqsize + codeword).  The caller passes
 * the file's contents: veclen[n_sv], featdim (concatenated), mean / var (concatenated
 * [size][veclen] blocks, raw variances), map [n_sen][max_comp][n_sv_file] as in the file.
 * n_sv_use = min(-svmax, n_sv_file) sub-vectors are kept (:228-243). */
int
orc_s3_set_svq(orc_s3_model_t *m, int n_sv_file, int n_sv_use, int vqsize, int vqeval, const int32_t *veclen,
               const int32_t *featdim, const float *mean, const float *var, const int32_t *map, double varfloor,
               double subvqbeam)
{
    int sv, r, c, c2, i, off_d = 0;
    size_t off_p = 0;
    if (n_sv_use < 0 || n_sv_use > n_sv_file) n_sv_use = n_sv_file;
    m->svq_n_sv = n_sv_use; m->svq_size = vqsize;
    m->svq_eval = n_sv_use < vqeval ? n_sv_use : vqeval;                 /* :240-241 */
    m->svq_veclen = calloc(n_sv_use, sizeof(int)); m->svq_featdim = calloc(n_sv_use, sizeof(int *));
    m->svq_mean = calloc(n_sv_use, sizeof(float *)); m->svq_var = calloc(n_sv_use, sizeof(float *));
    m->svq_lrd = calloc(n_sv_use, sizeof(float *));
    m->svq_distfloor = m->distfloor;                                    /* logmath_log_to_ln(S3_LOGPROB_ZERO), vector.c:549: the same value as mgau_init's */
    for (sv = 0; sv < n_sv_file; ++sv) {
        const int L = veclen[sv];
        if (sv < n_sv_use) {
            m->svq_veclen[sv] = L;
            m->svq_featdim[sv] = malloc(sizeof(int) * L);
            for (i = 0; i < L; ++i) m->svq_featdim[sv][i] = featdim[off_d + i];
            m->svq_mean[sv] = malloc(sizeof(float) * vqsize * L);
            m->svq_var[sv] = malloc(sizeof(float) * vqsize * L);
            m->svq_lrd[sv] = malloc(sizeof(float) * vqsize);
            memcpy(m->svq_mean[sv], mean + off_p, sizeof(float) * vqsize * L);
            memcpy(m->svq_var[sv], var + off_p, sizeof(float) * vqsize * L);
            for (r = 0; r < vqsize; ++r) {
                float *v = m->svq_var[sv] + (size_t)r * L;
                double det = 0.0;
                for (i = 0; i < L; ++i) if (v[i] < varfloor) v[i] = (float)varfloor;
                for (i = 0; i < L; ++i) { det -= (double)log(v[i]); v[i] = (float)(1.0 / (v[i] * 2.0)); }
                det -= log(2.0 * M_PI) * L;
                m->svq_lrd[sv][r] = (float)(det * 0.5);
            }
        }
        off_d += L; off_p += (size_t)vqsize * L;
    }
    m->svq_map = malloc(sizeof(int32_t) * m->n_sen * m->max_comp * n_sv_use);
    for (r = 0; r < m->n_sen; ++r) {
        int32_t *dst = m->svq_map + (size_t)r * m->max_comp * n_sv_use;
        const int32_t *src = map + (size_t)r * m->max_comp * n_sv_file;
        for (c = 0, c2 = 0; c < m->max_comp; ++c) {
            if (src[(size_t)c * n_sv_file] < 0) continue;
            for (sv = 0; sv < n_sv_use; ++sv) dst[(size_t)c2 * n_sv_use + sv] = sv * vqsize + src[(size_t)c * n_sv_file + sv];
            c2++;
        }
        if (c2 != m->n_comp[r]) return -1;                               /* :174-177 */
        for (; c2 < m->max_comp; ++c2) for (sv = 0; sv < n_sv_use; ++sv) dst[(size_t)c2 * n_sv_use + sv] = -1;
    }
    m->svq_dist = calloc((size_t)n_sv_use * vqsize, sizeof(int32_t));    /* unevaluated sub-vectors stay 0 (ckd_calloc_2d, :366) */
    m->svq_sl = malloc(sizeof(int32_t) * (m->max_comp + 1));
    m->svq_beam = subvqbeam <= 0.0 ? S3_ZERO : orc_logmath_log(m->lmath, subvqbeam);   /* logs3(), fast_algo_struct.c:454 */
    return 0;
}

void
orc_s3_svq_tables(const orc_s3_model_t *m, int sv, float *mean, float *var, float *lrd, double *scal)
{
    const int L = m->svq_veclen[sv];
    memcpy(mean, m->svq_mean[sv], sizeof(float) * m->svq_size * L);
    memcpy(var, m->svq_var[sv], sizeof(float) * m->svq_size * L);
    memcpy(lrd, m->svq_lrd[sv], sizeof(float) * m->svq_size);
    scal[0] = m->svq_distfloor;
}
void orc_s3_svq_map(const orc_s3_model_t *m, int32_t *out) { memcpy(out, m->svq_map, sizeof(int32_t) * m->n_sen * m->max_comp * m->svq_n_sv); }
int32_t orc_s3_svq_beam(const orc_s3_model_t *m) { return m->svq_beam; }

/* subvq_gautbl_eval_logs3 (subvq.c:488-506) + vector_gautbl_eval_logs3 (vector.c:590-650): only the
 * first VQ_EVAL sub-vectors are evaluated */
void
orc_s3_svq_eval(orc_s3_model_t *m, const float *x)
{
    int sv, r, i;
    for (sv = 0; sv < m->svq_n_sv && sv < m->svq_eval; ++sv) {
        const int L = m->svq_veclen[sv];
        for (r = 0; r < m->svq_size; ++r) {
            const float *mu = m->svq_mean[sv] + (size_t)r * L, *va = m->svq_var[sv] + (size_t)r * L;
            double dval = m->svq_lrd[sv][r], diff;
            for (i = 0; i < L; ++i) {
                diff = x[m->svq_featdim[sv][i]] - mu[i];      /* float32 subtraction, then widened */
                dval -= diff * diff * va[i];
            }
            if (dval < m->svq_distfloor) dval = m->svq_distfloor;
            m->svq_dist[sv * m->svq_size + r] = (int32_t)(m->f * dval);
        }
    }
}
void orc_s3_svq_dist(const orc_s3_model_t *m, int32_t *out) { memcpy(out, m->svq_dist, sizeof(int32_t) * m->svq_n_sv * m->svq_size); }

/* subvq_mgau_shortlist (subvq.c:383-468); returns the shortlist length, list in m->svq_sl */
int
orc_s3_svq_shortlist(orc_s3_model_t *m, int s)
{
    const int n = m->n_comp[s], nsv = m->svq_n_sv;
    const int32_t *map = m->svq_map + (size_t)s * m->max_comp * nsv, *vq = m->svq_dist;
    int32_t gs[256], v, bv = INT_MIN, th;
    int i, k, nc = 0;
    for (i = 0; i < n; ++i, map += nsv) {
        if (nsv == 3) {
            if (m->svq_eval == 1) v = vq[map[0]];
            else if (m->svq_eval == 2) v = vq[map[0]] + 2 * vq[map[1]];
            else v = vq[map[0]] + vq[map[1]] + vq[map[2]];
        } else {
            for (v = 0, k = 0; k < nsv; ++k) v += vq[map[k]];
        }
        gs[i] = v;
        if (bv < v) bv = v;
    }
    th = bv + m->svq_beam;
    for (i = 0; i < n; ++i) if (gs[i] >= th) m->svq_sl[nc++] = i;
    m->svq_sl[nc] = -1;
    return nc;
}

/* gs_read's result (gs.c:156-218) from arrays: codewords [n_code][featlen] and, per (senone, codeword), the
 * first 32-bit word of the file's bit vector */
void
orc_s3_set_gs(orc_s3_model_t *m, int n_code, int featlen, const float *codeword, const uint32_t *map)
{
    m->gs_n_code = n_code; m->gs_featlen = featlen;
    m->gs_codeword = malloc(sizeof(float) * n_code * featlen);
    memcpy(m->gs_codeword, codeword, sizeof(float) * n_code * featlen);
    m->gs_map = malloc(sizeof(uint32_t) * m->n_sen * n_code);
    memcpy(m->gs_map, map, sizeof(uint32_t) * m->n_sen * n_code);
    if (!m->svq_sl) m->svq_sl = malloc(sizeof(int32_t) * (m->max_comp + 1));
}

/* gc_compute_closest_cw (gs.c:221-259): squared Euclidean distance, float32 differences summed in float64,
 * the first minimum wins (the reference walks the codewords in pairs: an even count is assumed) */
int
orc_s3_gs_closest(const orc_s3_model_t *m, const float *x)
{
    int c, i, best = 0;
    double min = 1.7976931348623157e308, tmp, diff;
    for (c = 0; c < m->gs_n_code; ++c) {
        for (tmp = 0, i = 0; i < m->gs_featlen; ++i) {
            diff = x[i] - m->gs_codeword[(size_t)c * m->gs_featlen + i];
            tmp += diff * diff;
        }
        if (tmp < min) { min = tmp; best = c; }
    }
    return best;
}

/* approx_mgau_eval (approx_cont_mgau.c:187-284): Gaussian selector first (gs4gs), else sub-VQ; svq4svq off.
 * (The reference asserts best_cid > 0 before gs_mgau_shortlist, :209: a frame whose nearest codeword is number 0
 * aborts its debug build; the port -- like a release build -- goes on.) */
static int
s3_approx_mgau_eval(orc_s3_model_t *m, int s, int32_t *senscr, const float *x, int fr)
{
    const int32_t *sl = NULL;
    int ng = m->n_comp[s];
    if (m->gs_n_code) {                                  /* gs_mgau_shortlist, gs.c:263-300 */
        const uint32_t map = m->gs_map[(size_t)s * m->gs_n_code + m->gs_best];
        int b, n = m->n_comp[s];
        for (ng = 0, b = 0; b < n; ++b) if (map & (1u << b)) m->svq_sl[ng++] = b;
        if (ng == 0) for (b = 0; b < n; ++b) m->svq_sl[ng++] = b;
        m->svq_sl[ng] = -1;
        sl = m->svq_sl;
        if (ng == 0) { sl = NULL; ng = n; }
    } else if (m->svq_n_sv) {
        ng = orc_s3_svq_shortlist(m, s);
        sl = m->svq_sl;
        if (ng == 0) { sl = NULL; ng = m->n_comp[s]; }
    }
    senscr[s] = orc_s3_mgau_eval(m, s, sl, x, fr, 1);
    if (senscr[s] < S3_ZERO + 100000 && sl) {            /* :256-281: recompute with every component */
        ng += m->n_comp[s];
        senscr[s] = orc_s3_mgau_eval(m, s, NULL, x, fr, 1);
    }
    return ng;
}

/* approx_cont_mgau_ci_eval (S3/libam/approx_cont_mgau.c:368-431), no Gaussian selector */
void
orc_s3_ci_eval(orc_s3_model_t *m, const float *x, int32_t *ci_senscr, int32_t *best, int fr)
{
    int s;
    if (m->gs_n_code) m->gs_best = orc_s3_gs_closest(m, x);
    if (m->svq_n_sv) orc_s3_svq_eval(m, x);
    for (s = 0; s < m->n_ci_sen; ++s) s3_approx_mgau_eval(m, s, ci_senscr, x, fr);
    *best = INT_MIN;
    for (s = 0; s < m->n_ci_sen; ++s) if (ci_senscr[s] > *best) *best = ci_senscr[s];
}

static const int32_t *s3_sort_key;
static int s3_intcmp(const void *a, const void *b) { return s3_sort_key[*(const int32_t *)b] - s3_sort_key[*(const int32_t *)a]; }

/* approx_compute_dyn_ci_pbeam (approx_cont_mgau.c:302-358) */
static int32_t
s3_dyn_beam(orc_s3_model_t *m, const uint8_t *sen_active, const int32_t *ci)
{
    int s, total = 0;
    int32_t pbest, beam = m->ci_pbeam;
    for (s = 0; s < m->n_sen; ++s) {
        if (s < m->n_ci_sen) m->ci_occ[s] = 0;
        else if (!sen_active || sen_active[s]) m->ci_occ[m->cd2cisen[s]]++;
    }
    for (s = 0; s < m->n_ci_sen; ++s) m->idx[s] = s;
    s3_sort_key = ci;
    qsort(m->idx, m->n_ci_sen, sizeof(int32_t), s3_intcmp);
    pbest = ci[m->idx[0]];
    for (s = 0; s < m->n_ci_sen && ci[m->idx[s]] > pbest + m->ci_pbeam; ++s) {
        total += m->ci_occ[m->idx[s]];
        if (total > m->max_cd) { beam = ci[m->idx[s]] - pbest; break; }
    }
    return beam;
}

/* approx_cont_mgau_frame_eval (approx_cont_mgau.c:433-616), no GS/SVQ,
 * approx_isskip (:93-143) with plain -ds only.  sen_active is in/out (CI
 * entries are forced to 1); senscr entries of inactive CD senones are left
 * as they were. */
int32_t
orc_s3_frame_eval(orc_s3_model_t *m, const float *x, int frame, const int32_t *cache_ci_senscr,
                  uint8_t *sen_active, int32_t *senscr)
{
    int32_t best = INT_MIN, pbest = INT_MIN, dyn, single[2] = { -1, -1 };
    int s, is_skip;
    int64_t ns = 0, ng = 0;
    if (m->gs_n_code) m->gs_best = orc_s3_gs_closest(m, x);
    if (m->svq_n_sv) orc_s3_svq_eval(m, x);
    if (m->max_cd < m->n_sen - m->n_ci_sen) dyn = s3_dyn_beam(m, sen_active, cache_ci_senscr);
    else dyn = m->ci_pbeam;
    is_skip = (frame % m->ds_ratio) != 0;
    if (is_skip) dyn = (int32_t)((float)dyn * m->tighten_factor);
    for (s = 0; s < m->n_sen; ++s) {
        if (s < m->n_ci_sen) {
            senscr[s] = cache_ci_senscr[s];
            if (pbest < senscr[s]) pbest = senscr[s];
            if (best < senscr[s]) best = senscr[s];
            sen_active[s] = 1;
        } else if (sen_active[s]) {
            if (senscr[m->cd2cisen[s]] >= pbest + dyn) {
                ng += s3_approx_mgau_eval(m, s, senscr, x, frame);
                ns++;
            } else if (m->bstidx[s] == S3_NO_BSTIDX || m->updatetime[s] != frame - 1) {
                senscr[s] = senscr[m->cd2cisen[s]];
            } else {
                single[0] = m->bstidx[s];
                senscr[s] = orc_s3_mgau_eval(m, s, single, x, frame, is_skip ? 1 : 0);
                ng++;
            }
            if (best < senscr[s]) best = senscr[s];
        }
    }
    for (s = 0; s < m->n_sen; ++s) if (sen_active[s]) senscr[s] -= best;
    m->n_sen_eval += ns; m->n_gau_eval += ng;
    return best;
}

/* The decoder's per-utterance sequence (S3/libsearch/srch.c lv1 + lv2 with a
 * zero look-ahead window): for each frame the CI pass, then the frame eval.
 * sen_active [T][n_sen] in/out (NULL = all active); senscr_io [n_sen] holds
 * the score buffer before the first frame on entry; out [T][n_sen];
 * best [T]. */
void
orc_s3_eval_utt(orc_s3_model_t *m, const float *feat, int T, int frame0, uint8_t *sen_active,
                int32_t *senscr_io, int32_t *out, int32_t *best)
{
    int t;
    int32_t *ci = malloc(sizeof(int32_t) * (m->n_ci_sen > 0 ? m->n_ci_sen : 1)), cib;
    uint8_t *all = NULL;
    if (!sen_active) { all = malloc(m->n_sen); }
    for (t = 0; t < T; ++t) {
        const float *x = feat + (size_t)t * m->veclen;
        uint8_t *act = sen_active ? sen_active + (size_t)t * m->n_sen : all;
        if (all) memset(all, 1, m->n_sen);
        orc_s3_ci_eval(m, x, ci, &cib, frame0 + t);
        best[t] = orc_s3_frame_eval(m, x, frame0 + t, ci, act, senscr_io);
        memcpy(out + (size_t)t * m->n_sen, senscr_io, sizeof(int32_t) * m->n_sen);
    }
    free(ci); free(all);
}

void orc_s3_counts(const orc_s3_model_t *m, int64_t *c) { c[0] = m->n_sen_eval; c[1] = m->n_gau_eval; }

/* ======================================================================
 * Feature stage next to the scorer ("next" row of the scope table):
 * full-utterance 13-dim cepstra -> 39-dim 1s_c_d_dd features with
 * -cmn current, as acmod_process_cep(full_utt) computes them:
 *   feat_s2mfc2feat_block_utt  sphinxbase/src/libsphinxbase/feat/feat.c:1241-1265
 *     (the utterance is padded with `win` = 3 copies of its first and last
 *      frame BEFORE normalisation, so the padding enters the mean)
 *   cmn()                      sphinxbase/src/libsphinxbase/feat/cmn.c:150-186
 *     (float32 running sum in frame order, mean = sum / n, subtract)
 *   feat_1s_c_d_dd_cep2feat    feat.c:726-769
 * cep [T][cepsize] -> feat [T][3*cepsize].  cmn: 0 none, 1 current. */
void
orc_feat_1s_c_d_dd(const float *cep, int T, int cepsize, int cmn, float *feat)
{
    const int win = 3, n = T + 2 * win;
    float *buf = malloc(sizeof(float) * (size_t)n * cepsize), *mean = calloc(cepsize, sizeof(float));
    int t, i;
    if (T <= 0) { free(buf); free(mean); return; }
    for (t = 0; t < n; ++t) {
        int src = t - win;
        if (src < 0) src = 0;
        if (src > T - 1) src = T - 1;
        memcpy(buf + (size_t)t * cepsize, cep + (size_t)src * cepsize, sizeof(float) * cepsize);
    }
    if (cmn) {
        for (t = 0; t < n; ++t)
            for (i = 0; i < cepsize; ++i) mean[i] += buf[(size_t)t * cepsize + i];
        for (i = 0; i < cepsize; ++i) mean[i] /= n;
        for (t = 0; t < n; ++t)
            for (i = 0; i < cepsize; ++i) buf[(size_t)t * cepsize + i] -= mean[i];
    }
    for (t = 0; t < T; ++t) {
        const float *c = buf + (size_t)(t + win) * cepsize;
        float *f = feat + (size_t)t * 3 * cepsize;
        for (i = 0; i < cepsize; ++i) {
            float d1, d2;
            f[i] = c[i];
            f[cepsize + i] = c[2 * cepsize + i] - c[-2 * cepsize + i];
            d1 = c[3 * cepsize + i] - c[-1 * cepsize + i];
            d2 = c[1 * cepsize + i] - c[-3 * cepsize + i];
            f[2 * cepsize + i] = d1 - d2;
        }
    }
    free(buf); free(mean);
}


/* ---- general feature stage ------------------------------------------------
 * The same walk as the reference: pad (feat.c:1253-1259), cmn [+varnorm]
 * (cmn.c:150-213), agc max (agc.c:108-126), the type's compute_feat on every
 * frame (feat.c:559-849), lda (lda.c:141-160), subvectors (feat.c:334-355). */
static int orc_copy_window, orc_copy_streams, orc_copy_len[8];   /* type 6, set by orc_feat_set_copy */

/* feat_copy types "n[,n..][:w]" (feat.c:828-849, 952-1000) */
void
orc_feat_set_copy(int window, int n_streams, const int *len)
{
    int j;
    orc_copy_window = window; orc_copy_streams = n_streams;
    for (j = 0; j < n_streams && j < 8; ++j) orc_copy_len[j] = len[j];
}

static int
orc_feat_window(int type, int cs, int *k)
{
    switch (type) {
    case 6: {
        int j, tot = 0;
        for (j = 0; j < orc_copy_streams; ++j) tot += orc_copy_len[j];
        if (orc_copy_streams == 0) tot = cs;
        *k = tot * (2 * orc_copy_window + 1);
        return tot <= cs ? orc_copy_window : -1;
    }
    case 0: *k = 3 * cs; return 3;
    case 1: *k = 39; return cs == 13 ? 3 : -1;
    case 2: *k = 51; return cs == 13 ? 4 : -1;
    case 3: *k = 4 * cs; return 4;
    case 4: *k = cs; return 0;
    case 5: *k = 2 * cs; return 2;
    }
    return -1;
}

int
orc_feat_compute(int type, int cepsize, int cmn, int varnorm, int agc,
                 const float *lda, int lda_dim, const int *subvec, int n_subvec,
                 const float *cep, int T, float *out)
{
    int k, t, i, j, n, out_len;
    const int cs = cepsize, win = orc_feat_window(type, cepsize, &k);
    float *buf, *row, *tmp;
    if (win < 0) return -1;
    out_len = n_subvec > 0 ? n_subvec : (lda ? lda_dim : k);
    if (T <= 0) return out_len;
    n = T + 2 * win;
    buf = malloc(sizeof(float) * (size_t)n * cs);
    row = calloc(k, sizeof(float));
    tmp = calloc(k, sizeof(float));
    for (t = 0; t < n; ++t) {
        int src = t - win;
        if (src < 0) src = 0;
        if (src > T - 1) src = T - 1;
        memcpy(buf + (size_t)t * cs, cep + (size_t)src * cs, sizeof(float) * cs);
    }
    if (cmn) {
        float *mean = calloc(cs, sizeof(float)), *var = calloc(cs, sizeof(float));
        for (t = 0; t < n; ++t)
            for (i = 0; i < cs; ++i) mean[i] += buf[(size_t)t * cs + i];
        for (i = 0; i < cs; ++i) mean[i] /= n;
        if (!varnorm) {
            for (t = 0; t < n; ++t)
                for (i = 0; i < cs; ++i) buf[(size_t)t * cs + i] -= mean[i];
        }
        else {
            for (t = 0; t < n; ++t)
                for (i = 0; i < cs; ++i) {
                    float d = buf[(size_t)t * cs + i] - mean[i];
                    var[i] += d * d;
                }
            for (i = 0; i < cs; ++i) var[i] = (float)sqrt((double)n / var[i]);
            for (t = 0; t < n; ++t)
                for (i = 0; i < cs; ++i)
                    buf[(size_t)t * cs + i] = (buf[(size_t)t * cs + i] - mean[i]) * var[i];
        }
        free(mean); free(var);
    }
    if (agc) {
        float m = buf[0];
        for (t = 1; t < n; ++t)
            if (buf[(size_t)t * cs] > m) m = buf[(size_t)t * cs];
        for (t = 0; t < n; ++t) buf[(size_t)t * cs] -= m;
    }
#define C_(off, d) (c[(off) * cs + (d)])
#define DD_(d) ((C_(3, d) - C_(-1, d)) - (C_(1, d) - C_(-3, d)))
    for (t = 0; t < T; ++t) {
        const float *c = buf + (size_t)(t + win) * cs;
        float *f = row;
        switch (type) {
        case 0:
            for (i = 0; i < cs; ++i) *f++ = C_(0, i);
            for (i = 0; i < cs; ++i) *f++ = C_(2, i) - C_(-2, i);
            for (i = 0; i < cs; ++i) *f++ = DD_(i);
            break;
        case 1:
            for (i = 1; i < cs; ++i) *f++ = C_(0, i);
            for (i = 1; i < cs; ++i) *f++ = C_(2, i) - C_(-2, i);
            *f++ = C_(0, 0); *f++ = C_(2, 0) - C_(-2, 0); *f++ = DD_(0);
            for (i = 1; i < cs; ++i) *f++ = DD_(i);
            break;
        case 2:
            for (i = 1; i < cs; ++i) *f++ = C_(0, i);
            for (i = 1; i < cs; ++i) *f++ = C_(2, i) - C_(-2, i);
            for (i = 1; i < cs; ++i) *f++ = C_(4, i) - C_(-4, i);
            *f++ = C_(0, 0); *f++ = C_(2, 0) - C_(-2, 0); *f++ = DD_(0);
            for (i = 1; i < cs; ++i) *f++ = DD_(i);
            break;
        case 3:
            for (i = 0; i < cs; ++i) *f++ = C_(0, i);
            for (i = 0; i < cs; ++i) *f++ = C_(2, i) - C_(-2, i);
            for (i = 0; i < cs; ++i) *f++ = C_(4, i) - C_(-4, i);
            for (i = 0; i < cs; ++i) *f++ = DD_(i);
            break;
        case 4:
            for (i = 0; i < cs; ++i) *f++ = C_(0, i);
            break;
        case 5:
            for (i = 0; i < cs; ++i) *f++ = C_(0, i);
            for (i = 0; i < cs; ++i) *f++ = C_(2, i) - C_(-2, i);
            break;
        case 6: {
            int spos = 0, w;
            for (j = 0; j < (orc_copy_streams ? orc_copy_streams : 1); ++j) {
                const int len = orc_copy_streams ? orc_copy_len[j] : cs;
                for (w = -win; w <= win; ++w)
                    for (i = 0; i < len; ++i) *f++ = C_(w, spos + i);
                spos += len;
            }
            break;
        }
        }
        if (lda) {
            memset(tmp, 0, sizeof(float) * k);
            for (j = 0; j < lda_dim; ++j)
                for (i = 0; i < k; ++i) tmp[j] += row[i] * lda[(size_t)j * k + i];
            memcpy(row, tmp, sizeof(float) * k);
        }
        if (n_subvec > 0) {
            for (j = 0; j < n_subvec; ++j) tmp[j] = row[subvec[j]];
            memcpy(row, tmp, sizeof(float) * n_subvec);
        }
        memcpy(out + (size_t)t * out_len, row, sizeof(float) * out_len);
    }
#undef C_
#undef DD_
    free(buf); free(row); free(tmp);
    return out_len;
}

/* ======================================================================
 * sphinx3's hmm_vit_eval flavour (sphinx3/src/libs3decoder/libam/hmm.c:285-873):
 * int32 tp and senone scores that are added, WORST = S3_LOGPROB_ZERO
 * (s3types.h:192), int32 ssids (-1 = none), senone ids via sseq[ssid][state].
 * HMM-major arrays [n_hmm][n_emit]; non-mpx HMMs use ssid[i][0].
 * TEST INFRASTRUCTURE ONLY.
 * ====================================================================== */
#define S3W ((int32_t)0xc8000000)
#define S3TP(i, j) tp[(i) * (ne + 1) + (j)]
#define S3SEN(id, st) sen[sseq[(size_t)(id) * ne + (st)]]
/* new score of a state with a self loop (a), a one-step (b, from state fb) and a two-step
 * (c2, from state fc) entry: the reference's nested comparison -- on a == b the one-step path
 * wins, c2 must beat the winner strictly */
#define S3_PICK3(dst, a, b, c2, fb, fc)                                              \
    do {                                                                             \
        if ((a) > (b)) { if ((c2) > (a)) { dst = (c2); from = (fc); } else { dst = (a); from = -1; } } \
        else { if ((c2) > (b)) { dst = (c2); from = (fc); } else { dst = (b); from = (fb); } }          \
    } while (0)

static int32_t s3hmm_one(int ne, const int32_t *tp, const int16_t *sseq, const int32_t *sen, int32_t *sc, int32_t *hi,
                         int32_t *ssid, int mpx, int32_t *out_sc, int32_t *out_hi)
{
    int32_t best = S3W, from = -1;
    if (ne != 3 && ne != 5) {           /* hmm_vit_eval_anytopo, hmm.c:776-850 */
        int32_t st[5], nsc[5], nhi[5], nss[5], scr, bf;
        int s, f, to;
        for (s = 0; s < ne; ++s) {
            int32_t id = mpx ? ssid[s] : ssid[0];
            st[s] = sc[s] + (id == -1 ? S3W : S3SEN(id, s));
            if (s > 0 && st[s] < S3W) st[s] = S3W;
            nhi[s] = hi[s]; nss[s] = ssid[s];
        }
        scr = S3W; bf = -1;
        for (f = ne - 1; f >= 0; --f)
            if (S3TP(f, ne) > S3W && st[f] + S3TP(f, ne) > scr) { scr = st[f] + S3TP(f, ne); bf = f; }
        *out_sc = scr;
        if (bf >= 0) *out_hi = hi[bf];
        best = scr;
        for (to = ne - 1; to >= 0; --to) {
            scr = S3TP(to, to) > S3W ? st[to] + S3TP(to, to) : S3W;
            bf = -1;
            for (f = to - 1; f >= 0; --f)
                if (S3TP(f, to) > S3W && st[f] + S3TP(f, to) > scr) { scr = st[f] + S3TP(f, to); bf = f; }
            nsc[to] = scr;
            if (bf >= 0) { nhi[to] = hi[bf]; if (mpx) nss[to] = ssid[bf]; }
            if (best < scr) best = scr;
        }
        for (s = 0; s < ne; ++s) { sc[s] = nsc[s]; hi[s] = nhi[s]; if (mpx) ssid[s] = nss[s]; }
        return best;
    }
    if (ne == 3 && !mpx) {              /* hmm.c:592-671 */
        int32_t s3, s2, s1, s0, t0, t1, t2;
        s2 = sc[2] + S3SEN(ssid[0], 2); s1 = sc[1] + S3SEN(ssid[0], 1); s0 = sc[0] + S3SEN(ssid[0], 0);
        t0 = t1 = S3W; t2 = INT32_MIN;
        if (s2 > S3W) { t1 = s2 + S3TP(2, 3); t0 = s2 + S3TP(2, 2); }
        if (s1 > S3W && S3TP(1, 3) > S3W) t2 = s1 + S3TP(1, 3);
        if (t1 > t2) { s3 = t1; *out_hi = hi[2]; } else { s3 = t2; *out_hi = hi[1]; }
        if (s3 < S3W) s3 = S3W;
        *out_sc = s3; best = s3;
        t1 = t2 = S3W;
        if (s1 > S3W) t1 = s1 + S3TP(1, 2);
        if (S3TP(0, 2) > S3W) t2 = s0 + S3TP(0, 2);
        S3_PICK3(s2, t0, t1, t2, 1, 0);
        if (from >= 0) hi[2] = hi[from];
        if (s2 < S3W) s2 = S3W;
        if (s2 > best) best = s2;
        sc[2] = s2;
        t0 = t1 = S3W;
        if (s1 > S3W) t0 = s1 + S3TP(1, 1);
        if (s0 > S3W) t1 = s0 + S3TP(0, 1);
        if (t0 > t1) s1 = t0; else { s1 = t1; hi[1] = hi[0]; }
        if (s1 < S3W) s1 = S3W;
        if (s1 > best) best = s1;
        sc[1] = s1;
        s0 += S3TP(0, 0);
        if (s0 < S3W) s0 = S3W;
        if (s0 > best) best = s0;
        sc[0] = s0;
        return best;
    }
    if (ne == 3) {                      /* mpx, hmm.c:673-774 */
        int32_t s3, s2, s1, s0, t0, t1, t2 = INT32_MIN;
        if (ssid[2] == -1) s2 = t1 = S3W;
        else { s2 = sc[2] + S3SEN(ssid[2], 2); if (s2 < S3W) s2 = S3W; t1 = s2 + S3TP(2, 3); }
        if (ssid[1] == -1) s1 = S3W;
        else { s1 = sc[1] + S3SEN(ssid[1], 1); if (s1 < S3W) s1 = S3W; t2 = s1 + S3TP(1, 3); }
        if (t1 > t2) { s3 = t1; *out_hi = hi[2]; } else { s3 = t2; *out_hi = hi[1]; }
        if (s3 < S3W) s3 = S3W;
        *out_sc = s3; best = s3;
        s0 = sc[0] + S3SEN(ssid[0], 0);
        if (s0 < S3W) s0 = S3W;
        t0 = t1 = S3W;
        if (s2 != S3W) t0 = s2 + S3TP(2, 2);
        if (s1 != S3W) t1 = s1 + S3TP(1, 2);
        if (S3TP(0, 2) > S3W) t2 = s0 + S3TP(0, 2);
        S3_PICK3(s2, t0, t1, t2, 1, 0);
        if (from >= 0) { hi[2] = hi[from]; ssid[2] = ssid[from]; }
        if (s2 < S3W) s2 = S3W;
        if (s2 > best) best = s2;
        sc[2] = s2;
        t0 = S3W;
        if (s1 != S3W) t0 = s1 + S3TP(1, 1);
        t1 = s0 + S3TP(0, 1);
        if (t0 > t1) s1 = t0; else { s1 = t1; hi[1] = hi[0]; ssid[1] = ssid[0]; }
        if (s1 < S3W) s1 = S3W;
        if (s1 > best) best = s1;
        sc[1] = s1;
        s0 += S3TP(0, 0);
        if (s0 < S3W) s0 = S3W;
        if (s0 > best) best = s0;
        sc[0] = s0;
        return best;
    }
    if (!mpx) {                         /* 5 states, hmm.c:285-414 */
        int32_t s5, s4, s3, s2, s1, s0, t0, t1, t2;
        s4 = sc[4] + S3SEN(ssid[0], 4); s3 = sc[3] + S3SEN(ssid[0], 3);
        if (s3 > S3W) {
            t1 = s4 + S3TP(4, 5); t2 = s3 + S3TP(3, 5);
            if (t1 > t2) { s5 = t1; *out_hi = hi[4]; } else { s5 = t2; *out_hi = hi[3]; }
            if (s5 < S3W) s5 = S3W;
            *out_sc = s5; best = s5;
        }
        s2 = sc[2] + S3SEN(ssid[0], 2);
        if (s2 > S3W) {
            t0 = s4 + S3TP(4, 4); t1 = s3 + S3TP(3, 4); t2 = s2 + S3TP(2, 4);
            S3_PICK3(s4, t0, t1, t2, 3, 2);
            if (from >= 0) hi[4] = hi[from];
            if (s4 < S3W) s4 = S3W;
            if (s4 > best) best = s4;
            sc[4] = s4;
        }
        s1 = sc[1] + S3SEN(ssid[0], 1);
        if (s1 > S3W) {
            t0 = s3 + S3TP(3, 3); t1 = s2 + S3TP(2, 3); t2 = s1 + S3TP(1, 3);
            S3_PICK3(s3, t0, t1, t2, 2, 1);
            if (from >= 0) hi[3] = hi[from];
            if (s3 < S3W) s3 = S3W;
            if (s3 > best) best = s3;
            sc[3] = s3;
        }
        s0 = sc[0] + S3SEN(ssid[0], 0);
        t0 = s2 + S3TP(2, 2); t1 = s1 + S3TP(1, 2); t2 = s0 + S3TP(0, 2);
        S3_PICK3(s2, t0, t1, t2, 1, 0);
        if (from >= 0) hi[2] = hi[from];
        if (s2 < S3W) s2 = S3W;
        if (s2 > best) best = s2;
        sc[2] = s2;
        t0 = s1 + S3TP(1, 1); t1 = s0 + S3TP(0, 1);
        if (t0 > t1) s1 = t0; else { s1 = t1; hi[1] = hi[0]; }
        if (s1 < S3W) s1 = S3W;
        if (s1 > best) best = s1;
        sc[1] = s1;
        s0 += S3TP(0, 0);
        if (s0 < S3W) s0 = S3W;
        if (s0 > best) best = s0;
        sc[0] = s0;
        return best;
    }
    {                                   /* 5 states, mpx, hmm.c:416-586 */
        int32_t s5, s4, s3, s2, s1, s0, t0, t1, t2;
        if (ssid[4] == -1) s4 = t1 = S3W; else { s4 = sc[4] + S3SEN(ssid[4], 4); t1 = s4 + S3TP(4, 5); }
        if (ssid[3] == -1) s3 = t2 = S3W; else { s3 = sc[3] + S3SEN(ssid[3], 3); t2 = s3 + S3TP(3, 5); }
        if (t1 > t2) { s5 = t1; *out_hi = hi[4]; } else { s5 = t2; *out_hi = hi[3]; }
        if (s5 < S3W) s5 = S3W;
        *out_sc = s5; best = s5;
        if (ssid[2] == -1) s2 = t2 = S3W; else { s2 = sc[2] + S3SEN(ssid[2], 2); t2 = s2 + S3TP(2, 4); }
        t0 = t1 = S3W;
        if (s4 != S3W) t0 = s4 + S3TP(4, 4);
        if (s3 != S3W) t1 = s3 + S3TP(3, 4);
        S3_PICK3(s4, t0, t1, t2, 3, 2);
        if (from >= 0) { hi[4] = hi[from]; ssid[4] = ssid[from]; }
        if (s4 < S3W) s4 = S3W;
        if (s4 > best) best = s4;
        sc[4] = s4;
        if (ssid[1] == -1) s1 = t2 = S3W; else { s1 = sc[1] + S3SEN(ssid[1], 1); t2 = s1 + S3TP(1, 3); }
        t0 = t1 = S3W;
        if (s3 != S3W) t0 = s3 + S3TP(3, 3);
        if (s2 != S3W) t1 = s2 + S3TP(2, 3);
        S3_PICK3(s3, t0, t1, t2, 2, 1);
        if (from >= 0) { hi[3] = hi[from]; ssid[3] = ssid[from]; }
        if (s3 < S3W) s3 = S3W;
        if (s3 > best) best = s3;
        sc[3] = s3;
        s0 = sc[0] + S3SEN(ssid[0], 0);
        t0 = t1 = S3W;
        if (s2 != S3W) t0 = s2 + S3TP(2, 2);
        if (s1 != S3W) t1 = s1 + S3TP(1, 2);
        t2 = s0 + S3TP(0, 2);
        S3_PICK3(s2, t0, t1, t2, 1, 0);
        if (from >= 0) { hi[2] = hi[from]; ssid[2] = ssid[from]; }
        if (s2 < S3W) s2 = S3W;
        if (s2 > best) best = s2;
        sc[2] = s2;
        t0 = S3W;
        if (s1 != S3W) t0 = s1 + S3TP(1, 1);
        t1 = s0 + S3TP(0, 1);
        if (t0 > t1) s1 = t0; else { s1 = t1; hi[1] = hi[0]; ssid[1] = ssid[0]; }
        if (s1 < S3W) s1 = S3W;
        if (s1 > best) best = s1;
        sc[1] = s1;
        s0 += S3TP(0, 0);
        if (s0 < S3W) s0 = S3W;
        if (s0 > best) best = s0;
        sc[0] = s0;
        return best;
    }
}

int32_t orc_s3hmm_eval_batch(int ne, int n_hmm, const int32_t *tp, int n_tmat, const int16_t *sseq, int n_sseq,
                             const int32_t *sen, int32_t *score, int32_t *history, int32_t *out_score,
                             int32_t *out_history, int32_t *ssid, const int32_t *tmatid, const uint8_t *mpx,
                             int32_t *bestscore, int repeat)
{
    int32_t best = S3W;
    int r, i;
    (void)n_tmat; (void)n_sseq;
    for (r = 0; r < (repeat > 0 ? repeat : 1); ++r) {
        best = S3W;
        for (i = 0; i < n_hmm; ++i) {
            int32_t b = s3hmm_one(ne, tp + (size_t)tmatid[i] * ne * (ne + 1), sseq, sen, score + (size_t)i * ne,
                                  history + (size_t)i * ne, ssid + (size_t)i * ne, mpx[i], &out_score[i], &out_history[i]);
            bestscore[i] = b;
            if (b > best) best = b;
        }
    }
    return best;
}

/* ------------------------------------------------------------------------------------------------
 * prune_root_chan + prune_nonroot_chan (ngram_search_fwdtree.c:714-869), sequential restatement on
 * arrays (see sphinx_oracle.h).  WORST_SCORE = 0xE0000000 (hmm.h:74), BETTER_THAN is > (hmm.h:85).
 * hmm_enter (hmm.c:197-203) sets score[0], history[0] and frame; hmm_clear_scores (hmm.c:169-181)
 * sets every state score, the exit score and bestscore to WORST_SCORE and leaves histories and
 * frame alone. */
static void orc_prune_exits(int c, int32_t nps, int n_chan, const int32_t *pw_off, const int32_t *pw_wid,
                            const int32_t *pw_lastphone, const int32_t *par, const int32_t *pls_pen,
                            const int32_t *out_history, int32_t lastphn_thresh, int32_t *cand, int32_t *n_cand)
{
    int k;
    (void)n_chan;
    if (!(par[7] || nps > lastphn_thresh)) return;                 /* :763, :845 */
    for (k = pw_off[c]; k < pw_off[c + 1]; ++k) {
        const int32_t pl = nps + (par[7] ? pls_pen[pw_lastphone[k]] : 0);
        if (pl > lastphn_thresh) {                                 /* :771, :853 */
            int32_t *e = cand + 3 * (*n_cand)++;
            e[0] = pw_wid[k]; e[1] = pl - par[6]; e[2] = out_history[c];
        }
    }
}

void orc_fwdtree_prune(int n_root, int n_chan, int ne, const int32_t *child_off, const int32_t *child,
                       const int32_t *ciphone, const int32_t *pw_off, const int32_t *pw_wid,
                       const int32_t *pw_lastphone, const int32_t *par, const int32_t *pls_pen,
                       const int32_t *acl, int n_act, int32_t *score, int32_t *history, int32_t *out_score,
                       int32_t *out_history, int32_t *bestscore, int32_t *frame, int32_t *nacl, int32_t *n_nacl,
                       int32_t *cand, int32_t *n_cand)
{
    const int32_t WORST = (int32_t)0xE0000000;
    const int32_t fi = par[0], nf = fi + 1;
    const int32_t thresh = par[1] + par[2], newphone_thresh = par[1] + par[3], lastphn_thresh = par[1] + par[4];
    const int has_pls = par[7];
    int i, k, s, nn = 0;
    *n_cand = 0;
    /* prune_root_chan :733-789 */
    for (i = 0; i < n_root; ++i) {
        int32_t nps;
        if (frame[i] < fi) continue;                                   /* :737 */
        if (!(bestscore[i] > thresh)) continue;                        /* :740 */
        frame[i] = nf;
        nps = out_score[i] + par[5];
        if (has_pls || nps > newphone_thresh) {                        /* :746 */
            for (k = child_off[i]; k < child_off[i + 1]; ++k) {
                const int c = child[k];
                const int32_t pl = nps + (has_pls ? pls_pen[ciphone[c]] : 0);
                if (pl > newphone_thresh && (frame[c] < fi || pl > score[c])) {   /* :750-752 */
                    score[c] = pl; history[c] = out_history[i]; frame[c] = nf;
                    nacl[nn++] = c;
                }
            }
        }
        orc_prune_exits(i, nps, n_chan, pw_off, pw_wid, pw_lastphone, par, pls_pen, out_history, lastphn_thresh, cand, n_cand);
    }
    /* prune_nonroot_chan :811-868 */
    for (i = 0; i < n_act; ++i) {
        const int h = acl[i];
        if (bestscore[h] > thresh) {                                   /* :815 */
            int32_t nps;
            if (frame[h] != nf) { frame[h] = nf; nacl[nn++] = h; }     /* :817-820 */
            nps = out_score[h] + par[5];
            if (has_pls || nps > newphone_thresh) {                    /* :824 */
                for (k = child_off[h]; k < child_off[h + 1]; ++k) {
                    const int c = child[k];
                    const int32_t pl = nps + (has_pls ? pls_pen[ciphone[c]] : 0);
                    if (pl > newphone_thresh && (frame[c] < fi || pl > score[c])) {   /* :828-832 */
                        if (frame[c] != nf) nacl[nn++] = c;            /* :833-836 */
                        score[c] = pl; history[c] = out_history[h]; frame[c] = nf;
                    }
                }
            }
            orc_prune_exits(h, nps, n_chan, pw_off, pw_wid, pw_lastphone, par, pls_pen, out_history, lastphn_thresh, cand, n_cand);
        }
        else if (frame[h] != nf) {                                     /* :863-865 */
            for (s = 0; s < ne; ++s) score[(size_t)s * n_chan + h] = WORST;
            out_score[h] = WORST;
            bestscore[h] = WORST;
        }
    }
    *n_nacl = nn;
}

/* ------------------------------------------------------------------------------------------------
 * phone_loop_search_step (phone_loop_search.c:253-291), see sphinx_oracle.h. */
int32_t orc_phone_loop_step(int n_phones, int ne, const uint8_t *tp, const int16_t *senscr, int32_t *par,
                            int32_t *score, int32_t *history, int32_t *out_score, int32_t *out_history,
                            int32_t *bestscore, int32_t *frame, uint16_t *senid, const int16_t *tmatid,
                            int32_t *renorm)
{
    const int32_t fi = par[0], nf = fi + 1, beam = par[2], pbeam = par[3], pip = par[4];
    int32_t bs, thresh;
    int i, j, s;
    *renorm = 0;
    if (par[1] + 2 * beam < W) {                                   /* :273: WORSE_THAN WORST_SCORE */
        const int32_t norm = par[1];
        *renorm = 1;
        for (i = 0; i < n_phones; ++i) {                           /* hmm_normalize, hmm.c:205-216 */
            for (s = 0; s < ne; ++s)
                if (score[i * ne + s] > W) score[i * ne + s] -= norm;
            if (out_score[i] > W) out_score[i] -= norm;
        }
    }
    bs = W;                                                        /* evaluate_hmms :186-210 */
    for (i = 0; i < n_phones; ++i) {
        int32_t b;
        if (frame[i] < fi) continue;
        b = hmm_eval_one(ne, tp + (long)tmatid[i] * ne * (ne + 1), NULL, senscr, score + (long)i * ne,
                         history + (long)i * ne, &out_score[i], &out_history[i], senid + (long)i * ne, 0);
        bestscore[i] = b;
        if (b > bs) bs = b;
    }
    par[1] = bs;
    thresh = bs + beam;                                            /* prune_hmms :212-233 */
    for (i = 0; i < n_phones; ++i) {
        if (frame[i] < fi) continue;
        if (bestscore[i] > thresh) frame[i] = nf;
        else {
            for (s = 0; s < ne; ++s) score[i * ne + s] = W;
            out_score[i] = W; bestscore[i] = W;
        }
    }
    thresh = bs + pbeam;                                           /* phone_transition :235-268 */
    for (i = 0; i < n_phones; ++i) {
        int32_t nps;
        if (frame[i] != nf) continue;
        nps = out_score[i] + pip;
        if (nps > thresh)
            for (j = 0; j < n_phones; ++j)
                if (frame[j] < fi || nps > score[j * ne]) {
                    score[j * ne] = nps; history[j * ne] = out_history[i]; frame[j] = nf;
                }
    }
    return bs;
}
