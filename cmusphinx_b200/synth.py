"""Seeded synthetic acoustic models, features and HMM populations for the
BASELINE.json configurations (recipes: SURVEY.md section 8(d)).  Data
generation only -- no scoring here.
"""
import numpy as np


def cont_model(n_sen=5000, n_density=32, dim=39, seed=1234, n_feat=1):
    """Fully-continuous model (.cont.: one codebook per senone).  Returns raw
    (un-precomputed) float32 arrays: mean/var [n_sen][n_density][dim*n_feat laid
    out stream-major inside the row as the reference does], mixw [n_sen][n_feat][n_density]."""
    rng = np.random.default_rng(seed)
    veclen = dim * n_feat
    mean = rng.standard_normal((n_sen, n_density, veclen)).astype(np.float32)
    var = np.exp(rng.uniform(np.log(0.05), np.log(5.0), (n_sen, n_density, veclen))).astype(np.float32)
    mixw = rng.dirichlet(np.ones(n_density), (n_sen, n_feat)).astype(np.float32)
    return mean, var, mixw


def to_ref_layout(arr, n_feat, dim):
    """[mgau][density][n_feat*dim] (per-density concatenated streams) ->
    the reference's flat [mgau][feat][density][dim] order."""
    m, d, _ = arr.shape
    return np.ascontiguousarray(arr.reshape(m, d, n_feat, dim).transpose(0, 2, 1, 3)).reshape(m, -1)


def cont_features(mean, var, T, seed=5678):
    """Half of the frames sit near a random Gaussian (so the top-N is contested),
    half are N(0, 2)."""
    rng = np.random.default_rng(seed)
    n_sen, n_density, veclen = mean.shape
    x = (rng.standard_normal((T, veclen)) * np.sqrt(2.0)).astype(np.float32)
    near = rng.random(T) < 0.5
    s = rng.integers(0, n_sen, T)
    d = rng.integers(0, n_density, T)
    noise = rng.standard_normal((T, veclen)).astype(np.float32)
    xn = mean[s, d] + noise * np.sqrt(var[s, d])
    x[near] = xn[near]
    return np.ascontiguousarray(x, np.float32)


def bakis_tmat(n_tmat=50, n_emit=3, seed=7, skip_fraction=0.5):
    """Upper-triangular left-to-right matrices, about half with a 0->2 style
    skip so the hmm.c skip-transition quirks are exercised."""
    rng = np.random.default_rng(seed)
    tp = np.zeros((n_tmat, n_emit, n_emit + 1), np.float32)
    for t in range(n_tmat):
        skip = rng.random() < skip_fraction
        for i in range(n_emit):
            p = rng.dirichlet(np.ones(3) * 2.0)
            tp[t, i, i] = p[0]
            tp[t, i, i + 1] = p[1]
            if skip and i + 2 <= n_emit:
                tp[t, i, i + 2] = p[2]
            tp[t, i] /= tp[t, i].sum()
    return tp


def hmm_population(n_hmm, n_emit, n_sen, n_tmat, n_sseq, seed=42, mpx_fraction=0.1):
    """HMM-major arrays as a dict (score/history [n_hmm][n_emit], ...), plus an
    sseq table; 15 % of state scores are WORST_SCORE, 5 % of HMMs freshly entered."""
    rng = np.random.default_rng(seed)
    W = np.int32(-0x20000000)
    score = -rng.integers(0, 1 << 20, (n_hmm, n_emit)).astype(np.int32)
    score[rng.random((n_hmm, n_emit)) < 0.15] = W
    fresh = rng.random(n_hmm) < 0.05
    score[fresh, 1:] = W
    score[fresh, 0] = -rng.integers(0, 1 << 12, fresh.sum()).astype(np.int32)
    history = rng.integers(-1, 1 << 16, (n_hmm, n_emit)).astype(np.int32)
    out_score = -rng.integers(0, 1 << 20, n_hmm).astype(np.int32)
    out_score[rng.random(n_hmm) < 0.3] = W
    out_history = rng.integers(-1, 1 << 16, n_hmm).astype(np.int32)
    mpx = (rng.random(n_hmm) < mpx_fraction).astype(np.uint8)
    sseq = rng.integers(0, n_sen, (n_sseq, n_emit)).astype(np.uint16)
    senid = rng.integers(0, n_sen, (n_hmm, n_emit)).astype(np.uint16)
    m = mpx.astype(bool)
    ss = rng.integers(0, n_sseq, (int(m.sum()), n_emit)).astype(np.uint16)
    bad = rng.random(ss.shape) < 0.2
    bad[:, 0] = False
    ss[bad] = 0xFFFF
    senid[m] = ss
    tmatid = rng.integers(0, n_tmat, n_hmm).astype(np.int16)
    bestscore = np.full(n_hmm, W, np.int32)
    return dict(score=score, history=history, out_score=out_score, out_history=out_history, senid=senid,
                tmatid=tmatid, mpx=mpx, bestscore=bestscore, sseq=sseq)


def senscr_frames(n_frames, n_sen, seed=99):
    rng = np.random.default_rng(seed)
    s = np.abs(rng.standard_normal((n_frames, n_sen)) * 800.0)
    return np.clip(s, 0, 32767).astype(np.int16)


def s3_model(n_sen=600, n_ci_sen=30, n_density=8, dim=39, seed=77):
    """sphinx3-style fully continuous model: CI senones first, every CD senone
    has a CI parent (mdef cd2cisen); CD Gaussians are perturbations of their
    parent's so the CI beam is informative.  A few components are left
    uninitialised (zero variance vector) to exercise mgau_uninit_compact
    (S3/libam/cont_mgau.c:700-790) and one senone has an all-zero mixw row."""
    rng = np.random.default_rng(seed)
    cd2ci = np.arange(n_sen, dtype=np.int32)
    cd2ci[n_ci_sen:] = rng.integers(0, n_ci_sen, n_sen - n_ci_sen)
    mean = np.zeros((n_sen, n_density, dim), np.float32)
    var = np.zeros((n_sen, n_density, dim), np.float32)
    mean[:n_ci_sen] = rng.standard_normal((n_ci_sen, n_density, dim)) * 1.5
    var[:n_ci_sen] = np.exp(rng.uniform(np.log(0.2), np.log(4.0), (n_ci_sen, n_density, dim)))
    par = cd2ci[n_ci_sen:]
    mean[n_ci_sen:] = mean[par] + rng.standard_normal((n_sen - n_ci_sen, n_density, dim)) * 0.4
    var[n_ci_sen:] = var[par] * np.exp(rng.uniform(-0.7, 0.3, (n_sen - n_ci_sen, n_density, dim)))
    var[rng.random(var.shape) < 0.002] = 5e-5          # some values under the 1e-4 floor
    mixw = rng.dirichlet(np.ones(n_density), n_sen).astype(np.float32) * 1000.0   # un-normalised counts
    mixw[mixw < 1e-3] = 0.0
    for s in rng.integers(n_ci_sen, n_sen, max(1, n_sen // 50)):   # uninitialised components
        var[s, rng.integers(0, n_density)] = 0.0
    mixw[n_sen - 1] = 0.0
    return mean.astype(np.float32), var.astype(np.float32), mixw.astype(np.float32), cd2ci, n_ci_sen


def s3_features(mean, var, T, seed=88):
    """A slowly moving trajectory through the CI Gaussians plus noise, so
    consecutive frames keep similar best CI phones."""
    rng = np.random.default_rng(seed)
    n_sen, n_density, dim = mean.shape
    x = np.zeros((T, dim), np.float32)
    s = rng.integers(0, n_sen); d = rng.integers(0, n_density)
    for t in range(T):
        if rng.random() < 0.15:
            s = rng.integers(0, n_sen); d = rng.integers(0, n_density)
        v = np.where(var[s, d] > 0, var[s, d], 1.0)
        x[t] = mean[s, d] + rng.standard_normal(dim) * np.sqrt(v) * 0.9
    return x


def s3_active(n_sen, n_ci_sen, T, seed=99, p_on=0.12, p_off=0.25):
    """Bursty per-frame active-senone flags [T][n_sen] (the search keeps a
    senone active for a few frames); CI entries are left 0 -- the scorer
    forces them to 1."""
    rng = np.random.default_rng(seed)
    act = np.zeros((T, n_sen), np.uint8)
    cur = rng.random(n_sen) < 0.4
    for t in range(T):
        on = rng.random(n_sen) < p_on
        off = rng.random(n_sen) < p_off
        cur = (cur & ~off) | on
        act[t] = cur
    act[:, :n_ci_sen] = 0
    return act
