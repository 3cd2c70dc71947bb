"""Pins the sphinx3 half of the oracle (oracle/sphinx_oracle.c: orc_s3_*) to
the reference's own compiled functions (oracle/_ref/libref_shim_s3.so:
mgau_init, mgau_eval, approx_cont_mgau_ci_eval, approx_cont_mgau_frame_eval)
and to the committed golden vectors generated from them."""
import os

import numpy as np
import pytest

import orc
from cmusphinx_b200 import s3io, synth

needs_ref = pytest.mark.skipif(not orc.have_ref_s3(), reason="oracle/_ref (sphinx3) not built")
CONT = os.path.join(orc.DATA_DIR, "hmm", "cont")


def _synthetic(tmp_path, **kw):
    mean, var, mixw, cd2ci, n_ci = synth.s3_model(**kw)
    D = mean.shape[2]
    mf, vf, wf = (str(tmp_path / n) for n in ("means", "variances", "mixture_weights"))
    s3io.write_gauden(mf, mean, [D]); s3io.write_gauden(vf, var, [D]); s3io.write_mixw(wf, mixw[:, None, :])
    p = orc.PortS3(mean, var, mixw, cd2ci, n_ci)
    r = orc.RefS3(mf, vf, wf, None, cd2ci, n_ci)
    return mean, var, p, r, cd2ci, n_ci


@needs_ref
def test_s3_precompute_matches_reference(tmp_path):
    mean, var, p, r, cd2ci, n_ci = _synthetic(tmp_path)
    a, b = p.params(), r.params()
    assert np.array_equal(a[0], b[0]) and a[0].min() < a[0].max()      # compaction happened
    for x, y, name in zip(a[1:], b[1:], ("mean", "var", "lrd", "mixw", "scal")):
        S, M = a[0].shape[0], x.shape[1] if x.ndim > 1 else 0
        if x.ndim > 1:
            valid = np.arange(M)[None, :] < a[0][:, None]
            assert np.array_equal(x[valid], y[valid]), name
        else:
            assert np.array_equal(x, y), name
    assert orc.port.orc_s3_ci_pbeam(p.h) == r.ci_pbeam
    p.free(); r.free()


S3_CFGS = [
    dict(ci_pbeam=1e-80, max_cd=100000, ds_ratio=1),     # defaults: everything computed
    dict(ci_pbeam=1e-40, max_cd=100000, ds_ratio=1),     # CI beam prunes ~40 % -> CI back-off and best-Gaussian back-off
    dict(ci_pbeam=1e-40, max_cd=60, ds_ratio=1),         # dynamic beam (approx_compute_dyn_ci_pbeam)
    dict(ci_pbeam=1e-30, max_cd=100000, ds_ratio=3),     # down-sampling: tightened beam, best-index chains
    dict(ci_pbeam=1e-40, max_cd=80, ds_ratio=2, tighten=0.3),
]


@needs_ref
@pytest.mark.parametrize("cfg", S3_CFGS)
def test_s3_frame_eval_matches_reference(tmp_path, cfg):
    mean, var, p, r, cd2ci, n_ci = _synthetic(tmp_path)
    T = 60
    feat = synth.s3_features(mean, var, T)
    act = synth.s3_active(mean.shape[0], n_ci, T)
    for m in (p, r):
        m.set_fast(**cfg)
    stale0 = (np.arange(mean.shape[0]) * 7 - 1000).astype(np.int32)
    for rep, (active, f0) in enumerate([(act, 0), (None, 0), (act[::-1], 5)]):
        p.utt_reset(); r.utt_reset()
        a = p.eval_utt(feat, active, f0, stale0)
        b = r.eval_utt(feat, active, f0, stale0)
        assert np.array_equal(a[1], b[1]), f"best differs ({cfg}, rep {rep})"
        assert np.array_equal(a[0], b[0]), f"scores differ ({cfg}, rep {rep})"
        if active is not None:
            assert np.array_equal(a[2], b[2])
        sa, sb = p.state(), r.state()
        assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1])
    p.free(); r.free()


@needs_ref
def test_s3_real_model_matches_reference():
    """hub4_cd_continuous_8gau_1s_c_d_dd (6144 x 8 x 39) with its own mdef."""
    from cmusphinx_b200 import engine
    mf, vf, wf, md = (os.path.join(CONT, n) for n in ("means", "variances", "mixture_weights", "mdef"))
    if not os.path.exists(mf):
        pytest.skip("continuous model not bundled")
    r = orc.RefS3(mf, vf, wf, md, varfloor=1e-4, mixwfloor=1e-7)
    cd2ci = r.cd2cisen()
    mean, var, mixw = engine.read_s3_cont_arrays(mf, vf, wf)
    p = orc.PortS3(mean, var, mixw, cd2ci, r.n_ci_sen)
    rng = np.random.default_rng(3)
    T = 12
    idx = rng.integers(0, mean.shape[0], T)
    feat = (mean[idx, 0] + rng.standard_normal((T, mean.shape[2])) * np.sqrt(var[idx, 0])).astype(np.float32)
    act = synth.s3_active(mean.shape[0], r.n_ci_sen, T, p_on=0.1)
    for cfg in (dict(), dict(ci_pbeam=1e-40)):
        p.set_fast(**cfg); r.set_fast(**cfg); p.utt_reset(); r.utt_reset()
        a, b = p.eval_utt(feat, act), r.eval_utt(feat, act)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    p.free(); r.free()
