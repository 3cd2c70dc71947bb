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


# ---------------------------------------------------------------- sub-vector quantised shortlists (S3/libam/subvq.c)
def _svq_models(tmp_path, n_sv, vqsize):
    mean, var, p, r, cd2ci, n_ci = _synthetic(tmp_path)
    valid = ~np.all(var == 0, axis=2)                 # what mgau_uninit_compact keeps (synth.s3_model zeroes whole vectors)
    q = orc.synthetic_subvq(mean, var, valid, n_sv, vqsize)
    path = str(tmp_path / "synthetic.subvq")
    orc.write_subvq(path, q)
    return mean, var, p, r, n_ci, q, path


@needs_ref
@pytest.mark.parametrize("n_sv,vqeval,max_sv", [(3, 3, -1), (3, 2, -1), (3, 1, -1), (1, 3, -1), (2, 3, -1), (4, 2, -1), (3, 3, 2)])
def test_subvq_tables_distances_and_shortlists_match_reference(tmp_path, n_sv, vqeval, max_sv):
    mean, var, p, r, n_ci, q, path = _svq_models(tmp_path, n_sv, 24)
    beam = 1e-3
    orc.port_set_svq(p, orc.read_subvq(path), max_sv=max_sv, vqeval=vqeval, subvqbeam=beam)
    orc.ref_set_svq(r, path, max_sv=max_sv, vqeval=vqeval, subvqbeam=beam)
    r.set_fast(); p.set_fast()
    L = orc.ref_s3()
    use = r.svq_dims[0]
    assert use == p.n_sv_use and r.svq_dims[1] == 24
    d = np.zeros(6, np.int32); L.ref_s3_svq_dims(r.h, orc._p(d, orc.C.c_int32))
    assert int(d[5]) == orc.port.orc_s3_svq_beam(p.h)
    for sv in range(use):
        n = int(q["veclen"][sv])
        fd = np.zeros(n, np.int32); m1 = np.zeros((24, n), np.float32); v1 = np.zeros((24, n), np.float32)
        l1 = np.zeros(24, np.float32); s1 = np.zeros(1, np.float64)
        assert L.ref_s3_svq_tables(r.h, sv, orc._p(fd, orc.C.c_int32), orc._p(m1, orc.C.c_float), orc._p(v1, orc.C.c_float),
                                   orc._p(l1, orc.C.c_float), orc._p(s1, orc.C.c_double)) == n
        m2 = np.zeros_like(m1); v2 = np.zeros_like(v1); l2 = np.zeros_like(l1); s2 = np.zeros(1, np.float64)
        orc.port.orc_s3_svq_tables(p.h, sv, orc._p(m2, orc.C.c_float), orc._p(v2, orc.C.c_float), orc._p(l2, orc.C.c_float),
                                   orc._p(s2, orc.C.c_double))
        assert np.array_equal(fd, q["featdim"][sv])
        assert np.array_equal(m1, m2) and np.array_equal(v1, v2) and np.array_equal(l1, l2) and s1[0] == s2[0]
    S, M = mean.shape[:2]
    map_r = np.zeros((S, M, use), np.int32); map_p = np.zeros_like(map_r)
    L.ref_s3_svq_map(r.h, orc._p(map_r, orc.C.c_int32)); orc.port.orc_s3_svq_map(p.h, orc._p(map_p, orc.C.c_int32))
    assert np.array_equal(map_r, map_p) and (map_r < 0).any()
    T = 25
    feat = synth.s3_features(mean, var, T)
    vq_r = np.zeros((T, use * 24), np.int32)
    L.ref_s3_svq_vqdist(r.h, orc._p(feat, orc.C.c_float), T, orc._p(vq_r, orc.C.c_int32))
    nc = p.params()[0]
    for t in range(T):
        row = np.ascontiguousarray(feat[t])
        orc.port.orc_s3_svq_eval(p.h, orc._p(row, orc.C.c_float))
        vq_p = np.zeros(use * 24, np.int32); orc.port.orc_s3_svq_dist(p.h, orc._p(vq_p, orc.C.c_int32))
        assert np.array_equal(vq_p, vq_r[t]), f"vqdist, frame {t}"
    # shortlists against the last frame's distances (both sides hold them now)
    n_short = 0
    for s in range(0, S, 7):
        fl = np.zeros(M, np.uint8)
        ng_r = L.ref_s3_svq_shortlist(r.h, s, orc._p(fl, orc.C.c_uint8))
        ng_p = orc.port.orc_s3_svq_shortlist(p.h, s)
        assert ng_r == ng_p and 1 <= ng_p <= nc[s]
        n_short += ng_p < nc[s]
    assert n_short > 0          # the beam really prunes
    p.free(); r.free()


@needs_ref
@pytest.mark.parametrize("n_sv,vqeval,beam,cfg", [
    (3, 3, 1e-3, dict()), (3, 2, 1e-2, dict(ci_pbeam=1e-40)), (3, 1, 1e-1, dict(ci_pbeam=1e-40, max_cd=60)),
    (1, 3, 1e-3, dict(ci_pbeam=1e-30, ds_ratio=3)), (2, 3, 0.5, dict(ci_pbeam=1e-40, max_cd=80, ds_ratio=2, tighten=0.3))])
def test_subvq_frame_eval_matches_reference(tmp_path, n_sv, vqeval, beam, cfg):
    """approx_cont_mgau_ci_eval + approx_cont_mgau_frame_eval with a sub-VQ model: scores, frame bests, active
    flags and the best-index state, with the CI beam / dynamic beam / -ds layers on top."""
    mean, var, p, r, n_ci, q, path = _svq_models(tmp_path, n_sv, 16)
    orc.port_set_svq(p, orc.read_subvq(path), vqeval=vqeval, subvqbeam=beam)
    orc.ref_set_svq(r, path, vqeval=vqeval, subvqbeam=beam)
    T = 50
    feat = synth.s3_features(mean, var, T)
    act = synth.s3_active(mean.shape[0], n_ci, T)
    for m in (p, r):
        m.set_fast(**cfg)
    plain = orc.PortS3(mean, var, synth.s3_model()[2], synth.s3_model()[3], n_ci)
    plain.set_fast(**cfg)
    for active, f0 in [(act, 0), (None, 3)]:
        p.utt_reset(); r.utt_reset(); plain.utt_reset()
        a, b = p.eval_utt(feat, active, f0), r.eval_utt(feat, active, f0)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
        sa, sb = p.state(), r.state()
        assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1])
        c = plain.eval_utt(feat, active, f0)
        assert not np.array_equal(a[0], c[0])          # the shortlists change scores: the layer is really exercised
    p.free(); r.free(); plain.free()


@needs_ref
def test_subvq_real_model_matches_reference():
    """hub4_cd_continuous_8gau_1s_c_d_dd with the tree's own test.subvq (1 sub-vector x 16 codewords)."""
    from cmusphinx_b200 import engine
    mf, vf, wf, md, sv = (os.path.join(CONT, n) for n in ("means", "variances", "mixture_weights", "mdef", "test.subvq"))
    if not os.path.exists(sv):
        pytest.skip("test.subvq not bundled")
    r = orc.RefS3(mf, vf, wf, md, varfloor=1e-4, mixwfloor=1e-7)
    mean, var, mixw = engine.read_s3_cont_arrays(mf, vf, wf)
    p = orc.PortS3(mean, var, mixw, r.cd2cisen(), r.n_ci_sen)
    orc.port_set_svq(p, orc.read_subvq(sv), subvqbeam=3e-3)
    orc.ref_set_svq(r, sv, subvqbeam=3e-3)
    rng = np.random.default_rng(3)
    T = 10
    idx = rng.integers(0, mean.shape[0], T)
    feat = (mean[idx, 0] + rng.standard_normal((T, mean.shape[2])) * np.sqrt(var[idx, 0])).astype(np.float32)
    act = synth.s3_active(mean.shape[0], r.n_ci_sen, T, p_on=0.1)
    for cfg in (dict(), dict(ci_pbeam=1e-40)):
        p.set_fast(**cfg); r.set_fast(**cfg); p.utt_reset(); r.utt_reset()
        a, b = p.eval_utt(feat, act), r.eval_utt(feat, act)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    p.free(); r.free()


# ---------------------------------------------------------------- Gaussian selector (S3/libam/gs.c)
@needs_ref
@pytest.mark.parametrize("cfg", [dict(), dict(ci_pbeam=1e-40, max_cd=60), dict(ci_pbeam=1e-30, ds_ratio=3)])
def test_gaussian_selector_matches_reference(tmp_path, cfg):
    """gs_read / gc_compute_closest_cw / gs_mgau_shortlist inside approx_cont_mgau_ci_eval + _frame_eval."""
    mean, var, p, r, cd2ci, n_ci = _synthetic(tmp_path)
    cw, bits = orc.synthetic_gs(mean, 32)
    path = str(tmp_path / "synthetic.gs")
    orc.write_gs(path, cw, bits, mean.shape[1])
    orc.port_set_gs(p, cw, bits); orc.ref_set_gs(r, path)
    T = 50
    feat = synth.s3_features(mean, var, T)
    act = synth.s3_active(mean.shape[0], n_ci, T)
    want = np.zeros(T, np.int32)
    orc.ref_s3().ref_s3_gs_closest(r.h, orc._p(feat, orc.C.c_float), T, orc._p(want, orc.C.c_int32))
    got = np.array([orc.port.orc_s3_gs_closest(p.h, orc._p(np.ascontiguousarray(feat[t]), orc.C.c_float)) for t in range(T)])
    assert np.array_equal(got, want) and want.min() > 0 and len(set(want.tolist())) > 3
    for m in (p, r):
        m.set_fast(**cfg)
    plain = orc.PortS3(mean, var, synth.s3_model()[2], cd2ci, n_ci); plain.set_fast(**cfg)
    for active, f0 in [(act, 0), (None, 3)]:
        p.utt_reset(); r.utt_reset(); plain.utt_reset()
        a, b = p.eval_utt(feat, active, f0), r.eval_utt(feat, active, f0)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
        sa, sb = p.state(), r.state()
        assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1])
        assert not np.array_equal(a[0], plain.eval_utt(feat, active, f0)[0])
    p.free(); r.free(); plain.free()
