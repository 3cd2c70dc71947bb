"""sphinx3 flavour of the path (approx_cont_mgau_frame_eval, S3/libam):
 * CPU: the oracle port against the committed golden vectors generated from the
   reference (tests/golden/s3_synth.npz, made by make_golden.py:s3_case);
 * GPU (-m gpu): the CUDA path through the C ABI against the goldens and
   against the oracle on fresh seeded inputs -- bit-exact (int32 scores)."""
import os

import numpy as np
import pytest

import cases
import orc
import cmusphinx_b200 as b
from cmusphinx_b200 import synth


def _cfgs(g):
    return [dict(ci_pbeam=float(c[0]), max_cd=int(c[1]), ds_ratio=int(c[2]), tighten=float(c[3])) for c in g["cfgs"]]


def _check_golden(model, g, state=True):
    n_ci = int(g["n_ci"])
    for i, cfg in enumerate(_cfgs(g)):
        model.set_fast(**cfg); model.utt_reset()
        o, best, a = model.eval_utt(g["feat"], g["act"], int(g["frame0"]), g["stale0"])
        np.testing.assert_array_equal(best, g[f"best{i}"], err_msg=f"cfg {i}")
        np.testing.assert_array_equal(o, g[f"scr{i}"], err_msg=f"cfg {i}")
        np.testing.assert_array_equal(a, g[f"act{i}"])
        if state:
            bi, ut = model.state()
            np.testing.assert_array_equal(bi, g[f"bstidx{i}"]); np.testing.assert_array_equal(ut, g[f"upd{i}"])
    assert (g["scr1"] != g["scr0"]).mean() > 0.05      # the beam really pruned something
    assert n_ci > 0


def _check_params(model, g):
    nc, mean, var, lrd, mixw, scal = model.params()
    np.testing.assert_array_equal(nc, g["n_comp"])
    valid = np.arange(lrd.shape[1])[None, :] < nc[:, None]
    np.testing.assert_array_equal(lrd[valid], g["p_lrd"][valid])
    np.testing.assert_array_equal(mixw[valid], g["p_mixw"][valid])
    np.testing.assert_array_equal(var[:20][valid[:20]], g["p_var_head"][valid[:20]])
    np.testing.assert_array_equal(scal, g["scal"])


def test_s3_port_matches_golden():
    g = cases.load("s3_synth.npz")
    p = orc.PortS3(g["mean"], g["var"], g["mixw"], g["cd2ci"], int(g["n_ci"]))
    assert orc.port.orc_s3_ci_pbeam(p.h) == int(g["ci_pbeam_default"])
    _check_params(p, g)
    _check_golden(p, g)
    p.free()


def test_s3_without_gpu_fails_loudly():
    if b.device_count() > 0:
        pytest.skip("GPU present")
    g = cases.load("s3_synth.npz")
    with pytest.raises(b.B200Error, match="no CUDA device"):
        b.S3Mgau.from_arrays(g["mean"], g["var"], g["mixw"], g["cd2ci"], int(g["n_ci"]))


# ------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_s3_gpu_matches_golden():
    g = cases.load("s3_synth.npz")
    m = b.S3Mgau.from_arrays(g["mean"], g["var"], g["mixw"], g["cd2ci"], int(g["n_ci"]))
    assert m.ci_pbeam == int(g["ci_pbeam_default"])
    _check_params(m, g)
    np.testing.assert_array_equal(m.eval_dense(g["feat"]), g["dense"])
    _check_golden(m, g)
    m.free()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [
    dict(n_sen=600, n_ci_sen=30, n_density=8, dim=39, seed=1),
    dict(n_sen=300, n_ci_sen=21, n_density=5, dim=13, seed=2),     # odd component count (tail of mgau_eval_all)
    dict(n_sen=200, n_ci_sen=12, n_density=32, dim=39, seed=3),
    dict(n_sen=150, n_ci_sen=9, n_density=40, dim=25, seed=4),     # > 32 components: two rounds per lane
    dict(n_sen=97, n_ci_sen=7, n_density=1, dim=39, seed=5),
])
def test_s3_gpu_matches_oracle(shape):
    mean, var, mixw, cd2ci, n_ci = synth.s3_model(**shape)
    p = orc.PortS3(mean, var, mixw, cd2ci, n_ci)
    m = b.S3Mgau.from_arrays(mean, var, mixw, cd2ci, n_ci)
    T = 75
    feat = synth.s3_features(mean, var, T, seed=shape["seed"] + 100)
    act = synth.s3_active(mean.shape[0], n_ci, T, seed=shape["seed"] + 200)
    # dense mgau_eval over every senone
    want = np.zeros((T, mean.shape[0]), np.int32)
    for t in range(T):
        for s in range(mean.shape[0]):
            want[t, s] = orc.port.orc_s3_mgau_eval(p.h, s, None, orc._p(feat[t], orc.C.c_float), t, 1)
    np.testing.assert_array_equal(m.eval_dense(feat), want)
    # pick beams from the spread of the CI scores of this model
    ci = want[:, :n_ci]
    spread = float(np.median(ci.max(1) - np.median(ci, 1)))
    beam = float(np.float32(1.0003)) ** (-spread)
    for cfg in (dict(), dict(ci_pbeam=beam), dict(ci_pbeam=beam, max_cd=max(4, mean.shape[0] // 12)),
                dict(ci_pbeam=beam, ds_ratio=3), dict(ci_pbeam=beam * 1e-3, max_cd=mean.shape[0] // 8, ds_ratio=2, tighten=0.4)):
        for active in (act, None):
            p.set_fast(**cfg); m.set_fast(**cfg); p.utt_reset(); m.utt_reset()
            a, c = p.eval_utt(feat, active, 2), m.eval_utt(feat, active, 2)
            np.testing.assert_array_equal(c[1], a[1], err_msg=str(cfg))
            np.testing.assert_array_equal(c[0], a[0], err_msg=str(cfg))
            if active is not None:
                np.testing.assert_array_equal(c[2], a[2])
            np.testing.assert_array_equal(np.stack(m.state()), np.stack(p.state()))
    p.free(); m.free()


@pytest.mark.gpu
def test_s3_gpu_frame_by_frame_equals_batched():
    """The per-frame drop-in (gmm_compute_lv1+lv2) carries the same state as
    the batched call, also across a chunk boundary."""
    mean, var, mixw, cd2ci, n_ci = synth.s3_model(n_sen=200, n_ci_sen=12, seed=9)
    m = b.S3Mgau.from_arrays(mean, var, mixw, cd2ci, n_ci)
    p = orc.PortS3(mean, var, mixw, cd2ci, n_ci)
    T = 30
    feat = synth.s3_features(mean, var, T, seed=10)
    act = synth.s3_active(mean.shape[0], n_ci, T, seed=11)
    cfg = dict(ci_pbeam=1e-35, ds_ratio=2)
    p.set_fast(**cfg); m.set_fast(**cfg)
    want, wbest, wact = p.eval_utt(feat, act)
    senscr = np.zeros(mean.shape[0], np.int32)
    for t in range(T):
        a = act[t].copy()
        best = m.frame_eval(feat[t], t, a, senscr)
        assert best == wbest[t]
        np.testing.assert_array_equal(senscr, want[t]); np.testing.assert_array_equal(a, wact[t])
    # two halves with the handle's own score buffer == one call
    m.utt_reset()
    o1 = m.eval_utt(feat[:17], act[:17], 0)
    o2 = m.eval_utt(feat[17:], act[17:], 17, senscr0=m.last_row)
    np.testing.assert_array_equal(np.concatenate([o1[0], o2[0]]), want)
    p.free(); m.free()


@pytest.mark.gpu
def test_s3_gpu_real_model():
    """hub4_cd_continuous_8gau_1s_c_d_dd through the file loader; oracle as checker."""
    d = os.path.join(orc.DATA_DIR, "hmm", "cont")
    mf, vf, wf = (os.path.join(d, n) for n in ("means", "variances", "mixture_weights"))
    if not os.path.exists(mf):
        pytest.skip("continuous model not bundled")
    mean, var, mixw = b.read_s3_cont_arrays(mf, vf, wf)
    S = mean.shape[0]
    mm = b.mdef_maps(os.path.join(d, "mdef"))     # the model's own CD -> CI senone map (text mdef 0.3)
    n_ci, cd2ci = mm["n_ci_sen"], mm["cd2cisen"].astype(np.int32)
    assert n_ci == 144 and mm["n_sen"] == S      # 48 CI phones x 3 states
    rng = np.random.default_rng(4)
    m = b.S3Mgau.from_files(mf, vf, wf, cd2ci, n_ci)
    p = orc.PortS3(mean, var, mixw, cd2ci, n_ci)
    T = 40
    idx = rng.integers(0, S, T)
    feat = (mean[idx, 0] + rng.standard_normal((T, mean.shape[2])) * np.sqrt(var[idx, 0])).astype(np.float32)
    act = synth.s3_active(S, n_ci, T, seed=6)
    for cfg in (dict(), dict(ci_pbeam=1e-40), dict(ci_pbeam=1e-40, max_cd=400)):
        p.set_fast(**cfg); m.set_fast(**cfg); p.utt_reset(); m.utt_reset()
        a, c = p.eval_utt(feat, act), m.eval_utt(feat, act)
        np.testing.assert_array_equal(c[1], a[1]); np.testing.assert_array_equal(c[0], a[0])
    p.free(); m.free()


@pytest.mark.gpu
def test_s3_gpu_edge_sizes():
    """Empty and single-frame utterances, a chunk boundary (2048 frames), all senones inactive."""
    mean, var, mixw, cd2ci, n_ci = synth.s3_model(n_sen=90, n_ci_sen=6, n_density=8, dim=13, seed=12)
    p = orc.PortS3(mean, var, mixw, cd2ci, n_ci)
    m = b.S3Mgau.from_arrays(mean, var, mixw, cd2ci, n_ci)
    for mdl in (p, m):
        mdl.set_fast(ci_pbeam=1e-30, ds_ratio=2)
    out = m.eval_utt(np.zeros((0, 13), np.float32), np.zeros((0, 90), np.uint8))
    assert out[0].shape == (0, 90) and out[1].shape == (0,)
    assert m.eval_dense(np.zeros((0, 13), np.float32)).shape == (0, 90)
    feat = synth.s3_features(mean, var, 2100, seed=13)           # crosses the internal 2048-frame chunk
    act = synth.s3_active(90, n_ci, 2100, seed=14)
    act[5] = 0; act[2047] = 0; act[2048] = 0                     # frames with no active CD senone
    for T in (1, 2100):
        p.utt_reset(); m.utt_reset()
        a, c = p.eval_utt(feat[:T], act[:T]), m.eval_utt(feat[:T], act[:T])
        np.testing.assert_array_equal(c[0], a[0]); np.testing.assert_array_equal(c[1], a[1])
        np.testing.assert_array_equal(np.stack(m.state()), np.stack(p.state()))
    p.free(); m.free()


@pytest.mark.gpu
def test_s3_gpu_fuzz_against_oracle():
    """25 random model shapes / beams / down-sampling ratios / active-set densities."""
    for seed in range(25):
        rng = np.random.default_rng(1000 + seed)
        n_sen = int(rng.integers(40, 260)); n_ci = int(rng.integers(3, 24))
        M = int(rng.choice([1, 2, 3, 5, 8, 11, 16, 33])); D = int(rng.choice([5, 13, 39]))
        mean, var, mixw, cd2ci, n_ci = synth.s3_model(n_sen, n_ci, M, D, seed)
        p = orc.PortS3(mean, var, mixw, cd2ci, n_ci)
        m = b.S3Mgau.from_arrays(mean, var, mixw, cd2ci, n_ci)
        T = int(rng.integers(3, 70))
        feat = synth.s3_features(mean, var, T, seed + 1)
        act = synth.s3_active(n_sen, n_ci, T, seed + 2, p_on=float(rng.uniform(0.02, 0.5)))
        dense = m.eval_dense(feat)
        ci = dense[:, :n_ci]
        spread = float(np.median(ci.max(1) - np.median(ci, 1))) + 1
        cfg = dict(ci_pbeam=float(np.float32(1.0003)) ** (-spread * float(rng.uniform(0.3, 2))),
                   max_cd=int(rng.integers(1, n_sen)), ds_ratio=int(rng.integers(1, 5)), tighten=float(rng.uniform(0.1, 1.0)))
        f0 = int(rng.integers(0, 7))
        p.set_fast(**cfg); m.set_fast(**cfg); p.utt_reset(); m.utt_reset()
        a, c = p.eval_utt(feat, act, f0), m.eval_utt(feat, act, f0)
        np.testing.assert_array_equal(c[1], a[1], err_msg=f"seed {seed} {cfg}")
        np.testing.assert_array_equal(c[0], a[0], err_msg=f"seed {seed} {cfg}")
        np.testing.assert_array_equal(np.stack(m.state()), np.stack(p.state()), err_msg=f"seed {seed}")
        p.free(); m.free()


# ---------------------------------------------------------------- sub-vector quantised shortlists (S3/libam/subvq.c)
SVQ_GOLDEN_CFGS = [(3, 3, 1e-3, dict()), (3, 2, 1e-2, dict(ci_pbeam=1e-40, max_cd=60)), (1, 3, 1e-3, dict(ci_pbeam=1e-30, ds_ratio=3))]


def _svq_golden_model(tmp_path, i):
    mean, var, mixw, cd2ci, n_ci = synth.s3_model(n_sen=160, n_ci_sen=16)
    n_sv, vqeval, beam, cfg = SVQ_GOLDEN_CFGS[i]
    q = orc.synthetic_subvq(mean, var, ~np.all(var == 0, axis=2), n_sv, 16)
    path = str(tmp_path / ("m%d.subvq" % i))
    orc.write_subvq(path, q)
    return mean, var, mixw, cd2ci, n_ci, path, vqeval, beam, cfg


@pytest.mark.parametrize("i", range(len(SVQ_GOLDEN_CFGS)))
def test_s3_subvq_port_matches_golden(tmp_path, i):
    """The port's sub-VQ layer against outputs of the reference itself (tests/golden/s3_svq.npz,
    make_golden.py:s3_svq_case) -- holds where /root/reference is absent."""
    g = cases.load("s3_svq.npz")
    mean, var, mixw, cd2ci, n_ci, path, vqeval, beam, cfg = _svq_golden_model(tmp_path, i)
    p = orc.PortS3(mean, var, mixw, cd2ci, n_ci)
    orc.port_set_svq(p, orc.read_subvq(path), vqeval=vqeval, subvqbeam=beam)
    p.set_fast(**cfg); p.utt_reset()
    vq = np.zeros_like(g[f"vq{i}"])
    for t in range(g["feat"].shape[0]):
        row = np.ascontiguousarray(g["feat"][t])
        orc.port.orc_s3_svq_eval(p.h, orc._p(row, orc.C.c_float))
        orc.port.orc_s3_svq_dist(p.h, orc._p(vq[t], orc.C.c_int32))
    np.testing.assert_array_equal(vq, g[f"vq{i}"])
    o, best, a = p.eval_utt(g["feat"], g["act"], int(g["frame0"]))
    np.testing.assert_array_equal(best, g[f"best{i}"]); np.testing.assert_array_equal(o, g[f"scr{i}"])
    np.testing.assert_array_equal(a, g[f"act{i}"])
    bi, ut = p.state()
    np.testing.assert_array_equal(bi, g[f"bstidx{i}"]); np.testing.assert_array_equal(ut, g[f"upd{i}"])
    p.free()


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(SVQ_GOLDEN_CFGS)))
def test_s3_subvq_gpu_matches_golden(tmp_path, i):
    g = cases.load("s3_svq.npz")
    mean, var, mixw, cd2ci, n_ci, path, vqeval, beam, cfg = _svq_golden_model(tmp_path, i)
    m = b.S3Mgau.from_arrays(mean, var, mixw, cd2ci, n_ci)
    m.set_subvq(path, vqeval=vqeval, subvqbeam=beam)
    m.set_fast(**cfg); m.utt_reset()
    o, best, a = m.eval_utt(g["feat"], g["act"], int(g["frame0"]))
    np.testing.assert_array_equal(best, g[f"best{i}"]); np.testing.assert_array_equal(o, g[f"scr{i}"])
    np.testing.assert_array_equal(a, g[f"act{i}"])
    bi, ut = m.state()
    np.testing.assert_array_equal(bi, g[f"bstidx{i}"]); np.testing.assert_array_equal(ut, g[f"upd{i}"])
    m.free()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,n_sv,vqsize,vqeval,max_sv,beam", [
    (dict(n_sen=600, n_ci_sen=30, n_density=8, dim=39, seed=1), 3, 32, 3, -1, 1e-3),
    (dict(n_sen=600, n_ci_sen=30, n_density=8, dim=39, seed=1), 3, 32, 2, -1, 1e-1),
    (dict(n_sen=300, n_ci_sen=21, n_density=5, dim=13, seed=2), 3, 16, 1, -1, 1e-2),
    (dict(n_sen=200, n_ci_sen=12, n_density=32, dim=39, seed=3), 4, 64, 3, -1, 1e-4),     # generic #sub-vectors
    (dict(n_sen=150, n_ci_sen=9, n_density=40, dim=25, seed=4), 2, 24, 3, -1, 0.3),        # > 32 components
    (dict(n_sen=200, n_ci_sen=12, n_density=8, dim=39, seed=6), 3, 16, 3, 2, 1e-3),        # -svmax 2 of 3
    (dict(n_sen=97, n_ci_sen=7, n_density=1, dim=39, seed=5), 1, 8, 3, -1, 1e-3),
])
def test_s3_subvq_gpu_matches_oracle(tmp_path, shape, n_sv, vqsize, vqeval, max_sv, beam):
    mean, var, mixw, cd2ci, n_ci = synth.s3_model(**shape)
    q = orc.synthetic_subvq(mean, var, ~np.all(var == 0, axis=2), n_sv, vqsize, seed=shape["seed"])
    path = str(tmp_path / "m.subvq")
    orc.write_subvq(path, q)
    p = orc.PortS3(mean, var, mixw, cd2ci, n_ci)
    orc.port_set_svq(p, orc.read_subvq(path), max_sv=max_sv, vqeval=vqeval, subvqbeam=beam)
    m = b.S3Mgau.from_arrays(mean, var, mixw, cd2ci, n_ci)
    m.set_subvq(path, max_sv=max_sv, vqeval=vqeval, subvqbeam=beam)
    T = 70
    feat = synth.s3_features(mean, var, T, seed=shape["seed"] + 100)
    act = synth.s3_active(mean.shape[0], n_ci, T, seed=shape["seed"] + 200)
    for cfg in (dict(), dict(ci_pbeam=1e-40), dict(ci_pbeam=1e-40, max_cd=max(4, mean.shape[0] // 12)),
                dict(ci_pbeam=1e-30, ds_ratio=3)):
        for active in (act, None):
            p.set_fast(**cfg); m.set_fast(**cfg); p.utt_reset(); m.utt_reset()
            a, c = p.eval_utt(feat, active, 2), m.eval_utt(feat, active, 2)
            np.testing.assert_array_equal(c[1], a[1], err_msg=str(cfg))
            np.testing.assert_array_equal(c[0], a[0], err_msg=str(cfg))
            np.testing.assert_array_equal(np.stack(m.state()), np.stack(p.state()))
    # the dense call (mgau_eval on every senone) ignores the layer, and removing it restores the plain scores
    plain = orc.PortS3(mean, var, mixw, cd2ci, n_ci)
    m.set_subvq(None); m.set_fast(); m.utt_reset(); plain.set_fast()
    np.testing.assert_array_equal(m.eval_utt(feat, act, 0)[0], plain.eval_utt(feat, act, 0)[0])
    p.free(); m.free(); plain.free()


@pytest.mark.gpu
def test_s3_subvq_gpu_real_model_and_per_frame_binding():
    """hub4_cd_continuous_8gau_1s_c_d_dd with the tree's own test.subvq; also through the per-frame entry point."""
    d = os.path.join(orc.DATA_DIR, "hmm", "cont")
    mf, vf, wf, sv = (os.path.join(d, n) for n in ("means", "variances", "mixture_weights", "test.subvq"))
    if not os.path.exists(sv):
        pytest.skip("test.subvq not bundled")
    mean, var, mixw = b.read_s3_cont_arrays(mf, vf, wf)
    mm = b.mdef_maps(os.path.join(d, "mdef"))
    n_ci, cd2ci = mm["n_ci_sen"], mm["cd2cisen"].astype(np.int32)
    m = b.S3Mgau.from_files(mf, vf, wf, cd2ci, n_ci)
    p = orc.PortS3(mean, var, mixw, cd2ci, n_ci)
    m.set_subvq(sv, subvqbeam=3e-3); orc.port_set_svq(p, orc.read_subvq(sv), subvqbeam=3e-3)
    rng = np.random.default_rng(4)
    T = 24
    idx = rng.integers(0, mean.shape[0], T)
    feat = (mean[idx, 0] + rng.standard_normal((T, mean.shape[2])) * np.sqrt(var[idx, 0])).astype(np.float32)
    act = synth.s3_active(mean.shape[0], n_ci, T, seed=6)
    for cfg in (dict(), dict(ci_pbeam=1e-40, max_cd=400)):
        p.set_fast(**cfg); m.set_fast(**cfg); p.utt_reset(); m.utt_reset()
        a, c = p.eval_utt(feat, act), m.eval_utt(feat, act)
        np.testing.assert_array_equal(c[1], a[1]); np.testing.assert_array_equal(c[0], a[0])
    p.utt_reset(); m.utt_reset()
    want, wbest, wact = p.eval_utt(feat, act)
    senscr = np.zeros(mean.shape[0], np.int32)
    for t in range(T):
        a = act[t].copy()
        assert m.frame_eval(feat[t], t, a, senscr) == wbest[t]
        np.testing.assert_array_equal(senscr, want[t])
    with pytest.raises(b.B200Error, match="Model size conflict"):
        small = b.S3Mgau.from_arrays(*synth.s3_model(n_sen=97, n_ci_sen=7, n_density=1, dim=39, seed=5))
        small.set_subvq(sv)
    p.free(); m.free()


# ---------------------------------------------------------------- Gaussian selector (S3/libam/gs.c)
def _gs_golden_model(tmp_path):
    mean, var, mixw, cd2ci, n_ci = synth.s3_model(n_sen=160, n_ci_sen=16)
    cw, bits = orc.synthetic_gs(mean, 32)
    path = str(tmp_path / "m.gs")
    orc.write_gs(path, cw, bits, mean.shape[1])
    return mean, var, mixw, cd2ci, n_ci, cw, bits, path


def _check_gs_golden(model, g):
    model.set_fast(ci_pbeam=1e-40, max_cd=60); model.utt_reset()
    o, best, a = model.eval_utt(g["feat"], g["act"], int(g["frame0"]))
    np.testing.assert_array_equal(best, g["gs_best"]); np.testing.assert_array_equal(o, g["gs_scr"])
    np.testing.assert_array_equal(a, g["gs_act"])
    bi, ut = model.state()
    np.testing.assert_array_equal(bi, g["gs_bstidx"]); np.testing.assert_array_equal(ut, g["gs_upd"])


def test_s3_gs_port_matches_golden(tmp_path):
    g = cases.load("s3_svq.npz")
    mean, var, mixw, cd2ci, n_ci, cw, bits, path = _gs_golden_model(tmp_path)
    p = orc.PortS3(mean, var, mixw, cd2ci, n_ci)
    orc.port_set_gs(p, cw, bits)
    got = [orc.port.orc_s3_gs_closest(p.h, orc._p(np.ascontiguousarray(g["feat"][t]), orc.C.c_float)) for t in range(g["feat"].shape[0])]
    np.testing.assert_array_equal(got, g["gs_closest"])
    _check_gs_golden(p, g)
    p.free()


@pytest.mark.gpu
def test_s3_gs_gpu_matches_golden(tmp_path):
    g = cases.load("s3_svq.npz")
    mean, var, mixw, cd2ci, n_ci, cw, bits, path = _gs_golden_model(tmp_path)
    m = b.S3Mgau.from_arrays(mean, var, mixw, cd2ci, n_ci)
    m.set_gs(path)
    _check_gs_golden(m, g)
    m.free()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,n_code", [
    (dict(n_sen=600, n_ci_sen=30, n_density=8, dim=39, seed=1), 64),
    (dict(n_sen=300, n_ci_sen=21, n_density=5, dim=13, seed=2), 16),
    (dict(n_sen=200, n_ci_sen=12, n_density=32, dim=39, seed=3), 256),      # 32 densities: every map bit in use
    (dict(n_sen=97, n_ci_sen=7, n_density=1, dim=39, seed=5), 8),
])
def test_s3_gs_gpu_matches_oracle(tmp_path, shape, n_code):
    mean, var, mixw, cd2ci, n_ci = synth.s3_model(**shape)
    cw, bits = orc.synthetic_gs(mean, n_code, seed=shape["seed"])
    path = str(tmp_path / "m.gs")
    orc.write_gs(path, cw, bits, mean.shape[1])
    p = orc.PortS3(mean, var, mixw, cd2ci, n_ci); orc.port_set_gs(p, cw, bits)
    m = b.S3Mgau.from_arrays(mean, var, mixw, cd2ci, n_ci); m.set_gs(path)
    T = 70
    feat = synth.s3_features(mean, var, T, seed=shape["seed"] + 100)
    act = synth.s3_active(mean.shape[0], n_ci, T, seed=shape["seed"] + 200)
    for cfg in (dict(), dict(ci_pbeam=1e-40, max_cd=max(4, mean.shape[0] // 12)), dict(ci_pbeam=1e-30, ds_ratio=3)):
        for active in (act, None):
            p.set_fast(**cfg); m.set_fast(**cfg); p.utt_reset(); m.utt_reset()
            a, c = p.eval_utt(feat, active, 2), m.eval_utt(feat, active, 2)
            np.testing.assert_array_equal(c[1], a[1], err_msg=str(cfg))
            np.testing.assert_array_equal(c[0], a[0], err_msg=str(cfg))
            np.testing.assert_array_equal(np.stack(m.state()), np.stack(p.state()))
    # with both layers the selector wins (approx_mgau_eval: gs4gs first)
    q = orc.synthetic_subvq(mean, var, ~np.all(var == 0, axis=2), 1, 8, seed=3)
    sv = str(tmp_path / "m.subvq"); orc.write_subvq(sv, q)
    orc.port_set_svq(p, orc.read_subvq(sv)); m.set_subvq(sv)
    p.set_fast(); m.set_fast(); p.utt_reset(); m.utt_reset()
    np.testing.assert_array_equal(m.eval_utt(feat, act, 0)[0], p.eval_utt(feat, act, 0)[0])
    with pytest.raises(b.B200Error, match="odd codeword count"):
        orc.write_gs(path, cw[:7], bits[:, :7], mean.shape[1]); m.set_gs(path)
    p.free(); m.free()
