"""GPU parity tests for the GMM back-ends (through the C ABI) against the
oracle and the reference-generated goldens.  Bar: bit-exact for the exact
CUDA-core path; within +-1 (log-add-table quantisation, north_star) for the
tensor-core path, with the mismatch rate asserted small and printed."""
import os

import numpy as np
import pytest

import cases
import orc
import cmusphinx_b200 as b
from cmusphinx_b200 import s3io, synth

pytestmark = pytest.mark.gpu

MS_CASES = ["ms_small.npz", "ms_3stream.npz", "ms_allden.npz", "ms_cont32.npz"]


def _check_active(m, g, feat):
    n = g["act_scores"].shape[0]
    # (a) drop-in per-frame call
    for i in range(n):
        d = cases.deltas_of(g, i)
        ids = cases.active_ids(d)
        out = np.full(m.n_sen, 12345, np.int16)
        streams, off = [], 0
        for L in (m.cfg.featlen if m.cfg else [m.featdim]):
            streams.append(feat[i, off:off + L].copy())
            off += L
        m.frame_eval(out, d, d.size, streams, i, False)
        np.testing.assert_array_equal(out[ids], g["act_scores"][i][ids])
    # (b) utterance-batched serving
    m.utt_begin(feat[:n])
    for i in range(n):
        d = cases.deltas_of(g, i)
        ids = cases.active_ids(d)
        out = np.full(m.n_sen, 12345, np.int16)
        m.utt_frame(out, d, d.size, i, False)
        np.testing.assert_array_equal(out[ids], g["act_scores"][i][ids])
    out = np.zeros(m.n_sen, np.int16)
    m.utt_frame(out, None, 0, 0, True)
    np.testing.assert_array_equal(out, g["dense"][0])
    with pytest.raises(b.B200Error):
        m.utt_frame(out, None, 0, n + 3, True)


@pytest.mark.parametrize("name", MS_CASES)
def test_ms_exact_path_bit_exact_vs_golden(name):
    g = cases.load(name)
    m = cases.ms_product(g)
    assert m.name == "b200_ms"
    m.set_path(0)
    got = m.score(g["feat"])
    np.testing.assert_array_equal(got, g["dense"])
    _check_active(m, g, g["feat"])
    # ms non-compallsen leaves inactive entries of the caller's array untouched
    d = cases.deltas_of(g, 0)
    out = np.full(m.n_sen, 777, np.int16)
    m.frame_eval(out, d, d.size, [g["feat"][0][o:o + L] for o, L in zip(np.cumsum([0] + list(m.cfg.featlen[:-1])), m.cfg.featlen)], 0, False)
    inactive = np.setdiff1d(np.arange(m.n_sen), cases.active_ids(d))
    assert (out[inactive] == 777).all()
    m.free()


def test_ms_empty_and_ragged_batches():
    g = cases.load("ms_small.npz")
    m = cases.ms_product(g)
    m.set_path(0)
    assert m.score(np.zeros((0, m.featdim), np.float32)).shape == (0, m.n_sen)
    for T in (1, 2, 31, 33):
        np.testing.assert_array_equal(m.score(g["feat"][:T]), g["dense"][:T])
    m.free()


def test_ms_load_from_s3_files_matches_oracle(tmp_path):
    n_sen, n_density, dim = 70, 16, 39
    mean, var, mixw = synth.cont_model(n_sen, n_density, dim, 21)
    var[3, 2, :5] = 1e-7
    s3io.write_gauden(str(tmp_path / "means"), mean, [dim])
    s3io.write_gauden(str(tmp_path / "variances"), var, [dim])
    s3io.write_mixw(str(tmp_path / "mixture_weights"), mixw)
    m = b.ms_from_files(str(tmp_path / "means"), str(tmp_path / "variances"), str(tmp_path / "mixture_weights"),
                        ".cont.", topn=4, logbase=orc.LOGBASE)
    m.set_path(0)
    pv, pd = orc.port_precompute(var.reshape(-1, dim), dim, 1e-4, orc.LOGBASE)
    pm = orc.PortMs(n_sen, 1, [dim], n_density, n_sen, 4, 1, mean, pv, pd, orc.port_mixw_quantize(mixw, 1e-7, orc.LOGBASE),
                    np.arange(n_sen), orc.LOGBASE)
    feat = synth.cont_features(mean, var, 200, 22)
    np.testing.assert_array_equal(m.score(feat), pm.eval_all(feat))
    with pytest.raises(b.B200Error):
        b.ms_from_files(str(tmp_path / "means"), str(tmp_path / "nope"), str(tmp_path / "mixture_weights"))
    m.free()


def test_ms_shared_codebooks_semi_and_ptm_maps():
    """The generic ms back-end on shared codebooks (-senmgau .semi. / .ptm.): the
    two-kernel list path."""
    rng = np.random.default_rng(4)
    n_mgau, n_density, n_sen = 5, 24, 90
    vl = [5, 4]
    mean = rng.standard_normal((n_mgau, n_density * sum(vl))).astype(np.float32)
    var = np.exp(rng.uniform(-2, 1, mean.shape)).astype(np.float32)
    pv = np.zeros_like(var)
    pd = np.zeros((n_mgau, 2, n_density), np.float32)
    for mg in range(n_mgau):
        off = 0
        for f, L in enumerate(vl):
            blk = var[mg, n_density * off:n_density * (off + L)].reshape(n_density, L)
            a, d = orc.port_precompute(blk, L, 1e-4, orc.LOGBASE)
            pv[mg, n_density * off:n_density * (off + L)] = a.reshape(-1)
            pd[mg, f] = d
            off += L
    mixw = orc.port_mixw_quantize(rng.dirichlet(np.ones(n_density), (n_sen, 2)).astype(np.float32), 1e-7, orc.LOGBASE)
    s2m = rng.integers(0, n_mgau, n_sen).astype(np.uint32)
    feat = rng.standard_normal((77, sum(vl))).astype(np.float32)
    pm = orc.PortMs(n_mgau, 2, vl, n_density, n_sen, 3, 2, mean, pv, pd, mixw, s2m, orc.LOGBASE)
    cfg = b.MgauConfig(n_mgau, 2, n_density, n_sen, vl, topn=3, aw=2, logbase=orc.LOGBASE)
    m = b.ms_from_arrays(cfg, mean, pv, pd, mixw, s2m)
    assert m.path == 0
    np.testing.assert_array_equal(m.score(feat), pm.eval_all(feat))
    m.free()


@pytest.mark.parametrize("name", ["cont_hub4_topn4.npz", "cont_hub4_topn8.npz"])
def test_real_continuous_model_exact(name):
    if not cases.have_model(name):
        pytest.skip("model files (oracle/_ref/data) not present")
    g = cases.load(name)
    d = cases.model_dir(name)
    m = b.ms_from_files(d + "/means", d + "/variances", d + "/mixture_weights", ".cont.", topn=int(g["topn"]),
                        logbase=orc.LOGBASE)
    if int(g["topn"]) == 4:
        # the default (tcgen05) path on real speech frames against the reference's own scores
        assert m.path == 1
        np.testing.assert_array_equal(m.score(g["feat"]), g["dense"])
        assert m.cont_stats()["max_gemm_err"] <= 16
    m.set_path(0)
    np.testing.assert_array_equal(m.score(g["feat"]), g["dense"])
    for i in range(g["act_scores"].shape[0]):
        dl = cases.deltas_of(g, i)
        ids = cases.active_ids(dl)
        out = np.zeros(m.n_sen, np.int16)
        m.frame_eval(out, dl, dl.size, [g["feat"][i]], i, False)
        np.testing.assert_array_equal(out[ids], g["act_scores"][i][ids])
    m.free()


def _tied_product(name, g, kind):
    gm, gv, sd, n_sen = cases.tied_arrays(name, g)
    m = b.tied_from_model_dir(cases.model_dir(name), n_sen, sen2cb=g["sen2cb"] if kind == 1 else None, topn=4,
                              logbase=orc.LOGBASE)
    return m


@pytest.mark.parametrize("name,kind", [("semi_hub4wsj.npz", 2), ("ptm_hub4wsj.npz", 1)])
def test_tied_backends_vs_reference_golden(name, kind):
    """The device codebook stage is seed-free (every frame is evaluated with a
    fresh top-N list), the reference seeds each frame with the previous frame's
    list; results can differ only on exact integer-score ties at rank N
    (SURVEY.md section 7).  Frame 0 must be identical; the rest is required to
    match on >= 99.99 % of scores and within 2 units."""
    if not cases.have_model(name):
        pytest.skip("model files (oracle/_ref/data) not present")
    g = cases.load(name)
    m = _tied_product(name, g, kind)
    assert m.name == ("b200_ptm" if kind == 1 else "b200_semi")
    got = m.score(g["feat"])
    want = g["dense"]
    np.testing.assert_array_equal(got[0], want[0])
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    frac = float((diff != 0).mean())
    print(f"{name}: dense mismatch fraction {frac:.2e}, max |d| {diff.max()}")
    assert frac <= 1e-4 and diff.max() <= 2
    # active-list calls (codebook pruning + per-stream norm over active codebooks)
    bad = 0
    for i in range(g["act_scores"].shape[0]):
        dl = cases.deltas_of(g, i)
        streams, off = [], 0
        for L in g["streamlen"]:
            streams.append(g["feat"][i, off:off + int(L)].copy())
            off += int(L)
        out = np.full(m.n_sen, 12345, np.int16)
        m.frame_eval(out, dl, dl.size, streams, i, False)
        if i == 0:
            np.testing.assert_array_equal(out, g["act_scores"][0])
        bad += int((out != g["act_scores"][i]).sum())
    assert bad <= 2 * m.n_sen * 1e-3
    m.free()


def test_semi_topn_beam_vs_reference_golden_and_port():
    """s2_semi -topn_beam on the device (exact scan and tensor-core codebook stage): frame 0
    identical to the reference, the rest within the seed-free tolerance of the test above;
    and identical to the port on single frames (fresh lists on both sides)."""
    name = "semi_hub4wsj.npz"
    if not cases.have_model(name):
        pytest.skip("model files (oracle/_ref/data) not present")
    g, gb = cases.load(name), cases.load("semi_hub4wsj_beam.npz")
    gm, gv, sd, n_sen = cases.tied_arrays(name, g)
    pv, pd = orc.port_precompute(gv["data"].reshape(-1, 13), 13, 1e-4, orc.LOGBASE)
    pt = orc.PortTied(2, 1, 3, gm["veclen"], gm["n_density"], n_sen, 4, gm["data"], pv, pd, sd["mixw"], sd["n_clust"],
                      sd["mixw_cb"], None, orc.LOGBASE)
    for i, beam in enumerate(gb["beams"]):
        beam = [int(v) for v in beam]
        m = b.tied_from_model_dir(cases.model_dir(name), n_sen, topn=4, logbase=orc.LOGBASE, topn_beam=beam)
        want = gb[f"dense{i}"]
        for path in (0, 1):
            m.set_path(path)
            got = m.score(g["feat"])
            np.testing.assert_array_equal(got[0], want[0])
            diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
            assert float((diff != 0).mean()) <= 1e-4 and diff.max() <= 2, (beam, path)
        pt.set_topn_beam(beam)
        for t in (3, 11, 17):
            pt.reset()
            np.testing.assert_array_equal(m.score(g["feat"][t:t + 1])[0], pt.frame_eval(g["feat"][t], None, True, 0))
        m.free()
    # -ds: dense batch, a batch that crosses the internal list chunking, piecewise utterance
    # serving (blocks that start on a skipped frame continue from the block before) and
    # frame-by-frame calls all give the reference's scores
    for i, (ds, beam) in enumerate(gb["ds_cfg"]):
        ds, beam = int(ds), int(beam)
        want = gb[f"ds_dense{i}"]
        m = b.tied_from_model_dir(cases.model_dir(name), n_sen, topn=4, logbase=orc.LOGBASE, topn_beam=[beam] * 3,
                                  ds_ratio=ds)
        for path in (0, 1):
            m.set_path(path)
            got = m.score(g["feat"])
            np.testing.assert_array_equal(got[0], want[0])
            diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
            assert float((diff != 0).mean()) <= 1e-4 and diff.max() <= 2, (ds, beam, path)
        dense = m.score(g["feat"])
        T = g["feat"].shape[0]
        out = np.zeros(n_sen, np.int16)
        for t0, t1 in ((0, 5), (5, 6), (6, 13), (13, T)):
            m.utt_begin(g["feat"][t0:t1], frame0=t0)
            for t in range(t0, t1):
                m.utt_frame(out, None, 0, t - t0, True)
                np.testing.assert_array_equal(out, dense[t])
        for t in range(T):   # ps_mgaufuncs_t.frame_eval, one frame per call
            streams = [g["feat"][t, 13 * k:13 * k + 13].copy() for k in range(3)]
            m.frame_eval(out, None, 0, streams, t, True)
            np.testing.assert_array_equal(out, dense[t])
        with pytest.raises(b.B200Error):   # a skipped frame whose predecessor was never scored
            m.utt_begin(g["feat"][3:5], frame0=ds * 5 + 1)
        m.free()
        # a long batch (list chunking inside the library) == the port on the same frames
        pt.reset(); pt.set_topn_beam([beam] * 3); pt.set_ds(ds)
        long_feat = np.tile(g["feat"], (9, 1))
        m = b.tied_from_model_dir(cases.model_dir(name), n_sen, topn=4, logbase=orc.LOGBASE, topn_beam=[beam] * 3,
                                  ds_ratio=ds)
        got, want = m.score(long_feat), pt.eval_all(long_feat)
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        assert float((diff != 0).mean()) <= 1e-4 and diff.max() <= 2
        m.free()
    pt.set_ds(1)
    # ms ignores -ds, as ms_mgau.c does (it never reads the option)
    # ptm never reads -topn_beam: the setting must not change its scores
    name = "ptm_hub4wsj.npz"
    if cases.have_model(name):
        g = cases.load(name)
        n_sen = int(g["n_sen"])
        m0 = b.tied_from_model_dir(cases.model_dir(name), n_sen, sen2cb=g["sen2cb"], topn=4, logbase=orc.LOGBASE)
        m1 = b.tied_from_model_dir(cases.model_dir(name), n_sen, sen2cb=g["sen2cb"], topn=4, logbase=orc.LOGBASE,
                                   topn_beam=[5, 5, 5])
        np.testing.assert_array_equal(m0.score(g["feat"][:6]), m1.score(g["feat"][:6]))
        m0.free(); m1.free()


def test_config2_shape_subset_exact():
    """BASELINE config 2 shape (5000 senones x 32 Gaussians x 39 dims) on a frame
    subset the oracle finishes in seconds."""
    n_sen, M, D, T = 5000, 32, 39, 48
    mean, var, mixw = synth.cont_model(n_sen, M, D, 1234)
    pv, pd = orc.port_precompute(var.reshape(-1, D), D, 1e-4, orc.LOGBASE)
    q = orc.port_mixw_quantize(mixw, 1e-7, orc.LOGBASE)
    feat = synth.cont_features(mean, var, T, 5678)
    pm = orc.PortMs(n_sen, 1, [D], M, n_sen, 4, 1, mean, pv, pd, q, np.arange(n_sen), orc.LOGBASE)
    want = pm.eval_all(feat)
    cfg = b.MgauConfig(n_sen, 1, M, n_sen, [D], topn=4, logbase=orc.LOGBASE)
    m = b.ms_from_arrays(cfg, mean, pv, pd, q, np.arange(n_sen))
    m.set_path(0)
    np.testing.assert_array_equal(m.score(feat), want)
    if m.path == 0 and getattr(m, "_tc_checked", None) is None:
        try:
            m.set_path(1)
        except b.B200Error:
            m.free()
            return
        got = m.score(feat)
        diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
        print(f"tensor-core path: mismatch fraction {(diff != 0).mean():.2e}, max |d| {diff.max()}")
        assert diff.max() == 0
    m.free()


def test_full_size_properties():
    """Size-independent properties at a larger batch (no oracle): rows do not
    depend on how frames are batched; every row's best score is 0; scores are
    non-negative."""
    n_sen, M, D = 1000, 32, 39
    mean, var, mixw = synth.cont_model(n_sen, M, D, 77)
    pv, pd = orc.port_precompute(var.reshape(-1, D), D, 1e-4, orc.LOGBASE)
    q = orc.port_mixw_quantize(mixw, 1e-7, orc.LOGBASE)
    cfg = b.MgauConfig(n_sen, 1, M, n_sen, [D], topn=4, logbase=orc.LOGBASE)
    m = b.ms_from_arrays(cfg, mean, pv, pd, q, np.arange(n_sen))
    feat = synth.cont_features(mean, var, 20000, 78)
    for path in (0, 1):
        try:
            m.set_path(path)
        except b.B200Error:
            continue
        full = m.score(feat)
        assert (full.min(axis=1) == 0).all() and (full >= 0).all()
        perm = np.random.default_rng(1).permutation(20000)[:3000]
        np.testing.assert_array_equal(m.score(feat[perm]), full[perm])
        np.testing.assert_array_equal(m.score(feat[9000:9777]), full[9000:9777])
    m.free()


@pytest.mark.parametrize("S,M,D,T", [(64, 32, 39, 300), (250, 8, 39, 515), (100, 16, 13, 129), (37, 32, 20, 77),
                                     (4999, 32, 39, 130)])
def test_tensor_core_path_identical_to_exact(S, M, D, T):
    """tcgen05 path (GEMM + certificate / top-4 network + exact fix-ups, round 2)
    against the exact path, which is itself bit-exact against the oracle:
    IDENTICAL scores.  Shapes cover all template instantiations (M = 8/16/32,
    2/4/5 fp16 and 4/7/10 TF32 k-steps), padded tiles, ragged frame counts and an
    n_sen that forces the generic finish pass."""
    mean, var, mixw = synth.cont_model(S, M, D, 31)
    pv, pd = orc.port_precompute(var.reshape(-1, D), D, 1e-4, orc.LOGBASE)
    q = orc.port_mixw_quantize(mixw, 1e-7, orc.LOGBASE)
    cfg = b.MgauConfig(S, 1, M, S, [D], topn=4, logbase=orc.LOGBASE)
    m = b.ms_from_arrays(cfg, mean, pv, pd, q, np.arange(S))
    assert m.path == 1, "tensor-core path should be the default for this shape"
    feat = synth.cont_features(mean, var, T, 32)
    got = m.score(feat)
    m.set_path(0)
    want = m.score(feat)
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    print(f"S={S} M={M} D={D} T={T}: mismatch {float((diff != 0).mean()):.2e} max {diff.max()}")
    assert diff.max() == 0
    if S <= 250:
        pm = orc.PortMs(S, 1, [D], M, S, 4, 1, mean, pv, pd, q, np.arange(S), orc.LOGBASE)
        np.testing.assert_array_equal(want, pm.eval_all(feat))
    # un-normalised serving path (utt_begin/utt_frame) goes through the same kernels
    m.set_path(1)
    m.utt_begin(feat[:40])
    row = np.zeros(S, np.int16)
    m.utt_frame(row, None, 0, 7, True)
    np.testing.assert_array_equal(row, want[7])
    m.free()


def test_tensor_core_path_unsupported_shapes_fall_back_loudly():
    mean, var, mixw = synth.cont_model(40, 32, 39, 3)
    pv, pd = orc.port_precompute(var.reshape(-1, 39), 39, 1e-4, orc.LOGBASE)
    q = orc.port_mixw_quantize(mixw, 1e-7, orc.LOGBASE)
    cfg = b.MgauConfig(40, 1, 32, 40, [39], topn=3, logbase=orc.LOGBASE)   # topn != 4
    m = b.ms_from_arrays(cfg, mean, pv, pd, q, np.arange(40))
    assert m.path == 0
    with pytest.raises(b.B200Error, match="unavailable"):
        m.set_path(1)
    m.free()


def test_config3_ptm_shape_vs_oracle():
    """BASELINE config 3 shape: 256 codebooks x 4096 densities x 39 dims, 5000
    senones, 8-bit mixw, topn 4 -- a few frames against the (sequential) oracle.
    Every codebook holds the same Gaussians in a different order, so all
    codebooks share one top-1 score and the normalised scores stay inside the
    0..96 range that the reference's fast_logmath_add assumes (tied_mgau_common.h:
    82-85; random codebooks drive the REFERENCE itself out of its table)."""
    rng = np.random.default_rng(9)
    C, Mden, D, S, T = 256, 4096, 39, 5000, 3
    base_m = rng.standard_normal((Mden, D)).astype(np.float32)
    base_v = np.exp(rng.uniform(np.log(0.05), np.log(5.0), (Mden, D))).astype(np.float32)
    mean = np.empty((C, Mden * D), np.float32)
    var = np.empty((C, Mden * D), np.float32)
    for c in range(C):
        perm = rng.permutation(Mden)
        mean[c] = base_m[perm].reshape(-1)
        var[c] = base_v[perm].reshape(-1)
    pv, pd = orc.port_precompute(var.reshape(-1, D), D, 1e-4, orc.LOGBASE)
    pv = pv.reshape(C, -1)
    pd = pd.reshape(C, 1, Mden)
    mixw = rng.integers(0, 160, (1, Mden, S)).astype(np.uint8)
    s2c = (np.arange(S) * C // S).astype(np.uint8)
    feat = (base_m[rng.integers(0, Mden, T)] + rng.standard_normal((T, D)) * 0.3).astype(np.float32)
    pt = orc.PortTied(1, C, 1, [D], Mden, S, 4, mean, pv, pd, mixw, 0, None, s2c, orc.LOGBASE)
    want = []
    for t in range(T):
        pt.reset()
        want.append(pt.frame_eval(feat[t], None, True, 0))
    cfg = b.MgauConfig(C, 1, Mden, S, [D], topn=4, logbase=orc.LOGBASE)
    m = b.ptm_from_arrays(cfg, mean, pv, pd, mixw, s2c)
    got = m.score(feat)
    np.testing.assert_array_equal(got, np.stack(want))
    m.free()


def test_ptm_speech_frames_bit_exact_vs_oracle():
    """200 speech frames of the real 50-codebook PTM model: GPU == oracle port,
    bit for bit, dense and with codebook pruning (the reference itself is not a
    usable yardstick there, see test_reference_ptm_reads_past_its_logadd_table)."""
    name = "ptm_hub4wsj.npz"
    if not cases.have_model(name) or not orc.have_ref():
        pytest.skip("model files / oracle/_ref not present")
    g = cases.load(name)
    hmm = cases.model_dir(name)
    r = orc.RefAcmod(hmm)
    feat = r.cep2feat(orc.read_mfc(os.path.join(orc.DATA_DIR, "test", "wsj", "441c0201.mfc")))[:200]
    r.close()
    gm, gv, sd, n_sen = cases.tied_arrays(name, g)
    pv, pd = orc.port_precompute(gv["data"].reshape(-1, 13), 13, 1e-4, orc.LOGBASE)
    pt = orc.PortTied(1, 50, 3, [13, 13, 13], gm["n_density"], n_sen, 4, gm["data"], pv, pd, sd["mixw"], sd["n_clust"],
                      sd["mixw_cb"], g["sen2cb"], orc.LOGBASE)
    want = pt.eval_all(feat)
    m = _tied_product(name, g, 1)
    np.testing.assert_array_equal(m.score(feat), want)
    pt.reset()
    m.utt_begin(feat)
    rng = np.random.default_rng(8)
    for t in range(120, 160):
        mask = np.zeros((n_sen + 31) // 32, np.uint32)
        cbs = rng.choice(50, 9, replace=False)
        for s in np.nonzero(np.isin(g["sen2cb"], cbs))[0][::2]:
            mask[s // 32] |= np.uint32(1 << (s % 32))
        dl = orc.port_flags2list(mask, n_sen)
        got = np.zeros(n_sen, np.int16)
        m.utt_frame(got, dl, dl.size, t, False)
        np.testing.assert_array_equal(got, pt.frame_eval(feat[t], dl, False, t))
    m.free()


@pytest.mark.parametrize("C,Mden,streams,S,T,kind", [
    (6, 1024, [39], 300, 700, 1),          # ptm, 4 tiles per codebook, 10 k-steps
    (3, 256, [13, 13, 13], 200, 1500, 1),  # ptm, real-model shape: 3 streams x 13 dims (4 k-steps)
    (1, 256, [13, 13, 13], 150, 900, 2),   # s2_semi
    (4, 512, [25], 100, 300, 1),           # 7 k-steps
])
def test_tied_tensor_core_path_identical_to_exact(C, Mden, streams, S, T, kind):
    """ptm / s2_semi codebook stage on the tensor cores (GEMM candidates + exact
    re-scoring + exact-scan fallback) must return the same scores as the exact
    CUDA-core scan, bit for bit -- including frames built to produce integer
    score ties at rank 4/5 (duplicated Gaussians), which must go through the
    fallback."""
    rng = np.random.default_rng(C * 1000 + Mden)
    F, V = len(streams), sum(streams)
    mean = rng.standard_normal((C, F, Mden, max(streams))).astype(np.float32)
    var = np.exp(rng.uniform(np.log(0.05), np.log(5.0), mean.shape)).astype(np.float32)
    # duplicated Gaussians (identical distance for every frame) in codebook 0, stream 0
    mean[0, 0, 700 % Mden] = mean[0, 0, 3]; var[0, 0, 700 % Mden] = var[0, 0, 3]
    mean[0, 0, 200] = mean[0, 0, 100]; var[0, 0, 200] = var[0, 0, 100]
    # six near-identical Gaussians inside ONE 64-column group of the GEMM tile: the group keeps only
    # its 4 best keys, so the 5th/6th never reach the merge stage -- the bound must notice
    for k in range(1, 6):
        mean[0, 0, 64 + k] = mean[0, 0, 64] + rng.standard_normal(mean.shape[3]).astype(np.float32) * 1e-3
        var[0, 0, 64 + k] = var[0, 0, 64]
    flat_m = np.concatenate([mean[:, f, :, :L].reshape(C, -1) for f, L in enumerate(streams)], 1)
    flat_v = np.concatenate([var[:, f, :, :L].reshape(C, -1) for f, L in enumerate(streams)], 1)
    pv_parts, pd_parts = [], []
    for f, L in enumerate(streams):
        a, d = orc.port_precompute(var[:, f, :, :L].reshape(-1, L), L, 1e-4, orc.LOGBASE)
        pv_parts.append(a.reshape(C, -1)); pd_parts.append(d.reshape(C, 1, Mden))
    pv = np.concatenate(pv_parts, 1); pd = np.concatenate(pd_parts, 1)
    mixw = rng.integers(0, 160, (F, Mden, S)).astype(np.uint8)
    s2c = (np.arange(S) * C // S).astype(np.uint8)
    cfg = b.MgauConfig(C, F, Mden, S, streams, topn=4, logbase=orc.LOGBASE)
    m = (b.ptm_from_arrays(cfg, flat_m, pv, pd, mixw, s2c) if kind == 1 else b.semi_from_arrays(cfg, flat_m, pv, pd, mixw))
    assert m.path == 1, "tensor-core codebook stage should be the default for this shape"
    feat = (rng.standard_normal((T, V)) * 1.3).astype(np.float32)
    # frames sitting on the duplicated Gaussians: exact ties inside the top 4
    feat[5, :streams[0]] = mean[0, 0, 3, :streams[0]]
    feat[6, :streams[0]] = mean[0, 0, 100, :streams[0]] + 0.01
    feat[7, :streams[0]] = mean[0, 0, 64, :streams[0]]
    feat[8, :streams[0]] = mean[0, 0, 66, :streams[0]] + 0.003
    got_tc = m.score(feat)
    n_lists, n_fallback = m.tied_stats()
    assert n_lists == T * C * F
    assert m.tied_max_err <= 12, f"GEMM distance error {m.tied_max_err} raw units: too close to the eps=32 bound"
    assert 1 <= n_fallback <= max(4, n_lists // 10), (n_lists, n_fallback)   # the duplicates are in many top-5s
    m.set_path(0)
    got_exact = m.score(feat)
    np.testing.assert_array_equal(got_tc, got_exact)
    # utterance cache + per-frame serving goes through the same lists
    m.set_path(1)
    m.utt_begin(feat[:64])
    out = np.zeros(S, np.int16)
    m.utt_frame(out, None, 0, 17, True)
    np.testing.assert_array_equal(out, got_exact[17])
    print(f"tied TC: {n_lists} lists, {n_fallback} via exact fallback ({n_fallback / n_lists:.2e})")
    m.free()


def test_config2_identical_at_scale():
    """1e8 scores of BASELINE config 2 (frames 40000..79999 of the sweep in
    tools/tc_fullscale_check.py, which includes both pairs the round-1 kernel got
    wrong by more than 1: frame 43770 / senone 3943 and frame 64372): tcgen05
    path vs the exact path -- identical, no tolerance, no whitelist.  The
    run-time monitor of the GEMM error must stay inside the bound the
    certificates assume (eps0 = 5 raw units at |d| = 0)."""
    n_sen, M, D = 5000, 32, 39
    mean, var, mixw = synth.cont_model(n_sen, M, D, 1234)
    pv, pd = orc.port_precompute(var.reshape(-1, D), D, 1e-4, orc.LOGBASE)
    q = orc.port_mixw_quantize(mixw, 1e-7, orc.LOGBASE)
    cfg = b.MgauConfig(n_sen, 1, M, n_sen, [D], topn=4, logbase=orc.LOGBASE)
    m = b.ms_from_arrays(cfg, mean, pv, pd, q, np.arange(n_sen))
    for t0 in (40000, 60000):
        feat = synth.cont_features(mean, var, 20000, 5678 + t0)
        m.set_path(1)
        got = m.score(feat)
        st = m.cont_stats()
        m.set_path(0)
        want = m.score(feat)
        print(f"t0 {t0}: differing scores {int((got != want).sum())}; last chunk {st}")
        np.testing.assert_array_equal(got, want)
        assert st["overflow"] == 0 and st["max_gemm_err"] <= 6, st
        assert 0 < st["hard_pairs"] < 0.3 * st["pairs"] and st["rescored_pairs"] < 0.04 * st["pairs"], st
    m.free()


def test_tensor_core_fp16_operands_and_tf32_fallback():
    """The default operand format is fp16 hi/lo (kind::f16); a batch with a
    feature outside the scaled fp16 range must be detected by the prep kernel
    and scored by the TF32 kernel instead -- same tolerance either way."""
    S, M, D, T = 300, 32, 39, 400
    mean, var, mixw = synth.cont_model(S, M, D, 41)
    pv, pd = orc.port_precompute(var.reshape(-1, D), D, 1e-4, orc.LOGBASE)
    q = orc.port_mixw_quantize(mixw, 1e-7, orc.LOGBASE)
    cfg = b.MgauConfig(S, 1, M, S, [D], topn=4, logbase=orc.LOGBASE)
    m = b.ms_from_arrays(cfg, mean, pv, pd, q, np.arange(S))
    feat = synth.cont_features(mean, var, T, 42)
    got = m.score(feat)
    assert m.tc_last_format() == 1
    big = feat.copy()
    big[123, 7] = 3.0e4                     # x^2 = 9e8: far beyond what fp16 can hold at any tile's scale
    got_big = m.score(big)
    assert m.tc_last_format() == 0
    m.set_path(0)
    want, want_big = m.score(feat), m.score(big)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(got_big, want_big)
    m.set_path(1)
    assert np.array_equal(m.score(feat), got) and m.tc_last_format() == 1     # back on fp16 for clean batches
    m.free()


def test_tensor_core_tf32_operands_when_fp16_is_disabled(monkeypatch):
    """B200_TC_F16=0 at model creation keeps the TF32 hi/lo kernels as the only
    operand format (they are also the overflow fallback): same tolerance."""
    monkeypatch.setenv("B200_TC_F16", "0")
    S, M, D, T = 200, 16, 39, 300
    mean, var, mixw = synth.cont_model(S, M, D, 51)
    pv, pd = orc.port_precompute(var.reshape(-1, D), D, 1e-4, orc.LOGBASE)
    q = orc.port_mixw_quantize(mixw, 1e-7, orc.LOGBASE)
    cfg = b.MgauConfig(S, 1, M, S, [D], topn=4, logbase=orc.LOGBASE)
    m = b.ms_from_arrays(cfg, mean, pv, pd, q, np.arange(S))
    feat = synth.cont_features(mean, var, T, 52)
    got = m.score(feat)
    assert m.path == 1 and m.tc_last_format() == 0
    m.set_path(0)
    np.testing.assert_array_equal(got, m.score(feat))
    m.free()


def test_tensor_core_mixed_operand_formats_per_tile():
    """The operand format is decided per n-tile (8 senones) and per batch: senones
    with very sharp Gaussians (variances at the floor, as real models have for
    untrained densities) leave their tiles to the TF32 kernel while the rest of
    the model runs on fp16 -- both kernels write one consistent score matrix."""
    S, M, D, T = 400, 32, 39, 300
    mean, var, mixw = synth.cont_model(S, M, D, 61)
    sharp = [5, 6, 130, 390]                       # senones in 4 different tiles
    for s_ in sharp:
        var[s_, 3] = 1e-6                           # floored to 1e-4 -> 1/(2 var ln b) = 5e7
        mean[s_, 3] = 0.0
    pv, pd = orc.port_precompute(var.reshape(-1, D), D, 1e-4, orc.LOGBASE)
    q = orc.port_mixw_quantize(mixw, 1e-7, orc.LOGBASE)
    cfg = b.MgauConfig(S, 1, M, S, [D], topn=4, logbase=orc.LOGBASE)
    m = b.ms_from_arrays(cfg, mean, pv, pd, q, np.arange(S))
    feat = synth.cont_features(mean, var, T, 62)
    got = m.score(feat)
    assert m.tc_last_format() == 2, "expected some tiles on fp16 and some on TF32"
    m.set_path(0)
    want = m.score(feat)
    np.testing.assert_array_equal(got, want)
    m.free()


def test_tied_tensor_core_fuzz_identical_to_exact():
    """Random tied models: 1-3 streams of 3..39 dims, 256/512/768 densities, topn 1..4,
    ptm and s2_semi -- tensor-core lists must reproduce the exact scan bit for bit."""
    for seed in range(12):
        rng = np.random.default_rng(500 + seed)
        kind = int(rng.choice([1, 2]))
        C = 1 if kind == 2 else int(rng.integers(1, 6))
        Mden = int(rng.choice([256, 512, 768]))
        streams = [int(rng.integers(3, 40)) for _ in range(int(rng.integers(1, 4)))]
        topn = int(rng.integers(1, 5))
        S, T = int(rng.integers(20, 150)), int(rng.integers(1, 400))
        F, V = len(streams), sum(streams)
        mean = rng.standard_normal((C, F, Mden, max(streams))).astype(np.float32) * float(rng.uniform(0.5, 3))
        var = np.exp(rng.uniform(np.log(0.02), np.log(8.0), mean.shape)).astype(np.float32)
        flat_m = np.concatenate([mean[:, f, :, :L].reshape(C, -1) for f, L in enumerate(streams)], 1)
        pv_parts, pd_parts = [], []
        for f, L in enumerate(streams):
            a, d = orc.port_precompute(var[:, f, :, :L].reshape(-1, L), L, 1e-4, orc.LOGBASE)
            pv_parts.append(a.reshape(C, -1)); pd_parts.append(d.reshape(C, 1, Mden))
        pv = np.concatenate(pv_parts, 1); pd = np.concatenate(pd_parts, 1)
        mixw = rng.integers(0, 160, (F, Mden, S)).astype(np.uint8)
        s2c = (np.arange(S) * C // S).astype(np.uint8)
        cfg = b.MgauConfig(C, F, Mden, S, streams, topn=topn, logbase=orc.LOGBASE)
        m = (b.ptm_from_arrays(cfg, flat_m, pv, pd, mixw, s2c) if kind == 1 else b.semi_from_arrays(cfg, flat_m, pv, pd, mixw))
        assert m.path == 1
        feat = (rng.standard_normal((T, V)) * float(rng.uniform(0.5, 2.5))).astype(np.float32)
        got = m.score(feat)
        m.set_path(0)
        np.testing.assert_array_equal(got, m.score(feat), err_msg=f"seed {seed}: kind {kind} C {C} M {Mden} streams {streams} topn {topn}")
        m.free()


def test_tensor_core_ms_fuzz_identical_to_exact():
    """Random fully-continuous shapes (8/16/32 densities, 3..39 dims, variance spread up to
    4 decades, feature scale 0.3..6 -- sharp Gaussians far from the features, where the GEMM
    error is large): the tensor-core path stays IDENTICAL to the exact path; several of
    these models trip the error monitor / queue limits and are redone by the literal scan."""
    for seed in range(10):
        rng = np.random.default_rng(900 + seed)
        S, M, D, T = int(rng.integers(9, 700)), int(rng.choice([8, 16, 32])), int(rng.integers(3, 40)), int(rng.integers(1, 600))
        mean = (rng.standard_normal((S, M, D)) * float(rng.uniform(0.3, 4))).astype(np.float32)
        var = np.exp(rng.uniform(np.log(1e-3), np.log(10.0), (S, M, D))).astype(np.float32)
        mixw = rng.dirichlet(np.ones(M), (S, 1)).astype(np.float32)
        pv, pd = orc.port_precompute(var.reshape(-1, D), D, 1e-4, orc.LOGBASE)
        q = orc.port_mixw_quantize(mixw, 1e-7, orc.LOGBASE)
        cfg = b.MgauConfig(S, 1, M, S, [D], topn=4, logbase=orc.LOGBASE)
        m = b.ms_from_arrays(cfg, mean, pv, pd, q, np.arange(S))
        assert m.path == 1
        feat = (mean[rng.integers(0, S, T), rng.integers(0, M, T)] +
                rng.standard_normal((T, D)) * float(rng.uniform(0.3, 6))).astype(np.float32)
        got = m.score(feat)
        fmt, st = m.tc_last_format(), m.cont_stats()
        m.set_path(0)
        np.testing.assert_array_equal(got, m.score(feat), err_msg=f"seed {seed}: S {S} M {M} D {D} T {T} format {fmt} {st}")
        m.free()


def test_two_devices_in_one_process():
    """One process driving two GPUs (skipped on a single-GPU box): every handle carries its
    device, per-device kernel attributes are set on both."""
    if b.device_count() < 2:
        pytest.skip("needs two GPUs")
    S, M, D, T = 120, 32, 39, 260
    mean, var, mixw = synth.cont_model(S, M, D, 71)
    pv, pd = orc.port_precompute(var.reshape(-1, D), D, 1e-4, orc.LOGBASE)
    q = orc.port_mixw_quantize(mixw, 1e-7, orc.LOGBASE)
    feat = synth.cont_features(mean, var, T, 72)
    outs = []
    for dev in (0, 1):
        cfg = b.MgauConfig(S, 1, M, S, [D], topn=4, logbase=orc.LOGBASE, device=dev)
        m = b.ms_from_arrays(cfg, mean, pv, pd, q, np.arange(S))
        a = m.score(feat)
        m.set_path(0)
        e = m.score(feat)
        np.testing.assert_array_equal(a, e)
        outs.append((a, e))
        m.free()
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
