"""CPU suite: the oracle port and the product's host-side (load-time) code
against the committed golden vectors generated from the reference
(tests/golden/make_golden.py), plus the C-ABI export check."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cases
import orc
import cmusphinx_b200 as b

ROOT = orc.ROOT


# ----------------------------------------------------------------- logmath
def test_logadd_table_golden():
    g = cases.load("logmath.npz")
    base = float(g["base"])
    np.testing.assert_array_equal(orc.port_logadd_table(base, 10), g["table10"])
    np.testing.assert_array_equal(b.logadd_table(base, 10).astype(np.int32), g["table10"])
    assert list(g["table10"][:18]) == [7, 6, 6, 5, 5, 5, 4, 4, 4, 3, 3, 3, 3, 2, 2, 2, 2, 2]
    assert g["table10"].size == 256


# The reference's own logmath unit tests (sphinxbase/test/unit/test_logmath/): literal
# expected values and tolerance (LOG_EPSILON = 1500 raw units, compared after the shift).
_LOGMATH_KATS = [
    # (base, shift, [(p, expected log)])                      source
    (1.0001, 8, [(1e-150, -13493), (42.0, 146)]),          # test_log_shifted.c:13-24
    (1.003, 0, [(1e-48, -36896), (42.0, 1247)]),           # test_log_int8.c:13-20
    (1.0001, 0, [(1e-150, -3454050), (42.0, 37378)]),      # test_log_int16.c:13-23
]


@pytest.mark.parametrize("base,shift,kats", _LOGMATH_KATS)
def test_reference_logmath_unit_test_known_answers(base, shift, kats):
    """logmath_log / logmath_add of the oracle and of the product's host code reproduce the
    known answers the reference's unit tests assert, within the reference's own tolerance,
    and the two add identities those tests check (1e-48 + 5e-48 = 6e-48; 1e-48 + 42 = 42)."""
    eps = 1500
    lm = orc.port.orc_logmath_init(base, shift, 1)
    for p, want in kats:
        for got in (orc.port.orc_logmath_log(lm, p), b.lib.b200_logmath_log(base, shift, p)):
            assert abs(got - want) < eps, (base, shift, p, got, want)
    for log, add in ((lambda p: orc.port.orc_logmath_log(lm, p), lambda x, y: orc.port.orc_logmath_add(lm, x, y)),
                     (lambda p: b.lib.b200_logmath_log(base, shift, p),
                      lambda x, y: b.lib.b200_logmath_add(base, shift, x, y))):
        assert abs(add(log(1e-48), log(5e-48)) - log(6e-48)) < eps
        assert abs(add(log(1e-48), log(42.0)) - log(42.0)) < eps
    orc.port.orc_logmath_free(lm)


def test_logmath_log_add_golden():
    g = cases.load("logmath.npz")
    base = float(g["base"])
    for sh in (0, 10):
        lm = orc.port.orc_logmath_init(base, sh, 1)
        want = g[f"log_shift{sh}"]
        got_o = np.array([orc.port.orc_logmath_log(lm, float(p)) for p in g["p"]], np.int32)
        got_p = np.array([b.lib.b200_logmath_log(base, sh, float(p)) for p in g["p"]], np.int32)
        np.testing.assert_array_equal(got_o, want)
        np.testing.assert_array_equal(got_p, want)
        orc.port.orc_logmath_free(lm)
    lm = orc.port.orc_logmath_init(base, 10, 1)
    got_o = np.array([orc.port.orc_logmath_add(lm, int(x), int(y)) for x, y in zip(g["add_x"], g["add_y"])], np.int32)
    got_p = np.array([b.lib.b200_logmath_add(base, 10, int(x), int(y)) for x, y in zip(g["add_x"][:300], g["add_y"][:300])],
                     np.int32)
    np.testing.assert_array_equal(got_o, g["add_out"])
    np.testing.assert_array_equal(got_p, g["add_out"][:300])
    orc.port.orc_logmath_free(lm)
    # sphinxbase test_log_shifted.c known answers (shift 8, base 1.0001)
    assert abs(b.lib.b200_logmath_log(1.0001, 8, 1e-150) - (-13493)) <= 5
    assert abs(b.lib.b200_logmath_log(1.0001, 8, 42.0) - 146) <= 5


# ------------------------------------------------------- load-time tables
MS_CASES = ["ms_small.npz", "ms_3stream.npz", "ms_allden.npz", "ms_cont32.npz"]


@pytest.mark.parametrize("name", MS_CASES)
def test_precompute_and_quantiser_golden(name):
    g = cases.load(name)
    n_sen, n_density, dim, n_feat, topn, T = cases.ms_dims(g)
    for fn in (orc.port_precompute, b.gauden_precompute):
        pv, pd = fn(g["var_raw"].reshape(-1, dim), dim, 1e-4, orc.LOGBASE)
        np.testing.assert_array_equal(pv.reshape(-1), g["var"])
        np.testing.assert_array_equal(pd.reshape(-1), g["det"])
    np.testing.assert_array_equal(orc.port_mixw_quantize(g["mixw_raw"], 1e-7, orc.LOGBASE).reshape(-1), g["mixw"])
    np.testing.assert_array_equal(b.mixw_quantize_ms(g["mixw_raw"], 1e-7, orc.LOGBASE).reshape(-1), g["mixw"])


def test_tmat_quantiser_golden():
    g = cases.load("tmat_hmm.npz")
    np.testing.assert_array_equal(orc.port_tmat_quantize(g["tp_raw"], 1e-4, orc.LOGBASE), g["tp"])
    np.testing.assert_array_equal(b.tmat_quantize(g["tp_raw"], 1e-4, orc.LOGBASE), g["tp"])


def test_tied_quantiser_port_equals_product():
    rng = np.random.default_rng(1)
    mw = rng.dirichlet(np.ones(16) * 0.3, (37, 3)).astype(np.float32)
    mw[2, 1, :4] = 0
    np.testing.assert_array_equal(orc.port_mixw_quantize_tied(mw, 1e-7, orc.LOGBASE),
                                  b.mixw_quantize_tied(mw, 1e-7, orc.LOGBASE))


def test_flags2list_bridging():
    n_sen = 5150
    mask = np.zeros((n_sen + 31) // 32, np.uint32)
    for s in (3, 4, 700, 701, 2000, 5149):
        mask[s // 32] |= np.uint32(1 << (s % 32))
    want = orc.port_flags2list(mask, n_sen)
    got = b.flags2list(mask, n_sen)
    np.testing.assert_array_equal(got, want)
    ids = np.cumsum(got.astype(int))
    assert set((3, 4, 700, 701, 2000, 5149)) <= set(ids) and (got[got == 255].size > 0)
    assert b.flags2list(np.zeros(4, np.uint32), 100).size == 0


def test_s3_writers_readers_roundtrip(tmp_path):
    from cmusphinx_b200 import s3io
    rng = np.random.default_rng(2)
    a = rng.standard_normal((5, 4, 7)).astype(np.float32)
    c = rng.standard_normal((5, 4, 3)).astype(np.float32)
    s3io.write_gauden(str(tmp_path / "m"), [a, c], [7, 3])
    r = b.read_gauden(str(tmp_path / "m"))
    assert (r["n_mgau"], r["n_feat"], r["n_density"], r["veclen"]) == (5, 2, 4, [7, 3])
    want = np.concatenate([np.concatenate([a[m].reshape(-1), c[m].reshape(-1)]) for m in range(5)])
    np.testing.assert_array_equal(r["data"], want)
    w = rng.random((6, 2, 4)).astype(np.float32)
    s3io.write_mixw(str(tmp_path / "w"), w)
    np.testing.assert_array_equal(b.read_mixw(str(tmp_path / "w")), w)
    with pytest.raises(b.B200Error):
        b.read_mixw(str(tmp_path / "does_not_exist"))
    with open(tmp_path / "trunc", "wb") as fp:
        fp.write(open(tmp_path / "w", "rb").read()[:-5])
    with pytest.raises(b.B200Error):
        b.read_mixw(str(tmp_path / "trunc"))


# ------------------------------------------------------ oracle vs goldens
@pytest.mark.parametrize("name", MS_CASES)
def test_oracle_ms_golden(name):
    g = cases.load(name)
    pm = cases.ms_oracle(g)
    np.testing.assert_array_equal(pm.eval_all(g["feat"]), g["dense"])
    for i in range(g["act_scores"].shape[0]):
        out = np.full(pm.n_sen, 12345, np.int16)
        pm.frame_eval(g["feat"][i], cases.deltas_of(g, i), False, out)
        np.testing.assert_array_equal(out, g["act_scores"][i])


@pytest.mark.parametrize("ne", [3, 5, 1, 2, 4])   # 1, 2, 4: hmm_vit_eval_anytopo
def test_oracle_hmm_golden(ne):
    g = cases.load("tmat_hmm.npz" if ne in (3, 5) else "hmm_anytopo.npz")
    a = {k: g[f"h{ne}_in_{k}"].copy() for k in ("score", "history", "out_score", "out_history", "senid", "tmatid",
                                                "mpx", "bestscore", "sseq")}
    for f in range(g[f"h{ne}_senscr"].shape[0]):
        best = orc.hmm_eval(orc.port.orc_hmm_eval_batch, ne, g[f"h{ne}_tp"], a["sseq"], g[f"h{ne}_senscr"][f],
                            a["score"], a["history"], a["out_score"], a["out_history"], a["senid"], a["tmatid"],
                            a["mpx"], a["bestscore"])
        assert best == int(g[f"h{ne}_best"][f])
    for k in ("score", "history", "out_score", "out_history", "senid", "bestscore"):
        np.testing.assert_array_equal(a[k], g[f"h{ne}_out_{k}"], err_msg=k)


@pytest.mark.parametrize("name,kind", [("semi_hub4wsj.npz", 2), ("ptm_hub4wsj.npz", 1)])
def test_oracle_tied_golden(name, kind):
    if not cases.have_model(name):
        pytest.skip("model files (oracle/_ref/data) not present")
    g = cases.load(name)
    gm, gv, sd, n_sen = cases.tied_arrays(name, g)
    L = gm["veclen"][0]
    pv, pd = orc.port_precompute(gv["data"].reshape(-1, L), L, 1e-4, orc.LOGBASE)
    s2c = g["sen2cb"] if kind == 1 else None
    pt = orc.PortTied(kind, gm["n_mgau"], gm["n_feat"], gm["veclen"], gm["n_density"], n_sen, 4, gm["data"], pv, pd,
                      sd["mixw"], sd["n_clust"], sd["mixw_cb"], s2c, orc.LOGBASE)
    np.testing.assert_array_equal(pt.eval_all(g["feat"]), g["dense"])
    pt.reset()
    for i in range(g["act_scores"].shape[0]):
        got = pt.frame_eval(g["feat"][i], cases.deltas_of(g, i), False, i)
        np.testing.assert_array_equal(got, g["act_scores"][i])


def test_oracle_semi_topn_beam_golden():
    """-topn_beam (s2_semi_mgau.c:189-207) against the reference's scores for two beam settings."""
    name = "semi_hub4wsj.npz"
    if not cases.have_model(name):
        pytest.skip("model files (oracle/_ref/data) not present")
    g, gb = cases.load(name), cases.load("semi_hub4wsj_beam.npz")
    gm, gv, sd, n_sen = cases.tied_arrays(name, g)
    pv, pd = orc.port_precompute(gv["data"].reshape(-1, 13), 13, 1e-4, orc.LOGBASE)
    pt = orc.PortTied(2, 1, 3, gm["veclen"], gm["n_density"], n_sen, 4, gm["data"], pv, pd, sd["mixw"], sd["n_clust"],
                      sd["mixw_cb"], None, orc.LOGBASE)
    for i, beam in enumerate(gb["beams"]):
        pt.reset()
        pt.set_topn_beam([int(v) for v in beam])
        np.testing.assert_array_equal(pt.eval_all(g["feat"]), gb[f"dense{i}"])
        assert (gb[f"dense{i}"] != g["dense"]).mean() > 0.3   # the beam really changes the scores
    for i, (ds, beam) in enumerate(gb["ds_cfg"]):
        pt.reset()
        pt.set_topn_beam([int(beam)] * 3)
        pt.set_ds(int(ds))
        np.testing.assert_array_equal(pt.eval_all(g["feat"]), gb[f"ds_dense{i}"])
        assert (gb[f"ds_dense{i}"] != g["dense"]).mean() > 0.3


@pytest.mark.parametrize("name", ["cont_hub4_topn4.npz", "cont_hub4_topn8.npz"])
def test_oracle_real_cont_golden(name):
    if not cases.have_model(name):
        pytest.skip("model files (oracle/_ref/data) not present")
    g = cases.load(name)
    d = cases.model_dir(name)
    gm, gv, mw = b.read_gauden(d + "/means"), b.read_gauden(d + "/variances"), b.read_mixw(d + "/mixture_weights")
    pv, pd = orc.port_precompute(gv["data"].reshape(-1, 39), 39, 1e-4, orc.LOGBASE)
    q = orc.port_mixw_quantize(mw, 1e-7, orc.LOGBASE)
    n_sen = gm["n_mgau"]
    pm = orc.PortMs(n_sen, 1, [39], gm["n_density"], n_sen, int(g["topn"]), 1, gm["data"], pv, pd, q,
                    np.arange(n_sen), orc.LOGBASE)
    np.testing.assert_array_equal(pm.eval_all(g["feat"][:6]), g["dense"][:6])


# ------------------------------------------------------------- ABI surface
def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200sphinx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 45
    L = C.CDLL(b.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared but not exported: {missing}"
    assert b.lib.b200_abi_version() == 1


def test_no_gpu_means_loud_failure_not_fallback():
    if b.device_count() > 0:
        pytest.skip("GPU present")
    g = cases.load("ms_small.npz")
    with pytest.raises(b.B200Error, match="no CUDA device"):
        cases.ms_product(g)
    with pytest.raises(b.B200Error, match="no CUDA device"):
        b.HmmContext(3, np.zeros((1, 3, 4), np.uint8), None, 10)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "cmusphinx_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".c")):
                txt = open(os.path.join(dp, fn), errors="replace").read()
                assert "liboracle" not in txt and "sphinx_oracle" not in txt and "libref_shim" not in txt, fn


# ------------------------------------------------------------------- mdef maps
@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("hmm", ["ptm", "hub4wsj_sc_8k", "cont"])
def test_mdef_maps_match_reference(hmm):
    """Native model-definition reader (binary BMDF: ptm, hub4wsj_sc_8k; text 0.3: the
    sphinx3 continuous model) against the reference's bin_mdef (sen2cimap) and, for
    the text file, sphinx3's mdef_init (cd2cisen)."""
    d = os.path.join(orc.DATA_DIR, "hmm", hmm)
    if not os.path.exists(os.path.join(d, "mdef")):
        pytest.skip("model not bundled")
    mm = b.mdef_maps(os.path.join(d, "mdef"))
    r = orc.RefAcmod(d)
    assert mm["n_sen"] == r.n_sen
    np.testing.assert_array_equal(mm["sen2cimap"].astype(np.uint8), r.sen2cimap())
    r.close()
    assert (mm["cd2cisen"][:mm["n_ci_sen"]] == np.arange(mm["n_ci_sen"])).all()
    assert (mm["cd2cisen"] >= 0).all() and (mm["cd2cisen"] < mm["n_ci_sen"]).all()
    if hmm == "cont" and orc.have_ref_s3():
        r3 = orc.RefS3(*(os.path.join(d, n) for n in ("means", "variances", "mixture_weights", "mdef")))
        np.testing.assert_array_equal(mm["cd2cisen"], r3.cd2cisen())
        assert mm["n_ci_sen"] == r3.n_ci_sen
        r3.free()
