"""Feature stage (cepstra -> 1s_c_d_dd + cmn current): oracle port pinned to
the reference's feat_s2mfc2feat_live(full utterance) and to a golden vector;
GPU kernel bit-exact against both."""
import os

import numpy as np
import pytest

import cases
import orc
import cmusphinx_b200 as b

MFC = os.path.join(orc.DATA_DIR, "test", "wsj", "442c0201.mfc")


@pytest.mark.skipif(not (orc.have_ref() and os.path.exists(MFC)), reason="oracle/_ref not built")
@pytest.mark.parametrize("hmm", ["hub4wsj_sc_8k", "cont"])
def test_feature_port_matches_reference(hmm):
    """hub4wsj_sc_8k: 1s_c_d_dd + svspec 0-12/13-25/26-38 (3 streams); cont: one 39-dim stream."""
    r = orc.RefAcmod(os.path.join(orc.DATA_DIR, "hmm", hmm))
    cep = orc.read_mfc(MFC)
    for c in (cep, cep[:7], cep[:1], cep[100:105] * 3.0):
        want = r.cep2feat(c)
        got = orc.port_feat_1s_c_d_dd(c, True)
        assert want.shape == got.shape
        np.testing.assert_array_equal(got, want)
    r.close()


def test_feature_port_matches_golden():
    g = cases.load("feat_442.npz")
    np.testing.assert_array_equal(orc.port_feat_1s_c_d_dd(g["cep"], True), g["feat"])


@pytest.mark.gpu
def test_feature_kernel_bit_exact():
    g = cases.load("feat_442.npz")
    np.testing.assert_array_equal(b.feat_1s_c_d_dd(g["cep"]), g["feat"])
    # a ragged batch of utterances, including 1- and 2-frame ones, with and without cmn
    rng = np.random.default_rng(3)
    lens = [1, 2, 5, 300, 7, 1, 64, 129]
    cep = (rng.standard_normal((sum(lens), 13)) * 4 + rng.standard_normal(13) * 10).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    for cmn in (True, False):
        want = np.concatenate([orc.port_feat_1s_c_d_dd(cep[off[u]:off[u + 1]], cmn) for u in range(len(lens))])
        np.testing.assert_array_equal(b.feat_1s_c_d_dd(cep, off, cmn), want)


@pytest.mark.gpu
def test_feature_kernel_empty_batch():
    out = b.feat_1s_c_d_dd(np.zeros((0, 13), np.float32), np.array([0, 0], np.int32))
    assert out.shape == (0, 39)


# ------------------------------------------------------------------ general form
_CFGS = [(1, 0, 0), (0, 0, 0), (1, 1, 0), (1, 0, 1), (1, 1, 1), (0, 0, 1)]   # (cmn, varnorm, agc)


def _cep(T, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((T, 13)) * 4 + rng.standard_normal(13) * 10).astype(np.float32)


def _same(a, b):
    # 1-frame utterances with -varnorm divide by a zero variance: NaN in the reference too
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("ftype", orc.FEAT_TYPES)
def test_general_feature_port_matches_reference(ftype):
    """Every -feat type x cmn/varnorm/agc against the reference's own feat_t (feat_init +
    feat_s2mfc2feat_live(full utterance)), including 1- and 2-frame utterances."""
    cep = _cep(60)
    for cmn, vn, agc in _CFGS:
        for T in (60, 9, 2, 1):
            assert _same(orc.port_feat_compute(cep[:T], ftype, cmn, vn, agc),
                         orc.ref_feat_compute(cep[:T], ftype, cmn, vn, agc)), (ftype, cmn, vn, agc, T)


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")
def test_lda_and_subvector_port_matches_reference():
    """feat_lda_transform (-lda / -ldadim) and feat_subvec_project (-svspec), alone and chained."""
    cep = _cep(40, 1)
    rng = np.random.default_rng(2)
    for ftype in ("1s_c_d_dd", "s3_1x39", "1s_c_d_ld_dd", "1s_c_d"):
        k = orc.port_feat_compute(cep, ftype).shape[1]
        lda = rng.standard_normal((k, k)).astype(np.float32)
        for dim in (0, 20, k):
            for sv in (None, "0-9/10-19", "3,1,5/0"):
                assert _same(orc.port_feat_compute(cep, ftype, True, False, False, lda, dim, sv),
                             orc.ref_feat_compute(cep, ftype, True, False, False, lda, dim, sv)), (ftype, dim, sv)
        assert _same(orc.port_feat_compute(cep, ftype, svspec="0-12/13-25"),
                     orc.ref_feat_compute(cep, ftype, svspec="0-12/13-25"))


def test_general_feature_port_matches_golden():
    g = cases.load("feat_general.npz")
    lda = g["lda"]
    for i, (ftype, cmn, vn, agc, use_lda, dim, sv) in enumerate(cases.FEAT_GOLDEN_CASES):
        got = orc.port_feat_compute(g["cep"], ftype, cmn, vn, agc, lda[:_klen(ftype), :_klen(ftype)] if use_lda else None, dim, sv)
        assert _same(got, g[f"out{i}"]), (ftype, cmn, vn, agc, use_lda, dim, sv)
    # and the special case is the general one
    np.testing.assert_array_equal(orc.port_feat_compute(g["cep"]), orc.port_feat_1s_c_d_dd(g["cep"], True))


def _klen(ftype):
    return {"1s_c_d_dd": 39, "s3_1x39": 39, "s2_4x": 51, "1s_c_d_ld_dd": 52, "1s_c": 13, "1s_c_d": 26}[ftype]


@pytest.mark.gpu
def test_general_feature_kernels_match_golden():
    g = cases.load("feat_general.npz")
    lda = g["lda"]
    for i, (ftype, cmn, vn, agc, use_lda, dim, sv) in enumerate(cases.FEAT_GOLDEN_CASES):
        got = b.feat_compute(g["cep"], None, ftype, cmn, vn, agc, lda[:_klen(ftype), :_klen(ftype)] if use_lda else None, dim,
                             orc.parse_svspec(sv) if sv else None)
        assert _same(got, g[f"out{i}"]), (ftype, cmn, vn, agc, use_lda, dim, sv)


@pytest.mark.gpu
@pytest.mark.parametrize("ftype", orc.FEAT_TYPES)
def test_general_feature_kernels_bit_exact_ragged_batch(ftype):
    """A ragged batch (1-, 2-frame and long utterances) in one call == the port per utterance."""
    rng = np.random.default_rng(5)
    lens = [1, 2, 5, 300, 7, 1, 64, 129, 9]
    cep = _cep(sum(lens), 7)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    k = _klen(ftype)
    lda = rng.standard_normal((k, k)).astype(np.float32)
    for cmn, vn, agc in _CFGS:
        for use_lda, dim, sv in ((False, 0, None), (True, 0, None), (True, k - 7, "0-3/9,8"), (False, 0, "2,0/5-7")):
            if ftype == "s2_4x" and (use_lda or sv):
                continue
            want = np.concatenate([orc.port_feat_compute(cep[off[u]:off[u + 1]], ftype, cmn, vn, agc,
                                                         lda if use_lda else None, dim, sv) for u in range(len(lens))])
            got = b.feat_compute(cep, off, ftype, cmn, vn, agc, lda if use_lda else None, dim,
                                 orc.parse_svspec(sv) if sv else None)
            assert _same(got, want), (ftype, cmn, vn, agc, use_lda, dim, sv)


@pytest.mark.gpu
def test_general_feature_stage_refuses_what_the_reference_rejects():
    cep = _cep(10)
    with pytest.raises(b.B200Error):
        b.feat_compute(cep, None, "s2_4x", lda=np.eye(51, dtype=np.float32))          # lda.c:69-73
    with pytest.raises(b.B200Error):
        b.feat_compute(cep, None, "1s_c_d_dd", lda=np.eye(38, dtype=np.float32))      # lda.c:127-128
    with pytest.raises(b.B200Error):
        b.feat_compute(cep[:, :12], None, "s3_1x39")                                  # feat.c: cepsize must be 13
    with pytest.raises(b.B200Error):
        b.feat_compute(cep, None, "1s_c", subvec=list(range(14)))                     # feat.c:309-313
    assert b.feat_compute(np.zeros((0, 13), np.float32), np.array([0, 0], np.int32), "s2_4x").shape == (0, 51)


# ------------------------------------------------- the reference's own feature unit test
_UNIT = [("13", "res_13"), ("13:1", "res_13_1"), ("1s_c_d_dd", "res_1s_c_d_dd")]


def _as_printed(x):
    """what test_feat.c prints: "%.3f" of every value"""
    return np.array([[float("%.3f" % v) for v in row] for row in x])


@pytest.mark.parametrize("ftype,key", _UNIT)
def test_port_reproduces_the_references_feature_unit_test(ftype, key):
    """sphinxbase/test/unit/test_feat/test_feat.c + _test_feat.res (cmn none, agc none)."""
    g = cases.load("feat_unit_test.npz")
    np.testing.assert_array_equal(_as_printed(orc.port_feat_compute(g["cep"], ftype, False)), g[key])


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("ftype", ["13", "13:1", "13:3", "5,8:2", "4,4,5", "12"])
def test_numeric_feature_types_match_reference(ftype):
    """feat_copy types "n[,n..][:w]" (feat.c:828-849, 952-1000), with and without normalisation."""
    cep = _cep(40, 3)
    for cmn, vn, agc in ((0, 0, 0), (1, 0, 0), (1, 1, 1)):
        for T in (40, 2, 1):
            assert _same(orc.port_feat_compute(cep[:T], ftype, cmn, vn, agc),
                         orc.ref_feat_compute(cep[:T], ftype, cmn, vn, agc)), (ftype, cmn, vn, agc, T)


def test_window_copy_types_with_undefined_edges_are_not_offered():
    """`1s_3c` / `1s_4c`: feat_s3_cepwin copies contiguous memory across the non-contiguous padded
    utterance -- the reference's first and last w frames are undefined (checked against it below)."""
    with pytest.raises(ValueError, match="feat_s3_cepwin"):
        b.engine.parse_feat_type("1s_3c")
    if orc.have_ref():
        cep = _cep(30, 4)
        r, a = orc.ref_feat_compute(cep, "1s_3c", False), orc.port_feat_compute(cep, "13:3", False)
        assert np.array_equal(r[3:-3], a[3:-3])          # interior frames: the plain window copy


@pytest.mark.gpu
def test_feature_kernels_reproduce_the_references_unit_test_and_numeric_types():
    g = cases.load("feat_unit_test.npz")
    for ftype, key in _UNIT:
        np.testing.assert_array_equal(_as_printed(b.feat_compute(g["cep"], None, ftype, False)), g[key])
    lens = [1, 2, 40, 7, 130]
    cep = _cep(sum(lens), 9)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    for ftype in ("13", "13:1", "13:3", "5,8:2", "4,4,5", "12", "13:7"):
        for cmn, vn, agc in ((0, 0, 0), (1, 0, 0), (1, 1, 1)):
            want = np.concatenate([orc.port_feat_compute(cep[off[u]:off[u + 1]], ftype, cmn, vn, agc)
                                   for u in range(len(lens))])
            assert _same(b.feat_compute(cep, off, ftype, cmn, vn, agc), want), (ftype, cmn, vn, agc)
    with pytest.raises(b.B200Error):
        b.feat_compute(cep, off, "7,7")           # streams longer than the cepstral vector


def test_feature_plan_dims_and_refusals_host_only():
    """b200_feat_dims is host code (no device needed): window / stream length / output length of
    every type, and the configurations feat_init / feat_read_lda / feat_set_subvecs reject."""
    import ctypes as C
    from cmusphinx_b200.engine import _FeatCfg, parse_feat_type

    def dims(ftype, cs=13, lda=None, lda_dim=0, sv=None, **kw):
        cfg = _FeatCfg()
        tid, cw, clen = parse_feat_type(ftype)
        cfg.type, cfg.cepsize, cfg.cmn = tid, cs, kw.get("cmn", 1)
        cfg.agc = kw.get("agc", 0)
        cfg.copy_window, cfg.copy_streams = cw, len(clen)
        for j, l in enumerate(clen):
            cfg.copy_len[j] = l
        keep = []
        if lda is not None:
            lda = np.ascontiguousarray(lda, np.float32); keep.append(lda)
            cfg.lda_rows, cfg.lda_cols, cfg.lda_dim = lda.shape[0], lda.shape[1], lda_dim
            cfg.lda = lda.ctypes.data_as(C.POINTER(C.c_float))
        if sv is not None:
            sv = np.ascontiguousarray(sv, np.int32); keep.append(sv)
            cfg.n_subvec, cfg.subvec = sv.size, sv.ctypes.data_as(C.POINTER(C.c_int32))
        d = (C.c_int32 * 3)()
        rc = b.lib.b200_feat_dims(C.byref(cfg), d)
        return rc, list(d)

    assert dims("1s_c_d_dd") == (0, [3, 39, 39])
    assert dims("s3_1x39") == (0, [3, 39, 39])
    assert dims("s2_4x") == (0, [4, 51, 51])
    assert dims("1s_c_d_ld_dd") == (0, [4, 52, 52])
    assert dims("1s_c") == (0, [0, 13, 13])
    assert dims("1s_c_d") == (0, [2, 26, 26])
    assert dims("13:1") == (0, [1, 39, 39])
    assert dims("5,8:2") == (0, [2, 65, 65])
    assert dims("1s_c_d_dd", lda=np.eye(39), lda_dim=29) == (0, [3, 39, 29])
    assert dims("1s_c_d_dd", lda=np.eye(39), lda_dim=99) == (0, [3, 39, 39])      # lda.c:131-134
    assert dims("1s_c_d_dd", lda=np.eye(39), lda_dim=29, sv=[0, 5, 7]) == (0, [3, 39, 3])
    assert dims("s3_1x39", cs=12)[0] < 0                    # feat.c: s3_1x39 needs 13 cepstra
    assert dims("s2_4x", lda=np.eye(51))[0] < 0             # lda.c:69-73
    assert dims("1s_c_d_dd", lda=np.eye(38))[0] < 0         # lda.c:127-128
    assert dims("1s_c", sv=list(range(14)))[0] < 0          # feat.c:309-313
    assert dims("7,7")[0] < 0                               # more stream values than cepstra
    assert dims("1s_c_d_dd", cmn=2)[0] < 0                  # prior CMN: live mode only
    assert dims("1s_c_d_dd", agc=2)[0] < 0                  # emax AGC: live mode only
