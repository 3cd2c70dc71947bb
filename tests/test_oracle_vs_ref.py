"""Pins the oracle port (oracle/sphinx_oracle.c) against the reference's own
compiled code (oracle/_ref, built from /root/reference by oracle/Makefile).
Skipped when oracle/_ref has not been built; the committed goldens in
tests/golden (test_golden.py) pin the same functions without it."""
import ctypes as C
import os

import numpy as np
import pytest

import orc
from cmusphinx_b200 import s3io, synth

pytestmark = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")


def test_logadd_table_matches_reference():
    for base, shift in [(1.0001, 10), (1.0001, 8), (1.0003, 0)]:
        n = orc.ref().ref_logadd_table(base, shift, None, 0)
        t = np.zeros(n, np.int32)
        orc.ref().ref_logadd_table(base, shift, orc._p(t, C.c_int32), n)
        p = orc.port_logadd_table(base, shift)
        assert p.size == n
        np.testing.assert_array_equal(p, t)
    assert list(orc.port_logadd_table()[:18]) == [7, 6, 6, 5, 5, 5, 4, 4, 4, 3, 3, 3, 3, 2, 2, 2, 2, 2]


def test_logmath_log_add():
    rng = np.random.default_rng(0)
    p = np.concatenate([10.0 ** rng.uniform(-150, 3, 2000), [0.0, 1.0, 42.0, 1e-150]])
    for shift in (0, 8, 10):
        out = np.zeros(p.size, np.int32)
        orc.ref().ref_logmath_log(1.0001, shift, p.ctypes.data_as(C.POINTER(C.c_double)), p.size,
                                  orc._p(out, C.c_int32))
        lm = orc.port.orc_logmath_init(1.0001, shift, 1)
        mine = np.array([orc.port.orc_logmath_log(lm, float(v)) for v in p], np.int32)
        np.testing.assert_array_equal(mine, out)
        x = rng.integers(-600000 >> shift, 100, 5000).astype(np.int32)
        y = rng.integers(-600000 >> shift, 100, 5000).astype(np.int32)
        x[:50] = -2 ** 31 >> (shift + 2)
        y[50:80] = (-2 ** 31 >> (shift + 2)) - 5
        ra = np.zeros(x.size, np.int32)
        orc.ref().ref_logmath_add(1.0001, shift, orc._p(x, C.c_int32), orc._p(y, C.c_int32), x.size,
                                  orc._p(ra, C.c_int32))
        pa = np.array([orc.port.orc_logmath_add(lm, int(a), int(b)) for a, b in zip(x, y)], np.int32)
        np.testing.assert_array_equal(pa, ra)
        orc.port.orc_logmath_free(lm)


def _ms_files(tmp_path, n_sen, n_density, dim, n_feat, seed):
    mean, var, mixw = synth.cont_model(n_sen, n_density, dim, seed, n_feat)
    var[0, 0, :3] = 1e-6   # exercises -varfloor
    mixw[1, 0, :2] = 0.0   # exercises -mixwfloor
    vl = [dim] * n_feat
    mfile, vfile, wfile = (str(tmp_path / n) for n in ("means", "variances", "mixture_weights"))
    streams_m = [mean.reshape(n_sen, n_density, n_feat, dim)[:, :, f, :] for f in range(n_feat)]
    streams_v = [var.reshape(n_sen, n_density, n_feat, dim)[:, :, f, :] for f in range(n_feat)]
    s3io.write_gauden(mfile, streams_m, vl)
    s3io.write_gauden(vfile, streams_v, vl)
    s3io.write_mixw(wfile, mixw)
    return mean, var, mixw, mfile, vfile, wfile, vl


@pytest.mark.parametrize("n_sen,n_density,dim,n_feat,topn", [(48, 8, 13, 1, 4), (40, 16, 7, 3, 4), (33, 8, 39, 1, 8),
                                                            (21, 32, 39, 1, 4), (16, 4, 5, 2, 1)])
def test_ms_backend_matches_reference(tmp_path, n_sen, n_density, dim, n_feat, topn):
    mean, var, mixw, mfile, vfile, wfile, vl = _ms_files(tmp_path, n_sen, n_density, dim, n_feat, 11)
    h = orc.ref().ref_ms_init(mfile.encode(), vfile.encode(), wfile.encode(), b".cont.", 1e-4, 1e-7, topn, 1, orc.LOGBASE)
    assert h
    dims = (C.c_int32 * 6)()
    orc.ref().ref_ms_dims(h, dims, None)
    assert list(dims)[:4] == [n_sen, n_feat, n_density, n_sen]
    tot = n_sen * n_density * dim * n_feat
    rmean, rvar = np.zeros(tot, np.float32), np.zeros(tot, np.float32)
    rdet = np.zeros(n_sen * n_feat * n_density, np.float32)
    rmixw = np.zeros(n_sen * n_feat * n_density, np.uint8)
    orc.ref().ref_ms_params(h, orc._p(rmean, C.c_float), orc._p(rvar, C.c_float), orc._p(rdet, C.c_float),
                            orc._p(rmixw, C.c_uint8))
    # load-time KAT: the port's precompute / quantiser on the same raw arrays
    var_ref_layout = synth.to_ref_layout(var, n_feat, dim)
    pv, pd = orc.port_precompute(var_ref_layout.reshape(-1, dim), dim)
    np.testing.assert_array_equal(pv.reshape(-1), rvar)
    np.testing.assert_array_equal(pd.reshape(-1), rdet)
    pq = orc.port_mixw_quantize(mixw)
    np.testing.assert_array_equal(pq.reshape(-1), rmixw)
    # scoring: dense and with an active list
    feat = synth.cont_features(mean, var, 37, 5)
    out_ref = np.zeros((37, n_sen), np.int16)
    orc.ref().ref_ms_eval_all(h, orc._p(feat, C.c_float), 37, orc._p(out_ref, C.c_int16))
    pm = orc.PortMs(n_sen, n_feat, vl, n_density, n_sen, topn, 1, rmean, rvar, rdet, rmixw, np.arange(n_sen))
    np.testing.assert_array_equal(pm.eval_all(feat), out_ref)
    rng = np.random.default_rng(3)
    for t in range(5):
        mask = np.zeros((n_sen + 31) // 32, np.uint32)
        for s in rng.choice(n_sen, n_sen // 3, replace=False):
            mask[s // 32] |= np.uint32(1 << (s % 32))
        deltas = orc.port_flags2list(mask, n_sen)
        a = np.full(n_sen, 12345, np.int16)
        b = a.copy()
        orc.ref().ref_ms_eval_active(h, orc._p(feat[t], C.c_float), orc._p(deltas, C.c_uint8), deltas.size, t,
                                     orc._p(a, C.c_int16))
        pm.frame_eval(feat[t], deltas, False, b)
        np.testing.assert_array_equal(a, b)
    orc.ref().ref_ms_free(h)


def test_tmat_quantiser_matches_reference(tmp_path):
    tp = synth.bakis_tmat(23, 3, 5)
    f = str(tmp_path / "tmat")
    s3io.write_tmat(f, tp)
    out = np.zeros(tp.size, np.uint8)
    ns = C.c_int32()
    n = orc.ref().ref_tmat_load(f.encode(), 1e-4, orc.LOGBASE, orc._p(out, C.c_uint8), out.size, C.byref(ns))
    assert n == 23 and ns.value == 3
    np.testing.assert_array_equal(orc.port_tmat_quantize(tp).reshape(-1), out)


@pytest.mark.parametrize("n_emit", [3, 5, 1, 2, 4])   # 1, 2, 4 -> hmm_vit_eval_anytopo (hmm.c:711-786)
def test_hmm_eval_matches_reference(n_emit):
    n_sen, n_tmat, n_sseq, n_hmm = 500, 17, 300, 4000
    tp = orc.port_tmat_quantize(synth.bakis_tmat(n_tmat, n_emit, 3))
    # a few "zero" skip arcs and dead self loops to trigger the guard paths
    pop = synth.hmm_population(n_hmm, n_emit, n_sen, n_tmat, n_sseq, seed=1, mpx_fraction=0.3)
    sen = synth.senscr_frames(6, n_sen, 2)
    a = {k: v.copy() for k, v in pop.items()}
    b = {k: v.copy() for k, v in pop.items()}
    for f in range(6):
        ra = orc.hmm_eval(orc.ref().ref_hmm_eval_batch, n_emit, tp, pop["sseq"], sen[f], a["score"], a["history"],
                          a["out_score"], a["out_history"], a["senid"], a["tmatid"], a["mpx"], a["bestscore"])
        rb = orc.hmm_eval(orc.port.orc_hmm_eval_batch, n_emit, tp, pop["sseq"], sen[f], b["score"], b["history"],
                          b["out_score"], b["out_history"], b["senid"], b["tmatid"], b["mpx"], b["bestscore"])
        assert ra == rb
        for k in ("score", "history", "out_score", "out_history", "senid", "bestscore"):
            np.testing.assert_array_equal(a[k], b[k], err_msg=f"{k} frame {f}")


def _real_feats(model, mfc, n=60):
    cep = orc.read_mfc(os.path.join(orc.DATA_DIR, "test", mfc))
    return model.cep2feat(cep)[:n]


def test_s2_semi_matches_reference():
    hmm = os.path.join(orc.DATA_DIR, "hmm", "hub4wsj_sc_8k")
    r = orc.RefAcmod(hmm)
    assert r.backend == "s2_semi"
    feat = _real_feats(r, "wsj/440c0201.mfc", 80)
    want = r.score(feat)
    from cmusphinx_b200 import engine
    g, v = engine.read_gauden(hmm + "/means"), engine.read_gauden(hmm + "/variances")
    pv, pd = orc.port_precompute(v["data"].reshape(-1, 13), 13)
    sd = engine.read_sendump(hmm + "/sendump", 3, 256, r.n_sen)
    pt = orc.PortTied(2, 1, 3, [13, 13, 13], 256, r.n_sen, 4, g["data"], pv, pd, sd["mixw"], sd["n_clust"],
                      sd["mixw_cb"], None)
    np.testing.assert_array_equal(pt.eval_all(feat), want)
    # active-list calls, fresh decoder state on both sides
    r.close()
    r = orc.RefAcmod(hmm)
    pt.reset()
    rng = np.random.default_rng(5)
    for t in range(12):
        mask = np.zeros((r.n_sen + 31) // 32, np.uint32)
        for s in rng.choice(r.n_sen, 900, replace=False):
            mask[s // 32] |= np.uint32(1 << (s % 32))
        d = orc.port_flags2list(mask, r.n_sen)
        np.testing.assert_array_equal(pt.frame_eval(feat[t], d, False, t), r.frame_eval(feat[t], d, t, False))
    r.close()


def test_ptm_matches_reference():
    hmm = os.path.join(orc.DATA_DIR, "hmm", "ptm")
    r = orc.RefAcmod(hmm)
    assert r.backend == "ptm"
    feat = _real_feats(r, "wsj/442c0201.mfc", 50)
    want = r.score(feat)
    from cmusphinx_b200 import engine
    g, v = engine.read_gauden(hmm + "/means"), engine.read_gauden(hmm + "/variances")
    assert g["n_mgau"] == 50 and g["veclen"] == [13, 13, 13]
    pv, pd = orc.port_precompute(v["data"].reshape(-1, 13), 13)
    sd = engine.read_sendump(hmm + "/sendump", 3, g["n_density"], r.n_sen)
    s2c = r.sen2cimap()
    pt = orc.PortTied(1, 50, 3, [13, 13, 13], g["n_density"], r.n_sen, 4, g["data"], pv, pd, sd["mixw"],
                      sd["n_clust"], sd["mixw_cb"], s2c)
    np.testing.assert_array_equal(pt.eval_all(feat), want)
    r.close()
    r = orc.RefAcmod(hmm)
    pt.reset()
    rng = np.random.default_rng(6)
    for t in range(10):
        mask = np.zeros((r.n_sen + 31) // 32, np.uint32)
        # senones of a handful of phones only -> codebook pruning is exercised
        cbs = rng.choice(50, 7, replace=False)
        for s in np.nonzero(np.isin(s2c, cbs))[0][::3]:
            mask[s // 32] |= np.uint32(1 << (s % 32))
        d = orc.port_flags2list(mask, r.n_sen)
        np.testing.assert_array_equal(pt.frame_eval(feat[t], d, False, t), r.frame_eval(feat[t], d, t, False))
    r.close()


def test_reference_ptm_reads_past_its_logadd_table():
    """Documents a defect of the reference that bounds what "parity" can mean for
    ptm: fast_logmath_add (tied_mgau_common.h:104-121) indexes the 256-entry
    table with |x-y| unchecked while ptm_mgau.c:267-288 normalises with the MIN
    of the top-1 scores (negative normalised scores).  On speech frames the
    reference reads past the table, so ITS OWN scores depend on heap history: a
    fresh decoder and one that already decoded another utterance disagree.  The
    oracle (and the GPU kernels) return the intended value (correction 0 beyond
    the table); they equal the reference wherever it stays inside its table."""
    hmm = os.path.join(orc.DATA_DIR, "hmm", "ptm")
    r = orc.RefAcmod(hmm)
    f440 = r.cep2feat(orc.read_mfc(os.path.join(orc.DATA_DIR, "test", "wsj", "440c0201.mfc")))
    f441 = r.cep2feat(orc.read_mfc(os.path.join(orc.DATA_DIR, "test", "wsj", "441c0201.mfc")))[:200]
    fresh = r.score(f441)
    s2c, n_sen = r.sen2cimap(), r.n_sen
    r.close()
    r = orc.RefAcmod(hmm)
    r.score(f440)
    after = r.score(f441)
    r.close()
    from cmusphinx_b200 import engine
    g, v = engine.read_gauden(hmm + "/means"), engine.read_gauden(hmm + "/variances")
    pv, pd = orc.port_precompute(v["data"].reshape(-1, 13), 13)
    sd = engine.read_sendump(hmm + "/sendump", 3, g["n_density"], n_sen)
    pt = orc.PortTied(1, 50, 3, [13, 13, 13], g["n_density"], n_sen, 4, g["data"], pv, pd, sd["mixw"], sd["n_clust"],
                      sd["mixw_cb"], s2c)
    port = pt.eval_all(f441)
    same_rows = (fresh == after).all(axis=1)
    # the leading (quiet) frames are inside the table for everyone ...
    assert same_rows[:100].all() and (port[:100] == fresh[:100]).all()
    # ... later the reference disagrees with itself; where it is self-consistent
    # AND inside its table it still equals the oracle on most rows
    if not same_rows.all():
        print(f"reference self-disagreement on {int((~same_rows).sum())} of {len(same_rows)} frames")


@pytest.mark.parametrize("beam,per_stream", [("20", [20, 20, 20]), ("10,40", [10, 40, 40]), ("5,0,60", [5, 0, 60]), ("1", [1, 1, 1])])
def test_s2_semi_topn_beam_matches_reference(beam, per_stream):
    """-topn_beam: mgau_norm cuts each stream's list at the first normalised score above the
    beam (s2_semi_mgau.c:189-207); split_topn fills missing streams with the largest value."""
    hmm = os.path.join(orc.DATA_DIR, "hmm", "hub4wsj_sc_8k")
    r = orc.RefAcmod(hmm, topn_beam=beam)
    feat = _real_feats(r, "wsj/440c0201.mfc", 60)
    want = r.score(feat)
    from cmusphinx_b200 import engine
    g, v = engine.read_gauden(hmm + "/means"), engine.read_gauden(hmm + "/variances")
    pv, pd = orc.port_precompute(v["data"].reshape(-1, 13), 13)
    sd = engine.read_sendump(hmm + "/sendump", 3, 256, r.n_sen)
    pt = orc.PortTied(2, 1, 3, [13, 13, 13], 256, r.n_sen, 4, g["data"], pv, pd, sd["mixw"], sd["n_clust"],
                      sd["mixw_cb"], None)
    pt.set_topn_beam(per_stream)
    np.testing.assert_array_equal(pt.eval_all(feat), want)
    r.close()


@pytest.mark.parametrize("ds,beam", [(2, 0), (3, 0), (4, 15), (7, 0)])
def test_s2_semi_frame_downsampling_matches_reference(ds, beam):
    """-ds: frames that are not a multiple of ds_ratio stop after eval_topn (s2_semi_mgau.c:176-186)."""
    hmm = os.path.join(orc.DATA_DIR, "hmm", "hub4wsj_sc_8k")
    r = orc.RefAcmod(hmm, ds=ds, topn_beam=str(beam) if beam else "")
    feat = _real_feats(r, "wsj/440c0201.mfc", 60)
    want = r.score(feat)
    from cmusphinx_b200 import engine
    g, v = engine.read_gauden(hmm + "/means"), engine.read_gauden(hmm + "/variances")
    pv, pd = orc.port_precompute(v["data"].reshape(-1, 13), 13)
    sd = engine.read_sendump(hmm + "/sendump", 3, 256, r.n_sen)
    pt = orc.PortTied(2, 1, 3, [13, 13, 13], 256, r.n_sen, 4, g["data"], pv, pd, sd["mixw"], sd["n_clust"],
                      sd["mixw_cb"], None)
    pt.set_topn_beam([beam] * 3)
    pt.set_ds(ds)
    np.testing.assert_array_equal(pt.eval_all(feat), want)
    r.close()


@pytest.mark.parametrize("topn", [1, 2, 3, 6, 8])
def test_tied_backends_other_topn_match_reference(topn):
    """-topn other than the default 4: the reference runs different unrolled scorers
    (get_scores_{4b,8b}_feat_1 .. _6 and _any, s2_semi_mgau.c:209-700; ptm loops over max_topn)."""
    from cmusphinx_b200 import engine
    hmm = os.path.join(orc.DATA_DIR, "hmm", "hub4wsj_sc_8k")
    r = orc.RefAcmod(hmm, topn=topn)
    feat = _real_feats(r, "wsj/440c0201.mfc", 30)
    g, v = engine.read_gauden(hmm + "/means"), engine.read_gauden(hmm + "/variances")
    pv, pd = orc.port_precompute(v["data"].reshape(-1, 13), 13)
    sd = engine.read_sendump(hmm + "/sendump", 3, 256, r.n_sen)
    pt = orc.PortTied(2, 1, 3, [13, 13, 13], 256, r.n_sen, topn, g["data"], pv, pd, sd["mixw"], sd["n_clust"],
                      sd["mixw_cb"], None)
    np.testing.assert_array_equal(pt.eval_all(feat), r.score(feat))
    r.close()
    hmm = os.path.join(orc.DATA_DIR, "hmm", "ptm")
    r = orc.RefAcmod(hmm, topn=topn)
    feat = _real_feats(r, "wsj/442c0201.mfc", 10)
    g, v = engine.read_gauden(hmm + "/means"), engine.read_gauden(hmm + "/variances")
    pv, pd = orc.port_precompute(v["data"].reshape(-1, 13), 13)
    sd = engine.read_sendump(hmm + "/sendump", 3, g["n_density"], r.n_sen)
    pt = orc.PortTied(1, 50, 3, [13, 13, 13], g["n_density"], r.n_sen, topn, g["data"], pv, pd, sd["mixw"],
                      sd["n_clust"], sd["mixw_cb"], r.sen2cimap())
    np.testing.assert_array_equal(pt.eval_all(feat), r.score(feat))
    r.close()


@pytest.mark.parametrize("aw", [2, 3])
def test_ms_backend_acoustic_weight_matches_reference(tmp_path, aw):
    """-aw (ms_mgau.c:105, senone_eval's `scr /= aw`, ms_senone.c:408-409) on a 2-stream model."""
    n_sen, n_density, dim, n_feat, topn = 40, 16, 7, 2, 3
    mean, var, mixw, mfile, vfile, wfile, vl = _ms_files(tmp_path, n_sen, n_density, dim, n_feat, 23)
    h = orc.ref().ref_ms_init(mfile.encode(), vfile.encode(), wfile.encode(), b".cont.", 1e-4, 1e-7, topn, aw, orc.LOGBASE)
    assert h
    tot = n_sen * n_density * dim * n_feat
    rmean, rvar = np.zeros(tot, np.float32), np.zeros(tot, np.float32)
    rdet = np.zeros(n_sen * n_feat * n_density, np.float32)
    rmixw = np.zeros(n_sen * n_feat * n_density, np.uint8)
    orc.ref().ref_ms_params(h, orc._p(rmean, C.c_float), orc._p(rvar, C.c_float), orc._p(rdet, C.c_float),
                            orc._p(rmixw, C.c_uint8))
    feat = synth.cont_features(mean, var, 25, 6)
    out_ref = np.zeros((25, n_sen), np.int16)
    orc.ref().ref_ms_eval_all(h, orc._p(feat, C.c_float), 25, orc._p(out_ref, C.c_int16))
    pm = orc.PortMs(n_sen, n_feat, vl, n_density, n_sen, topn, aw, rmean, rvar, rdet, rmixw, np.arange(n_sen))
    np.testing.assert_array_equal(pm.eval_all(feat), out_ref)
    orc.ref().ref_ms_free(h)
