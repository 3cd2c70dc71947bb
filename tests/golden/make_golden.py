#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref: the
unmodified reference sources compiled by oracle/Makefile, driven through
oracle/ref_shim.c).  Run in the build container (needs /root/reference for
`make -C oracle ref`):   python tests/golden/make_golden.py
The .npz files are committed; tests never need /root/reference.
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import orc  # noqa: E402
from cmusphinx_b200 import s3io, synth  # noqa: E402

R = orc.ref()
LB = orc.LOGBASE


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **kw)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def logmath():
    n = R.ref_logadd_table(LB, 10, None, 0)
    tab = np.zeros(n, np.int32)
    R.ref_logadd_table(LB, 10, orc._p(tab, C.c_int32), n)
    rng = np.random.default_rng(0)
    p = np.concatenate([10.0 ** rng.uniform(-150, 3, 500), [0.0, 1.0, 42.0, 1e-150]])
    logs = {}
    for sh in (0, 10):
        o = np.zeros(p.size, np.int32)
        R.ref_logmath_log(LB, sh, p.ctypes.data_as(C.POINTER(C.c_double)), p.size, orc._p(o, C.c_int32))
        logs[f"log_shift{sh}"] = o
    x = rng.integers(-600, 100, 2000).astype(np.int32)
    y = rng.integers(-600, 100, 2000).astype(np.int32)
    x[:20] = -2 ** 31 >> 12
    add = np.zeros(x.size, np.int32)
    R.ref_logmath_add(LB, 10, orc._p(x, C.c_int32), orc._p(y, C.c_int32), x.size, orc._p(add, C.c_int32))
    save("logmath.npz", base=LB, table10=tab, p=p, add_x=x, add_y=y, add_out=add, **logs)


def ms_case(name, n_sen, n_density, dim, n_feat, topn, T, seed):
    mean, var, mixw = synth.cont_model(n_sen, n_density, dim, seed, n_feat)
    var[0, 0, :3] = 1e-6
    mixw[1, 0, :2] = 0.0
    vl = [dim] * n_feat
    with tempfile.TemporaryDirectory() as d:
        mf, vf, wf = (os.path.join(d, n) for n in ("means", "variances", "mixture_weights"))
        s3io.write_gauden(mf, [mean.reshape(n_sen, n_density, n_feat, dim)[:, :, f, :] for f in range(n_feat)], vl)
        s3io.write_gauden(vf, [var.reshape(n_sen, n_density, n_feat, dim)[:, :, f, :] for f in range(n_feat)], vl)
        s3io.write_mixw(wf, mixw)
        h = R.ref_ms_init(mf.encode(), vf.encode(), wf.encode(), b".cont.", 1e-4, 1e-7, topn, 1, LB)
    tot = mean.size
    rmean, rvar = np.zeros(tot, np.float32), np.zeros(tot, np.float32)
    rdet = np.zeros(n_sen * n_feat * n_density, np.float32)
    rmixw = np.zeros(n_sen * n_feat * n_density, np.uint8)
    R.ref_ms_params(h, orc._p(rmean, C.c_float), orc._p(rvar, C.c_float), orc._p(rdet, C.c_float),
                    orc._p(rmixw, C.c_uint8))
    feat = synth.cont_features(mean, var, T, seed + 1)
    dense = np.zeros((T, n_sen), np.int16)
    R.ref_ms_eval_all(h, orc._p(feat, C.c_float), T, orc._p(dense, C.c_int16))
    rng = np.random.default_rng(seed + 2)
    act_deltas, act_scores = [], []
    for t in range(min(T, 6)):
        mask = np.zeros((n_sen + 31) // 32, np.uint32)
        for s in rng.choice(n_sen, max(1, n_sen // 3), replace=False):
            mask[s // 32] |= np.uint32(1 << (s % 32))
        dl = orc.port_flags2list(mask, n_sen)
        o = np.full(n_sen, 12345, np.int16)
        R.ref_ms_eval_active(h, orc._p(feat[t], C.c_float), orc._p(dl, C.c_uint8), dl.size, t, orc._p(o, C.c_int16))
        pad = np.zeros(2 * n_sen, np.uint8)
        pad[:dl.size] = dl
        act_deltas.append(np.concatenate([[dl.size], pad]).astype(np.int32))
        act_scores.append(o)
    R.ref_ms_free(h)
    save(name, dims=np.array([n_sen, n_density, dim, n_feat, topn, T]), mean_raw=synth.to_ref_layout(mean, n_feat, dim),
         var_raw=synth.to_ref_layout(var, n_feat, dim), mixw_raw=mixw, mean=rmean, var=rvar, det=rdet, mixw=rmixw,
         feat=feat, dense=dense, act_deltas=np.array(act_deltas), act_scores=np.array(act_scores))


def tmat_hmm():
    tp_raw = synth.bakis_tmat(23, 3, 5)
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "tmat")
        s3io.write_tmat(f, tp_raw)
        out = np.zeros(tp_raw.size, np.uint8)
        ns = C.c_int32()
        R.ref_tmat_load(f.encode(), 1e-4, LB, orc._p(out, C.c_uint8), out.size, C.byref(ns))
    keep = dict(tp_raw=tp_raw, tp=out.reshape(tp_raw.shape))
    for ne in (3, 5):
        n_sen, n_tmat, n_sseq, n_hmm, nfr = 400, 13, 200, 1500, 4
        tp = orc.port_tmat_quantize(synth.bakis_tmat(n_tmat, ne, 3))
        pop = synth.hmm_population(n_hmm, ne, n_sen, n_tmat, n_sseq, seed=ne, mpx_fraction=0.3)
        sen = synth.senscr_frames(nfr, n_sen, ne + 10)
        a = {k: v.copy() for k, v in pop.items()}
        bests = []
        for f in range(nfr):
            bests.append(orc.hmm_eval(R.ref_hmm_eval_batch, ne, tp, pop["sseq"], sen[f], a["score"], a["history"],
                                      a["out_score"], a["out_history"], a["senid"], a["tmatid"], a["mpx"],
                                      a["bestscore"]))
        for k, v in pop.items():
            keep[f"h{ne}_in_{k}"] = v
        for k in ("score", "history", "out_score", "out_history", "senid", "bestscore"):
            keep[f"h{ne}_out_{k}"] = a[k]
        keep[f"h{ne}_tp"] = tp
        keep[f"h{ne}_senscr"] = sen
        keep[f"h{ne}_best"] = np.array(bests, np.int32)
    save("tmat_hmm.npz", **keep)


def hmm_anytopo():
    """hmm_vit_eval for n_emit_state 1, 2 and 4, i.e. hmm_vit_eval_anytopo (hmm.c:711-786), run by
    the reference on synthetic populations (30 % multiplex HMMs, some with BAD_SSID states)."""
    keep = {}
    for ne in (1, 2, 4):
        n_sen, n_tmat, n_sseq, n_hmm, nfr = 400, 13, 200, 1500, 4
        tp = orc.port_tmat_quantize(synth.bakis_tmat(n_tmat, ne, 3))
        pop = synth.hmm_population(n_hmm, ne, n_sen, n_tmat, n_sseq, seed=ne, mpx_fraction=0.3)
        sen = synth.senscr_frames(nfr, n_sen, ne + 10)
        a = {k: v.copy() for k, v in pop.items()}
        bests = []
        for f in range(nfr):
            bests.append(orc.hmm_eval(R.ref_hmm_eval_batch, ne, tp, pop["sseq"], sen[f], a["score"], a["history"],
                                      a["out_score"], a["out_history"], a["senid"], a["tmatid"], a["mpx"],
                                      a["bestscore"]))
        for k, v in pop.items():
            keep[f"h{ne}_in_{k}"] = v
        for k in ("score", "history", "out_score", "out_history", "senid", "bestscore"):
            keep[f"h{ne}_out_{k}"] = a[k]
        keep[f"h{ne}_tp"] = tp
        keep[f"h{ne}_senscr"] = sen
        keep[f"h{ne}_best"] = np.array(bests, np.int32)
    save("hmm_anytopo.npz", **keep)


def real_model(name, hmmdir, mfc, n_frames, senmgau="", topn=4):
    r = orc.RefAcmod(os.path.join(orc.DATA_DIR, "hmm", hmmdir), senmgau, topn)
    cep = orc.read_mfc(os.path.join(orc.DATA_DIR, "test", mfc))
    feat = r.cep2feat(cep)[:n_frames]
    dense = r.score(feat)
    extra = {}
    if r.backend == "ptm":
        extra["sen2cb"] = r.sen2cimap()
    tp, sseq = r.tables()
    r.close()
    # active-list calls on a fresh decoder
    r = orc.RefAcmod(os.path.join(orc.DATA_DIR, "hmm", hmmdir), senmgau, topn)
    rng = np.random.default_rng(5)
    dl_all, sc_all = [], []
    for t in range(min(n_frames, 8)):
        mask = np.zeros((r.n_sen + 31) // 32, np.uint32)
        if r.backend == "ptm":
            cbs = rng.choice(50, 7, replace=False)
            sel = np.nonzero(np.isin(extra["sen2cb"], cbs))[0][::3]
        else:
            sel = rng.choice(r.n_sen, 700, replace=False)
        for s in sel:
            mask[s // 32] |= np.uint32(1 << (s % 32))
        dl = orc.port_flags2list(mask, r.n_sen)
        o = np.full(r.n_sen, 12345, np.int16)
        r.frame_eval(feat[t], dl, t, False, o)
        pad = np.zeros(2 * r.n_sen, np.uint8)
        pad[:dl.size] = dl
        dl_all.append(np.concatenate([[dl.size], pad]).astype(np.int32))
        sc_all.append(o)
    save(name, backend=r.backend, n_sen=r.n_sen, streamlen=np.array(r.streamlen), feat=feat, dense=dense,
         act_deltas=np.array(dl_all), act_scores=np.array(sc_all), tp=tp, topn=topn, **extra)
    r.close()


S3_CFGS = [
    dict(ci_pbeam=1e-80, max_cd=100000, ds_ratio=1, tighten=0.5),
    dict(ci_pbeam=1e-40, max_cd=100000, ds_ratio=1, tighten=0.5),
    dict(ci_pbeam=1e-40, max_cd=60, ds_ratio=1, tighten=0.5),
    dict(ci_pbeam=1e-30, max_cd=100000, ds_ratio=3, tighten=0.5),
    dict(ci_pbeam=1e-40, max_cd=80, ds_ratio=2, tighten=0.3),
]


def s3_case(name="s3_synth.npz", T=40):
    """sphinx3 flavour: the reference's mgau_init / mgau_eval /
    approx_cont_mgau_frame_eval (oracle/_ref/libs3am.so via ref_shim_s3.c) on
    the seeded synthetic CI/CD model of synth.s3_model()."""
    mean, var, mixw, cd2ci, n_ci = synth.s3_model(n_sen=160, n_ci_sen=16)
    D = mean.shape[2]
    feat = synth.s3_features(mean, var, T)
    act = synth.s3_active(mean.shape[0], n_ci, T)
    stale0 = (np.arange(mean.shape[0]) * 7 - 1000).astype(np.int32)
    with tempfile.TemporaryDirectory() as d:
        mf, vf, wf = (os.path.join(d, n) for n in ("means", "variances", "mixture_weights"))
        s3io.write_gauden(mf, mean, [D]); s3io.write_gauden(vf, var, [D]); s3io.write_mixw(wf, mixw[:, None, :])
        r = orc.RefS3(mf, vf, wf, None, cd2ci, n_ci)
    nc, pm, pv, lrd, pw, scal = r.params()
    dense = r.eval_dense(feat)
    out = {}
    for i, cfg in enumerate(S3_CFGS):
        r.set_fast(**cfg); r.utt_reset()
        o, best, a = r.eval_utt(feat, act, 3, stale0)
        bi, ut = r.state()
        out.update({f"scr{i}": o, f"best{i}": best, f"act{i}": a, f"bstidx{i}": bi, f"upd{i}": ut})
    save(name, mean=mean, var=var, mixw=mixw, cd2ci=cd2ci, n_ci=n_ci, feat=feat, act=act, stale0=stale0, frame0=3,
         n_comp=nc, p_var_head=pv[:20], p_lrd=lrd, p_mixw=pw, scal=scal, dense=dense, ci_pbeam_default=r.ci_pbeam,
         cfgs=np.array([[c["ci_pbeam"], c["max_cd"], c["ds_ratio"], c["tighten"]] for c in S3_CFGS]), **out)
    r.free()


S3_SVQ_CFGS = [   # (n_sv, vqeval, subvqbeam, fast-GMM settings)
    (3, 3, 1e-3, dict()),
    (3, 2, 1e-2, dict(ci_pbeam=1e-40, max_cd=60)),
    (1, 3, 1e-3, dict(ci_pbeam=1e-30, ds_ratio=3)),
]


def s3_svq_case(name="s3_svq.npz", T=30):
    """sphinx3 sub-vector quantised shortlists: the reference's subvq_init / subvq_gautbl_eval_logs3 /
    approx_cont_mgau_frame_eval(svq) on the synthetic model of s3_case with the synthetic sub-VQ
    models of orc.synthetic_subvq (the .subvq text is regenerated from the seeds by the test)."""
    mean, var, mixw, cd2ci, n_ci = synth.s3_model(n_sen=160, n_ci_sen=16)
    D = mean.shape[2]
    feat = synth.s3_features(mean, var, T)
    act = synth.s3_active(mean.shape[0], n_ci, T)
    valid = ~np.all(var == 0, axis=2)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        mf, vf, wf = (os.path.join(d, n) for n in ("means", "variances", "mixture_weights"))
        s3io.write_gauden(mf, mean, [D]); s3io.write_gauden(vf, var, [D]); s3io.write_mixw(wf, mixw[:, None, :])
        for i, (n_sv, vqeval, beam, cfg) in enumerate(S3_SVQ_CFGS):
            q = orc.synthetic_subvq(mean, var, valid, n_sv, 16)
            path = os.path.join(d, "m%d.subvq" % i)
            orc.write_subvq(path, q)
            r = orc.RefS3(mf, vf, wf, None, cd2ci, n_ci)
            orc.ref_set_svq(r, path, vqeval=vqeval, subvqbeam=beam)
            r.set_fast(**cfg); r.utt_reset()
            vq = np.zeros((T, r.svq_dims[0] * 16), np.int32)
            orc.ref_s3().ref_s3_svq_vqdist(r.h, orc._p(feat, orc.C.c_float), T, orc._p(vq, orc.C.c_int32))
            o, best, a = r.eval_utt(feat, act, 1)
            bi, ut = r.state()
            out.update({f"vq{i}": vq, f"scr{i}": o, f"best{i}": best, f"act{i}": a, f"bstidx{i}": bi, f"upd{i}": ut})
            r.free()
        # Gaussian selector (gs.c) on the same model
        cw, bits = orc.synthetic_gs(mean, 32)
        path = os.path.join(d, "m.gs")
        orc.write_gs(path, cw, bits, mean.shape[1])
        r = orc.RefS3(mf, vf, wf, None, cd2ci, n_ci)
        orc.ref_set_gs(r, path)
        r.set_fast(ci_pbeam=1e-40, max_cd=60); r.utt_reset()
        closest = np.zeros(T, np.int32)
        orc.ref_s3().ref_s3_gs_closest(r.h, orc._p(feat, orc.C.c_float), T, orc._p(closest, orc.C.c_int32))
        o, best, a = r.eval_utt(feat, act, 1)
        bi, ut = r.state()
        out.update(gs_closest=closest, gs_scr=o, gs_best=best, gs_act=a, gs_bstidx=bi, gs_upd=ut)
        r.free()
    save(name, n_sen=160, n_ci=n_ci, feat=feat, act=act, frame0=1, **out)


def semi_beam():
    """s2_semi with -topn_beam (mgau_norm's list cut, s2_semi_mgau.c:189-207): the reference's
    dense scores on the frames of semi_hub4wsj.npz for two beam settings."""
    g = np.load(os.path.join(HERE, "semi_hub4wsj.npz"))
    out = {}
    for i, beam in enumerate(("20", "10,40")):
        r = orc.RefAcmod(os.path.join(orc.DATA_DIR, "hmm", "hub4wsj_sc_8k"), topn_beam=beam)
        out[f"dense{i}"] = r.score(g["feat"])
        r.close()
    # -ds 2 and -ds 3 (+ a beam): skipped frames re-score the previous frame's codewords (s2_semi_mgau.c:176-186)
    for i, (ds, beam) in enumerate(((2, ""), (3, "15"))):
        r = orc.RefAcmod(os.path.join(orc.DATA_DIR, "hmm", "hub4wsj_sc_8k"), ds=ds, topn_beam=beam)
        out[f"ds_dense{i}"] = r.score(g["feat"])
        r.close()
    save("semi_hub4wsj_beam.npz", beams=np.array([[20, 20, 20], [10, 40, 40]], np.int32),
         ds_cfg=np.array([[2, 0], [3, 15]], np.int32), **out)


def feat_unit_test():
    """The reference's own feature unit test (sphinxbase/test/unit/test_feat/test_feat.c): its 6 x 13
    input cepstra (parsed out of the C source) and the expected printout _test_feat.res, 6 lines each
    for -feat "13", "13:1" and "1s_c_d_dd" (cmn none, agc none), values printed with %.3f."""
    import re
    d = "/root/reference/sphinxbase/test/unit/test_feat"
    src = open(os.path.join(d, "test_feat.c")).read()
    body = src[src.index("const mfcc_t data[6][13]"):src.index("};") + 2]
    vals = [float(v) for v in re.findall(r"FLOAT2MFCC\(([-0-9.]+)\)", body)]
    cep = np.array(vals, np.float32).reshape(6, 13)
    lines = [l.split() for l in open(os.path.join(d, "_test_feat.res")).read().strip().splitlines()]
    assert len(lines) == 18
    save("feat_unit_test.npz", cep=cep, res_13=np.array(lines[0:6], np.float64), res_13_1=np.array(lines[6:12], np.float64),
         res_1s_c_d_dd=np.array(lines[12:18], np.float64))


def feat_general():
    """General feature stage: the reference's own feat_t (ref_feat_compute in oracle/ref_shim.c)
    on 80 frames of test/data/wsj/442c0201.mfc for the configurations of cases.FEAT_GOLDEN_CASES."""
    import cases
    cep = orc.read_mfc(os.path.join(orc.DATA_DIR, "test", "wsj", "442c0201.mfc"))[40:120].copy()
    lda = np.random.default_rng(77).standard_normal((52, 52)).astype(np.float32)
    klen = {"1s_c_d_dd": 39, "s3_1x39": 39, "s2_4x": 51, "1s_c_d_ld_dd": 52, "1s_c": 13, "1s_c_d": 26}
    outs = {}
    for i, (ftype, cmn, vn, agc, use_lda, dim, sv) in enumerate(cases.FEAT_GOLDEN_CASES):
        outs[f"out{i}"] = orc.ref_feat_compute(cep, ftype, cmn, vn, agc, lda[:klen[ftype], :klen[ftype]] if use_lda else None, dim, sv)
    save("feat_general.npz", cep=cep, lda=lda, **outs)


if __name__ == "__main__":
    assert orc.have_ref(), "build oracle/_ref first: make -C oracle ref"
    if len(sys.argv) > 1:      # regenerate only the named fixtures, e.g. `make_golden.py feat_general`
        for fn in sys.argv[1:]:
            globals()[fn]()
        sys.exit(0)
    logmath()
    ms_case("ms_small.npz", 48, 8, 13, 1, 4, 33, 11)
    ms_case("ms_3stream.npz", 40, 16, 7, 3, 4, 21, 12)
    ms_case("ms_allden.npz", 33, 8, 39, 1, 8, 17, 13)
    ms_case("ms_cont32.npz", 64, 32, 39, 1, 4, 130, 14)
    tmat_hmm()
    real_model("semi_hub4wsj.npz", "hub4wsj_sc_8k", "wsj/440c0201.mfc", 24)
    real_model("ptm_hub4wsj.npz", "ptm", "wsj/442c0201.mfc", 16)
    real_model("cont_hub4_topn4.npz", "cont", "pittsburgh.littleendian.mfc", 20, ".cont.", 4)
    real_model("cont_hub4_topn8.npz", "cont", "pittsburgh.littleendian.mfc", 12, ".cont.", 8)
    s3_case()
    s3_svq_case()
