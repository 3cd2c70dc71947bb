"""The reference-side binding (cmusphinx_b200/plugin/b200_mgau.c, LD_PRELOADed
into the UNMODIFIED reference decoder built under oracle/_ref): identical
1-best hypotheses on the bundled regression utterances (BASELINE configs[0]).
"""
import os
import subprocess

import pytest

import orc

PLUGIN = os.path.join(orc.ROOT, "cmusphinx_b200", "_plugin", "libb200_ps_plugin.so")
BATCH = os.path.join(orc.REF_DIR, "pocketsphinx_batch")
D = orc.DATA_DIR

needs = pytest.mark.skipif(not (os.path.exists(PLUGIN) and os.path.exists(BATCH)),
                           reason="oracle/_ref or the plug-in not built")


def _decode(tmp_path, tag, ctl_lines, cepdir, cepext, extra, env_extra, hmm="hub4wsj_sc_8k"):
    ctl = tmp_path / f"{tag}.ctl"
    ctl.write_text("\n".join(ctl_lines) + "\n")
    hyp = tmp_path / f"{tag}.hyp"
    cmd = [BATCH, "-hmm", os.path.join(D, "hmm", hmm), "-lm", os.path.join(D, "lm", "wsj0vp.5000.DMP"),
           "-dict", os.path.join(D, "lm", "cmu07a.dic"), "-ctl", str(ctl), "-cepdir", cepdir, "-cepext", cepext,
           "-hyp", str(hyp), "-logfn", str(tmp_path / f"{tag}.log")] + extra
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = orc.REF_DIR + ":" + env.get("LD_LIBRARY_PATH", "")
    env.update(env_extra)
    subprocess.run(cmd, env=env, check=True, timeout=900, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return hyp.read_text().strip().splitlines(), (tmp_path / f"{tag}.log").read_text(errors="replace")


RAW = (["goforward", "numbers", "something"], os.path.join(D, "test"), ".raw", ["-adcin", "yes", "-samprate", "16000"])
MFC = (["440c0201", "442c0201"], os.path.join(D, "test", "wsj"), ".mfc", [])


@needs
def test_plugin_passthrough_is_transparent(tmp_path):
    """With B200_PLUGIN_DISABLE=1 the interposed constructors forward to the
    reference's own: proves the LD_PRELOAD hook is where acmod_init_am looks."""
    ref, _ = _decode(tmp_path, "ref", RAW[0][:1], RAW[1], RAW[2], RAW[3] + ["-fwdflat", "no", "-bestpath", "no"], {})
    got, _ = _decode(tmp_path, "pass", RAW[0][:1], RAW[1], RAW[2], RAW[3] + ["-fwdflat", "no", "-bestpath", "no"],
                     {"LD_PRELOAD": PLUGIN, "B200_PLUGIN_DISABLE": "1"})
    assert ref == got
    assert ref[0].startswith("go forward ten years")   # pocketsphinx test_ps_fwdtree.c:24


@pytest.mark.gpu
@needs
@pytest.mark.parametrize("passes", [["-fwdflat", "no", "-bestpath", "no"], []])
@pytest.mark.parametrize("inp", [RAW, MFC], ids=["raw", "mfc"])
def test_identical_hypotheses_semi_continuous(tmp_path, passes, inp):
    utts, cepdir, ext, extra = inp
    ref, _ = _decode(tmp_path, "ref", utts, cepdir, ext, extra + passes, {})
    got, log = _decode(tmp_path, "gpu", utts, cepdir, ext, extra + passes, {"LD_PRELOAD": PLUGIN})
    assert "b200_semi back-end on GPU" in log
    assert [l.rsplit("(", 1)[0] for l in got] == [l.rsplit("(", 1)[0] for l in ref]
    # path scores: the s2_semi kernels are bit-exact except on integer ties
    assert got == ref
    if not passes and inp is RAW:
        assert ref[0].startswith("go forward and users")  # test_ps_simple.c:26


@pytest.mark.gpu
@needs
@pytest.mark.parametrize("opts", [["-ds", "2"], ["-topn_beam", "20"], ["-ds", "3", "-topn_beam", "30,20,40"]],
                         ids=["ds2", "beam20", "ds3+beams"])
def test_identical_hypotheses_semi_with_ds_and_topn_beam(tmp_path, opts):
    """The s2_semi fast-evaluation options reach the device back-end through the plug-in:
    -ds (s2_semi_mgau.c:176-186) and -topn_beam (:189-207); default 3-pass decode, so the
    second pass re-scores from frame 0 after acmod_rewind."""
    utts, cepdir, ext, extra = MFC
    ref, _ = _decode(tmp_path, "ref", utts, cepdir, ext, extra + opts, {})
    got, log = _decode(tmp_path, "gpu", utts, cepdir, ext, extra + opts, {"LD_PRELOAD": PLUGIN})
    assert "b200_semi back-end on GPU" in log
    assert got == ref


@pytest.mark.gpu
@needs
def test_identical_hypotheses_ptm_and_compallsen(tmp_path):
    utts, cepdir, ext, extra = MFC
    for more in ([], ["-compallsen", "yes"]):
        ref, _ = _decode(tmp_path, "ref", utts[1:], cepdir, ext, extra + more, {}, hmm="ptm")
        got, log = _decode(tmp_path, "gpu", utts[1:], cepdir, ext, extra + more, {"LD_PRELOAD": PLUGIN}, hmm="ptm")
        assert "b200_ptm back-end on GPU" in log
        assert [l.rsplit("(", 1)[0] for l in got] == [l.rsplit("(", 1)[0] for l in ref]
    assert "bids totaling five hundred twenty five" in ref[0]


@needs
@pytest.mark.gpu
@pytest.mark.parametrize("passes", [[], ["-fwdflat", "no", "-bestpath", "no"]], ids=["3pass", "fwdtree"])
def test_identical_hypotheses_fully_continuous(tmp_path, passes):
    """hub4_cd_continuous_8gau_1s_c_d_dd (6144 senones x 8 Gaussians x 39 dims) goes
    through the reference's generic ms back-end; the plug-in serves it from the
    exact kernels (path 0) and from the tcgen05 Mahalanobis GEMM (path 1, the
    default; bit-identical senone scores since round 2): identical words AND
    identical path score on both."""
    an4 = os.path.join(D, "lm", "an4")
    if not os.path.exists(os.path.join(an4, "an4.dict")):
        pytest.skip("an4 LM/dictionary not copied (make -C oracle ref)")
    ctl = tmp_path / "c.ctl"
    ctl.write_text("pittsburgh.littleendian\n")

    def run(tag, env_extra):
        hyp = tmp_path / f"{tag}.hyp"
        cmd = [BATCH, "-hmm", os.path.join(D, "hmm", "cont"), "-lm", os.path.join(an4, "an4.ug.lm.DMP"), "-dict",
               os.path.join(an4, "an4.dict"), "-fdict", os.path.join(an4, "filler.dict"), "-ctl", str(ctl), "-cepdir",
               os.path.join(D, "test"), "-cepext", ".mfc", "-hyp", str(hyp), "-logfn", str(tmp_path / f"{tag}.log")] + passes
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = orc.REF_DIR + ":" + env.get("LD_LIBRARY_PATH", "")
        env.update(env_extra)
        subprocess.run(cmd, env=env, check=True, timeout=900, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        line = hyp.read_text().strip()
        words, rest = line.rsplit("(", 1)
        return words.strip(), int(rest.split()[-1].rstrip(")")), (tmp_path / f"{tag}.log").read_text(errors="replace")

    w_cpu, s_cpu, _ = run("cpu", {})
    w_ex, s_ex, log = run("exact", {"LD_PRELOAD": PLUGIN, "B200_MS_PATH": "0"})
    assert "b200" in log.lower()
    w_tc, s_tc, _ = run("tc", {"LD_PRELOAD": PLUGIN, "B200_MS_PATH": "1"})
    if not passes:
        assert w_cpu.split() == "P I T T S B U R G H".split()
    assert (w_ex, s_ex) == (w_cpu, s_cpu)
    assert (w_tc, s_tc) == (w_cpu, s_cpu)
    w_def, s_def, _ = run("default", {"LD_PRELOAD": PLUGIN})
    assert (w_def, s_def) == (w_cpu, s_cpu)


def _write_mllr(path, veclens, seed):
    """An MLLR file in the text format ps_mllr_read parses (ps_mllr.c:55-125): one class,
    A = I + small rotation, a bias and a variance scale per stream."""
    import numpy as np
    rng = np.random.default_rng(seed)
    with open(path, "w") as fh:
        fh.write(f"1\n{len(veclens)}\n")
        for n in veclens:
            A = np.eye(n) + 0.03 * rng.standard_normal((n, n))
            fh.write(f"{n}\n")
            for row in A:
                fh.write(" ".join(f"{v:.6f}" for v in row) + " \n")
            fh.write(" ".join(f"{v:.6f}" for v in 0.2 * rng.standard_normal(n)) + " \n")
            fh.write(" ".join(f"{v:.6f}" for v in 1.0 + 0.1 * rng.random(n)) + " \n")


@pytest.mark.gpu
@needs
def test_mllr_transform_through_the_vtable(tmp_path):
    """ps_mgaufuncs_t.transform (acmod_update_mllr, acmod.c:337-346 -> gauden_mllr_transform,
    ms_gauden.c:551-605): with -mllr the plug-in re-reads, transforms, re-precomputes and
    re-uploads the Gaussians.  ptm: identical words (and the transform does change the
    scores); fully continuous on the exact kernels: identical path score."""
    utts, cepdir, ext, extra = MFC
    mllr = tmp_path / "ptm.mllr"
    _write_mllr(mllr, [13, 13, 13], 1)
    opts = extra + ["-mllr", str(mllr)]
    plain, _ = _decode(tmp_path, "plain", utts[1:], cepdir, ext, extra, {}, hmm="ptm")
    ref, _ = _decode(tmp_path, "ref", utts[1:], cepdir, ext, opts, {}, hmm="ptm")
    got, log = _decode(tmp_path, "gpu", utts[1:], cepdir, ext, opts, {"LD_PRELOAD": PLUGIN}, hmm="ptm")
    assert "b200_ptm back-end on GPU" in log
    assert ref != plain                                   # the transform is not a no-op
    assert [l.rsplit("(", 1)[0] for l in got] == [l.rsplit("(", 1)[0] for l in ref]

    an4 = os.path.join(D, "lm", "an4")
    if not os.path.exists(os.path.join(an4, "an4.dict")):
        pytest.skip("an4 LM/dictionary not copied (make -C oracle ref)")
    mllr = tmp_path / "cont.mllr"
    _write_mllr(mllr, [39], 2)
    ctl = tmp_path / "c.ctl"
    ctl.write_text("pittsburgh.littleendian\n")

    def run(tag, env_extra, more):
        hyp = tmp_path / f"{tag}.hyp"
        cmd = [BATCH, "-hmm", os.path.join(D, "hmm", "cont"), "-lm", os.path.join(an4, "an4.ug.lm.DMP"), "-dict",
               os.path.join(an4, "an4.dict"), "-fdict", os.path.join(an4, "filler.dict"), "-ctl", str(ctl), "-cepdir",
               os.path.join(D, "test"), "-cepext", ".mfc", "-hyp", str(hyp), "-logfn", str(tmp_path / f"{tag}.log")] + more
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = orc.REF_DIR + ":" + env.get("LD_LIBRARY_PATH", "")
        env.update(env_extra)
        subprocess.run(cmd, env=env, check=True, timeout=900, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return hyp.read_text().strip()

    base = run("c0", {}, [])
    want = run("c1", {}, ["-mllr", str(mllr)])
    got = run("c2", {"LD_PRELOAD": PLUGIN, "B200_MS_PATH": "0"}, ["-mllr", str(mllr)])
    assert want != base
    assert got == want


@pytest.mark.gpu
@needs
def test_s2_semi_mllr_is_the_references_no_op(tmp_path):
    """The reference's s2_semi back-end transforms s->g but keeps scoring with the aliases
    taken at init (s2_semi_mgau.c:1267-1269, 1338-1343): -mllr does not change its scores.
    The plug-in mirrors that (B200_SEMI_MLLR=1 applies the intended transform instead)."""
    if not os.path.exists(os.path.join(TIDIGITS, "tidigits.ctl")):
        pytest.skip("tidigits fixtures not copied (make -C oracle ref)")
    mllr = os.path.join(D, "test", "wsj", "s1.mllr")      # the bundled 4-stream (12/24/3/12) transform

    def run(tag, env_extra, more):
        hyp = tmp_path / f"{tag}.hyp"
        cmd = [BATCH, "-hmm", os.path.join(D, "hmm", "tidigits"), "-lm", os.path.join(D, "lm", "tidigits.DMP"), "-dict",
               os.path.join(D, "lm", "tidigits.dic"), "-ctl", os.path.join(TIDIGITS, "tidigits.ctl"), "-ctlcount", "8",
               "-cepdir", TIDIGITS, "-hyp", str(hyp), "-logfn", str(tmp_path / f"{tag}.log")] + more
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = orc.REF_DIR + ":" + env.get("LD_LIBRARY_PATH", "")
        env.update(env_extra)
        subprocess.run(cmd, env=env, check=True, timeout=900, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return hyp.read_text().strip().splitlines()

    plain = run("p", {}, [])
    ref = run("r", {}, ["-mllr", mllr])
    assert ref == plain                                    # the reference's own behaviour
    got = run("g", {"LD_PRELOAD": PLUGIN}, ["-mllr", mllr])
    assert got == ref
    forced = run("f", {"LD_PRELOAD": PLUGIN, "B200_SEMI_MLLR": "1"}, ["-mllr", mllr])
    assert [l.rsplit(" ", 1)[1] for l in forced] != [l.rsplit(" ", 1)[1] for l in ref]   # scores move when applied


@needs
def test_sen_dump_roundtrip_against_the_reference(tmp_path):
    """CPU-only: the reference writes a senone dump (-senlogdir); our reader
    parses it (dense and delta-compressed frames), our writer re-emits it dense,
    and the unmodified decoder fed with OUR file (-senin yes) returns the same
    hypothesis as from the cepstra."""
    import numpy as np
    import cmusphinx_b200 as b
    senlog = tmp_path / "senlog"
    senlog.mkdir()
    hyp0, _ = _decode(tmp_path, "a", ["442c0201"], os.path.join(D, "test", "wsj"), ".mfc",
                      ["-senlogdir", str(senlog), "-fwdflat", "no", "-bestpath", "no"], {})
    sc, na, lb = b.sen_read(str(senlog / "442c0201.sen"))
    assert sc.shape[1] == 5150 and sc.shape[0] > 100 and abs(lb - 1.0001) < 1e-3
    assert (na < 5150).any() and (sc[na < 5150] == 0x7fff).any()         # compressed frames, dummies filled in
    for t in (0, sc.shape[0] // 2):
        listed = sc[t] != 0x7fff
        assert listed.sum() == na[t] and (sc[t][listed] >= 0).all()
    # dense scores for every senone (compallsen) through the reference, re-written by us
    senlog2 = tmp_path / "senlog2"
    senlog2.mkdir()
    _decode(tmp_path, "b", ["442c0201"], os.path.join(D, "test", "wsj"), ".mfc",
            ["-senlogdir", str(senlog2), "-compallsen", "yes", "-fwdflat", "no", "-bestpath", "no"], {})
    sc2, na2, _ = b.sen_read(str(senlog2 / "442c0201.sen"))
    assert (na2 == 5150).all()
    ours = tmp_path / "ours"
    ours.mkdir()
    b.sen_write(str(ours / "442c0201.sen"), sc2, lb)
    assert (ours / "442c0201.sen").read_bytes()[-sc2.nbytes // sc2.shape[0]:] == \
        (senlog2 / "442c0201.sen").read_bytes()[-sc2.nbytes // sc2.shape[0]:]
    hyp1, _ = _decode(tmp_path, "c", ["442c0201"], str(ours), ".sen", ["-senin", "yes", "-fwdflat", "no", "-bestpath", "no"], {})
    assert hyp1 == hyp0


@needs
@pytest.mark.gpu
def test_gpu_scores_via_sen_file_drive_the_unmodified_decoder(tmp_path):
    """No plug-in at all: GPU senone scores -> .sen file -> `pocketsphinx_batch
    -senin yes`.  Same hypothesis and path score as the reference decoding the
    cepstra itself (s2_semi model, fwdtree + fwdflat + bestpath)."""
    import numpy as np
    import cases
    import cmusphinx_b200 as b
    name = "semi_hub4wsj.npz"
    if not cases.have_model(name) or not orc.have_ref():
        pytest.skip("model files not present")
    g = cases.load(name)
    r = orc.RefAcmod(cases.model_dir(name))
    cep = orc.read_mfc(os.path.join(D, "test", "wsj", "442c0201.mfc"))
    feat_ref = r.cep2feat(cep)
    r.close()
    feat = b.feat_1s_c_d_dd(cep)                       # device feature stage
    np.testing.assert_array_equal(feat, feat_ref)
    m = b.tied_from_model_dir(cases.model_dir(name), int(g["n_sen"]), topn=4, logbase=orc.LOGBASE)
    scores = m.score(feat)
    m.free()
    d = tmp_path / "sen"
    d.mkdir()
    b.sen_write(str(d / "442c0201.sen"), scores, orc.LOGBASE)
    hyp_gpu, _ = _decode(tmp_path, "g", ["442c0201"], str(d), ".sen", ["-senin", "yes"], {})
    # the yardstick is the reference's own dump fed back to itself: its -senin path scores the
    # utterance 10 units differently from decoding the cepstra directly (-52953 vs -52963)
    refdump = tmp_path / "refdump"
    refdump.mkdir()
    hyp_direct, _ = _decode(tmp_path, "r", ["442c0201"], os.path.join(D, "test", "wsj"), ".mfc",
                            ["-compallsen", "yes", "-senlogdir", str(refdump)], {})
    hyp_cpu, _ = _decode(tmp_path, "s", ["442c0201"], str(refdump), ".sen", ["-senin", "yes"], {})
    assert hyp_gpu == hyp_cpu
    assert [l.rsplit("(", 1)[0] for l in hyp_gpu] == [l.rsplit("(", 1)[0] for l in hyp_direct]
    ref_sc, _, _ = b.sen_read(str(refdump / "442c0201.sen"))
    # frame 0 identical; later frames up to the documented rank-N tie deviation of the seed-free lists
    np.testing.assert_array_equal(scores[0], ref_sc[0])
    assert (scores != ref_sc[:scores.shape[0]]).mean() < 1e-4


TIDIGITS = os.path.join(D, "test", "tidigits")


def _tidigits(tmp_path, tag, env_extra):
    ctl = [l.strip() for l in open(os.path.join(TIDIGITS, "tidigits.ctl")) if l.strip()]
    hyp = tmp_path / f"{tag}.hyp"
    cmd = [BATCH, "-hmm", os.path.join(D, "hmm", "tidigits"), "-lm", os.path.join(D, "lm", "tidigits.DMP"), "-dict",
           os.path.join(D, "lm", "tidigits.dic"), "-ctl", os.path.join(TIDIGITS, "tidigits.ctl"), "-cepdir", TIDIGITS,
           "-hyp", str(hyp), "-logfn", str(tmp_path / f"{tag}.log")]
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = orc.REF_DIR + ":" + env.get("LD_LIBRARY_PATH", "")
    env.update(env_extra)
    subprocess.run(cmd, env=env, check=True, timeout=900, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return hyp.read_text().strip().splitlines(), len(ctl), (tmp_path / f"{tag}.log").read_text(errors="replace")


@needs
def test_reference_build_reproduces_the_bundled_tidigits_goldens(tmp_path):
    """The reference's own regression (test/regression/test-tidigits-simple.sh): the
    decoder built by oracle/Makefile must give the word strings of the bundled
    test-tidigits-simple.match on all utterances (its path scores differ by about 1 %
    across compilers, which the reference's own compare_table tolerates)."""
    if not os.path.exists(os.path.join(TIDIGITS, "tidigits.ctl")):
        pytest.skip("tidigits fixtures not copied (make -C oracle ref)")
    got, n, _ = _tidigits(tmp_path, "cpu", {})
    want = open(os.path.join(TIDIGITS, "test-tidigits-simple.match")).read().strip().splitlines()
    assert len(got) == len(want) == n
    assert [l.rsplit("(", 1)[0] for l in got] == [l.rsplit("(", 1)[0] for l in want]
    for g, w in zip(got, want):
        sg, sw = int(g.rsplit(" ", 1)[1].rstrip(")")), int(w.rsplit(" ", 1)[1].rstrip(")"))
        assert abs(sg - sw) <= 0.03 * abs(sw)


@needs
@pytest.mark.gpu
def test_identical_hypotheses_tidigits_four_stream_model(tmp_path):
    """hmm/en/tidigits: s2_4x features (4 streams of 12/24/3/12 dims), semi-continuous,
    through the plug-in: every hypothesis line (words and path score) identical to
    the reference's, and the words equal to the bundled goldens."""
    if not os.path.exists(os.path.join(TIDIGITS, "tidigits.ctl")):
        pytest.skip("tidigits fixtures not copied (make -C oracle ref)")
    cpu, n, _ = _tidigits(tmp_path, "cpu", {})
    gpu, _, log = _tidigits(tmp_path, "gpu", {"LD_PRELOAD": PLUGIN})
    assert "b200" in log.lower()
    assert gpu == cpu and len(gpu) == n


@needs
@pytest.mark.gpu
@pytest.mark.parametrize("inp", [RAW, MFC], ids=["raw", "wsj_mfc"])
def test_hmm_boundary_evaluate_channels_on_the_gpu(tmp_path, inp):
    """SURVEY.md section 8(b), HMM boundary: with B200_HMM_PLUGIN=1 the binding evaluates every
    frame's active channels -- what eval_root_chan / eval_nonroot_chan / eval_word_chan
    (ngram_search_fwdtree.c:598-691) would pass to hmm_vit_eval one by one -- in one batched
    b200_hmm_eval_host call through b200_hmm_pack / b200_hmm_unpack_one.  The unmodified decoder
    must produce the same hypothesis lines (words and path scores), with the reference's own GMM
    back-end and with the GPU one."""
    utts, cepdir, ext, extra = inp
    passes = ["-fwdflat", "no", "-bestpath", "no"]
    ref, _ = _decode(tmp_path, "ref", utts, cepdir, ext, extra + passes, {})
    for tag, env in (("hmm", {"LD_PRELOAD": PLUGIN, "B200_HMM_PLUGIN": "1", "B200_PLUGIN_DISABLE": "1"}),
                     ("hmm_gmm", {"LD_PRELOAD": PLUGIN, "B200_HMM_PLUGIN": "1"})):
        ctl = tmp_path / f"{tag}.ctl"
        ctl.write_text("\n".join(utts) + "\n")
        hyp = tmp_path / f"{tag}.hyp"
        cmd = [BATCH, "-hmm", os.path.join(D, "hmm", "hub4wsj_sc_8k"), "-lm", os.path.join(D, "lm", "wsj0vp.5000.DMP"),
               "-dict", os.path.join(D, "lm", "cmu07a.dic"), "-ctl", str(ctl), "-cepdir", cepdir, "-cepext", ext,
               "-hyp", str(hyp), "-logfn", str(tmp_path / f"{tag}.log")] + extra + passes
        e = dict(os.environ)
        e["LD_LIBRARY_PATH"] = orc.REF_DIR + ":" + e.get("LD_LIBRARY_PATH", "")
        e.update(env)
        p = subprocess.run(cmd, env=e, check=True, timeout=1800, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        got = hyp.read_text().strip().splitlines()
        assert got == ref, tag
        rep = [l for l in p.stderr.splitlines() if l.startswith("b200 hmm:")]
        assert rep, "the HMM binding did not run"
        frames, evals, fell = (int(x) for x in __import__("re").findall(r"(\d+) frames, (\d+) HMM evaluations on the GPU, (\d+) calls", rep[-1])[0])
        assert frames > 100 and evals > 20 * frames and fell == 0, rep[-1]


# ---------------------------------------------------------------- sphinx3 boundary (last row)
S3_PLUGIN = os.path.join(orc.ROOT, "cmusphinx_b200", "_plugin", "libb200_s3_plugin.so")
S3_DECODE = os.path.join(orc.REF_DIR, "sphinx3_decode")
needs_s3 = pytest.mark.skipif(not (os.path.exists(S3_PLUGIN) and os.path.exists(S3_DECODE)),
                              reason="sphinx3_decode (oracle/_ref) or the sphinx3 plug-in not built")


def _s3_decode(tmp_path, tag, extra, env_extra):
    """The reference's own regression decode (sphinx3/src/tests/regression/test-decode-s3cont.sh): hub4_cd_continuous
    + the an4 unigram LM on pittsburgh.littleendian.mfc through the UNMODIFIED sphinx3_decode."""
    am, lm = os.path.join(D, "hmm", "cont"), os.path.join(D, "lm", "an4")
    cep = tmp_path / "cep"
    cep.mkdir(exist_ok=True)
    dst = cep / "pittsburgh.littleendian.mfc"
    if not dst.exists():
        os.symlink(os.path.join(D, "test", "pittsburgh.littleendian.mfc"), dst)
    cmd = [S3_DECODE, "-mdef", os.path.join(am, "mdef"), "-fdict", os.path.join(lm, "filler.dict"), "-dict", os.path.join(lm, "an4.dict"),
           "-mean", os.path.join(am, "means"), "-var", os.path.join(am, "variances"), "-mixw", os.path.join(am, "mixture_weights"),
           "-tmat", os.path.join(am, "transition_matrices"), "-ctl", os.path.join(lm, "an4.ctl"), "-cepdir", str(cep),
           "-agc", "none", "-varnorm", "no", "-cmn", "current", "-maxwpf", "1", "-beam", "1e-40", "-pbeam", "1e-30",
           "-wbeam", "1e-20", "-maxhmmpf", "1500", "-wend_beam", "1e-1", "-feat", "1s_c_d_dd",
           "-lm", os.path.join(lm, "an4.ug.lm.DMP")] + extra
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = orc.REF_DIR + ":" + env.get("LD_LIBRARY_PATH", "")
    env.update(env_extra)
    p = subprocess.run(cmd, env=env, timeout=900, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, errors="replace")
    assert p.returncode == 0, p.stdout[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("FWDVIT:") or l.startswith("FWDXCT:")]
    return lines, p.stdout


@needs_s3
def test_sphinx3_reference_build_reproduces_its_regression_result(tmp_path):
    lines, _ = _s3_decode(tmp_path, "ref", ["-senmgau", ".s3cont."], {})      # the regression's own setting (ms_mgau path)
    assert lines and all("P I T G S B U R G H" in l for l in lines if l.startswith("FWDVIT:"))     # test-decode-s3cont.sh


@needs_s3
@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["-ci_pbeam", "1e-5", "-maxcdsenpf", "1000"], ["-ds", "2", "-ci_pbeam", "1e-8"],
                                   ["-subvq", os.path.join(D, "hmm", "cont", "test.subvq"), "-subvqbeam", "1e-2"]],
                         ids=["default", "ci_beam_dyn", "ds2", "subvq"])
def test_sphinx3_boundary_frame_eval_on_the_gpu(tmp_path, extra):
    """approx_cont_mgau_frame_eval of the unmodified sphinx3 decoder served by b200_s3_frame_eval
    (plugin/b200_s3_mgau.c): the hypothesis AND the per-word acoustic / LM scores of the FWDXCT line
    must be identical, with the fast-GMM layers switched on as well.  (-senmgau .cont. selects the
    mgau_init / approx_cont_mgau path, kbcore.c:296-310; the regression script's .s3cont. is the ms_mgau one.)"""
    extra = ["-senmgau", ".cont."] + extra
    ref, _ = _s3_decode(tmp_path, "ref", extra, {})
    gpu, log = _s3_decode(tmp_path, "gpu", extra, {"LD_PRELOAD": S3_PLUGIN})
    assert "approx_cont_mgau_frame_eval is served by libb200sphinx" in log
    assert ref and gpu == ref
    off, log2 = _s3_decode(tmp_path, "off", extra, {"LD_PRELOAD": S3_PLUGIN, "B200_S3_PLUGIN_DISABLE": "1"})
    assert off == ref and "served by libb200sphinx" not in log2
