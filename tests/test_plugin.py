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
    exact kernels (path 0: identical path score required) and from the tcgen05
    Mahalanobis GEMM (path 1, scores within +-1: identical words required, and
    the path score may move by at most a few units)."""
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
    assert w_tc == w_cpu and abs(s_tc - s_cpu) <= 40, (s_tc, s_cpu)
