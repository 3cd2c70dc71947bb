"""sphinx3's flavour of hmm_vit_eval (SURVEY.md section 2.2b / VERDICT r1 row a21;
sphinx3/src/libs3decoder/libam/hmm.c:285-873): int32 transition log-probabilities and
senone scores that are added, WORST_SCORE = 0xc8000000, int32 ssids.

  * not gpu: the oracle port against sphinx3's own hmm.c compiled into oracle/_ref
    (all five evaluators: 3/5-state, mpx and not, and the generic topology for 1, 2, 4
    states), several frames in a row so that state written by one call feeds the next;
  * gpu: b200_s3hmm_eval_host against the port (and the reference when present).
The bundled KAT (_testhmm_tidigits.res) needs the tidigits model files of sphinx3's own
test tree, which are not in /root/reference: read-only reference, not reproducible here."""
import numpy as np
import pytest

import orc

W = np.int32(-0x38000000)            # S3_LOGPROB_ZERO = (int32)0xc8000000


def _case(ne, n_hmm, seed, mpx_fraction=0.3, n_tmat=7, n_sseq=300, n_sen=500):
    rng = np.random.default_rng(seed)
    # upper-triangular log transition matrices, at most one skip; "no transition" = W (or below)
    tp = np.full((n_tmat, ne, ne + 1), W, np.int32)
    for t in range(n_tmat):
        skip = rng.random() < 0.5
        for i in range(ne):
            for j in range(i, min(ne, i + (2 if skip else 1)) + 1):
                tp[t, i, j] = -rng.integers(100, 60000)
    sseq = rng.integers(0, n_sen, (n_sseq, ne)).astype(np.int16)
    score = -rng.integers(0, 1 << 22, (n_hmm, ne)).astype(np.int32)
    score[rng.random((n_hmm, ne)) < 0.25] = W
    score[rng.random((n_hmm, ne)) < 0.05] = W - 12345            # below WORST: the clamps and `!=` guards differ
    history = rng.integers(-1, 1 << 16, (n_hmm, ne)).astype(np.int32)
    out_score = -rng.integers(0, 1 << 22, n_hmm).astype(np.int32)
    out_history = rng.integers(-1, 1 << 16, n_hmm).astype(np.int32)
    mpx = (rng.random(n_hmm) < mpx_fraction).astype(np.uint8)
    ssid = rng.integers(0, n_sseq, (n_hmm, ne)).astype(np.int32)
    miss = (rng.random((n_hmm, ne)) < 0.3) & (mpx[:, None] == 1)
    miss[:, 0] = False
    ssid[miss] = -1
    if ne not in (3, 5):
        ssid[(rng.random(n_hmm) < 0.05) & (mpx == 0), 0] = -1    # the generic evaluator tolerates a missing ssid
    tmatid = rng.integers(0, n_tmat, n_hmm).astype(np.int32)
    best = np.full(n_hmm, W, np.int32)
    sen = -rng.integers(0, 200000, (4, n_sen)).astype(np.int32)
    return dict(tp=tp, sseq=sseq, score=score, history=history, out_score=out_score, out_history=out_history,
                ssid=ssid, tmatid=tmatid, mpx=mpx, bestscore=best, sen=sen, n_sen=n_sen)


def _run(which, ne, c):
    st = {k: c[k].copy() for k in ("score", "history", "out_score", "out_history", "ssid", "bestscore")}
    bests = []
    for f in range(c["sen"].shape[0]):
        bests.append(orc.s3hmm_eval(which, ne, c["tp"], c["sseq"], c["sen"][f], st["score"], st["history"],
                                    st["out_score"], st["out_history"], st["ssid"], c["tmatid"], c["mpx"], st["bestscore"]))
    return st, np.array(bests, np.int32)


@pytest.mark.parametrize("ne", [3, 5, 1, 2, 4])
def test_port_matches_sphinx3_hmm_vit_eval(ne):
    if not orc.have_ref_s3():
        pytest.skip("oracle/_ref/libref_shim_s3.so not built (make -C oracle ref)")
    c = _case(ne, 4000, 100 + ne)
    a, ba = _run("port", ne, c)
    r, br = _run("ref", ne, c)
    np.testing.assert_array_equal(ba, br)
    for k in a:
        if k == "ssid":                      # only mpx HMMs carry per-state ssids
            m = c["mpx"] == 1
            np.testing.assert_array_equal(a[k][m], r[k][m], err_msg=k)
            np.testing.assert_array_equal(a[k][~m][:, 0], r[k][~m][:, 0], err_msg=k)
        else:
            np.testing.assert_array_equal(a[k], r[k], err_msg=k)
    assert (a["score"] != c["score"]).any() and (ne == 1 or (a["history"] != c["history"]).any())


@pytest.mark.gpu
@pytest.mark.parametrize("ne", [3, 5, 1, 2, 4])
def test_gpu_matches_port_and_reference(ne):
    import cmusphinx_b200 as b
    c = _case(ne, 30000, 200 + ne)
    want, bw = _run("port", ne, c)
    g = {k: c[k].T.copy(order="C") if c[k].ndim == 2 else c[k].copy()
         for k in ("score", "history", "ssid", "out_score", "out_history", "bestscore")}
    best = b.s3hmm_vit_eval(ne, c["tp"], c["sseq"], c["n_sen"], c["sen"], g["score"], g["history"], g["out_score"],
                            g["out_history"], g["ssid"], c["tmatid"], c["mpx"], g["bestscore"])
    np.testing.assert_array_equal(best, bw)
    m = c["mpx"] == 1
    for k in ("score", "history"):
        np.testing.assert_array_equal(g[k].T, want[k], err_msg=k)
    np.testing.assert_array_equal(g["ssid"].T[m], want["ssid"][m])
    for k in ("out_score", "out_history", "bestscore"):
        np.testing.assert_array_equal(g[k], want[k], err_msg=k)
    if orc.have_ref_s3():
        r, br = _run("ref", ne, c)
        np.testing.assert_array_equal(best, br)
        np.testing.assert_array_equal(g["score"].T, r["score"])
        np.testing.assert_array_equal(g["out_score"], r["out_score"])


@pytest.mark.gpu
def test_gpu_rejects_bad_ids():
    import cmusphinx_b200 as b
    c = _case(3, 64, 9)
    g = {k: c[k].T.copy(order="C") if c[k].ndim == 2 else c[k].copy()
         for k in ("score", "history", "ssid", "out_score", "out_history", "bestscore")}
    g["ssid"][0, 5] = 10 ** 6
    with pytest.raises(b.B200Error):
        b.s3hmm_vit_eval(3, c["tp"], c["sseq"], c["n_sen"], c["sen"], g["score"], g["history"], g["out_score"],
                         g["out_history"], g["ssid"], c["tmatid"], c["mpx"], g["bestscore"])
