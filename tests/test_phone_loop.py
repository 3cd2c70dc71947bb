"""The phone-loop look-ahead search, one frame: phone_loop_search_step (pocketsphinx/src/libpocketsphinx/
phone_loop_search.c:253-291) = renormalize_hmms + evaluate_hmms + prune_hmms + phone_transition.

Pins: tests/golden/phone_loop.npz and phone_loop_tight.npz (narrow beams: phones are pruned and re-entered) hold 60 consecutive frames each of a real decode of the UNMODIFIED reference
(numbers.raw, -pl_window 5) recorded by oracle/ref_pls_trace.c -- the reference's own phone_loop_search.c compiled
in place with one macro hook; make_pls_golden.py checks the oracle port on all 400 frames of the utterance.  The
GPU tests go through the C ABI (b200_phone_loop_*)."""
import os

import numpy as np
import pytest

import orc

GOLDENS = ("phone_loop.npz", "phone_loop_tight.npz")       # default beams (never prune) / -pl_beam 0.985 -pl_pbeam 0.99
KEYS = ("score", "history", "out_score", "out_history", "bestscore", "frame_of")
WORST = -0x20000000


def golden(name="phone_loop.npz"):
    z = np.load(os.path.join(orc.GOLDEN_DIR, name))
    n_fr = z["score"].shape[0]
    recs = []
    for f in range(n_fr):
        r = {k: z[k][f] for k in KEYS}
        r.update(frame=int(z["frame0"]) + f, best_score=int(z["best_score"][f]), beam=int(z["par"][0]), pbeam=int(z["par"][1]),
                 pip=int(z["par"][2]), tmatid=z["tmatid"], senscr=z["senscr"][f])
        recs.append(r)
    return z["tp"], recs


@pytest.mark.parametrize("name", GOLDENS)
def test_port_matches_reference_golden(name):
    tp, recs = golden(name)
    assert len(recs) == 61 and tp.shape == (50, 3, 4)
    n_pruned = n_entered = 0
    for a, b in zip(recs[:-1], recs[1:]):
        r = orc.port_phone_loop_step(tp, a)
        for k in KEYS:
            assert np.array_equal(r[k], b[k]), (a["frame"], k)
        assert r["best_score"] == b["best_score"]
        pruned = (a["frame_of"] >= a["frame"]) & (b["bestscore"] == WORST)        # prune_hmms cleared it ...
        n_pruned += int(pruned.sum())
        n_entered += int((pruned & (b["frame_of"] == a["frame"] + 1)).sum())        # ... and phone_transition re-entered it
    if name == "phone_loop_tight.npz":
        assert n_pruned > 1000 and n_entered > 0 and sum(int((r["frame_of"] < r["frame"]).sum()) for r in recs) > 100


def synthetic(rng, n, ne, n_utt, n_sen=200, very_negative=False):
    """Random phone-loop states: active / pruned-earlier phones, exact ties on a coarse grid; with very_negative the
    previous best is low enough for renormalize_hmms to fire."""
    from cmusphinx_b200 import synth
    tp = orc.port_tmat_quantize(synth.bakis_tmat(7, ne, 3), 1e-4, orc.LOGBASE)
    senid = rng.integers(0, n_sen, (ne, n)).astype(np.uint16)
    tmatid = rng.integers(0, 7, n).astype(np.int16)
    st = []
    for u in range(n_utt):
        off = -530_000_000 if (very_negative and u % 2 == 0) else 0
        sc = (off - rng.integers(0, 60, (n, ne)) * 50).astype(np.int32)
        sc[rng.random((n, ne)) < 0.2] = WORST
        fr = np.where(rng.random(n) < 0.7, 5, rng.integers(-1, 5, n)).astype(np.int32)
        sc[fr < 5] = WORST
        out = np.where(fr < 5, WORST, off - rng.integers(0, 60, n) * 50).astype(np.int32)
        st.append(dict(score=sc, history=rng.integers(-1, 40, (n, ne)).astype(np.int32), out_score=out,
                       out_history=rng.integers(-1, 40, n).astype(np.int32), bestscore=np.where(fr < 5, WORST, sc.max(1)).astype(np.int32),
                       frame_of=fr, best_score=int(off - 100), frame=5))
    return tp, senid, tmatid, st


def port_steps(tp, senid, tmatid, st, senscr, beam, pbeam, pip, n_frames):
    """The port on one utterance for n_frames frames; senscr [n_frames][n_sen]."""
    cur = dict(st)
    out = []
    for f in range(n_frames):
        rec = dict(cur, beam=beam, pbeam=pbeam, pip=pip, tmatid=tmatid, senscr=senscr[f][senid.T.astype(np.int64)], frame=st["frame"] + f)
        r = orc.port_phone_loop_step(tp, rec)
        cur = dict(r, frame=rec["frame"])
        out.append(r)
    return out


def test_abi_exports_the_phone_loop_entry_points():
    import ctypes as C
    lib = C.CDLL(os.path.join(orc.ROOT, "cmusphinx_b200", "libb200sphinx.so"))
    for name in ("b200_phone_loop_create", "b200_phone_loop_free", "b200_phone_loop_start", "b200_phone_loop_set_state",
                 "b200_phone_loop_get_state", "b200_phone_loop_step_dev", "b200_phone_loop_step_host"):
        assert hasattr(lib, name), name


def _set(pl, states, ne):
    n_utt, n = len(states), states[0]["out_score"].shape[0]
    cat = lambda k: np.concatenate([s[k] for s in states])
    pl.set_state(np.ascontiguousarray(np.concatenate([s["score"] for s in states]).T), np.ascontiguousarray(np.concatenate([s["history"] for s in states]).T),
                 cat("out_score"), cat("out_history"), cat("bestscore"), cat("frame_of"), np.array([s["best_score"] for s in states], np.int32))


def _check(pl, want, ne, tag):
    got = pl.state()
    n = want[0]["out_score"].shape[0]
    for u, w in enumerate(want):
        sl = slice(u * n, (u + 1) * n)
        assert np.array_equal(got["score"][:, sl].T, w["score"]), (tag, u, "score")
        assert np.array_equal(got["history"][:, sl].T, w["history"]), (tag, u, "history")
        for k, gk in (("out_score", "out_score"), ("out_history", "out_history"), ("bestscore", "bestscore"), ("frame_of", "frame")):
            assert np.array_equal(got[gk][sl], w[k]), (tag, u, k)
        assert got["best"][u] == w["best_score"], (tag, u, "best")


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDENS)
def test_gpu_phone_loop_matches_reference_golden(name):
    import cmusphinx_b200 as b
    tp, recs = golden(name)
    n, ne = recs[0]["score"].shape
    # compact senone numbering: senone of (phone i, state s) = i * ne + s, as the recording stores the scores
    senid = np.arange(n * ne, dtype=np.uint16).reshape(n, ne).T.copy()
    pl = b.PhoneLoop(tp, senid, recs[0]["tmatid"], n * ne, recs[0]["beam"], recs[0]["pbeam"], recs[0]["pip"], n_utt=1)
    _set(pl, [recs[0]], ne)
    for a, nxt in zip(recs[:-1], recs[1:]):
        best, pen = pl.step(a["senscr"].reshape(1, -1), a["frame"])
        _check(pl, [nxt], ne, a["frame"])
        assert best[0] == nxt["best_score"]
        assert np.array_equal(pen[0], nxt["bestscore"] - nxt["best_score"])       # phone_loop_search_score
    # phone_loop_search_start: every phone entered with score 0, history -1, frame 0; best 0
    pl.start()
    s = pl.state()
    assert (s["score"][0] == 0).all() and (s["score"][1:] == WORST).all() and (s["history"] == -1).all()
    assert (s["frame"] == 0).all() and (s["out_score"] == WORST).all() and s["best"][0] == 0
    pl.free()


@pytest.mark.gpu
@pytest.mark.parametrize("n,ne,n_utt,neg", [(50, 3, 5, False), (50, 3, 4, True), (33, 5, 3, False), (7, 1, 2, True), (130, 4, 2, False)])
def test_gpu_phone_loop_matches_port_on_synthetic_states(n, ne, n_utt, neg):
    import cmusphinx_b200 as b
    from cmusphinx_b200 import synth
    rng = np.random.default_rng(n * 7 + ne)
    n_sen, T = 200, 6
    tp, senid, tmatid, st = synthetic(rng, n, ne, n_utt, n_sen, very_negative=neg)
    # renormalize_hmms fires when best + 2 * beam is below WORST_SCORE: the very_negative case needs a wide beam
    beam, pbeam, pip = (-4_000_000 if neg else -600), -300, -20
    sen = np.stack([synth.senscr_frames(n_utt, n_sen, 70 + f) for f in range(T)])       # [T][n_utt][n_sen]
    want = [port_steps(tp, senid, tmatid, st[u], sen[:, u], beam, pbeam, pip, T) for u in range(n_utt)]
    if neg:
        assert any(w[0]["renorm"] for w in want) and not all(w[0]["renorm"] for w in want)
    pl = b.PhoneLoop(tp, senid, tmatid, n_sen, beam, pbeam, pip, n_utt=n_utt)
    _set(pl, st, ne)
    for f in range(T):
        best, pen = pl.step(sen[f], 5 + f)
        _check(pl, [w[f] for w in want], ne, f)
        assert np.array_equal(pl.state()["renorm"] != 0, np.array([w[f]["renorm"] for w in want]) != 0)
        for u in range(n_utt):
            assert np.array_equal(pen[u], want[u][f]["bestscore"] - want[u][f]["best_score"])
    pl.free()
