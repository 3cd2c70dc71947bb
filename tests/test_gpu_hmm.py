"""GPU parity tests for the batched hmm_vit_eval kernel, the beam/compaction
step and the active-senone gather: bit-exact against the oracle and the
reference-generated goldens."""
import numpy as np
import pytest

import cases
import orc
import cmusphinx_b200 as b
from cmusphinx_b200 import synth

pytestmark = pytest.mark.gpu

FIELDS = ("score", "history", "out_score", "out_history", "senid", "bestscore")


def _to_pop(d, ne):
    n = d["out_score"].shape[0]
    p = b.HmmPopulation(n, ne)
    p.score[:] = d["score"].T
    p.history[:] = d["history"].T
    p.senid[:] = d["senid"].T
    p.out_score[:] = d["out_score"]
    p.out_history[:] = d["out_history"]
    p.tmatid[:] = d["tmatid"]
    p.mpx[:] = d["mpx"]
    p.bestscore[:] = d["bestscore"]
    return p


def _assert_pop(p, d):
    np.testing.assert_array_equal(p.score.T, d["score"])
    np.testing.assert_array_equal(p.history.T, d["history"])
    np.testing.assert_array_equal(p.senid.T, d["senid"])
    np.testing.assert_array_equal(p.out_score, d["out_score"])
    np.testing.assert_array_equal(p.out_history, d["out_history"])
    np.testing.assert_array_equal(p.bestscore, d["bestscore"])


@pytest.mark.parametrize("ne", [3, 5, 1, 2, 4])   # 1, 2, 4: hmm_vit_eval_anytopo
def test_hmm_vit_eval_golden(ne):
    g = cases.load("tmat_hmm.npz" if ne in (3, 5) else "hmm_anytopo.npz")
    src = {k: g[f"h{ne}_in_{k}"] for k in ("score", "history", "out_score", "out_history", "senid", "tmatid", "mpx",
                                           "bestscore")}
    sen = g[f"h{ne}_senscr"]
    ctx = b.HmmContext(ne, g[f"h{ne}_tp"], g[f"h{ne}_in_sseq"], sen.shape[1])
    pop = _to_pop(src, ne)
    best = ctx.vit_eval(pop, sen)
    np.testing.assert_array_equal(best, g[f"h{ne}_best"])
    _assert_pop(pop, {k: g[f"h{ne}_out_{k}"] for k in FIELDS})
    ctx.free()


@pytest.mark.parametrize("ne,n_hmm", [(3, 50000), (5, 20011), (3, 1), (3, 255), (5, 257),
                                      (1, 3001), (2, 4099), (4, 20000)])   # 1, 2, 4: hmm_vit_eval_anytopo
def test_hmm_vit_eval_vs_oracle_config4(ne, n_hmm):
    """BASELINE config 4 population (50k HMMs, 10 % mpx, both skip / no-skip
    transition matrices) over several frames."""
    n_sen, n_tmat, n_sseq, nfr = 5000, 50, 27000, 5
    tp = orc.port_tmat_quantize(synth.bakis_tmat(n_tmat, ne, 7), 1e-4, orc.LOGBASE)
    d = synth.hmm_population(n_hmm, ne, n_sen, n_tmat, n_sseq, seed=42, mpx_fraction=0.1)
    sen = synth.senscr_frames(nfr, n_sen, 99)
    o = {k: v.copy() for k, v in d.items()}
    bests = [orc.hmm_eval(orc.port.orc_hmm_eval_batch, ne, tp, d["sseq"], sen[f], o["score"], o["history"],
                          o["out_score"], o["out_history"], o["senid"], o["tmatid"], o["mpx"], o["bestscore"])
             for f in range(nfr)]
    ctx = b.HmmContext(ne, tp, d["sseq"], n_sen)
    pop = _to_pop(d, ne)
    best = ctx.vit_eval(pop, sen)
    np.testing.assert_array_equal(best, np.array(bests, np.int32))
    _assert_pop(pop, o)
    ctx.free()


def test_empty_population_and_bad_arguments():
    tp = np.zeros((2, 3, 4), np.uint8)
    ctx = b.HmmContext(3, tp, None, 100)
    pop = b.HmmPopulation(0, 3)
    best = ctx.vit_eval(pop, np.zeros((2, 100), np.int16))
    assert (best == b.engine.WORST_SCORE).all()
    pop = b.HmmPopulation(4, 3)
    pop.tmatid[:] = 9
    with pytest.raises(b.B200Error, match="tmatid"):
        ctx.vit_eval(pop, np.zeros((1, 100), np.int16))
    with pytest.raises(b.B200Error):
        b.HmmContext(6, np.zeros((1, 6, 7), np.uint8), None, 10)   # HMM_MAX_NSTATE is 5 (PS/hmm.h:90)
    ctx.free()


@pytest.mark.parametrize("ne", [3, 5, 4])
def test_step_beam_compaction_and_active_senones(ne):
    n_sen, n_tmat, n_sseq, n_hmm = 5000, 50, 27000, 50000
    tp = orc.port_tmat_quantize(synth.bakis_tmat(n_tmat, ne, 7), 1e-4, orc.LOGBASE)
    d = synth.hmm_population(n_hmm, ne, n_sen, n_tmat, n_sseq, seed=3, mpx_fraction=0.1)
    sen = synth.senscr_frames(3, n_sen, 5)
    beam = -60000   # wide enough to keep a good fraction of this synthetic population
    ctx = b.HmmContext(ne, tp, d["sseq"], n_sen)
    pop = _to_pop(d, ne)
    ctx.upload(pop)
    o = {k: v.copy() for k, v in d.items()}
    for f in range(3):
        best_o = orc.hmm_eval(orc.port.orc_hmm_eval_batch, ne, tp, d["sseq"], sen[f], o["score"], o["history"],
                              o["out_score"], o["out_history"], o["senid"], o["tmatid"], o["mpx"], o["bestscore"])
        best, idx, mask = ctx.step(sen[f], beam, n_hmm)
        assert best == best_o
        keep = np.nonzero(o["bestscore"] > best_o + beam)[0]   # PS/ngram_search_fwdtree.c:741
        assert keep.size > 0 and keep.size < n_hmm
        np.testing.assert_array_equal(idx, keep.astype(np.int32))
        # acmod_activate_hmm over the survivors
        want = np.zeros((n_sen + 31) // 32, np.uint32)
        sid = o["senid"][keep]
        mp = o["mpx"][keep].astype(bool)
        for st in range(ne):
            ids = sid[:, st].astype(np.int64)
            nm = ids[~mp]
            ss = ids[mp]
            ss = ss[ss != 0xFFFF]
            allid = np.concatenate([nm, d["sseq"][ss, st].astype(np.int64)])
            np.bitwise_or.at(want, allid // 32, (np.uint32(1) << (allid % 32).astype(np.uint32)))
        np.testing.assert_array_equal(mask, want)
        np.testing.assert_array_equal(b.flags2list(mask, n_sen), orc.port_flags2list(want, n_sen))
    got = b.HmmPopulation(n_hmm, ne)
    got.tmatid[:] = d["tmatid"]
    got.mpx[:] = d["mpx"]
    ctx.download(got)
    _assert_pop(got, o)
    ctx.free()


def test_batched_utterances_match_independent_decoders():
    """B utterances sharing one resident population (b200_hmm_pop_set_utts): every
    utterance must behave exactly like its own decoder -- own senone scores, own
    best score / beam threshold, own active-senone mask; survivors ordered by
    (utterance, index)."""
    ne, n_sen, n_tmat, n_sseq = 3, 2000, 20, 5000
    sizes = [3000, 1, 0, 777, 2560]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    n = int(off[-1])
    tp = orc.port_tmat_quantize(synth.bakis_tmat(n_tmat, ne, 7), 1e-4, orc.LOGBASE)
    d = synth.hmm_population(n, ne, n_sen, n_tmat, n_sseq, seed=11, mpx_fraction=0.2)
    sen = synth.senscr_frames(2 * len(sizes), n_sen, 12).reshape(2, len(sizes), n_sen)
    beam = -80000
    ctx = b.HmmContext(ne, tp, d["sseq"], n_sen)
    ctx.upload(_to_pop(d, ne))
    ctx.set_utts(off)
    o = {k: v.copy() for k, v in d.items()}
    for f in range(2):
        want_best, want_idx, want_mask = [], [], []
        for u, sz in enumerate(sizes):
            sl = slice(int(off[u]), int(off[u + 1]))
            part = {k: (o[k][sl].copy() if k != "sseq" else o[k]) for k in o}
            bb = orc.hmm_eval(orc.port.orc_hmm_eval_batch, ne, tp, d["sseq"], sen[f, u], part["score"], part["history"],
                              part["out_score"], part["out_history"], part["senid"], part["tmatid"], part["mpx"],
                              part["bestscore"]) if sz else int(b.engine.WORST_SCORE)
            for k in FIELDS:
                o[k][sl] = part[k]
            want_best.append(bb)
            keep = np.nonzero(part["bestscore"] > bb + beam)[0] if sz else np.zeros(0, np.int64)
            want_idx.append(keep + int(off[u]))
            m = np.zeros((n_sen + 31) // 32, np.uint32)
            for st in range(ne):
                ids = part["senid"][keep, st].astype(np.int64)
                mp = part["mpx"][keep].astype(bool)
                ss = ids[mp]
                ss = ss[ss != 0xFFFF]
                allid = np.concatenate([ids[~mp], d["sseq"][ss, st].astype(np.int64)])
                np.bitwise_or.at(m, allid // 32, (np.uint32(1) << (allid % 32).astype(np.uint32)))
            want_mask.append(m)
        best, idx, mask = ctx.step(sen[f], beam, n)
        np.testing.assert_array_equal(best, np.array(want_best, np.int32))
        np.testing.assert_array_equal(idx, np.concatenate(want_idx).astype(np.int32))
        np.testing.assert_array_equal(mask, np.array(want_mask))
    ctx.free()


def test_hmm_maintenance_ops_match_sequential_reference():
    """hmm_normalize, hmm_clear_scores on pruned HMMs and batched hmm_enter
    (PS/hmm.c:169-218 + the callers' `if better` test) against the literal
    sequential loops."""
    ne, n_hmm, n_sen = 3, 6000, 400
    tp = orc.port_tmat_quantize(synth.bakis_tmat(8, ne, 3), 1e-4, orc.LOGBASE)
    d = synth.hmm_population(n_hmm, ne, n_sen, 8, 200, seed=5)
    sen = synth.senscr_frames(1, n_sen, 6)
    ctx = b.HmmContext(ne, tp, d["sseq"], n_sen)
    pop = b.HmmPopulation(n_hmm, ne)
    pop.score[:], pop.history[:], pop.senid[:] = d["score"].T, d["history"].T, d["senid"].T
    pop.out_score[:], pop.out_history[:], pop.tmatid[:], pop.mpx[:] = d["out_score"], d["out_history"], d["tmatid"], d["mpx"]
    ctx.upload(pop)
    ctx.set_utts(np.array([0, 2500, n_hmm], np.int32))
    best, idx, mask = ctx.step(np.stack([sen[0], sen[0]]), -60000, n_hmm)
    after = b.HmmPopulation(n_hmm, ne)
    ctx.download(after)
    W = int(b.engine.WORST_SCORE)
    keep = np.zeros(n_hmm, bool); keep[idx] = True
    # --- hmm_clear_scores for the pruned ones
    want_s, want_o, want_b = after.score.copy(), after.out_score.copy(), after.bestscore.copy()
    want_s[:, ~keep] = W; want_o[~keep] = W; want_b[~keep] = W
    ctx.clear_pruned()
    got = b.HmmPopulation(n_hmm, ne); ctx.download(got)
    np.testing.assert_array_equal(got.score, want_s); np.testing.assert_array_equal(got.out_score, want_o)
    np.testing.assert_array_equal(got.bestscore, want_b); np.testing.assert_array_equal(got.history, after.history)
    # --- hmm_normalize with the per-utterance best of the step
    per = np.where(np.arange(n_hmm) < 2500, best[0], best[1]).astype(np.int64)
    ws = want_s.astype(np.int64); wo = want_o.astype(np.int64)
    ws = np.where(ws > W, ws - per[None, :], ws); wo = np.where(wo > W, wo - per, wo)
    ctx.normalize()
    ctx.download(got)
    np.testing.assert_array_equal(got.score, ws.astype(np.int32)); np.testing.assert_array_equal(got.out_score, wo.astype(np.int32))
    # --- batched hmm_enter: duplicates, ties, entries that do not beat the resident score
    rng = np.random.default_rng(7)
    n = 5000
    eidx = rng.integers(0, 900, n).astype(np.int32)           # many duplicates
    escore = (rng.integers(-40, 5, n) * 1000).astype(np.int32)  # many ties
    escore[::7] = W
    ehist = np.arange(n, dtype=np.int32) + 100000
    s0, h0 = got.score[0].copy(), got.history[0].copy()
    for k in range(n):                                           # the reference's order
        if escore[k] > s0[eidx[k]]:
            s0[eidx[k]] = escore[k]; h0[eidx[k]] = ehist[k]
    ctx.enter(eidx, escore, ehist)
    fin = b.HmmPopulation(n_hmm, ne); ctx.download(fin)
    np.testing.assert_array_equal(fin.score[0], s0); np.testing.assert_array_equal(fin.history[0], h0)
    np.testing.assert_array_equal(fin.score[1:], got.score[1:]); np.testing.assert_array_equal(fin.history[1:], got.history[1:])
    ctx.free()
    if orc.have_ref():
        # the same three operations by the reference's own hmm.c (oracle/_ref), HMM-major
        r0 = orc.ref_hmm_maint(0, after.score.T, after.history.T, after.out_score, after.out_history, after.bestscore,
                               sel=(~keep).astype(np.uint8))
        np.testing.assert_array_equal(r0[0].T, want_s); np.testing.assert_array_equal(r0[2], want_o)
        np.testing.assert_array_equal(r0[4], want_b); np.testing.assert_array_equal(r0[1].T, after.history)
        r1 = orc.ref_hmm_maint(1, r0[0], r0[1], r0[2], r0[3], r0[4], arg=per.astype(np.int32))
        np.testing.assert_array_equal(r1[0].T, got.score); np.testing.assert_array_equal(r1[2], got.out_score)
        r2 = orc.ref_hmm_maint(2, r1[0], r1[1], r1[2], r1[3], r1[4], lidx=eidx, lscore=escore, lhist=ehist)
        np.testing.assert_array_equal(r2[0].T, fin.score); np.testing.assert_array_equal(r2[1].T, fin.history)
        np.testing.assert_array_equal(r2[2], fin.out_score)


@pytest.mark.parametrize("n_utt", [1, 3])
def test_run_of_frames_as_graph_replays_equals_single_steps(n_utt):
    """b200_hmm_run_dev (one persistent launch per run: hmm_resident_kernel for these populations, state in
    registers and one barrier per frame; the single steps go through hmm_run_kernel) leaves exactly the
    population and the last-frame results of the same number of b200_hmm_step_dev calls."""
    ne, n_sen, n_tmat, n_sseq, n, cyc, beam = 3, 800, 10, 2000, 4000, 8, -60000
    tp = orc.port_tmat_quantize(synth.bakis_tmat(n_tmat, ne, 7), 1e-4, orc.LOGBASE)
    d = synth.hmm_population(n, ne, n_sen, n_tmat, n_sseq, seed=21, mpx_fraction=0.15)
    off = np.linspace(0, n, n_utt + 1).astype(np.int32)
    sen = np.ascontiguousarray(synth.senscr_frames(cyc * n_utt, n_sen, 31).reshape(cyc, n_utt, n_sen))
    d_sen = b.lib.b200_dev_alloc(sen.nbytes, 0)
    assert d_sen
    b.engine.check(b.lib.b200_dev_upload(d_sen, sen.ctypes.data, sen.nbytes), "upload")
    stride = n_utt * n_sen
    ctxs = []
    for _ in range(2):
        c = b.HmmContext(ne, tp, d["sseq"], n_sen)
        c.upload(_to_pop(d, ne))
        c.set_utts(off)
        ctxs.append(c)
    a, g = ctxs
    done = 0
    for n_frames in (100, 75, 3, 2):
        for f in range(n_frames):
            a.step_dev_async(d_sen + ((f % cyc) * stride) * 2, beam)
        g.run_dev(d_sen, stride, cyc, n_frames, beam)
        ra, rg = a.step_results(n), g.step_results(n)
        for x, y in zip(ra, rg):
            np.testing.assert_array_equal(np.asarray(x), np.asarray(y))
        pa, pg = b.HmmPopulation(n, ne), b.HmmPopulation(n, ne)
        a.download(pa); g.download(pg)
        for k in ("score", "history", "out_score", "out_history", "bestscore", "senid"):
            np.testing.assert_array_equal(getattr(pa, k), getattr(pg, k), err_msg=k)
        done += n_frames
    a.free(); g.free()
    b.lib.b200_dev_free(d_sen)



def test_cluster_form_of_the_run_kernel_gives_the_same_results():
    """hmm_run_kernel<NE, 1024, true> (thread-block clusters, one utterance at a time through the
    whole run, B200_HMM_CLUSTER=1; opt-in because it measured slower) must pass the same
    beam / compaction / batched-utterance / run-of-frames tests.  The switch is read once per
    process, hence the sub-process."""
    import os
    import subprocess
    import sys
    if os.environ.get("B200_HMM_CLUSTER") == "1":
        pytest.skip("already inside the cluster-form run")
    env = dict(os.environ, B200_HMM_CLUSTER="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                        "step_beam or batched_utterances or run_of_frames"],
                       env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "passed" in r.stdout


@pytest.mark.parametrize("ne", [1, 2, 4, 5])
def test_run_of_frames_resident_kernel_every_topology(ne):
    """hmm_resident_kernel<NE> for the other numbers of emitting states (eval_any / eval5): a run of frames against
    the same number of single steps, two utterances with a ragged last tile each."""
    n_sen, n_tmat, n_sseq, n, cyc, beam, n_frames = 700, 9, 1500, 3001, 3, -50000, 7
    tp = orc.port_tmat_quantize(synth.bakis_tmat(n_tmat, ne, 5), 1e-4, orc.LOGBASE)
    d = synth.hmm_population(n, ne, n_sen, n_tmat, n_sseq, seed=40 + ne, mpx_fraction=0.2)
    off = np.array([0, 1301, n], np.int32)
    sen = np.ascontiguousarray(synth.senscr_frames(cyc * 2, n_sen, 33).reshape(cyc, 2, n_sen))
    d_sen = b.lib.b200_dev_alloc(sen.nbytes, 0)
    assert d_sen
    b.engine.check(b.lib.b200_dev_upload(d_sen, sen.ctypes.data, sen.nbytes), "upload")
    ctxs = []
    for _ in range(2):
        c = b.HmmContext(ne, tp, d["sseq"], n_sen)
        c.upload(_to_pop(d, ne))
        c.set_utts(off)
        ctxs.append(c)
    a, g = ctxs
    for f in range(n_frames):
        a.step_dev_async(d_sen + ((f % cyc) * 2 * n_sen) * 2, beam)
    g.run_dev(d_sen, 2 * n_sen, cyc, n_frames, beam)
    for x, y in zip(a.step_results(n), g.step_results(n)):
        np.testing.assert_array_equal(np.asarray(x), np.asarray(y))
    pa, pg = b.HmmPopulation(n, ne), b.HmmPopulation(n, ne)
    a.download(pa); g.download(pg)
    for k in ("score", "history", "out_score", "out_history", "bestscore", "senid"):
        np.testing.assert_array_equal(getattr(pa, k), getattr(pg, k), err_msg=k)
    a.free(); g.free()
    b.lib.b200_dev_free(d_sen)
