"""The beam / scatter passes on their fat tiles (8 x 256 HMMs per CTA), which only populations of
several hundred thousand HMMs reach -- the bench_hmm.py configuration.  Kept in its own file, after
the other suites, because it is the largest GPU test."""
import numpy as np
import pytest

import orc
import cmusphinx_b200 as b
from cmusphinx_b200 import synth
from test_gpu_hmm import _to_pop

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ne,per", [(3, 16384), (3, 8191), (1, 8192), (2, 8192), (4, 4096), (5, 4096), (2, 6001)])
def test_big_batched_population_uses_fat_tiles_and_matches_independent_decoders(ne, per):
    """40 utterances x 16 384 HMMs: large enough that every CTA of hmm_run_kernel owns several 256-HMM tiles of an
    utterance (the bench_hmm.py configuration), so phase A runs its pair form (two adjacent HMMs per thread, NE <= 3),
    the beam pass its 8-tile batches and the scatter its 4-tile batches; best scores, survivor lists in (utterance, index)
    order and the per-utterance active-senone masks against the oracle.  The other shapes: every NE (1, 2 pair form through
    eval_any; 4, 5 single loop), and odd utterance lengths (odd utterance starts switch the pair form off for those
    utterances, the last tile is ragged)."""
    n_sen, n_tmat, n_sseq, n_utt = 3000, 20, 6000, 40
    n = n_utt * per
    off = (np.arange(n_utt + 1) * per).astype(np.int32)
    tp = orc.port_tmat_quantize(synth.bakis_tmat(n_tmat, ne, 7), 1e-4, orc.LOGBASE)
    d = synth.hmm_population(n, ne, n_sen, n_tmat, n_sseq, seed=17, mpx_fraction=0.1)
    sen = synth.senscr_frames(n_utt, n_sen, 18)
    beam = -30000
    ctx = b.HmmContext(ne, tp, d["sseq"], n_sen)
    ctx.upload(_to_pop(d, ne))
    ctx.set_utts(off)
    best, idx, mask = ctx.step(sen, beam, n)
    want_best, want_idx, want_mask = [], [], []
    for u in range(n_utt):
        sl = slice(u * per, (u + 1) * per)
        part = {k: (d[k][sl].copy() if k != "sseq" else d[k]) for k in d}
        bb = orc.hmm_eval(orc.port.orc_hmm_eval_batch, ne, tp, d["sseq"], sen[u], part["score"], part["history"],
                          part["out_score"], part["out_history"], part["senid"], part["tmatid"], part["mpx"],
                          part["bestscore"])
        want_best.append(bb)
        keep = np.nonzero(part["bestscore"] > bb + beam)[0]
        want_idx.append(keep + u * per)
        m = np.zeros((n_sen + 31) // 32, np.uint32)
        for st in range(ne):
            ids = part["senid"][keep, st].astype(np.int64)
            mp = part["mpx"][keep].astype(bool)
            ss = ids[mp]
            ss = ss[ss != 0xFFFF]
            allid = np.concatenate([ids[~mp], d["sseq"][ss, st].astype(np.int64)])
            np.bitwise_or.at(m, allid // 32, (np.uint32(1) << (allid % 32).astype(np.uint32)))
        want_mask.append(m)
    np.testing.assert_array_equal(best, np.array(want_best, np.int32))
    want_idx = np.concatenate(want_idx).astype(np.int32)
    assert 0.02 * n < want_idx.size < 0.98 * n          # the beam really splits the population
    np.testing.assert_array_equal(idx, want_idx)
    np.testing.assert_array_equal(mask, np.array(want_mask))
    ctx.free()


@pytest.mark.parametrize("ne,per", [(3, 8192), (3, 6001), (5, 4096)])
def test_big_run_of_frames_through_the_general_kernel(ne, per):
    """A population too large for hmm_resident_kernel (40 utterances, > 592 tiles): b200_hmm_run_dev goes through
    hmm_run_kernel with barriers per grid row, the pair form of phase A and per-utterance survivor lists packed after
    the run.  Five frames in one launch against five single steps (grid barriers, one frame per launch) on a second
    context, and the final state against the oracle."""
    n_sen, n_tmat, n_sseq, n_utt, cyc, n_frames, beam = 3000, 20, 6000, 40, 4, 5, -40000
    n = n_utt * per
    off = (np.arange(n_utt + 1) * per).astype(np.int32)
    tp = orc.port_tmat_quantize(synth.bakis_tmat(n_tmat, ne, 7), 1e-4, orc.LOGBASE)
    d = synth.hmm_population(n, ne, n_sen, n_tmat, n_sseq, seed=23, mpx_fraction=0.1)
    sen = np.ascontiguousarray(synth.senscr_frames(cyc * n_utt, n_sen, 19).reshape(cyc, n_utt, n_sen))
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
    for f in range(n_frames):
        a.step_dev_async(d_sen + ((f % cyc) * stride) * 2, beam)
    g.run_dev(d_sen, stride, cyc, n_frames, beam)
    ra, rg = a.step_results(n), g.step_results(n)
    for x, y in zip(ra, rg):
        np.testing.assert_array_equal(np.asarray(x), np.asarray(y))
    assert 0.02 * n < np.asarray(rg[1]).size < 0.98 * n
    pa, pg = b.HmmPopulation(n, ne), b.HmmPopulation(n, ne)
    a.download(pa); g.download(pg)
    for k in ("score", "history", "out_score", "out_history", "bestscore", "senid"):
        np.testing.assert_array_equal(getattr(pa, k), getattr(pg, k), err_msg=k)
    o = {k: v.copy() for k, v in d.items()}
    for f in range(n_frames):
        for u in range(n_utt):
            sl = slice(u * per, (u + 1) * per)
            part = {k: (o[k][sl] if k != "sseq" else o[k]) for k in o}
            orc.hmm_eval(orc.port.orc_hmm_eval_batch, ne, tp, d["sseq"], sen[f % cyc, u], part["score"], part["history"],
                         part["out_score"], part["out_history"], part["senid"], part["tmatid"], part["mpx"], part["bestscore"])
    np.testing.assert_array_equal(pg.score.T, o["score"])
    np.testing.assert_array_equal(pg.history.T, o["history"])
    np.testing.assert_array_equal(pg.bestscore, o["bestscore"])
    a.free(); g.free()
    b.lib.b200_dev_free(d_sen)
