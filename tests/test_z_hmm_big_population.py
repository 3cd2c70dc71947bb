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


def test_big_batched_population_uses_fat_tiles_and_matches_independent_decoders():
    """40 utterances x 16 384 HMMs: large enough for the 8 x 256-HMM tiles of the beam and
    scatter passes (the bench_hmm.py configuration runs on them); best scores, survivor lists in
    (utterance, index) order and the per-utterance active-senone masks against the oracle."""
    ne, n_sen, n_tmat, n_sseq, n_utt, per = 3, 3000, 20, 6000, 40, 16384
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
