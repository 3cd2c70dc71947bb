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
