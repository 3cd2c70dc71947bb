"""Builders shared by the CPU and GPU tests: golden fixtures -> oracle models
(tests only) and product models (cmusphinx_b200)."""
import os

import numpy as np

import orc

G = orc.GOLDEN_DIR


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def ms_dims(g):
    n_sen, n_density, dim, n_feat, topn, T = [int(v) for v in g["dims"]]
    return n_sen, n_density, dim, n_feat, topn, T


def ms_oracle(g):
    n_sen, n_density, dim, n_feat, topn, T = ms_dims(g)
    return orc.PortMs(n_sen, n_feat, [dim] * n_feat, n_density, n_sen, topn, 1, g["mean"], g["var"], g["det"],
                      g["mixw"], np.arange(n_sen))


def ms_product(g, device=0):
    import cmusphinx_b200 as b
    n_sen, n_density, dim, n_feat, topn, T = ms_dims(g)
    cfg = b.MgauConfig(n_sen, n_feat, n_density, n_sen, [dim] * n_feat, topn=topn, logbase=orc.LOGBASE, device=device)
    return b.ms_from_arrays(cfg, g["mean"], g["var"], g["det"], g["mixw"].reshape(n_sen, n_feat, n_density),
                            np.arange(n_sen))


def deltas_of(g, i):
    row = g["act_deltas"][i]
    n = int(row[0])
    return row[1:1 + n].astype(np.uint8)


def active_ids(deltas):
    return np.cumsum(deltas.astype(np.int64))


MODEL_DIRS = {"semi_hub4wsj.npz": "hub4wsj_sc_8k", "ptm_hub4wsj.npz": "ptm",
              "cont_hub4_topn4.npz": "cont", "cont_hub4_topn8.npz": "cont"}


def model_dir(name):
    return os.path.join(orc.DATA_DIR, "hmm", MODEL_DIRS[name])


def have_model(name):
    return os.path.exists(os.path.join(model_dir(name), "means"))


def tied_arrays(name, g):
    """Reads a tied-mixture model directory with the PRODUCT's readers and
    precomputes with the ORACLE (tests) -- returns everything both sides need."""
    from cmusphinx_b200 import engine
    d = model_dir(name)
    gm, gv = engine.read_gauden(d + "/means"), engine.read_gauden(d + "/variances")
    n_sen = int(g["n_sen"])
    sd = engine.read_sendump(d + "/sendump", gm["n_feat"], gm["n_density"], n_sen)
    return gm, gv, sd, n_sen


# (ftype, cmn, varnorm, agc, use_lda, lda_dim, svspec) of tests/golden/feat_general.npz
# (outputs of the reference's own feat_t; tests/golden/make_golden.py feat_general)
FEAT_GOLDEN_CASES = [
    ("1s_c_d_dd", 1, 0, 0, False, 0, None),
    ("1s_c_d_dd", 1, 1, 1, False, 0, None),
    ("1s_c_d_dd", 1, 0, 0, True, 29, None),
    ("1s_c_d_dd", 1, 0, 0, True, 0, "0-9/10-19/20-28"),
    ("1s_c_d_dd", 0, 0, 1, False, 0, "0-12/13-25/26-38"),
    ("s3_1x39", 1, 0, 1, False, 0, None),
    ("s3_1x39", 1, 1, 0, True, 32, "3,1,5/0"),
    ("s2_4x", 1, 0, 0, False, 0, None),
    ("s2_4x", 0, 0, 1, False, 0, None),
    ("1s_c_d_ld_dd", 1, 1, 0, False, 0, None),
    ("1s_c_d_ld_dd", 1, 0, 0, True, 40, None),
    ("1s_c", 1, 0, 0, False, 0, None),
    ("1s_c_d", 1, 0, 1, True, 0, None),
]
