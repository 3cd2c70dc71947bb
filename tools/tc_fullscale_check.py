#!/usr/bin/env python
"""Full BASELINE-config-2 parity sweep on the GPU: tcgen05 path against the exact
CUDA-core path (itself bit-exact against the reference) over T frames x 5000
senones.  Prints the histogram of |difference| -- the north_star tolerance is 1."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cmusphinx_b200 as b  # noqa: E402
from cmusphinx_b200 import synth  # noqa: E402
from cmusphinx_b200.engine import LOGBASE  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
n_sen, M, D = 5000, 32, 39
mean, var, mixw = synth.cont_model(n_sen, M, D, 1234)
pv, pd = b.gauden_precompute(var.reshape(-1, D), D, 1e-4, LOGBASE)
q = b.mixw_quantize_ms(mixw, 1e-7, LOGBASE)
cfg = b.MgauConfig(n_sen, 1, M, n_sen, [D], topn=4, logbase=LOGBASE)
m = b.ms_from_arrays(cfg, mean, pv.reshape(mean.shape), pd.reshape(n_sen, 1, M), q, np.arange(n_sen))
hist = np.zeros(8, np.int64)
worst = []
stats = {}
for t0 in range(0, T, 20000):
    feat = synth.cont_features(mean, var, min(20000, T - t0), 5678 + t0)
    m.set_path(1); a = m.score(feat).astype(np.int32)
    for k, v in m.cont_stats().items():
        stats[k] = max(stats.get(k, 0), v) if k in ("overflow", "max_gemm_err") else stats.get(k, 0) + v
    m.set_path(0); e = m.score(feat).astype(np.int32)
    d = np.abs(a - e)
    hist += np.bincount(np.minimum(d, 7).ravel(), minlength=8)
    if d.max() > 1:
        idx = np.argwhere(d > 1)[:5]
        worst += [(int(t0 + i), int(j), int(a[i, j]), int(e[i, j])) for i, j in idx]
print(json.dumps({"frames": T, "scores": int(hist.sum()), "abs_diff_histogram_0_to_7plus": hist.tolist(),
                  "mismatch_fraction": float(hist[1:].sum() / hist.sum()), "beyond_tolerance": int(hist[2:].sum()),
                  "examples_frame_senone_tc_exact": worst[:10], "tensor_core_path_stats": stats,
                  "eps": [os.environ.get("B200_TC_EPS0", "default"), os.environ.get("B200_TC_EPS_SHIFT", "default")]}))
