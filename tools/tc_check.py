#!/usr/bin/env python
"""Development aid: compares the tcgen05 path with the exact CUDA-core path
(itself bit-exact vs the oracle) on the GPU and prints mismatch statistics."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cmusphinx_b200 as b
from cmusphinx_b200 import synth
from cmusphinx_b200.engine import LOGBASE


def run(n_sen, M, D, T, seed=5):
    mean, var, mixw = synth.cont_model(n_sen, M, D, seed)
    pv, pd = b.gauden_precompute(var.reshape(-1, D), D, 1e-4, LOGBASE)
    q = b.mixw_quantize_ms(mixw, 1e-7, LOGBASE)
    feat = synth.cont_features(mean, var, T, seed + 1)
    cfg = b.MgauConfig(n_sen, 1, M, n_sen, [D], topn=4, logbase=LOGBASE)
    m = b.ms_from_arrays(cfg, mean, pv, pd, q, np.arange(n_sen))
    print(f"S={n_sen} M={M} D={D} T={T}: default path {m.path}", flush=True)
    m.set_path(0)
    t0 = time.time(); want = m.score(feat); t_exact = time.time() - t0
    try:
        m.set_path(1)
    except b.B200Error as e:
        print("  tc path unavailable:", e); m.free(); return
    t0 = time.time(); got = m.score(feat); t_tc = time.time() - t0
    diff = got.astype(np.int32) - want.astype(np.int32)
    ad = np.abs(diff)
    print(f"  exact {t_exact*1e3:.1f} ms, tc {t_tc*1e3:.1f} ms; mismatch frac {(ad != 0).mean():.3e}, max|d| {ad.max()}, "
          f">1: {(ad > 1).sum()}  hist {np.bincount(np.minimum(ad.ravel(), 5))}", flush=True)
    if ad.max() > 1:
        idx = np.argwhere(ad > 1)[:5]
        for t, s in idx:
            print("   bad", t, s, got[t, s], want[t, s])
    m.free()


if __name__ == "__main__":
    run(64, 32, 39, 300)
    run(256, 32, 39, 1000)
    run(250, 8, 39, 515)
    run(100, 16, 13, 129)
    run(5000, 32, 39, 4096)
