cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_gmm.py -x -q -m gpu -k fuzz_within 2>&1 | grep -E "AssertionError|seed|passed|failed" | cut -c1-300
python - <<'PY'
import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import orc, cmusphinx_b200 as b
for seed in range(10):
    rng = np.random.default_rng(900 + seed)
    S, M, D, T = int(rng.integers(9, 700)), int(rng.choice([8, 16, 32])), int(rng.integers(3, 40)), int(rng.integers(1, 600))
    mean = (rng.standard_normal((S, M, D)) * float(rng.uniform(0.3, 4))).astype(np.float32)
    var = np.exp(rng.uniform(np.log(1e-3), np.log(10.0), (S, M, D))).astype(np.float32)
    mixw = rng.dirichlet(np.ones(M), (S, 1)).astype(np.float32)
    pv, pd = orc.port_precompute(var.reshape(-1, D), D, 1e-4, orc.LOGBASE)
    q = orc.port_mixw_quantize(mixw, 1e-7, orc.LOGBASE)
    cfg = b.MgauConfig(S, 1, M, S, [D], topn=4, logbase=orc.LOGBASE)
    m = b.ms_from_arrays(cfg, mean, pv, pd, q, np.arange(S))
    feat = (mean[rng.integers(0, S, T), rng.integers(0, M, T)] + rng.standard_normal((T, D)) * float(rng.uniform(0.3, 6))).astype(np.float32)
    got = m.score_raw(feat).astype(np.int32) if hasattr(m,'score_raw') else None
    g2 = m.score(feat).astype(np.int32); st = m.cont_stats(); fmt = m.tc_last_format()
    m.set_path(0); e = m.score(feat).astype(np.int32)
    d = np.abs(g2 - e)
    bad = np.argwhere(d > 0)
    print("seed", seed, "S", S, "M", M, "D", D, "T", T, "fmt", fmt, "max", d.max(), "nbad", len(bad), "frames bad", len(set(bad[:,0])) if len(bad) else 0, st)
    if len(bad):
        t, s = bad[0]
        print("   first bad t", t, "s", s, "got", g2[t, s], "exact", e[t, s], " row min got", g2[t].min(), "argmin-exact", e[t].argmin(), "got there", g2[t, e[t].argmin()])
    m.free()
PY
