#!/usr/bin/env python
"""bench_s3.py -- sphinx3's flavour of the scoring path (north_star:
approx_cont_mgau_frame_eval) on one B200.  Secondary benchmark (bench.py
carries the headline metric); prints one JSON line.

Workload: the shape of sphinx3's bundled hub4_cd_continuous_8gau_1s_c_d_dd
model (6144 senones x 8 diagonal Gaussians x 39 dims, 144 CI senones,
SURVEY.md Appendix B) with seeded synthetic parameters and features.
  dense : mgau_eval for every senone and frame (b200_s3_dense_dev)
  approx: CI pass + approx_cont_mgau_frame_eval with bursty active-senone sets
          and a CI beam (b200_s3_score_utt_dev)

Roofline: the reference's arithmetic forces float64 multiplies/subtracts that
cannot be fused (S3/libam/cont_mgau.c:1062-1068), so the kernel is bound by the
FP64 pipe: algorithmic work = 3 float64 instructions per (Gaussian, dimension)
+ 1 per Gaussian = M*(3D+1) per frame*senone; peak = the FP64 issue rate
measured on the same device by b200_fp64_issue_rate().

  python bench_s3.py [--frames T] [--steps K] [--warmup W]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S, N_CI, M, D = 6144, 144, 8, 39


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-frames", type=int, default=2000)
    args = ap.parse_args()
    print(json.dumps(run(args)))


def cpu_reference(mean, var, mixw, cd2ci, n_ci, feat_h, act_h, beam, n):
    """The reference's own mgau_eval / approx_cont_mgau_frame_eval (oracle/_ref/libs3am.so) on one
    core, model handed over as S3 files; falls back to the oracle port when _ref is not built."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    from cmusphinx_b200 import s3io
    have = os.path.exists(os.path.join(orc.REF_DIR, "libref_shim_s3.so"))
    tmp = tempfile.TemporaryDirectory(prefix="b200s3_")
    if have:
        f = [os.path.join(tmp.name, k) for k in ("means", "variances", "mixture_weights")]
        s3io.write_gauden(f[0], mean, [D]); s3io.write_gauden(f[1], var, [D]); s3io.write_mixw(f[2], mixw.reshape(S, 1, M))
        p = orc.RefS3(f[0], f[1], f[2], None, cd2ci, n_ci)
    else:
        p = orc.PortS3(mean, var, mixw, cd2ci, n_ci)
    t0 = time.perf_counter()
    p.eval_utt(feat_h[:n], None)
    dt = time.perf_counter() - t0
    p.set_fast(ci_pbeam=beam); p.utt_reset()
    t0 = time.perf_counter()
    o = p.eval_utt(feat_h[:4 * n], act_h[:4 * n])
    dt2 = time.perf_counter() - t0
    p.free()
    tmp.cleanup()
    return o, {"value": n * S / dt, "unit": "frame*senones/s", "cores": 1, "kind": "reference" if have else "port",
               "sample": f"{n} frames dense ({dt:.1f} s), {4 * n} frames approx ({dt2:.1f} s), one core; GPU == this on the sample",
               "approx_frames_per_s": 4 * n / dt2}


def run(args):
    import torch
    import cmusphinx_b200 as b
    from cmusphinx_b200 import synth
    from cmusphinx_b200._lib import check
    assert b.device_count() > 0, "bench_s3.py needs a CUDA device"
    mean, var, mixw, cd2ci, n_ci = synth.s3_model(S, N_CI, M, D, seed=31)
    m = b.S3Mgau.from_arrays(mean, var, mixw, cd2ci, n_ci)
    T = args.frames
    feat_h = synth.s3_features(mean, var, T, seed=32)
    act_h = synth.s3_active(S, n_ci, T, seed=33, p_on=0.05, p_off=0.2)
    feat = torch.from_numpy(feat_h).cuda()
    act = torch.from_numpy(act_h).cuda()
    out = torch.empty((T, S), dtype=torch.int32, device="cuda")
    best = torch.empty(T, dtype=torch.int32, device="cuda")
    prev_stream = torch.cuda.current_stream()
    torch.cuda.set_stream(torch.cuda.Stream())     # time and launch on the same non-default stream
    st = torch.cuda.current_stream().cuda_stream
    fp64_rate = b.lib.b200_fp64_issue_rate(0)
    flop_unit = M * (3 * D + 1)

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    l0 = b.launch_count()
    ms_dense = timed(lambda: check(b.lib.b200_s3_dense_dev(m.h, feat.data_ptr(), T, out.data_ptr(), st), "dense"))
    dense_rate = T * S / (ms_dense * 1e-3)
    res = {"dense": {"ms_per_step": ms_dense, "frame_senones_per_s": dense_rate,
                     "fp64_instr_per_s": dense_rate * flop_unit, "frac_of_fp64_peak": dense_rate * flop_unit / fp64_rate}}
    # approx path: beam chosen so that roughly half of the active CD senones stay inside it
    dense_h = out[:256, :n_ci].cpu().numpy()
    spread = float(np.median(dense_h.max(1) - np.median(dense_h, 1)))
    beam = float(np.float32(1.0003)) ** (-spread)
    m.set_fast(ci_pbeam=beam)

    def approx():
        m.utt_reset()
        check(b.lib.b200_s3_score_utt_dev(m.h, feat.data_ptr(), T, 0, act.data_ptr(), out.data_ptr(),
                                            best.data_ptr(), st), "score_utt")

    ms_approx = timed(approx)
    act_after = act.cpu().numpy()
    n_active = int(act_after.sum())
    res["approx"] = {"ms_per_step": ms_approx, "frames_per_s": T / (ms_approx * 1e-3),
                     "active_frame_senones_per_s": n_active / (ms_approx * 1e-3),
                     "active_fraction": n_active / (T * S), "ci_pbeam": beam}
    launches = b.launch_count() - l0
    # CPU baseline: the reference itself on a bounded sample, one core; parity on the sample while we are here
    n = args.cpu_frames
    o, cpu = cpu_reference(mean, var, mixw, cd2ci, n_ci, feat_h, act_h, beam, n)
    m.utt_reset()
    got = m.eval_utt(feat_h[:4 * n], act_h[:4 * n])
    assert np.array_equal(got[0], o[0]) and np.array_equal(got[1], o[1]), "GPU != reference on the bench sample"
    torch.cuda.synchronize()
    torch.cuda.set_stream(prev_stream)
    return ({
        "metric": "frames_x_senones_scored_per_sec (sphinx3 mgau_eval, float64)", "value": dense_rate,
        "unit": "frame*senones/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dense,
        "higher_is_better": True, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"sphinx3 cont_mgau {S} senones x {M} Gaussians x {D} dims ({n_ci} CI), {T} frames/step",
                   "l2": "score matrix 4 B x T x S = %.0f MB per step streams through L2" % (4e-6 * T * S)},
        "roofline": {"bound": "fp64", "achieved": dense_rate * flop_unit / 1e12, "peak": fp64_rate / 1e12,
                     "unit": "T f64-instr/s", "frac": dense_rate * flop_unit / fp64_rate,
                     "algorithmic_f64_instr_per_unit": flop_unit,
                     "peak_source": "b200_fp64_issue_rate(): DMUL/DADD chains measured on this device"},
        "paths": res, "gpu_launches": int(launches),
        "cpu_baseline": cpu,
    })


if __name__ == "__main__":
    main()
