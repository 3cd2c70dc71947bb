#!/usr/bin/env python
"""bench_hmm.py -- BASELINE.json configs[3]: hmm_vit_eval (3-state) over 50 000
active HMMs per frame with the fwdtree beam test and the active-senone gather,
on one B200.  Secondary benchmark (bench.py carries the headline metric); prints
one JSON line with the HBM roofline of the step (SURVEY.md section 8(d):
76 algorithmic bytes per HMM*frame).

  python bench_hmm.py [--utts B] [--frames F]

B > 1 batches B utterances x 50k HMMs in one resident population
(b200_hmm_pop_set_utts): per-frame launch latency, not bandwidth, bounds the
single-utterance case.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_HMM, N_SEN, N_TMAT, N_SSEQ, NE = 50_000, 5000, 50, 27_000, 3
BYTES_PER_UNIT = 76
BEAM = -1080 * 40      # synthetic scores are far more spread than real ones: keep ~half


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=64)
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--cpu-frames", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=80)
    ap.add_argument("--single-steps", action="store_true", help="one b200_hmm_step_dev call per frame (no graph)")
    args = ap.parse_args()
    import torch
    import cmusphinx_b200 as b
    from cmusphinx_b200 import synth
    from cmusphinx_b200.engine import LOGBASE
    assert b.device_count() > 0
    B = args.utts
    tp = b.tmat_quantize(synth.bakis_tmat(N_TMAT, NE, 7), 1e-4, LOGBASE)
    d = synth.hmm_population(N_HMM * B, NE, N_SEN, N_TMAT, N_SSEQ, seed=42, mpx_fraction=0.1)
    pop = b.HmmPopulation(N_HMM * B, NE)
    pop.score[:], pop.history[:], pop.senid[:] = d["score"].T, d["history"].T, d["senid"].T
    pop.out_score[:], pop.out_history[:], pop.tmatid[:], pop.mpx[:] = d["out_score"], d["out_history"], d["tmatid"], d["mpx"]
    ctx = b.HmmContext(NE, tp, d["sseq"], N_SEN)
    ctx.upload(pop)
    ctx.set_utts(np.arange(B + 1, dtype=np.int32) * N_HMM)
    n_sets = 8
    sen = torch.from_numpy(synth.senscr_frames(n_sets * B, N_SEN, 99).reshape(n_sets, B, N_SEN)).cuda()
    # The frames of a run are issued by b200_hmm_run_dev: replays of one instantiated CUDA graph of
    # 32 frames (5 kernels each) -- per-frame launch latency is what bounds the single-utterance case.
    # --single-steps times the same frames as individual b200_hmm_step_dev calls.
    side = torch.cuda.Stream()          # the legacy default stream cannot be captured
    stream = side.cuda_stream
    torch.cuda.synchronize()

    def run(n):
        if args.single_steps:
            for f in range(n):
                b.lib.b200_hmm_step_dev(ctx._h, sen[f % n_sets].data_ptr(), BEAM, stream)
        else:
            ctx.run_dev(sen.data_ptr(), B * N_SEN, n_sets, n, BEAM, stream)

    run(args.warmup)                    # warm-up; >= 72 frames builds the graph
    torch.cuda.synchronize()
    l0 = b.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(side)
    run(args.frames)
    e1.record(side)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.frames
    launches = b.launch_count() - l0
    units = N_HMM * B
    value = units / (ms / 1e3)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    achieved = units * BYTES_PER_UNIT / (ms / 1e3) / 1e9
    best, nk, _ = ctx.step(sen[0].cpu().numpy(), BEAM, units, want_idx=False)
    # CPU: the reference's own hmm_vit_eval on one utterance's population
    cpu = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orc
        fn = orc.ref().ref_hmm_eval_batch if orc.have_ref() else orc.port.orc_hmm_eval_batch
        one = {k: (v[:N_HMM].copy() if k != "sseq" else v) for k, v in d.items()}
        s0 = synth.senscr_frames(1, N_SEN, 99)[0]
        t0 = time.perf_counter()
        orc.hmm_eval(fn, NE, tp, d["sseq"], s0, one["score"], one["history"], one["out_score"], one["out_history"],
                     one["senid"], one["tmatid"], one["mpx"], one["bestscore"], repeat=args.cpu_frames)
        dt = time.perf_counter() - t0
        cpu = {"value": N_HMM * args.cpu_frames / dt, "unit": "HMM*frames/s", "cores": 1,
               "kind": "reference" if orc.have_ref() else "port",
               "sample": f"{args.cpu_frames} frames x {N_HMM} HMMs, hmm_vit_eval only (no beam / gather), includes AoS marshalling"}
    except Exception as ex:
        cpu = {"value": None, "sample": f"failed: {ex!r}"}
    print(json.dumps({
        "metric": "hmm_frames_evaluated_per_sec", "value": value, "unit": "HMM*frames/s", "n_gpus": 1,
        "steps": args.frames, "ms_per_step": ms, "higher_is_better": True, "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"hmm_vit_eval_3st + beam + compaction + active-senone gather, {B} utterances x {N_HMM} "
                               "HMMs per frame (BASELINE configs[3])", "n_sen": N_SEN, "mpx_fraction": 0.1},
        "gpu_launches": int(launches),
        "issue": "b200_hmm_step_dev per frame" if args.single_steps else "b200_hmm_run_dev (CUDA-graph replays of 32 frames)",
        "survivor_fraction": float(np.sum(nk)) / units,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "algorithmic_bytes_per_unit": BYTES_PER_UNIT, "traffic": None},
        "cpu_baseline": cpu}))
    ctx.free()


if __name__ == "__main__":
    main()
