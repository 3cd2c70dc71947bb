#!/usr/bin/env python
"""bench_hmm.py -- BASELINE.json configs[3]: hmm_vit_eval (3-state) over 50 000
active HMMs per frame with the fwdtree beam test and the active-senone gather,
on one B200.  Secondary benchmark: `python bench.py` runs it after the headline
(the literal 1 x 50 000 configuration and the batched 64 x 50 000 one) and
attaches the results under `secondary.hmm_single` / `secondary.hmm`; stand-alone
it prints one JSON line with the HBM roofline of the step (SURVEY.md section
8(d): 76 algorithmic bytes per HMM*frame).

  python bench_hmm.py [--utts B] [--frames F]

The beam is chosen so that about half of the population survives every frame
of the timed region (the synthetic state scores are uniform over 2^20, so the
spread is stationary): the order-preserving compaction and the active-senone
gather are exercised at the survivor rate the launch list was profiled at.
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
BEAM = -(1 << 19)      # state scores are spread uniformly over 2^20: about half of the HMMs stay inside


def cpu_reference(d, tp, cpu_frames):
    """The reference's own hmm_vit_eval (oracle/_ref) on one utterance's population, one core."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    from cmusphinx_b200 import synth
    fn = orc.ref().ref_hmm_eval_batch if orc.have_ref() else orc.port.orc_hmm_eval_batch
    one = {k: (v[:N_HMM].copy() if k != "sseq" else v) for k, v in d.items()}
    s0 = synth.senscr_frames(1, N_SEN, 99)[0]
    t0 = time.perf_counter()
    orc.hmm_eval(fn, NE, tp, d["sseq"], s0, one["score"], one["history"], one["out_score"], one["out_history"],
                 one["senid"], one["tmatid"], one["mpx"], one["bestscore"], repeat=cpu_frames)
    dt = time.perf_counter() - t0
    return {"value": N_HMM * cpu_frames / dt, "unit": "HMM*frames/s", "cores": 1,
            "kind": "reference" if orc.have_ref() else "port",
            "sample": f"{cpu_frames} frames x {N_HMM} HMMs ({dt:.1f} s), hmm_vit_eval only (no beam / gather), one core"}


def _measure(b, torch, synth, d, tp, B, frames, warmup, single_steps, roots_first):
    """One timed run of `frames` frames on the population d (roots_first: every utterance's multiplex HMMs first)."""
    if roots_first:
        order = np.concatenate([u * N_HMM + np.argsort(d["mpx"][u * N_HMM:(u + 1) * N_HMM] == 0, kind="stable") for u in range(B)])
        d = {k: (np.ascontiguousarray(v[order]) if k != "sseq" else v) for k, v in d.items()}
    pop = b.HmmPopulation(N_HMM * B, NE)
    pop.score[:], pop.history[:], pop.senid[:] = d["score"].T, d["history"].T, d["senid"].T
    pop.out_score[:], pop.out_history[:], pop.tmatid[:], pop.mpx[:] = d["out_score"], d["out_history"], d["tmatid"], d["mpx"]
    ctx = b.HmmContext(NE, tp, d["sseq"], N_SEN)
    ctx.upload(pop)
    ctx.set_utts(np.arange(B + 1, dtype=np.int32) * N_HMM)
    n_sets = 8
    sen = torch.from_numpy(synth.senscr_frames(n_sets * B, N_SEN, 99).reshape(n_sets, B, N_SEN)).cuda()
    side = torch.cuda.Stream()          # the legacy default stream cannot be captured
    stream = side.cuda_stream
    torch.cuda.synchronize()

    def go(n):
        if single_steps:
            for f in range(n):
                b.lib.b200_hmm_step_dev(ctx._h, sen[f % n_sets].data_ptr(), BEAM, stream)
        else:
            ctx.run_dev(sen.data_ptr(), B * N_SEN, n_sets, n, BEAM, stream)

    go(warmup)
    torch.cuda.synchronize()
    _, nk0, _ = ctx.step_results(N_HMM * B, want_idx=False)
    l0 = b.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(side)
    go(frames)
    e1.record(side)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / frames
    launches = b.launch_count() - l0
    _, nk1, _ = ctx.step_results(N_HMM * B, want_idx=False)
    ctx.free()
    del sen
    torch.cuda.empty_cache()
    return ms, launches, float(np.sum(nk0)) / (N_HMM * B), float(np.sum(nk1)) / (N_HMM * B)


def run(utts=64, frames=200, warmup=80, single_steps=False, cpu=True, cpu_frames=200, both_layouts=True):
    import torch
    import cmusphinx_b200 as b
    from cmusphinx_b200 import synth
    from cmusphinx_b200.engine import LOGBASE
    assert b.device_count() > 0
    B = utts
    tp = b.tmat_quantize(synth.bakis_tmat(N_TMAT, NE, 7), 1e-4, LOGBASE)
    d = synth.hmm_population(N_HMM * B, NE, N_SEN, N_TMAT, N_SSEQ, seed=42, mpx_fraction=0.1)
    # The headline layout keeps every utterance's multiplex HMMs (the 10 % "roots") ahead of the others, as the reference does:
    # root channels live in their own array (ngram_search.h root_chan) and eval_root_chan / eval_nonroot_chan
    # (ngram_search_fwdtree.c:598-634) are separate loops.  The same population with the two kinds interleaved at random
    # (every warp then runs both hmm_vit_eval flavours) is timed beside it.
    ms, launches, s0, s1 = _measure(b, torch, synth, d, tp, B, frames, warmup, single_steps, True)
    units = N_HMM * B
    value = units / (ms / 1e3)
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    achieved = units * BYTES_PER_UNIT / (ms / 1e3) / 1e9
    res = {
        "metric": "hmm_frames_evaluated_per_sec", "value": value, "unit": "HMM*frames/s", "n_gpus": 1,
        "steps": frames, "ms_per_step": ms, "us_per_frame": ms * 1e3, "higher_is_better": True, "dtype": "int32",
        "data": "synthetic",
        "config": {"workload": f"hmm_vit_eval_3st + beam + compaction + active-senone gather, {B} utterance(s) x {N_HMM} "
                               "HMMs per frame (BASELINE configs[3])", "n_sen": N_SEN, "mpx_fraction": 0.1, "beam": BEAM,
                   "mpx_layout": "roots first within each utterance (the reference's root_chan array)"},
        "gpu_launches": int(launches),
        "issue": "b200_hmm_step_dev per frame" if single_steps else "b200_hmm_run_dev",
        "survivor_fraction": {"first_timed_frame": s0, "last_timed_frame": s1},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "algorithmic_bytes_per_unit": BYTES_PER_UNIT, "traffic": None,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if os.path.exists(pk) else "fallback"}}
    prof = os.path.join(ROOT, "profiles", "r2_ncu_hmm_run.json")
    if B == 64 and os.path.exists(prof):       # the profiled shape: DRAM bytes per frame of one ncu --set full capture
        res["roofline"]["traffic"] = json.load(open(prof))["per_frame"]["dram_bytes"]
        res["roofline"]["traffic_source"] = "static: profiles/r2_ncu_hmm_run.json (dram__bytes_read + write of one 8-frame launch / 8)"
    if both_layouts:
        ms2, _, _, _ = _measure(b, torch, synth, d, tp, B, max(20, frames // 4), max(10, warmup // 4), single_steps, False)
        res["interleaved_layout"] = {"us_per_frame": ms2 * 1e3, "frac": units * BYTES_PER_UNIT / (ms2 / 1e3) / 1e9 / peak,
                                     "note": "multiplex and plain HMMs interleaved at random: every warp runs both flavours"}
    if cpu:
        try:
            res["cpu_baseline"] = cpu_reference(d, tp, cpu_frames)
        except Exception as ex:
            res["cpu_baseline"] = {"value": None, "kind": "reference", "cores": 0, "sample": f"failed: {ex!r}"}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=64)
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--cpu-frames", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=80)
    ap.add_argument("--single-steps", action="store_true", help="one b200_hmm_step_dev call per frame")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    print(json.dumps(run(args.utts, args.frames, args.warmup, args.single_steps, not args.no_cpu_baseline, args.cpu_frames)))


if __name__ == "__main__":
    main()
