#!/usr/bin/env python
"""bench_e2e.py -- BASELINE.json configs[0]/[4] on what the tree actually
bundles: the UNMODIFIED reference decoder (oracle/_ref/pocketsphinx_batch) on the
7 WSJ .mfc regression utterances (57.7 s of audio), once with its own CPU
back-ends and once with the GPU plug-in LD_PRELOADed.  Reports the decoder's own
xRT lines (pocketsphinx/src/programs/batch.c:774-776) and whether the hypothesis
files are identical (words, and words + path scores; for ptm the reference's own
scores are heap dependent -- it reads past its log-add table -- so only the
words can be expected to match, see tests/test_oracle_vs_ref.py).  The search (lextree, LM, lattice) stays on one host core
in both runs, so this measures integration overhead, not kernel throughput; the
100k-utterance / WSJ-20k-LM configuration is not reproducible (LM not in tree).
"""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(ROOT, "oracle", "_ref")
D = os.path.join(REF, "data")
PLUGIN = os.path.join(ROOT, "cmusphinx_b200", "_plugin", "libb200_ps_plugin.so")
UTTS = ["440c0201", "441c0201", "442c0201", "443c0201", "444c0201", "446c0201", "447c0201"]


def run(tmp, tag, hmm, preload, extra):
    ctl = os.path.join(tmp, f"{tag}.ctl")
    open(ctl, "w").write("\n".join(UTTS) + "\n")
    hyp, log = os.path.join(tmp, f"{tag}.hyp"), os.path.join(tmp, f"{tag}.log")
    cmd = [os.path.join(REF, "pocketsphinx_batch"), "-hmm", os.path.join(D, "hmm", hmm), "-lm",
           os.path.join(D, "lm", "wsj0vp.5000.DMP"), "-dict", os.path.join(D, "lm", "cmu07a.dic"), "-ctl", ctl,
           "-cepdir", os.path.join(D, "test", "wsj"), "-cepext", ".mfc", "-hyp", hyp, "-logfn", log] + extra
    env = dict(os.environ, LD_LIBRARY_PATH=REF + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    if preload:
        env["LD_PRELOAD"] = PLUGIN
    subprocess.run(cmd, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=1800)
    txt = open(log, errors="replace").read()
    m = re.findall(r"TOTAL\s+([\d.]+) seconds speech, ([\d.]+) seconds CPU, ([\d.]+) seconds wall", txt)
    x = re.findall(r"AVERAGE\s+([\d.]+) xRT \(CPU\), ([\d.]+) xRT \(elapsed\)", txt)
    speech, cpu, wall = (float(v) for v in m[-1]) if m else (None, None, None)
    return {"speech_s": speech, "cpu_s": cpu, "wall_s": wall, "xrt_cpu": float(x[-1][0]) if x else None,
            "xrt_elapsed": float(x[-1][1]) if x else None, "hyp": open(hyp).read()}


def run_sharded(tmp, tag, hmm, ctl_lines, cepdir, cepext, extra, preload, procs):
    """The reference's own batch sharding (-ctloffset/-ctlcount, batch.c:560-640): `procs` decoder
    processes over one control file; returns wall seconds and the concatenated hypothesis lines."""
    import time
    ctl = os.path.join(tmp, f"{tag}.ctl")
    open(ctl, "w").write("\n".join(ctl_lines) + "\n")
    n = len(ctl_lines)
    per = (n + procs - 1) // procs
    env = dict(os.environ, LD_LIBRARY_PATH=REF + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    if preload:
        env["LD_PRELOAD"] = PLUGIN
    ps, hyps = [], []
    t0 = time.perf_counter()
    for i in range(procs):
        if i * per >= n:
            break
        hyp = os.path.join(tmp, f"{tag}.{i}.hyp")
        hyps.append(hyp)
        cmd = [os.path.join(REF, "pocketsphinx_batch"), "-hmm", os.path.join(D, "hmm", hmm), "-lm",
               os.path.join(D, "lm", "wsj0vp.5000.DMP"), "-dict", os.path.join(D, "lm", "cmu07a.dic"), "-ctl", ctl,
               "-ctloffset", str(i * per), "-ctlcount", str(per), "-cepdir", cepdir, "-cepext", cepext, "-hyp", hyp,
               "-logfn", os.path.join(tmp, f"{tag}.{i}.log")] + extra
        ps.append(subprocess.Popen(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    for p in ps:
        if p.wait(timeout=1800) != 0:
            raise RuntimeError(f"decoder process failed ({tag})")
    wall = time.perf_counter() - t0
    lines = []
    for h in hyps:
        lines += open(h).read().splitlines()
    return wall, lines


def batch_pipeline(replicas, procs):
    """BASELINE configs[4] scaled to what the tree bundles: the 7 WSJ utterances x `replicas`,
    whole-box wall-clock xRT with `procs` decoder processes, three ways:
      cpu     the reference decodes the cepstra (GMM + search on the host cores)
      plugin  same processes with the GPU plug-in LD_PRELOADed (scoring on one shared GPU)
      senin   one GPU stage for the whole batch (features + scoring of all frames in a few
              launches, .sen files) followed by search-only decoder processes (-senin yes)"""
    import time
    import numpy as np
    import cmusphinx_b200 as b

    def read_mfc(path):                      # batch.c:185-233: int32 count, float32 data, either byte order
        n = int(np.fromfile(path, dtype="<i4", count=1)[0])
        data = np.fromfile(path, dtype="<f4", offset=4)
        if data.size != n:
            data = np.fromfile(path, dtype=">f4", offset=4).astype(np.float32)
        return data.reshape(-1, 13).astype(np.float32)
    res = {"utterances": 7 * replicas, "decoder_processes": procs, "models": []}
    with tempfile.TemporaryDirectory() as tmp:
        cepdir = os.path.join(tmp, "cep")
        os.makedirs(cepdir)
        names, speech_frames = [], 0
        ceps = {}
        for r in range(replicas):
            for u in UTTS:
                nm = f"{u}_{r:03d}"
                os.symlink(os.path.join(D, "test", "wsj", u + ".mfc"), os.path.join(cepdir, nm + ".mfc"))
                names.append(nm)
                if u not in ceps:
                    ceps[u] = read_mfc(os.path.join(D, "test", "wsj", u + ".mfc"))
                speech_frames += ceps[u].shape[0]
        speech_s = speech_frames / 100.0
        res["speech_s"] = speech_s
        for hmm, kind in (("ptm", 1), ("hub4wsj_sc_8k", 2)):
            w_cpu, h_cpu = run_sharded(tmp, f"cpu_{hmm}", hmm, names, cepdir, ".mfc", [], False, procs)
            # Every decoder process creates its own CUDA context on the one GPU; since round 2 a frame_eval
            # of the ms / s2_semi back-ends costs no GPU work (the utterance's rows are on the host), so
            # what the plug-in arm pays per process is that start-up.  Two shapes: one process per core,
            # and 4 processes (fewer contexts, longer queues).
            plg_procs = procs
            w_plg, h_plg = run_sharded(tmp, f"plg_{hmm}", hmm, names, cepdir, ".mfc", [], True, plg_procs)
            w_plg4, h_plg4 = run_sharded(tmp, f"plg4_{hmm}", hmm, names, cepdir, ".mfc", [], True, min(4, procs))
            # start-up alone: one process, one utterance, both arms
            w_one_cpu, _ = run_sharded(tmp, f"one_cpu_{hmm}", hmm, names[:1], cepdir, ".mfc", [], False, 1)
            w_one_plg, _ = run_sharded(tmp, f"one_plg_{hmm}", hmm, names[:1], cepdir, ".mfc", [], True, 1)
            # ---- GPU stage for the whole batch
            t0 = time.perf_counter()
            mm = b.mdef_maps(os.path.join(D, "hmm", hmm, "mdef"))
            s2c = mm["sen2cimap"].astype(np.uint8) if kind == 1 else None   # ptm: senone -> codebook
            n_sen = mm["n_sen"]
            m = b.tied_from_model_dir(os.path.join(D, "hmm", hmm), n_sen, sen2cb=s2c, topn=4)
            t_load = time.perf_counter() - t0
            t0 = time.perf_counter()
            sendir = os.path.join(tmp, f"sen_{hmm}")
            os.makedirs(sendir)
            order = [n.rsplit("_", 1)[0] for n in names]
            cep_all = np.concatenate([ceps[u] for u in order])
            off = np.concatenate([[0], np.cumsum([ceps[u].shape[0] for u in order])]).astype(np.int32)
            feat = b.feat_1s_c_d_dd(cep_all, off)
            scores = m.score(feat)
            t_score = time.perf_counter() - t0
            t0 = time.perf_counter()
            for i, nm in enumerate(names):
                b.sen_write(os.path.join(sendir, nm + ".sen"), scores[off[i]:off[i + 1]])
            t_write = time.perf_counter() - t0
            m.free()
            w_sen, h_sen = run_sharded(tmp, f"sen_{hmm}", hmm, names, sendir, ".sen", ["-senin", "yes"], False, procs)
            words = lambda hs: [l.rsplit("(", 1)[0] for l in hs]
            res["models"].append({
                "model": hmm, "frames": int(off[-1]),
                "cpu": {"wall_s": w_cpu, "xrt_wall": w_cpu / speech_s},
                "plugin": {"decoder_processes": plg_procs, "wall_s": w_plg, "xrt_wall": w_plg / speech_s, "identical_hyp_lines": h_plg == h_cpu,
                           "identical_words": words(h_plg) == words(h_cpu),
                           "with_4_processes": {"wall_s": w_plg4, "identical_words": words(h_plg4) == words(h_cpu)},
                           "one_process_one_utterance_wall_s": {"cpu": w_one_cpu, "plugin": w_one_plg,
                                                                "note": "the difference is CUDA context creation + parameter upload"}},
                "senin_pipeline": {"gpu_stage_s": t_score, "sen_write_s": t_write, "search_wall_s": w_sen,
                                   "wall_s": t_score + t_write + w_sen, "xrt_wall": (t_score + t_write + w_sen) / speech_s,
                                   "model_load_s_excluded": t_load, "identical_words": words(h_sen) == words(h_cpu),
                                   "gpu_frame_senones_per_s": int(off[-1]) * n_sen / t_score}})
    return res


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=16, help="copies of the 7 WSJ utterances in the batch arm (0 = skip)")
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--mps", action="store_true",
                    help="run under a CUDA MPS daemon (started and stopped here): the decoder processes of the plug-in arm "
                         "then share the GPU concurrently instead of being time-sliced")
    args = ap.parse_args()
    out = {"metric": "batch_decode_xRT", "unit": "xRT (lower is better)", "higher_is_better": False, "data": "bundled WSJ .mfc x7",
           "runs": []}
    with tempfile.TemporaryDirectory() as tmp:
        for hmm in ("hub4wsj_sc_8k", "ptm"):
            for extra, name in (([], "default 3-pass"), (["-fwdflat", "no", "-bestpath", "no"], "fwdtree only")):
                cpu = run(tmp, "cpu", hmm, False, extra)
                gpu = run(tmp, "gpu", hmm, True, extra)
                out["runs"].append({"model": hmm, "passes": name,
                                    "cpu_reference": {k: v for k, v in cpu.items() if k != "hyp"},
                                    "gpu_plugin": {k: v for k, v in gpu.items() if k != "hyp"},
                                    "identical_words": [l.rsplit("(", 1)[0] for l in cpu["hyp"].splitlines()] ==
                                                       [l.rsplit("(", 1)[0] for l in gpu["hyp"].splitlines()],
                                    "identical_path_scores": cpu["hyp"] == gpu["hyp"]})
    mps = None
    if args.mps:
        import shutil
        if shutil.which("nvidia-cuda-mps-control"):
            os.environ["CUDA_MPS_PIPE_DIRECTORY"] = "/tmp/b200_mps_pipe"
            os.environ["CUDA_MPS_LOG_DIRECTORY"] = "/tmp/b200_mps_log"
            for d in (os.environ["CUDA_MPS_PIPE_DIRECTORY"], os.environ["CUDA_MPS_LOG_DIRECTORY"]):
                os.makedirs(d, exist_ok=True)
            mps = subprocess.run(["nvidia-cuda-mps-control", "-d"], timeout=60).returncode == 0
        out["mps"] = bool(mps)
    try:
        if args.replicas > 0:
            out["batch"] = batch_pipeline(args.replicas, args.procs)
    finally:
        if mps:
            subprocess.run(["nvidia-cuda-mps-control"], input=b"quit\n", timeout=60)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
