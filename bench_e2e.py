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


def main():
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
    print(json.dumps(out))


if __name__ == "__main__":
    main()
