#!/usr/bin/env python
"""bench_e2e_decode.py -- BASELINE.json configs[4]: end-to-end
batch decode of synthetic 5 s utterances against a 20 000-word trigram, utterances sharded
over the GPUs of one box, whole-box xRT next to the reference's own pocketsphinx_batch on all
host cores (its xRT definition: batch.c:774-776 = elapsed / seconds of speech).

    python bench.py --workload e2e_decode [--gpus N] [--utts-per-gpu U]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --workload e2e_decode --gpus N

What the tree does not have is substituted, and named in the JSON line:
  * WSJ 20k trigram + lexicon  -> a synthetic 20 000-word trigram (Zipf unigrams, 200k bigrams,
    100k trigrams, seeded) over 20 000 words of the bundled cmu07a.dic with their real
    pronunciations; written as ARPA and converted to the decoder's .DMP by the reference's own
    ngram_model_write (oracle/_ref/libsphinxbase.so -- data preparation, not scoring);
  * the config-2 synthetic continuous model has no model definition / dictionary, so the
    acoustic model is the bundled hub4wsj_sc_8k (5150 senones, semi-continuous);
  * utterances: 500-frame windows of the bundled WSJ cepstra at seeded offsets plus seeded
    noise (speech-like, so the search sees realistic beams).
GPU arm, per rank: utterances r, r + N, ... (shard.shard_strided = -ctloffset r -ctlincr N) in
waves: one batched feature + scoring pass on the GPU, .sen files, then search-only decoder
processes (`-senin yes`; the lextree / LM search stays on the host, out of scope).  No
collective on the path; ranks meet at a barrier at both ends and the slowest rank's wall is
the job's.  CPU arm (rank 0, after the timed region): the unmodified decoder, one process per
host core, on a bounded sample of the same utterances; hypotheses of the sample must be
identical in both arms.
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")
D = os.path.join(REF, "data")
HMM = os.path.join(D, "hmm", "hub4wsj_sc_8k")
UTTS = ["440c0201", "441c0201", "442c0201", "443c0201", "444c0201", "446c0201", "447c0201"]
N_WORDS, N_BIGRAMS, N_TRIGRAMS, FRAMES = 20_000, 200_000, 100_000, 500


def read_mfc(path):
    n = int(np.fromfile(path, dtype="<i4", count=1)[0])
    data = np.fromfile(path, dtype="<f4", offset=4)
    if data.size != n:
        data = np.fromfile(path, dtype=">f4", offset=4).astype(np.float32)
    return data.reshape(-1, 13).astype(np.float32)


def build_lm_and_dict(out_dir, seed=20):
    """-> (dict path, DMP path).  Deterministic; every rank builds the same files."""
    rng = np.random.default_rng(seed)
    entries = {}
    for line in open(os.path.join(D, "lm", "cmu07a.dic"), errors="replace"):
        p = line.split()
        if len(p) >= 2 and re.fullmatch(r"[a-z]+", p[0]) and p[0] not in entries:
            entries[p[0]] = " ".join(p[1:])
    allw = sorted(entries)
    words = sorted(rng.choice(len(allw), N_WORDS, replace=False))
    words = [allw[i] for i in words]
    dic = os.path.join(out_dir, "synth20k.dic")
    with open(dic, "w") as fh:
        for w in words:
            fh.write(f"{w}\t{entries[w]}\n")
    V = len(words)
    zipf = 1.0 / (np.arange(V) + 10.0)
    perm = rng.permutation(V)                      # frequency rank -> word id
    p1 = np.empty(V); p1[perm] = zipf / zipf.sum()
    cdf = np.cumsum(p1)
    draw = lambda n: np.minimum(np.searchsorted(cdf, rng.random(n)), V - 1)
    bg = np.unique(np.stack([draw(N_BIGRAMS * 2), draw(N_BIGRAMS * 2)], 1), axis=0)[:N_BIGRAMS * 2]
    bg = bg[rng.permutation(len(bg))[:N_BIGRAMS]]
    bg = bg[np.lexsort((bg[:, 1], bg[:, 0]))]
    succ = {}
    for a, b_ in bg:
        succ.setdefault(int(a), []).append(int(b_))
    tg = set()
    pick = rng.integers(0, len(bg), N_TRIGRAMS * 3)
    for k in pick:
        a, b_ = int(bg[k, 0]), int(bg[k, 1])
        s = succ.get(b_)
        if s:
            tg.add((a, b_, s[int(rng.integers(0, len(s)))]))
        if len(tg) >= N_TRIGRAMS:
            break
    tg = sorted(tg)
    arpa = os.path.join(out_dir, "synth20k.arpa")
    with open(arpa, "w") as fh:
        fh.write(f"\\data\\\nngram 1={V + 2}\nngram 2={len(bg)}\nngram 3={len(tg)}\n\n\\1-grams:\n")
        fh.write("-99.0000 <s> -0.3000\n-1.5000 </s> 0.0000\n")
        for i, w in enumerate(words):
            fh.write(f"{np.log10(p1[i]) - 0.05:.4f} {w} -0.3000\n")
        fh.write("\n\\2-grams:\n")
        lp2 = -rng.uniform(0.5, 2.5, len(bg))
        for (a, b_), lp in zip(bg, lp2):
            fh.write(f"{lp:.4f} {words[a]} {words[b_]} -0.2000\n")
        fh.write("\n\\3-grams:\n")
        lp3 = -rng.uniform(0.3, 1.5, len(tg))
        for (a, b_, c), lp in zip(tg, lp3):
            fh.write(f"{lp:.4f} {words[a]} {words[b_]} {words[c]}\n")
        fh.write("\n\\end\\\n")
    dmp = os.path.join(out_dir, "synth20k.DMP")
    sb = C.CDLL(os.path.join(REF, "libsphinxbase.so"))
    sb.logmath_init.restype = C.c_void_p
    sb.logmath_init.argtypes = [C.c_double, C.c_int, C.c_int]
    sb.ngram_model_read.restype = C.c_void_p
    sb.ngram_model_read.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
    sb.ngram_model_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    sb.err_set_logfp.argtypes = [C.c_void_p]
    sb.err_set_logfp(None)
    lm = sb.logmath_init(1.0001, 0, 0)
    m = sb.ngram_model_read(None, arpa.encode(), 1, lm)      # NGRAM_ARPA
    assert m, "the reference's ARPA reader rejected the synthetic LM"
    assert sb.ngram_model_write(m, dmp.encode(), 2) == 0     # NGRAM_DMP
    return dic, dmp, {"words": V, "bigrams": int(len(bg)), "trigrams": int(len(tg))}


def synth_utterances(n, seed=77):
    """n x FRAMES x 13 cepstra: windows of the concatenated bundled WSJ cepstra + noise."""
    rng = np.random.default_rng(seed)
    base = np.concatenate([read_mfc(os.path.join(D, "test", "wsj", u + ".mfc")) for u in UTTS])
    sd = base.std(0)
    offs = rng.integers(0, base.shape[0] - FRAMES, n)
    return base, sd, offs


def utterance(base, sd, offs, i):
    rng = np.random.default_rng(1000003 * 7 + i)
    return (base[offs[i]:offs[i] + FRAMES] + rng.standard_normal((FRAMES, 13)).astype(np.float32) * (0.05 * sd)).astype(np.float32)


def run_decoders(tag, tmp, ctl_names, cepdir, cepext, dic, dmp, extra, procs, passes):
    """`procs` decoder processes over one control file (the reference's own -ctloffset / -ctlcount
    sharding); returns (wall seconds, {utt: hypothesis line})."""
    ctl = os.path.join(tmp, f"{tag}.ctl")
    open(ctl, "w").write("\n".join(ctl_names) + "\n")
    n = len(ctl_names)
    procs = max(1, min(procs, n))
    per = (n + procs - 1) // procs
    env = dict(os.environ, LD_LIBRARY_PATH=REF + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    env.pop("LD_PRELOAD", None)
    ps, hyps = [], []
    t0 = time.perf_counter()
    for i in range(procs):
        if i * per >= n:
            break
        hyp = os.path.join(tmp, f"{tag}.{i}.hyp")
        hyps.append(hyp)
        cmd = [os.path.join(REF, "pocketsphinx_batch"), "-hmm", HMM, "-lm", dmp, "-dict", dic, "-ctl", ctl,
               "-ctloffset", str(i * per), "-ctlcount", str(per), "-cepdir", cepdir, "-cepext", cepext, "-hyp", hyp,
               "-logfn", os.path.join(tmp, f"{tag}.{i}.log")] + extra + passes
        ps.append(subprocess.Popen(cmd, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    for p in ps:
        if p.wait(timeout=3600) != 0:
            raise RuntimeError(f"decoder process failed ({tag}); see {tmp}")
    wall = time.perf_counter() - t0
    out = {}
    for h in hyps:
        for line in open(h).read().splitlines():
            m = re.match(r"(.*)\((\S+) (-?\d+)\)$", line)
            if m:
                out[m.group(2)] = line
    return wall, out


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--utts-per-gpu", type=int, default=1024)
    ap.add_argument("--wave", type=int, default=128, help="utterances per GPU stage / search wave")
    ap.add_argument("--cpu-sample", type=int, default=256, help="utterances the reference arm decodes (rank 0)")
    ap.add_argument("--passes", default="fwdtree", choices=["fwdtree", "3pass"])
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    args = ap.parse_args(argv)
    import torch
    import torch.distributed as dist
    import cmusphinx_b200 as b
    from cmusphinx_b200 import shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available() and b.device_count() > local
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)
    procs_rank = max(1, cores // world)
    passes = ["-fwdflat", "no", "-bestpath", "no"] if args.passes == "fwdtree" else []
    n_total = args.utts_per_gpu * world
    tmpd = tempfile.TemporaryDirectory(prefix=f"b200cfg5_r{rank}_")
    tmp = tmpd.name
    dic, dmp, lm_info = build_lm_and_dict(tmp)
    base, sd, offs = synth_utterances(n_total)
    mine = list(shard.shard_strided(n_total, rank, world))
    mm = b.mdef_maps(os.path.join(HMM, "mdef"))
    n_sen = mm["n_sen"]
    m = b.tied_from_model_dir(HMM, n_sen, sen2cb=None, topn=4, device=local)
    # warm-up: one small wave through the GPU stage (allocations, kernel attributes)
    w = b.feat_1s_c_d_dd(np.concatenate([utterance(base, sd, offs, mine[0])] * 2), np.array([0, FRAMES, 2 * FRAMES], np.int32), device=local)
    m.score(w)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    t_start = time.perf_counter()
    t_gpu = t_write = t_search = 0.0
    hyp_all = {}
    sendir = os.path.join(tmp, "sen")
    os.makedirs(sendir)
    for w0 in range(0, len(mine), args.wave):
        ids = mine[w0:w0 + args.wave]
        t0 = time.perf_counter()
        cep = np.concatenate([utterance(base, sd, offs, i) for i in ids])
        off = (np.arange(len(ids) + 1) * FRAMES).astype(np.int32)
        feat = b.feat_1s_c_d_dd(cep, off, device=local)      # 13 -> 39 dims on the GPU (CMN per utterance)
        scores = m.score(feat)                               # [frames][n_sen] int16, H2D + D2H inside
        t1 = time.perf_counter()
        names = [f"u{i:07d}" for i in ids]
        for k, nm in enumerate(names):
            b.sen_write(os.path.join(sendir, nm + ".sen"), scores[off[k]:off[k + 1]])
        t2 = time.perf_counter()
        _, hy = run_decoders(f"g{w0}", tmp, names, sendir, ".sen", dic, dmp, ["-senin", "yes"], procs_rank, passes)
        t3 = time.perf_counter()
        for nm in names:
            os.remove(os.path.join(sendir, nm + ".sen"))
        hyp_all.update(hy)
        t_gpu += t1 - t0; t_write += t2 - t1; t_search += t3 - t2
    barrier()
    wall = time.perf_counter() - t_start
    tw = torch.tensor([wall, t_gpu, t_write, t_search], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    wall, t_gpu, t_write, t_search = (float(x) for x in tw.tolist())
    speech_s = n_total * FRAMES / 100.0
    if rank == 0:
        # ---- reference arm on a bounded sample: rank 0's first utterances, one decoder per host core
        k = min(args.cpu_sample, len(mine))
        cepdir = os.path.join(tmp, "cep")
        os.makedirs(cepdir)
        names = []
        for i in mine[:k]:
            nm = f"u{i:07d}"
            c = utterance(base, sd, offs, i)
            with open(os.path.join(cepdir, nm + ".mfc"), "wb") as fh:
                fh.write(np.int32(c.size).tobytes()); fh.write(c.astype("<f4").tobytes())
            names.append(nm)
        w_cpu, hy_cpu = run_decoders("cpu", tmp, names, cepdir, ".mfc", dic, dmp, [], cores, passes)
        same = sum(1 for nm in names if hy_cpu.get(nm) == hyp_all.get(nm))
        words_same = sum(1 for nm in names if hy_cpu.get(nm, "a").rsplit("(", 1)[0] == hyp_all.get(nm, "b").rsplit("(", 1)[0])
        xrt = wall / speech_s
        xrt_cpu = w_cpu / (k * FRAMES / 100.0)
        line = {
            "metric": "batch_decode_xRT", "value": xrt, "unit": "xRT (elapsed / seconds of speech, whole box)",
            "higher_is_better": False, "n_gpus": world, "steps": 1, "warmup": 0, "ms_per_step": wall * 1e3,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32/int16", "data": "synthetic",
            "config": {"workload": f"end-to-end batch decode, {n_total} synthetic 5 s utterances ({args.utts_per_gpu} per GPU), "
                                   f"hub4wsj_sc_8k + synthetic {lm_info['words']}-word trigram ({lm_info['bigrams']} bigrams, "
                                   f"{lm_info['trigrams']} trigrams), {args.passes}, -senin pipeline (BASELINE configs[4], scaled)",
                       "substitutions": "WSJ-20k LM / lexicon and a model definition for the config-2 synthetic model are "
                                        "not in the tree: synthetic trigram over 20 000 cmu07a words, bundled acoustic model",
                       "sharding": f"utterance i -> rank i % {world} (shard.shard_strided = -ctloffset r -ctlincr N), "
                                   f"{procs_rank} search processes per rank, waves of {args.wave} utterances",
                       "host_cores": cores},
            "speech_seconds": speech_s,
            "extrapolated_wall_s_for_100k_utterances": xrt * 100_000 * FRAMES / 100.0,
            "stage_seconds_max_over_ranks": {"features_and_scoring_gpu_incl_copies": t_gpu, "sen_files": t_write, "search_processes": t_search},
            "scoring_stage_frame_senones_per_s": n_total * FRAMES * n_sen / max(t_gpu, 1e-9),
            "cpu_baseline": {"value": xrt_cpu, "unit": "xRT", "cores": cores, "kind": "reference",
                             "sample": f"{k} of the same utterances, unmodified pocketsphinx_batch, {cores} processes ({w_cpu:.1f} s)"},
            "speedup_vs_reference_xrt": xrt_cpu / xrt,
            "parity_on_sample": {"utterances": k, "identical_hypothesis_lines": same, "identical_words": words_same},
            "gpu_launches": int(b.launch_count()),
        }
        print(json.dumps(line), flush=True)
    m.free()
    tmpd.cleanup()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
