#!/usr/bin/env python
"""bench_prune.py -- the prune / phone-transition stage of the forward tree search
(prune_root_chan + prune_nonroot_chan, ngram_search_fwdtree.c:714-869) on one B200:
the lexical tree and the recorded frames of tests/golden/fwdtree_prune.npz (hub4wsj_sc_8k + wsj0vp.5000: 443
roots, 14 331 channels; REAL pre-prune states of the reference decoder), replicated to a batch of utterances,
through b200_fwdtree_prune_dev with every array resident in HBM.  Secondary benchmark of `python bench.py`
(`secondary.prune`); stand-alone it prints one JSON line.

  python bench_prune.py [--utts B] [--steps K]

Unit: walk elements (active roots + active-list entries) per second.  HBM model per element: list id 4 + own
{bestscore, out_score, out_history, in-score, frame} 20 + parent {position, bestscore, out_score} 12 + the
kernel's scratch (position, count, flag; written and read) 26 + frame stamp / list append 8 = 70 B, plus 22 B per
child of a surviving element {id, position, frame, in-score, bestscore, decision byte twice}.
The state is restored from a pristine device copy before every timed launch (outside the timed events)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
TOPO_KEYS = ("child_off", "child", "ciphone", "pw_off", "pw_wid", "pw_lastphone")
PAR = ("frame", "best_score", "beam", "pbeam", "lpbeam", "pip", "nwpen", "has_pls")


def load():
    z = np.load(os.path.join(ROOT, "tests", "golden", "fwdtree_prune.npz"))
    topo = {k: z["topo_" + k] for k in TOPO_KEYS}
    topo["n_root"], topo["n_chan"], topo["n_ci"] = int(z["topo_n_root"]), int(z["topo_n_chan"]), 50
    cases = []
    for name in z["cases"]:
        k = str(name) + "_"
        cases.append(dict(par=z[k + "par"], pls_pen=z[k + "pls_pen"], acl=z[k + "acl"], state=z[k + "state"],
                          n_nacl=len(z[k + "nacl"]), n_cand=len(z[k + "cand"])))
    return topo, [c for c in cases if len(c["acl"]) > 500]


def cpu_port(topo, cases, budget_s=4.0):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    par = [dict(zip(PAR, (int(x) for x in c["par"]))) for c in cases]
    soa = [orc.prune_rows_to_soa(c["state"]) for c in cases]
    n, el, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        for c, p, s in zip(cases, par, soa):
            orc.port_fwdtree_prune(topo, p, c["pls_pen"], c["acl"], s)
            el += len(c["acl"]) + int((c["state"][:topo["n_root"], 9] >= p["frame"]).sum())
            n += 1
    dt = time.perf_counter() - t0
    return {"value": el / dt, "unit": "walk elements/s", "cores": 1, "kind": "port",
            "sample": f"{n} frames ({dt:.1f} s) of the same recorded states through oracle/sphinx_oracle.c orc_fwdtree_prune "
                      "(pinned to the reference's own functions on every frame of three decodes), one core, incl. the "
                      "ctypes copy of the state per call"}


def run(utts=128, steps=30, warmup=5, cpu=True):
    import torch
    import cmusphinx_b200 as b
    from cmusphinx_b200 import _lib
    assert b.device_count() > 0
    topo, cases = load()
    nc, nr, ne = topo["n_chan"], topo["n_root"], 3
    tree = b.ChanTree(nr, nc, *[topo[k] for k in TOPO_KEYS], topo["n_ci"])
    B = utts
    pick = [cases[u % len(cases)] for u in range(B)]
    rows = np.stack([c["state"] for c in pick])                       # [B][nc][10]
    cap, ccap = nc - nr, tree.cand_cap
    acl = np.zeros((B, cap), np.int32)
    for u, c in enumerate(pick):
        acl[u, :len(c["acl"])] = c["acl"]
    dev = torch.device("cuda", 0)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    # state-major over the whole batch: [st][B * nc]
    pristine = dict(score=T(rows[:, :, 0:3].transpose(2, 0, 1).reshape(3, B * nc)), history=T(rows[:, :, 3:6].transpose(2, 0, 1).reshape(3, B * nc)),
                    out_score=T(rows[:, :, 6].reshape(-1)), out_history=T(rows[:, :, 7].reshape(-1)),
                    bestscore=T(rows[:, :, 8].reshape(-1)), frame=T(rows[:, :, 9].reshape(-1)))
    work = {k: v.clone() for k, v in pristine.items()}
    d_par = T(np.stack([c["par"] for c in pick]).astype(np.int32))
    d_pen = T(np.stack([c["pls_pen"] for c in pick]).astype(np.int32))
    d_acl, d_nact = T(acl), T(np.array([len(c["acl"]) for c in pick], np.int32))
    d_nacl, d_nn = torch.zeros((B, cap), dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)
    d_cand, d_ncand = torch.zeros((B, ccap, 3), dtype=torch.int32, device=dev), torch.zeros(B, dtype=torch.int32, device=dev)
    p = _lib.PruneDev()
    for k in ("score", "history", "out_score", "out_history", "bestscore", "frame"):
        setattr(p, k, work[k].data_ptr())
    p.state_stride = B * nc
    p.par, p.pls_pen, p.acl, p.n_act, p.list_cap = d_par.data_ptr(), d_pen.data_ptr(), d_acl.data_ptr(), d_nact.data_ptr(), cap
    p.nacl, p.n_nacl, p.cand, p.n_cand, p.cand_cap = d_nacl.data_ptr(), d_nn.data_ptr(), d_cand.data_ptr(), d_ncand.data_ptr(), ccap
    side = torch.cuda.Stream()
    ms = []
    n0 = b.launch_count()
    with torch.cuda.stream(side):
        for it in range(warmup + steps):
            for k in work:
                work[k].copy_(pristine[k])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            _lib.check(b.lib.b200_fwdtree_prune_dev(tree._h, B, C.byref(p), side.cuda_stream), "prune_dev")
            e1.record(side)
            side.synchronize()
            if it >= warmup:
                ms.append(e0.elapsed_time(e1))
    launches = b.launch_count() - n0
    # the batch reproduces the recorded outcome of every frame
    nn, ncd = d_nn.cpu().numpy(), d_ncand.cpu().numpy()
    assert all(nn[u] == c["n_nacl"] and ncd[u] == c["n_cand"] for u, c in enumerate(pick)), "prune result differs from the recording"
    elements = sum(len(c["acl"]) + int((c["state"][:nr, 9] >= int(c["par"][0])).sum()) for c in pick)
    thresh = [int(c["par"][1]) + int(c["par"][2]) for c in pick]
    nchild = np.diff(topo["child_off"])
    edges = 0
    for c, th in zip(pick, thresh):
        ids = np.concatenate([np.nonzero(c["state"][:nr, 9] >= int(c["par"][0]))[0], c["acl"]])
        edges += int(nchild[ids[c["state"][ids, 8] > th]].sum())
    t = float(np.mean(ms)) * 1e-3
    bytes_model = 70 * elements + 22 * edges
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = None
    for k in ("hbm_copy_gbs_burst", "hbm_gbs_burst", "hbm_copy_gbs", "hbm_gbs"):
        if isinstance(peaks.get(k), (int, float)):
            peak = float(peaks[k]); break
    if peak is None:
        peak = 6530.0
    out = {"metric": "fwdtree prune walk elements/s", "value": elements / t, "unit": "walk elements/s", "ms_per_step": t * 1e3,
           "us_per_utterance_frame": t * 1e6 / B, "dtype": "int32", "gpu_launches": int(launches),
           "config": {"workload": f"prune_root_chan + prune_nonroot_chan, {B} utterances x one frame, recorded decoder states "
                                  f"(hub4wsj_sc_8k / wsj0vp.5000 tree: {nr} roots, {nc} channels)",
                      "elements_per_step": elements, "child_edges_per_step": edges, "steps": steps, "warmup": warmup,
                      "l2": "state restored from a pristine copy (a 2 x 73 MB stream through L2) before every timed launch"},
           "roofline": {"bound": "hbm", "achieved": bytes_model / t / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": bytes_model / t / 1e9 / peak, "traffic": None,
                        "bytes_model": "70 B per walk element + 22 B per child of a surviving element (bench_prune.py docstring)",
                        "note": "one CTA per utterance: latency bound by the three dependent phases, not by HBM"}}
    # ---- the stage in front of it on the same lists: eval_root_chan + eval_nonroot_chan (b200_hmm_eval_list_dev) on the
    # resident population of an HmmContext (synthetic senone ids / transition matrices for the tree's channels)
    try:
        from cmusphinx_b200 import synth
        from cmusphinx_b200.engine import LOGBASE
        n_sen, n_tmat, n_sseq = 5000, 50, 27000
        tp = b.tmat_quantize(synth.bakis_tmat(n_tmat, ne, 7), 1e-4, LOGBASE)
        dd = synth.hmm_population(B * nc, ne, n_sen, n_tmat, n_sseq, seed=42, mpx_fraction=0.0)
        pop = b.HmmPopulation(B * nc, ne)
        pop.score[:] = rows[:, :, 0:3].transpose(2, 0, 1).reshape(3, B * nc)
        pop.history[:] = rows[:, :, 3:6].transpose(2, 0, 1).reshape(3, B * nc)
        pop.out_score[:], pop.out_history[:], pop.bestscore[:] = rows[:, :, 6].reshape(-1), rows[:, :, 7].reshape(-1), rows[:, :, 8].reshape(-1)
        pop.senid[:], pop.tmatid[:] = dd["senid"].T, dd["tmatid"]
        mp = np.zeros((B, nc), np.uint8); mp[:, :nr] = 1                    # the roots are multiplex HMMs
        pop.mpx[:] = mp.reshape(-1)
        sid = pop.senid.reshape(ne, B, nc); sid[:, :, :nr] = np.random.default_rng(3).integers(0, n_sseq, (ne, B, nr))
        ctx = b.HmmContext(ne, tp, dd["sseq"], n_sen)
        d_sen = T(synth.senscr_frames(B, n_sen, 99))
        d_best = torch.zeros(B, dtype=torch.int32, device=dev)
        ems = []
        n_eval = sum(len(c["acl"]) + int((c["state"][:nr, 9] == int(c["par"][0])).sum()) for c in pick)
        for it in range(3 + 10):
            ctx.upload(pop)
            ctx.set_utts(np.arange(B + 1, dtype=np.int32) * nc)
            with torch.cuda.stream(side):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(side)
                _lib.check(b.lib.b200_hmm_eval_list_dev(ctx._h, nr, nc, pristine["frame"].data_ptr(), d_par.data_ptr(), d_acl.data_ptr(),
                                                         d_nact.data_ptr(), cap, d_sen.data_ptr(), d_best.data_ptr(), side.cuda_stream), "eval_list")
                e1.record(side)
                side.synchronize()
            if it >= 3:
                ems.append(e0.elapsed_time(e1))
        te = float(np.mean(ems)) * 1e-3
        out["eval_list"] = {"what": "eval_root_chan + eval_nonroot_chan on the same lists (b200_hmm_eval_list_dev), resident population of "
                                    f"{B} x {nc} channels", "hmms_per_step": n_eval, "ms_per_step": te * 1e3,
                            "value": n_eval / te, "unit": "HMM*frames/s",
                            "roofline": {"bound": "hbm", "achieved": 72 * n_eval / te / 1e9, "peak": peak, "unit": "GB/s",
                                         "frac": 72 * n_eval / te / 1e9 / peak,
                                         "bytes_model": "72 B per evaluated HMM (68 B of hmm_t fields + the list id), gathered by index"}}
        out["tree_frame_us_per_utterance"] = (t + te) * 1e6 / B
        ctx.free()
    except Exception as ex:
        out["eval_list"] = {"value": None, "note": f"failed: {ex!r}"}
    if cpu:
        try:
            out["cpu_baseline"] = cpu_port(topo, cases)
        except Exception as ex:
            out["cpu_baseline"] = {"value": None, "note": f"failed: {ex!r}"}
    tree.free()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=128)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    print(json.dumps(run(a.utts, a.steps, cpu=not a.no_cpu)))
