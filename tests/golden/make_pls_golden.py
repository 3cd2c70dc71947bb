"""Golden vectors of the phone-loop look-ahead search (phone_loop_search_step, pocketsphinx/src/libpocketsphinx/
phone_loop_search.c:253-291) from a REAL decode of the unmodified reference: oracle/_ref/libref_pls_trace.so (the
reference's own source file compiled in place with one macro hook, oracle/ref_pls_trace.c) preloaded into
pocketsphinx_batch, numbers.raw with -pl_window 5.  Run in the build container:
    python tests/golden/make_pls_golden.py
Writes tests/golden/phone_loop.npz: the transition table and 60 consecutive (before, senone scores, after) frames;
the oracle port is checked against ALL recorded frames first."""
import os
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import orc  # noqa: E402
from fwdtree_trace import read_pls_trace  # noqa: E402

KEYS = ("score", "history", "out_score", "out_history", "bestscore", "frame_of")


def record(extra, path, lo):
    D, R = orc.DATA_DIR, orc.REF_DIR
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "a.ctl"), "w").write("numbers\n")
        out = os.path.join(tmp, "pls.bin")
        env = dict(os.environ, LD_LIBRARY_PATH=R, LD_PRELOAD=os.path.join(R, "libref_pls_trace.so"), B200_PLS_TRACE=out)
        subprocess.run([os.path.join(R, "pocketsphinx_batch"), "-hmm", os.path.join(D, "hmm", "hub4wsj_sc_8k"), "-lm",
                        os.path.join(D, "lm", "wsj0vp.5000.DMP"), "-dict", os.path.join(D, "lm", "cmu07a.dic"), "-ctl",
                        os.path.join(tmp, "a.ctl"), "-cepdir", os.path.join(D, "test"), "-cepext", ".raw", "-adcin", "yes", "-samprate",
                        "16000", "-hyp", os.path.join(tmp, "a.hyp"), "-logfn", os.path.join(tmp, "a.log"), "-fwdflat", "no", "-bestpath",
                        "no", "-pl_window", "5"] + extra, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        tp, recs = read_pls_trace(out)
    n = 0
    for a, b in zip(recs[:-1], recs[1:]):
        if b["frame"] != a["frame"] + 1:
            continue
        r = orc.port_phone_loop_step(tp, a)
        assert all(np.array_equal(r[k], b[k]) for k in KEYS) and r["best_score"] == b["best_score"], a["frame"]
        n += 1
    print(n, "frames: port == reference")
    sel = recs[lo:lo + 61]
    assert all(b["frame"] == a["frame"] + 1 for a, b in zip(sel[:-1], sel[1:]))
    g = dict(tp=tp, frame0=np.int32(sel[0]["frame"]), par=np.array([sel[0]["beam"], sel[0]["pbeam"], sel[0]["pip"]], np.int32),
             best_score=np.array([r["best_score"] for r in sel], np.int32), tmatid=sel[0]["tmatid"].astype(np.int16),
             senscr=np.stack([r["senscr"] for r in sel]).astype(np.int16))
    for k in KEYS:
        g[k] = np.stack([r[k] for r in sel]).astype(np.int32)
    n_pruned = sum(int(((a["frame_of"] >= a["frame"]) & (b["bestscore"] == -0x20000000)).sum()) for a, b in zip(sel[:-1], sel[1:]))
    n_idle = sum(int((r["frame_of"] < r["frame"]).sum()) for r in sel)
    print("in the selection:", n_pruned, "phones pruned,", n_idle, "phone-frames inactive")
    np.savez_compressed(path, **g)
    print(path, os.path.getsize(path), "bytes")


def main():
    here = os.path.dirname(__file__)
    # the default beams (1e-10, unshifted log: -230231) never prune a phone; the second recording uses beams
    # narrow enough that prune_hmms clears phones and phone_transition re-enters them
    record([], os.path.join(here, "phone_loop.npz"), 100)
    record(["-pl_beam", "0.985", "-pl_pbeam", "0.99"], os.path.join(here, "phone_loop_tight.npz"), 100)


if __name__ == "__main__":
    main()
