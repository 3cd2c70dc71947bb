"""Golden vectors of the prune / phone-transition stage (prune_root_chan + prune_nonroot_chan,
pocketsphinx/src/libpocketsphinx/ngram_search_fwdtree.c:714-869) taken from REAL decodes of the
unmodified reference: oracle/_ref/libref_fwdtree_trace.so (the reference's own source file, compiled
in place with two macro hooks, oracle/ref_fwdtree_trace.c) is preloaded into pocketsphinx_batch and
records the lexical tree before and after the two functions.  Run in the build container (needs
oracle/_ref):  python tests/golden/make_fwdtree_golden.py
Writes tests/golden/fwdtree_prune.npz: the tree's topology plus a few frames of goforward.raw
(no look-ahead), numbers.raw (-pl_window 5: phone-loop look-ahead penalties) and something.raw with the
active list shuffled before prune_root_chan (B200_FWDTREE_TRACE_SHUFFLE)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import orc  # noqa: E402
from fwdtree_trace import read_trace  # noqa: E402

PICK = {"goforward": ([], {}, [1, 20, 60, 100, 150, 200, 264]), "numbers": (["-pl_window", "5"], {}, [30, 120, 250, 330]),
        # the active list permuted before the reference's functions see it: pins their dependence on the list order
        "something": (["-pl_window", "2"], {"B200_FWDTREE_TRACE_SHUFFLE": "7"}, [15, 90, 170])}


def trace(utt, extra, env_extra, tmp):
    D, R = orc.DATA_DIR, orc.REF_DIR
    ctl = os.path.join(tmp, utt + ".ctl")
    open(ctl, "w").write(utt + "\n")
    out = os.path.join(tmp, utt + ".trace")
    env = dict(os.environ, LD_LIBRARY_PATH=R, LD_PRELOAD=os.path.join(R, "libref_fwdtree_trace.so"), B200_FWDTREE_TRACE=out, **env_extra)
    subprocess.run([os.path.join(R, "pocketsphinx_batch"), "-hmm", os.path.join(D, "hmm", "hub4wsj_sc_8k"), "-lm",
                    os.path.join(D, "lm", "wsj0vp.5000.DMP"), "-dict", os.path.join(D, "lm", "cmu07a.dic"), "-ctl", ctl,
                    "-cepdir", os.path.join(D, "test"), "-cepext", ".raw", "-adcin", "yes", "-samprate", "16000", "-hyp",
                    os.path.join(tmp, utt + ".hyp"), "-logfn", os.path.join(tmp, utt + ".log"), "-fwdflat", "no",
                    "-bestpath", "no"] + extra, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return read_trace(out)


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for utt, (extra, env_extra, frames) in PICK.items():
            tr = trace(utt, extra, env_extra, tmp)
            topo = tr[0][0]
            for k in ("child_off", "child", "ciphone", "pw_off", "pw_wid", "pw_lastphone"):
                if "topo_" + k in out:
                    assert np.array_equal(out["topo_" + k], topo[k])      # same dictionary + LM: same tree
                out["topo_" + k] = topo[k]
            out["topo_n_root"] = np.int32(topo["n_root"])
            out["topo_n_chan"] = np.int32(topo["n_chan"])
            # the port must agree with the reference on EVERY frame before any of them becomes a golden
            for t, b, a in tr:
                s, nacl, cand = orc.port_fwdtree_prune(t, b, b["pls_pen"], b["acl"], orc.prune_rows_to_soa(b["state"]))
                assert np.array_equal(orc.prune_soa_to_rows(s), a["state"]) and np.array_equal(nacl, a["nacl"])
                assert not a["cand_valid"] or np.array_equal(cand, a["cand"])
            print(utt, len(tr), "frames: port == reference")
            for f in frames:
                t, b, a = tr[f]
                assert b["frame"] == f and a["cand_valid"]
                key = f"{utt}_{f}_"
                out[key + "par"] = np.array([b[k] for k in orc.PRUNE_PAR], np.int32)
                out[key + "pls_pen"] = b["pls_pen"]
                out[key + "acl"] = b["acl"]
                out[key + "state"] = b["state"]
                ch = np.nonzero((a["state"] != b["state"]).any(axis=1))[0].astype(np.int32)
                out[key + "after_idx"] = ch
                out[key + "after_rows"] = a["state"][ch]
                out[key + "nacl"] = a["nacl"]
                out[key + "cand"] = a["cand"]
    out["cases"] = np.array([f"{u}_{f}" for u, (_, _, fr) in PICK.items() for f in fr])
    path = os.path.join(os.path.dirname(__file__), "fwdtree_prune.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
