"""Reader of the records oracle/ref_fwdtree_trace.c writes (TEST INFRASTRUCTURE): the lexical tree of the
reference's forward-tree search before prune_root_chan and after prune_nonroot_chan
(pocketsphinx/src/libpocketsphinx/ngram_search_fwdtree.c:714-869) of a real decode."""
import numpy as np


def read_trace(path):
    """-> list of (topology dict, before dict, after dict) per recorded frame."""
    a = np.fromfile(path, dtype=np.int32)
    pos, topo, before, out = 0, None, None, []

    def take(n):
        nonlocal pos
        v = a[pos:pos + n]
        pos += n
        return v

    while pos < a.size:
        tag = chr(int(take(1)[0]))
        if tag == "T":
            n_root, n_chan, n_edge, n_pw, n_ci = (int(x) for x in take(5))
            topo = dict(n_root=n_root, n_chan=n_chan, n_ci=n_ci, child_off=take(n_chan + 1).copy(), child=take(n_edge).copy(),
                        ciphone=take(n_chan).copy(), pw_off=take(n_chan + 1).copy(), pw_wid=take(n_pw).copy(),
                        pw_lastphone=take(n_pw).copy())
        elif tag == "B":
            frame, best, dyn_beam, pbeam, lpbeam, pip, nwpen, has_pls, n_act = (int(x) for x in take(9))
            before = dict(frame=frame, best_score=best, beam=dyn_beam, pbeam=pbeam, lpbeam=lpbeam, pip=pip, nwpen=nwpen,
                          has_pls=has_pls, pls_pen=take(topo["n_ci"]).copy(), acl=take(n_act).copy(),
                          state=take(topo["n_chan"] * 10).reshape(-1, 10).copy())
        elif tag == "A":
            frame, cand_valid, n_nacl, n_cand = (int(x) for x in take(4))
            after = dict(frame=frame, cand_valid=cand_valid, nacl=take(n_nacl).copy(), cand=take(3 * n_cand).reshape(-1, 3).copy(),
                         state=take(topo["n_chan"] * 10).reshape(-1, 10).copy())
            assert before is not None and before["frame"] == frame
            out.append((topo, before, after))
            before = None
        else:
            raise ValueError(f"bad record tag {tag!r} at {pos}")
    return out
