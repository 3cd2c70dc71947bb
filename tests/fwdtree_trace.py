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


def read_pls_trace(path):
    """Records of oracle/ref_pls_trace.c -> (tp [n_tmat][ne][ne+1] uint8, list of frame dicts).  The state in
    record f+1 of an utterance is the reference's result for frame f."""
    a = np.fromfile(path, dtype=np.int32)
    pos, tp, out = 0, None, []
    while pos < a.size:
        tag = chr(int(a[pos])); pos += 1
        if tag == "M":
            n_tmat, ne = int(a[pos]), int(a[pos + 1]); pos += 2
            tp = a[pos:pos + n_tmat * ne * (ne + 1)].reshape(n_tmat, ne, ne + 1).astype(np.uint8); pos += n_tmat * ne * (ne + 1)
        elif tag == "P":
            frame, best, beam, pbeam, pip, n, ne = (int(x) for x in a[pos:pos + 7]); pos += 7
            w = 3 * ne + 5
            rows = a[pos:pos + n * w].reshape(n, w).copy(); pos += n * w
            sen = a[pos:pos + n * ne].reshape(n, ne).copy(); pos += n * ne
            out.append(dict(frame=frame, best_score=best, beam=beam, pbeam=pbeam, pip=pip, ne=ne, score=rows[:, 0:ne], history=rows[:, ne:2 * ne],
                            out_score=rows[:, 2 * ne], out_history=rows[:, 2 * ne + 1], bestscore=rows[:, 2 * ne + 2], frame_of=rows[:, 2 * ne + 3],
                            senid=rows[:, 2 * ne + 4:3 * ne + 4], tmatid=rows[:, 3 * ne + 4], senscr=sen))
        else:
            raise ValueError(f"bad record tag {tag!r}")
    return tp, out
