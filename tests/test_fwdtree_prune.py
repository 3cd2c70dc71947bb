"""Prune / phone-transition stage of the forward tree search:
prune_root_chan + prune_nonroot_chan, pocketsphinx/src/libpocketsphinx/ngram_search_fwdtree.c:714-869.

Pins: tests/golden/fwdtree_prune.npz holds the lexical tree (hub4wsj_sc_8k + wsj0vp.5000 + cmu07a: 443 roots,
14 331 channels) and fourteen frames of real decodes of the UNMODIFIED reference, recorded before and after the two
functions by oracle/ref_fwdtree_trace.c (the reference's own source file compiled in place with two macro hooks);
tests/golden/make_fwdtree_golden.py also checks the oracle port on every frame of the three utterances.  Where
oracle/_ref exists the whole decode is re-traced here.  The GPU tests go through the C ABI
(b200_chantree_create / b200_fwdtree_prune_host / _dev) and ask for identical states, identical ORDER of the
next active list and identical order of the last-phone candidates."""
import os
import subprocess

import numpy as np
import pytest

import orc
from fwdtree_trace import read_trace

G = os.path.join(orc.GOLDEN_DIR, "fwdtree_prune.npz")
TOPO_KEYS = ("child_off", "child", "ciphone", "pw_off", "pw_wid", "pw_lastphone")
WORST = np.int32(-0x20000000)


def golden():
    z = np.load(G)
    topo = {k: z["topo_" + k] for k in TOPO_KEYS}
    topo["n_root"], topo["n_chan"], topo["n_ci"] = int(z["topo_n_root"]), int(z["topo_n_chan"]), 50
    cases = []
    for name in z["cases"]:
        k = str(name) + "_"
        par = dict(zip(orc.PRUNE_PAR, (int(x) for x in z[k + "par"])))
        after = z[k + "state"].copy()
        after[z[k + "after_idx"]] = z[k + "after_rows"]
        cases.append(dict(name=str(name), par=par, pls_pen=z[k + "pls_pen"], acl=z[k + "acl"], state=z[k + "state"],
                          after=after, nacl=z[k + "nacl"], cand=z[k + "cand"]))
    return topo, cases


def test_port_matches_reference_golden():
    topo, cases = golden()
    assert len(cases) == 14 and any(c["par"]["has_pls"] for c in cases)
    for c in cases:
        s, nacl, cand = orc.port_fwdtree_prune(topo, c["par"], c["pls_pen"], c["acl"], orc.prune_rows_to_soa(c["state"]))
        assert np.array_equal(orc.prune_soa_to_rows(s), c["after"]), c["name"]
        assert np.array_equal(nacl, c["nacl"]), c["name"]
        assert np.array_equal(cand, c["cand"]), c["name"]
    # the goldens exercise the order-dependent cases: a pruned channel entered by a LATER parent (cleared, then
    # restarted with only state 0 alive) and one entered by an EARLIER parent (not cleared).  A real decode always
    # lists a parent ahead of its children; the "something" frames were recorded with the list permuted before the
    # reference's own functions ran (B200_FWDTREE_TRACE_SHUFFLE)
    n_later = n_earlier = 0
    for c in cases:
        pos = np.full(topo["n_chan"], -1)
        pos[c["acl"]] = np.arange(len(c["acl"]))
        parent = np.full(topo["n_chan"], -1)
        for p in range(topo["n_chan"]):
            parent[topo["child"][topo["child_off"][p]:topo["child_off"][p + 1]]] = p
        b, a = c["state"], c["after"]
        thresh = c["par"]["best_score"] + c["par"]["beam"]
        pruned = c["acl"][b[c["acl"], 8] <= thresh]
        ent = pruned[a[pruned, 9] == c["par"]["frame"] + 1]
        for ch in ent:
            p = parent[ch]
            if p >= topo["n_root"] and pos[p] > pos[ch]:
                n_later += 1
                assert a[ch, 1] == WORST and a[ch, 2] == WORST and a[ch, 8] == WORST
            else:
                n_earlier += 1
                assert a[ch, 1] == b[ch, 1] and a[ch, 2] == b[ch, 2]
    assert n_later > 0 and n_earlier > 0


@pytest.mark.skipif(not os.path.exists(os.path.join(orc.REF_DIR, "libref_fwdtree_trace.so")), reason="oracle/_ref not built")
def test_port_matches_reference_on_a_whole_decode(tmp_path):
    """Every frame of goforward.raw (-pl_window 3: with the phone-loop look-ahead) through the reference's own
    prune_root_chan / prune_nonroot_chan, against the port."""
    D, R = orc.DATA_DIR, orc.REF_DIR
    (tmp_path / "a.ctl").write_text("goforward\n")
    out = tmp_path / "trace.bin"
    env = dict(os.environ, LD_LIBRARY_PATH=R, LD_PRELOAD=os.path.join(R, "libref_fwdtree_trace.so"), B200_FWDTREE_TRACE=str(out),
               B200_FWDTREE_TRACE_EVERY="3")
    subprocess.run([os.path.join(R, "pocketsphinx_batch"), "-hmm", os.path.join(D, "hmm", "hub4wsj_sc_8k"), "-lm",
                    os.path.join(D, "lm", "wsj0vp.5000.DMP"), "-dict", os.path.join(D, "lm", "cmu07a.dic"), "-ctl",
                    str(tmp_path / "a.ctl"), "-cepdir", os.path.join(D, "test"), "-cepext", ".raw", "-adcin", "yes", "-samprate",
                    "16000", "-hyp", str(tmp_path / "a.hyp"), "-logfn", str(tmp_path / "a.log"), "-fwdflat", "no", "-bestpath",
                    "no", "-pl_window", "3"], env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert (tmp_path / "a.hyp").read_text().startswith("go forward ten years")
    tr = read_trace(str(out))
    assert len(tr) > 80
    for t, b, a in tr:
        s, nacl, cand = orc.port_fwdtree_prune(t, b, b["pls_pen"], b["acl"], orc.prune_rows_to_soa(b["state"]))
        assert np.array_equal(orc.prune_soa_to_rows(s), a["state"]) and np.array_equal(nacl, a["nacl"])
        assert not a["cand_valid"] or np.array_equal(cand, a["cand"])


# ----------------------------------------------------------------------------- synthetic trees
def random_tree(rng, n_root, n_chan, n_ci=12, n_word=40):
    """A random forest: channel c > n_root hangs under a random earlier channel; children lists in random order."""
    parent = np.full(n_chan, -1)
    kids = [[] for _ in range(n_chan)]
    for c in range(n_root, n_chan):
        p = int(rng.integers(0, c))
        parent[c] = p
        kids[p].append(c)
    child_off, child = [0], []
    for c in range(n_chan):
        k = kids[c]
        rng.shuffle(k)
        child += k
        child_off.append(len(child))
    pw_off, pw_wid, pw_lp = [0], [], []
    for c in range(n_chan):
        for _ in range(int(rng.integers(0, 3)) if rng.random() < 0.4 else 0):
            pw_wid.append(int(rng.integers(0, n_word)))
            pw_lp.append(int(rng.integers(0, n_ci)))
        pw_off.append(len(pw_wid))
    return dict(n_root=n_root, n_chan=n_chan, n_ci=n_ci, child_off=np.array(child_off, np.int32), child=np.array(child, np.int32),
                ciphone=rng.integers(0, n_ci, n_chan).astype(np.int32), pw_off=np.array(pw_off, np.int32),
                pw_wid=np.array(pw_wid, np.int32), pw_lastphone=np.array(pw_lp, np.int32)), parent


def random_frame(rng, topo, ne=3, frame=7, p_active=0.5, has_pls=False):
    """A consistent pre-prune state: active channels carry frame == frame_idx and evaluated scores near the beam
    (so keep / prune, enter / not-enter and both walk orders all occur, ties included), the rest are cleared."""
    n_chan, n_root = topo["n_chan"], topo["n_root"]
    rows = np.zeros((n_chan, 2 * ne + 4), np.int32)
    sc, hi = rows[:, 0:ne], rows[:, ne:2 * ne]
    sc[:] = WORST
    hi[:] = rng.integers(-1, 50, (n_chan, ne))
    rows[:, 2 * ne] = WORST
    rows[:, 2 * ne + 1] = rng.integers(-1, 50, n_chan)
    rows[:, 2 * ne + 2] = WORST
    rows[:, 2 * ne + 3] = rng.integers(-1, frame, n_chan)
    act = rng.random(n_chan) < p_active
    n = int(act.sum())
    s = -rng.integers(0, 40, (n, ne + 1)).astype(np.int32) * 25            # coarse grid: exact ties happen
    s[rng.random((n, ne + 1)) < 0.15] = WORST
    rows[act, 0:ne] = s[:, :ne]
    rows[act, 2 * ne] = s[:, ne]
    rows[act, 2 * ne + 2] = s.max(axis=1)
    rows[act, 2 * ne + 3] = frame
    acl = np.nonzero(act)[0]
    acl = acl[acl >= n_root].astype(np.int32)
    rng.shuffle(acl)
    par = dict(frame=frame, best_score=int(rows[act, 2 * ne + 2].max()) if n else 0, beam=-500, pbeam=-450, lpbeam=-400,
               pip=-5, nwpen=-3, has_pls=int(has_pls))
    pen = (-rng.integers(0, 6, topo["n_ci"]) * 25).astype(np.int32) if has_pls else np.zeros(topo["n_ci"], np.int32)
    return rows, acl, par, pen


def rows_to_soa(rows, ne):
    return dict(score=np.ascontiguousarray(rows[:, 0:ne].T), history=np.ascontiguousarray(rows[:, ne:2 * ne].T),
                out_score=rows[:, 2 * ne].copy(), out_history=rows[:, 2 * ne + 1].copy(), bestscore=rows[:, 2 * ne + 2].copy(),
                frame=rows[:, 2 * ne + 3].copy())


def test_port_order_dependence_on_synthetic_trees():
    """The port (= the reference's walk) really depends on the list order: reversing the active list changes the
    state of some pruned-and-entered channel.  This is the effect the GPU kernel has to reproduce."""
    rng = np.random.default_rng(3)
    topo, _ = random_tree(rng, 6, 400)
    differs = 0
    for _ in range(20):
        rows, acl, par, pen = random_frame(rng, topo)
        s1, n1, c1 = orc.port_fwdtree_prune(topo, par, pen, acl, rows_to_soa(rows, 3))
        s2, n2, c2 = orc.port_fwdtree_prune(topo, par, pen, acl[::-1].copy(), rows_to_soa(rows, 3))
        assert sorted(n1) == sorted(n2)                         # the SET of survivors does not depend on the order
        differs += not np.array_equal(s1["score"], s2["score"])
    assert differs > 0


def parallel_formulation(topo, par, pen, acl, soa, ne):
    """The kernel's order-free formulation (csrc/fwdtree_prune.cu), phase by phase, in Python."""
    n_root, n_chan = topo['n_root'], topo['n_chan']
    co, ch, ci, po, pw, pl_ = (topo[k] for k in ('child_off','child','ciphone','pw_off','pw_wid','pw_lastphone'))
    parent = np.full(n_chan, -1)
    for p in range(n_chan): parent[ch[co[p]:co[p+1]]] = p
    s = {k: v.copy() for k, v in soa.items()}
    score0, hist0 = s['score'][0], s['history'][0]
    fi = par['frame']; nf = fi + 1
    thresh = par['best_score'] + par['beam']; newphone = par['best_score'] + par['pbeam']; lastphn = par['best_score'] + par['lpbeam']
    pip, nwpen, pls = par['pip'], par['nwpen'], bool(par['has_pls'])
    E = n_root + len(acl)
    tau = np.full(n_chan, -1)
    for e in range(E):
        if e < n_root: tau[e] = e if s['frame'][e] >= fi else -1
        else: tau[acl[e - n_root]] = e
    def edge(nps, pn, p_tau, c_tau, c_frame, c_in, c_best):
        pl = nps + pn
        if not (pls or nps > newphone) or not (pl > newphone): return 0
        if c_tau < 0 or p_tau < c_tau: return 3 if (c_frame < fi or pl > c_in) else 0
        if c_best > thresh: return 1 if pl > c_in else 0
        return 3 if pl > WORST else 0
    pre = {k: v.copy() for k, v in s.items()}
    P0, PB, PO, PF = pre['score'][0], pre['bestscore'], pre['out_score'], pre['frame']
    dec = np.zeros(n_chan, int); cnt = np.zeros(E, int); ccnt = np.zeros(E, int); flag = np.zeros(E, int)
    for e in range(E):
        c = e if e < n_root else acl[e - n_root]
        if tau[c] != e: continue
        keep = PB[c] > thresh
        fl = 0; na = 0; ncd = 0
        if e >= n_root:
            p = parent[c]; pt = tau[p]; d = 0
            if pt >= 0 and PB[p] > thresh:
                d = edge(PO[p] + pip, pen[ci[c]] if pls else 0, pt, e, PF[c], P0[c], PB[c])
            eb = (d & 1) and pt < e
            if keep: fl = 1 | (0 if eb else 2)
            elif not eb: fl = 4 | (8 if d & 1 else 0)
            na = 1 if fl & 2 else 0
        elif keep: fl = 1
        if keep:
            nps = PO[c] + pip
            for k in range(co[c], co[c+1]):
                c2 = ch[k]
                d = edge(nps, pen[ci[c2]] if pls else 0, e, tau[c2], PF[c2], P0[c2], PB[c2])
                dec[c2] = d; na += d >> 1
            if pls or nps > lastphn:
                for k in range(po[c], po[c+1]):
                    ncd += (nps + (pen[pl_[k]] if pls else 0)) > lastphn
        cnt[e], ccnt[e], flag[e] = na, ncd, fl
    off = np.concatenate([[0], np.cumsum(cnt)]); coff = np.concatenate([[0], np.cumsum(ccnt)])
    nacl = np.zeros(off[-1], np.int32); cand = np.zeros((coff[-1], 3), np.int32)
    for e in range(E):
        c = e if e < n_root else acl[e - n_root]
        fl = flag[e]; o = off[e]; q = coff[e]
        if fl & 2: nacl[o] = c; o += 1
        if fl & 1:
            s['frame'][c] = nf
            nps = PO[c] + pip; oh = pre['out_history'][c]
            for k in range(co[c], co[c+1]):
                c2 = ch[k]; d = dec[c2]
                if d & 1:
                    score0[c2] = nps + (pen[ci[c2]] if pls else 0); hist0[c2] = oh; s['frame'][c2] = nf
                    if d & 2: nacl[o] = c2; o += 1
            if pls or nps > lastphn:
                for k in range(po[c], po[c+1]):
                    pl = nps + (pen[pl_[k]] if pls else 0)
                    if pl > lastphn: cand[q] = (pw[k], pl - nwpen, oh); q += 1
        elif fl & 4:
            for st in range(1 if fl & 8 else 0, ne): s['score'][st][c] = WORST
            s['out_score'][c] = WORST; s['bestscore'][c] = WORST
    return s, nacl, cand


def test_order_free_formulation_equals_the_sequential_walk():
    """The kernel's design, checked without a GPU: every decision taken from the PRE-prune state plus walk
    positions, appends placed by a prefix sum -- same states, same list order, same candidate order as the
    reference's sequential walk (the port), shuffled lists and exact ties included."""
    rng = np.random.default_rng(1)
    for ne, nr, nc, hp in [(3, 6, 400, False), (3, 10, 800, True), (5, 3, 300, True), (1, 2, 64, False), (3, 1, 2, False)]:
        topo, _ = random_tree(rng, nr, nc)
        for u in range(12):
            rows, acl, par, pen = random_frame(rng, topo, ne, frame=3 + u, p_active=[0.05, 0.5, 0.95][u % 3], has_pls=hp)
            soa = rows_to_soa(rows, ne)
            ws, wn, wc = orc.port_fwdtree_prune(topo, par, pen, acl, soa)
            gs, gn, gc = parallel_formulation(topo, par, pen, acl, soa, ne)
            assert all(np.array_equal(gs[k], ws[k]) for k in ws) and np.array_equal(gn, wn) and np.array_equal(gc, wc)
    topo, cases = golden()
    for c in cases[::4]:
        gs, gn, gc = parallel_formulation(topo, c["par"], c["pls_pen"], c["acl"], orc.prune_rows_to_soa(c["state"]), 3)
        assert np.array_equal(orc.prune_soa_to_rows(gs), c["after"]) and np.array_equal(gn, c["nacl"]) and np.array_equal(gc, c["cand"])


def test_abi_exports_the_prune_entry_points():
    import ctypes as C
    lib = C.CDLL(os.path.join(orc.ROOT, "cmusphinx_b200", "libb200sphinx.so"))
    for name in ("b200_chantree_create", "b200_chantree_free", "b200_chantree_cand_cap", "b200_fwdtree_prune_dev",
                 "b200_fwdtree_prune_host"):
        assert hasattr(lib, name), name


def test_chantree_rejects_a_graph_that_is_not_a_tree():
    import cmusphinx_b200 as b
    if b.device_count() > 0:
        pytest.skip("argument checks come before the device check; covered on the GPU box by the tests below")
    with pytest.raises(b.B200Error, match="two parents"):
        b.ChanTree(1, 3, [0, 2, 3, 3], [1, 2, 2], [0, 0, 0], [0, 0, 0, 0], [], [], 1)
    with pytest.raises(b.B200Error, match="no CUDA device"):
        b.ChanTree(1, 3, [0, 2, 2, 2], [1, 2], [0, 0, 0], [0, 0, 0, 0], [], [], 1)


# ----------------------------------------------------------------------------- GPU
def _gpu_prune(tree, items, ne):
    """items: list of (rows, acl, par, pen); one batched call."""
    import cmusphinx_b200 as b  # noqa: F401
    n_utt, n_chan = len(items), tree.n_chan
    score = np.empty((ne, n_utt * n_chan), np.int32)
    history = np.empty((ne, n_utt * n_chan), np.int32)
    flat = {k: np.empty(n_utt * n_chan, np.int32) for k in ("out_score", "out_history", "bestscore", "frame")}
    for u, (rows, _, _, _) in enumerate(items):
        s = rows_to_soa(rows, ne)
        score[:, u * n_chan:(u + 1) * n_chan] = s["score"]
        history[:, u * n_chan:(u + 1) * n_chan] = s["history"]
        for k in flat:
            flat[k][u * n_chan:(u + 1) * n_chan] = s[k]
    par = np.array([[it[2][k] for k in orc.PRUNE_PAR] for it in items], np.int32)
    pen = np.stack([it[3] for it in items]).astype(np.int32)
    nacl, cand = tree.prune(par, pen, [it[1] for it in items], score, history, flat["out_score"], flat["out_history"],
                            flat["bestscore"], flat["frame"])
    out = []
    for u in range(n_utt):
        sl = slice(u * n_chan, (u + 1) * n_chan)
        out.append((dict(score=score[:, sl], history=history[:, sl], **{k: v[sl] for k, v in flat.items()}), nacl[u], cand[u]))
    return out


def _same(got, want, tag):
    gs, gn, gc = got
    ws, wn, wc = want
    for k in ("score", "history", "out_score", "out_history", "bestscore", "frame"):
        assert np.array_equal(gs[k], ws[k]), (tag, k)
    assert np.array_equal(gn, wn), (tag, "next active list")
    assert np.array_equal(gc, wc), (tag, "last-phone candidates")


@pytest.mark.gpu
def test_gpu_prune_matches_reference_golden():
    import cmusphinx_b200 as b
    topo, cases = golden()
    tree = b.ChanTree(topo["n_root"], topo["n_chan"], *[topo[k] for k in TOPO_KEYS], topo["n_ci"])
    n0 = b.launch_count()
    # one utterance per call, then all the frames as one batch of "utterances"
    items = [(c["state"], c["acl"], c["par"], c["pls_pen"]) for c in cases]
    for c, it in zip(cases, items):
        got = _gpu_prune(tree, [it], 3)[0]
        _same(got, (orc.prune_rows_to_soa(c["after"]), c["nacl"], c["cand"]), c["name"])
    for c, got in zip(cases, _gpu_prune(tree, items, 3)):
        _same(got, (orc.prune_rows_to_soa(c["after"]), c["nacl"], c["cand"]), "batched " + c["name"])
    assert b.launch_count() > n0
    # an argument the reference would trip an assert on
    with pytest.raises(b.B200Error, match="not a non-root channel"):
        _gpu_prune(tree, [(cases[0]["state"], np.array([0], np.int32), cases[0]["par"], cases[0]["pls_pen"])], 3)
    tree.free()


@pytest.mark.gpu
@pytest.mark.parametrize("ne,n_root,n_chan,has_pls", [(3, 6, 400, False), (3, 40, 5000, True), (5, 3, 1500, True), (1, 2, 64, False),
                                                        (3, 1, 2, False)])
def test_gpu_prune_matches_port_on_synthetic_trees(ne, n_root, n_chan, has_pls):
    import cmusphinx_b200 as b
    rng = np.random.default_rng(100 + n_chan)
    topo, _ = random_tree(rng, n_root, n_chan)
    tree = b.ChanTree(n_root, n_chan, *[topo[k] for k in TOPO_KEYS], topo["n_ci"], n_emit=ne)
    items = []
    for u in range(24):
        items.append(random_frame(rng, topo, ne, frame=3 + u, p_active=[0.05, 0.5, 0.95][u % 3], has_pls=has_pls))
    items.append((items[0][0], np.zeros(0, np.int32), items[0][2], items[0][3]))      # empty active list, active roots only
    want = []
    for rows, acl, par, pen in items:
        want.append(orc.port_fwdtree_prune(topo, par, pen, acl, rows_to_soa(rows, ne)))
    for u, (g, w) in enumerate(zip(_gpu_prune(tree, items, ne), want)):
        _same(g, w, f"utt {u}")
    # frame after frame on the same tree: the scratch of one call must not leak into the next
    for u, (g, w) in enumerate(zip(_gpu_prune(tree, items[::-1], ne), want[::-1])):
        _same(g, w, f"second call utt {u}")
    tree.free()


@pytest.mark.gpu
def test_eval_then_prune_without_leaving_the_device():
    """hmm_vit_eval for every channel of the tree (the resident population of an HmmContext, three utterances) and
    the prune / transition stage chained behind it on the context's stream (b200_hmm_pop_device +
    b200_fwdtree_prune_dev): the channel states stay in HBM between the two stages; only the frame stamps and the
    lists travel.  Against the oracle's hmm_vit_eval followed by its sequential prune walk."""
    import cmusphinx_b200 as b
    from cmusphinx_b200 import synth
    ne, n_sen, n_tmat, n_sseq, n_root, n_chan, n_utt = 3, 800, 12, 900, 40, 5000, 3
    rng = np.random.default_rng(11)
    topo, _ = random_tree(rng, n_root, n_chan)
    tree = b.ChanTree(n_root, n_chan, *[topo[k] for k in TOPO_KEYS], topo["n_ci"], n_emit=ne)
    tp = orc.port_tmat_quantize(synth.bakis_tmat(n_tmat, ne, 7), 1e-4, orc.LOGBASE)
    d = synth.hmm_population(n_utt * n_chan, ne, n_sen, n_tmat, n_sseq, seed=4, mpx_fraction=0.0)
    frames, acls, pens = [], [], []
    for u in range(n_utt):
        rows, acl, par, pen = random_frame(rng, topo, ne, frame=20, p_active=0.4, has_pls=(u == 1))
        sl = slice(u * n_chan, (u + 1) * n_chan)
        d["score"][sl], d["history"][sl] = rows[:, 0:ne], rows[:, ne:2 * ne]
        d["out_score"][sl], d["out_history"][sl], d["bestscore"][sl] = rows[:, 2 * ne], rows[:, 2 * ne + 1], rows[:, 2 * ne + 2]
        d["mpx"][u * n_chan:u * n_chan + n_root] = 1                     # the roots are multiplex HMMs (ngram_search_fwdtree.c:239)
        d["senid"][u * n_chan:u * n_chan + n_root] = rng.integers(0, n_sseq, (n_root, ne))
        frames.append(rows[:, 2 * ne + 3].copy()); acls.append(acl); pens.append(pen)
    sen = np.stack([synth.senscr_frames(1, n_sen, 30 + u)[0] for u in range(n_utt)])
    # oracle: evaluate, then walk
    o = {k: v.copy() for k, v in d.items()}
    want = []
    for u in range(n_utt):
        sl = slice(u * n_chan, (u + 1) * n_chan)
        v = {k: np.ascontiguousarray(o[k][sl]) for k in ("score", "history", "out_score", "out_history", "senid", "tmatid", "mpx", "bestscore")}
        orc.hmm_eval(orc.port.orc_hmm_eval_batch, ne, tp, d["sseq"], sen[u], v["score"], v["history"], v["out_score"],
                     v["out_history"], v["senid"], v["tmatid"], v["mpx"], v["bestscore"])
        act = np.concatenate([np.nonzero(frames[u][:n_root] >= 20)[0], acls[u]])
        best = int(v["bestscore"][act].max())
        par = dict(frame=20, best_score=best, beam=-600, pbeam=-500, lpbeam=-450, pip=-5, nwpen=-3, has_pls=int(u == 1))
        soa = dict(score=np.ascontiguousarray(v["score"].T), history=np.ascontiguousarray(v["history"].T), out_score=v["out_score"],
                   out_history=v["out_history"], bestscore=v["bestscore"], frame=frames[u])
        want.append((par,) + orc.port_fwdtree_prune(topo, par, pens[u], acls[u], soa))
    # device: one resident population, evaluate, prune on the same stream
    ctx = b.HmmContext(ne, tp, d["sseq"], n_sen)
    pop = b.HmmPopulation(n_utt * n_chan, ne)
    pop.score[:], pop.history[:], pop.senid[:] = d["score"].T, d["history"].T, d["senid"].T
    pop.out_score[:], pop.out_history[:], pop.tmatid[:], pop.mpx[:] = d["out_score"], d["out_history"], d["tmatid"], d["mpx"]
    pop.bestscore[:] = d["bestscore"]
    ctx.upload(pop)
    ctx.set_utts(np.arange(n_utt + 1) * n_chan)
    ctx.step(sen, -600, n_utt * n_chan)
    frame = np.concatenate(frames).astype(np.int32)
    nacl, cand = tree.prune_resident(ctx, frame, [[w[0][k] for k in orc.PRUNE_PAR] for w in want], np.stack(pens),
                                     acls)
    ctx.download(pop)
    for u in range(n_utt):
        sl = slice(u * n_chan, (u + 1) * n_chan)
        _, ws, wn, wc = want[u]
        assert np.array_equal(pop.score[:, sl], ws["score"]) and np.array_equal(pop.history[:, sl], ws["history"])
        assert np.array_equal(pop.out_score[sl], ws["out_score"]) and np.array_equal(pop.bestscore[sl], ws["bestscore"])
        assert np.array_equal(frame[sl], ws["frame"])
        assert np.array_equal(nacl[u], wn) and np.array_equal(cand[u], wc)
        assert len(wn) > 100 and len(wc) > 10
    ctx.free()
    tree.free()


@pytest.mark.gpu
def test_frames_of_evaluate_and_prune_resident_on_the_device():
    """Eight frames of the tree-internal part of the forward tree search for three utterances, device resident
    (FwdtreeDevice: b200_hmm_eval_list_dev -> b200_fwdtree_prune_dev, the list one stage writes is the list the
    next one reads), against the oracle doing the same with hmm_vit_eval on the listed channels
    (eval_root_chan + eval_nonroot_chan, ngram_search_fwdtree.c:598-634) and the sequential prune walk.  The host
    plays word_transition: it re-enters a few root channels every frame."""
    import cmusphinx_b200 as b
    from cmusphinx_b200 import synth
    ne, n_sen, n_tmat, n_sseq, n_root, n_chan, n_utt, T = 3, 600, 10, 700, 30, 3000, 3, 8
    rng = np.random.default_rng(21)
    topo, _ = random_tree(rng, n_root, n_chan)
    tree = b.ChanTree(n_root, n_chan, *[topo[k] for k in TOPO_KEYS], topo["n_ci"], n_emit=ne)
    tp = orc.port_tmat_quantize(synth.bakis_tmat(n_tmat, ne, 7), 1e-4, orc.LOGBASE)
    d = synth.hmm_population(n_utt * n_chan, ne, n_sen, n_tmat, n_sseq, seed=5, mpx_fraction=0.0)
    frame = np.zeros(n_utt * n_chan, np.int32)
    lists = []
    for u in range(n_utt):
        rows, acl, _, _ = random_frame(rng, topo, ne, frame=0, p_active=0.3)
        sl = slice(u * n_chan, (u + 1) * n_chan)
        d["score"][sl], d["history"][sl] = rows[:, 0:ne], rows[:, ne:2 * ne]
        d["out_score"][sl], d["out_history"][sl], d["bestscore"][sl] = rows[:, 2 * ne], rows[:, 2 * ne + 1], rows[:, 2 * ne + 2]
        d["mpx"][u * n_chan:u * n_chan + n_root] = 1
        d["senid"][u * n_chan:u * n_chan + n_root] = rng.integers(0, n_sseq, (n_root, ne))
        frame[sl] = rows[:, 2 * ne + 3]
        lists.append(acl)
    sen = np.stack([synth.senscr_frames(n_utt, n_sen, 50 + f) for f in range(T)])          # [T][n_utt][n_sen]
    entries = [[(rng.choice(n_root, 4, replace=False), -rng.integers(0, 200, 4).astype(np.int32), rng.integers(0, 99, 4).astype(np.int32))
                for _ in range(n_utt)] for _ in range(T)]
    # ---- device
    ctx = b.HmmContext(ne, tp, d["sseq"], n_sen)
    pop = b.HmmPopulation(n_utt * n_chan, ne)
    pop.score[:], pop.history[:], pop.senid[:] = d["score"].T, d["history"].T, d["senid"].T
    pop.out_score[:], pop.out_history[:], pop.tmatid[:], pop.mpx[:] = d["out_score"], d["out_history"], d["tmatid"], d["mpx"]
    pop.bestscore[:] = d["bestscore"]
    ctx.upload(pop)
    ctx.set_utts(np.arange(n_utt + 1) * n_chan)
    dev = b.FwdtreeDevice(tree, ctx, n_utt, frame)
    dev.set_lists(lists)
    got_best, got_cand, got_lists = [], [], []
    for f in range(T):
        best = dev.evaluate(sen[f], f)
        got_best.append(best.copy())
        par = [[f, int(best[u]), -700, -600, -500, -5, -3, 0] for u in range(n_utt)]
        got_cand.append(dev.prune(par))
        got_lists.append(dev.lists())
        # word_transition stand-in: hmm_enter of a few roots for the next frame (only if better, PS/ngram_search_fwdtree.c:1383)
        fr = dev.frame_stamps()
        idx = np.concatenate([u * n_chan + e[0] for u, e in enumerate(entries[f])]).astype(np.int32)
        ctx.enter(idx, np.concatenate([e[1] for e in entries[f]]), np.concatenate([e[2] for e in entries[f]]))
        for u, e in enumerate(entries[f]):
            fr[u * n_chan + e[0]] = f + 1
        dev.set_frame_stamps(fr)
    ctx.download(pop)
    frame_dev = dev.frame_stamps()
    # ---- oracle
    o = {k: v.copy() for k, v in d.items()}
    fr_o = frame.copy()
    cur = [l.copy() for l in lists]
    n_eval = 0
    for f in range(T):
        for u in range(n_utt):
            base = u * n_chan
            roots = np.nonzero(fr_o[base:base + n_root] == f)[0]
            ids = base + np.concatenate([roots, cur[u]]).astype(np.int64)
            best = int(WORST)
            if ids.size:
                v = {k: np.ascontiguousarray(o[k][ids]) for k in ("score", "history", "out_score", "out_history", "senid", "tmatid", "mpx", "bestscore")}
                orc.hmm_eval(orc.port.orc_hmm_eval_batch, ne, tp, d["sseq"], sen[f][u], v["score"], v["history"], v["out_score"],
                             v["out_history"], v["senid"], v["tmatid"], v["mpx"], v["bestscore"])
                for k in ("score", "history", "out_score", "out_history", "senid", "bestscore"):
                    o[k][ids] = v[k]
                best = int(v["bestscore"].max())
                n_eval += ids.size
            assert best == got_best[f][u], (f, u)
            par = dict(frame=f, best_score=best, beam=-700, pbeam=-600, lpbeam=-500, pip=-5, nwpen=-3, has_pls=0)
            sl = slice(base, base + n_chan)
            soa = dict(score=np.ascontiguousarray(o["score"][sl].T), history=np.ascontiguousarray(o["history"][sl].T),
                       out_score=o["out_score"][sl].copy(), out_history=o["out_history"][sl].copy(), bestscore=o["bestscore"][sl].copy(),
                       frame=fr_o[sl].copy())
            s2, nacl, cand = orc.port_fwdtree_prune(topo, par, np.zeros(topo["n_ci"], np.int32), cur[u], soa)
            o["score"][sl], o["history"][sl] = s2["score"].T, s2["history"].T
            o["out_score"][sl], o["bestscore"][sl], fr_o[sl] = s2["out_score"], s2["bestscore"], s2["frame"]
            assert np.array_equal(nacl, got_lists[f][u]), (f, u, "list")
            assert np.array_equal(cand, got_cand[f][u]), (f, u, "candidates")
            cur[u] = nacl
            r, sc, hi = entries[f][u]
            for k in range(len(r)):                              # hmm_enter if better, list order
                i = base + r[k]
                if sc[k] > o["score"][i, 0]:
                    o["score"][i, 0], o["history"][i, 0] = sc[k], hi[k]
                fr_o[i] = f + 1
    assert n_eval > 2000
    assert np.array_equal(pop.score.T, o["score"]) and np.array_equal(pop.history.T, o["history"])
    assert np.array_equal(pop.out_score, o["out_score"]) and np.array_equal(pop.bestscore, o["bestscore"])
    assert np.array_equal(frame_dev, fr_o)
    dev.free(); ctx.free(); tree.free()


@pytest.mark.gpu
@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built")
def test_renormalize_and_deactivate_on_the_device():
    """renormalize_scores' tree part (ngram_search_fwdtree.c:557-576) and deactivate_channels' root loop
    (:1418-1431) on the resident population, against the reference's OWN hmm_normalize / hmm_clear_scores
    (oracle/ref_shim.c ref_hmm_maint) applied to the channels those loops visit."""
    import cmusphinx_b200 as b
    from cmusphinx_b200 import synth
    ne, n_sen, n_root, n_chan, n_utt, f = 3, 300, 25, 2000, 2, 9
    rng = np.random.default_rng(31)
    topo, _ = random_tree(rng, n_root, n_chan)
    tree = b.ChanTree(n_root, n_chan, *[topo[k] for k in TOPO_KEYS], topo["n_ci"], n_emit=ne)
    tp = orc.port_tmat_quantize(synth.bakis_tmat(4, ne, 7), 1e-4, orc.LOGBASE)
    pop = b.HmmPopulation(n_utt * n_chan, ne)
    frame = np.zeros(n_utt * n_chan, np.int32)
    lists = []
    for u in range(n_utt):
        rows, acl, _, _ = random_frame(rng, topo, ne, frame=f, p_active=0.5)
        sl = slice(u * n_chan, (u + 1) * n_chan)
        pop.score[:, sl], pop.history[:, sl] = rows[:, 0:ne].T, rows[:, ne:2 * ne].T
        pop.out_score[sl], pop.out_history[sl], pop.bestscore[sl], frame[sl] = rows[:, 2 * ne], rows[:, 2 * ne + 1], rows[:, 2 * ne + 2], rows[:, 2 * ne + 3]
        lists.append(acl)
    before = {k: getattr(pop, k).copy() for k in ("score", "history", "out_score", "out_history", "bestscore")}
    ctx = b.HmmContext(ne, tp, None, n_sen)
    ctx.upload(pop)
    ctx.set_utts(np.arange(n_utt + 1) * n_chan)
    dev = b.FwdtreeDevice(tree, ctx, n_utt, frame)
    dev.set_lists(lists)
    norm = np.array([-777, -31], np.int32)

    def reference(op, visited, arg):
        sel = np.zeros(n_utt * n_chan, np.uint8)
        sel[visited] = 1
        a = np.zeros(n_utt * n_chan, np.int32)
        a[visited] = arg[visited // n_chan]                      # hmm_normalize by 0 leaves the others alone
        return orc.ref_hmm_maint(op, np.ascontiguousarray(cur["score"].T), np.ascontiguousarray(cur["history"].T), cur["out_score"],
                                 cur["out_history"], cur["bestscore"], sel=sel, arg=a)

    cur = before
    roots = np.concatenate([u * n_chan + np.nonzero(frame[u * n_chan:u * n_chan + n_root] == f)[0] for u in range(n_utt)])
    listed = np.concatenate([u * n_chan + lists[u] for u in range(n_utt)])
    assert roots.size > 5 and listed.size > 500
    dev.renormalize(f, norm)
    sc, hi, os_, oh, bs = reference(1, np.concatenate([roots, listed]), norm)
    ctx.download(pop)
    assert np.array_equal(pop.score.T, sc) and np.array_equal(pop.out_score, os_) and np.array_equal(pop.bestscore, bs)
    assert not np.array_equal(pop.score, before["score"])
    cur = dict(score=pop.score.copy(), history=pop.history.copy(), out_score=pop.out_score.copy(), out_history=pop.out_history.copy(),
               bestscore=pop.bestscore.copy())
    dev.deactivate(f)
    sc, hi, os_, oh, bs = reference(0, roots, norm)
    ctx.download(pop)
    assert np.array_equal(pop.score.T, sc) and np.array_equal(pop.out_score, os_) and np.array_equal(pop.bestscore, bs)
    assert np.array_equal(pop.history.T, hi) and (pop.score[:, roots] == WORST).all()
    dev.free(); ctx.free(); tree.free()
