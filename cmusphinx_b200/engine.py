"""Host-side mirror of the reference's back-end interface over the C ABI.

Names follow the reference: an ``Mgau`` is a ``ps_mgau_t`` (PS/acmod.h:95-124)
with ``frame_eval`` / ``transform`` / ``free``; ``HmmContext`` is
``hmm_context_t`` (PS/hmm.h:136-151) and ``HmmPopulation`` a structure-of-arrays
batch of ``hmm_t`` (PS/hmm.h:156-173).  All arithmetic happens inside
libb200sphinx.so on the GPU; numpy is used only to hold buffers.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import lib, check, B200Error, MgauCfg, HmmSoa

# -logbase is read with cmd_ln_float32_r (PS/pocketsphinx.c:223), so the
# reference's effective base is (double)(float)1.0001.
LOGBASE = float(np.float32(1.0001))
WORST_SCORE = np.int32(-0x20000000)   # (int)0xE0000000, PS/hmm.h:74
BAD_SSID = 0xFFFF


def _p(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def device_count() -> int:
    return lib.b200_device_count()


def launch_count() -> int:
    return lib.b200_launch_count()


# ------------------------------------------------------------ host utilities
def logadd_table(base: float = 1.0001, shift: int = 10) -> np.ndarray:
    n = lib.b200_logadd_table(base, shift, None, 0)
    check(n, "logadd_table")
    out = np.zeros(n, np.uint32)
    lib.b200_logadd_table(base, shift, _p(out, C.c_uint32), n)
    return out


def gauden_precompute(var: np.ndarray, length: int, varfloor: float = 1e-4, logbase: float = LOGBASE):
    """var: float32 [..., length] raw variances -> (scaled 1/(2var), det[...])."""
    v = _c(var, np.float32).copy().reshape(-1, length)
    det = np.zeros(v.shape[0], np.float32)
    check(lib.b200_gauden_precompute(_p(v, C.c_float), _p(det, C.c_float), v.shape[0], length,
                                     varfloor, logbase), "gauden_precompute")
    return v.reshape(var.shape), det.reshape(var.shape[:-1])


def mixw_quantize_ms(mixw: np.ndarray, mixwfloor: float = 1e-7, logbase: float = LOGBASE) -> np.ndarray:
    m = _c(mixw, np.float32).copy()
    n_sen, n_feat, n_cw = m.shape
    out = np.zeros(m.shape, np.uint8)
    check(lib.b200_mixw_quantize_ms(_p(m, C.c_float), _p(out, C.c_uint8), n_sen, n_feat, n_cw,
                                    mixwfloor, logbase), "mixw_quantize_ms")
    return out


def mixw_quantize_tied(mixw: np.ndarray, mixwfloor: float = 1e-7, logbase: float = LOGBASE) -> np.ndarray:
    m = _c(mixw, np.float32).copy()
    n_sen, n_feat, n_cw = m.shape
    out = np.zeros((n_feat, n_cw, n_sen), np.uint8)
    check(lib.b200_mixw_quantize_tied(_p(m, C.c_float), _p(out, C.c_uint8), n_sen, n_feat, n_cw,
                                      mixwfloor, logbase), "mixw_quantize_tied")
    return out


def tmat_quantize(tp: np.ndarray, tmatfloor: float = 1e-4, logbase: float = LOGBASE) -> np.ndarray:
    t = _c(tp, np.float32).copy()
    n_tmat, n_src, n_dst = t.shape
    assert n_dst == n_src + 1
    out = np.zeros(t.shape, np.uint8)
    check(lib.b200_tmat_quantize(_p(t, C.c_float), _p(out, C.c_uint8), n_tmat, n_src, tmatfloor, logbase),
          "tmat_quantize")
    return out


def flags2list(mask: np.ndarray, n_sen: int) -> np.ndarray:
    m = _c(mask, np.uint32)
    out = np.zeros(n_sen * 2 + 8, np.uint8)
    n = check(lib.b200_flags2list(_p(m, C.c_uint32), n_sen, _p(out, C.c_uint8), out.size), "flags2list")
    return out[:n].copy()


def read_gauden(path: str):
    dims = (C.c_int32 * 4)()
    vl = (C.c_int32 * 64)()
    check(lib.b200_s3_read_gauden(path.encode(), dims, vl, None), "read_gauden")
    data = np.zeros(dims[3], np.float32)
    check(lib.b200_s3_read_gauden(path.encode(), dims, vl, _p(data, C.c_float)), "read_gauden")
    return dict(n_mgau=dims[0], n_feat=dims[1], n_density=dims[2], veclen=[vl[i] for i in range(dims[1])],
                data=data)


def _read4(fn, path):
    dims = (C.c_int32 * 4)()
    check(fn(path.encode(), dims, None), path)
    data = np.zeros(dims[3], np.float32)
    check(fn(path.encode(), dims, _p(data, C.c_float)), path)
    return data.reshape(dims[0], dims[1], dims[2])


def read_mixw(path: str) -> np.ndarray:
    return _read4(lib.b200_s3_read_mixw, path)


def read_tmat(path: str) -> np.ndarray:
    return _read4(lib.b200_s3_read_tmat, path)


def read_s3_cont_arrays(meanfile: str, varfile: str, mixwfile: str):
    """Single-stream continuous model files -> raw (mean, var [S][M][D], mixw [S][M])
    as mgau_file_read / mgau_mixw_read see them (S3/libam/cont_mgau.c:160-400, 480-680)."""
    gm, gv = read_gauden(meanfile), read_gauden(varfile)
    if gm["n_feat"] != 1:
        raise B200Error("continuous model must have one feature stream")
    S, M, D = gm["n_mgau"], gm["n_density"], gm["veclen"][0]
    w = read_mixw(mixwfile)
    return gm["data"].reshape(S, M, D), gv["data"].reshape(S, M, D), w.reshape(S, M)


def read_sendump(path: str, n_feat: int, n_density: int, n_sen: int):
    dims = (C.c_int32 * 5)(n_feat, n_density, n_sen, 0, 0)
    check(lib.b200_s3_read_sendump(path.encode(), dims, None, None), "read_sendump")
    mixw = np.zeros((dims[0], dims[1], dims[4]), np.uint8)
    cb = np.zeros(16, np.uint8)
    dims2 = (C.c_int32 * 5)(n_feat, n_density, n_sen, 0, 0)
    check(lib.b200_s3_read_sendump(path.encode(), dims2, _p(mixw, C.c_uint8), _p(cb, C.c_uint8)), "read_sendump")
    return dict(n_feat=dims[0], n_density=dims[1], n_sen=dims[2], n_clust=dims[3], row_bytes=dims[4],
                mixw=mixw, mixw_cb=cb)


def mdef_maps(path: str):
    """Model definition (binary BMDF or text 0.3) -> dict(n_sen, n_ci_sen, n_ciphone, n_emit,
    sen2cimap int16[n_sen], cd2cisen int16[n_sen]) as bin_mdef_read builds them (PS/bin_mdef.c:462-497)."""
    dims = (C.c_int32 * 4)()
    check(lib.b200_mdef_read_maps(path.encode(), dims, None, None), "mdef_maps")
    a = np.zeros(dims[0], np.int16); c = np.zeros(dims[0], np.int16)
    check(lib.b200_mdef_read_maps(path.encode(), dims, _p(a, C.c_int16), _p(c, C.c_int16)), "mdef_maps")
    return dict(n_sen=dims[0], n_ci_sen=dims[1], n_ciphone=dims[2], n_emit=dims[3], sen2cimap=a, cd2cisen=c)


def sen_write(path: str, scores: np.ndarray, logbase: float = LOGBASE, mdef_file: str = "(null)"):
    """Dense senone-score dump readable by `-senin yes` / ps_decode_senscr (PS/acmod.c:349-361,885-923)."""
    sc = _c(scores, np.int16)
    check(lib.b200_sen_write(path.encode(), mdef_file.encode(), sc.shape[1], logbase, _p(sc, C.c_int16), sc.shape[0],
                             None, None), "sen_write")


def sen_read(path: str):
    """-> (scores [T][n_sen] int16 with 0x7fff for unlisted senones, n_active [T], logbase)."""
    dims = (C.c_int32 * 2)()
    lb = C.c_double(0)
    check(lib.b200_sen_read(path.encode(), dims, C.byref(lb), None, None), "sen_read")
    sc = np.zeros((dims[1], dims[0]), np.int16)
    na = np.zeros(dims[1], np.int32)
    check(lib.b200_sen_read(path.encode(), dims, C.byref(lb), _p(sc, C.c_int16), _p(na, C.c_int32)), "sen_read")
    return sc[:dims[1]], na[:dims[1]], lb.value


# --------------------------------------------------------------------- GMM
@dataclass
class MgauConfig:
    n_mgau: int
    n_feat: int
    n_density: int
    n_sen: int
    featlen: Sequence[int]
    topn: int = 4
    aw: int = 1
    ds_ratio: int = 1
    logbase: float = LOGBASE
    device: int = 0
    topn_beam: Sequence[int] = ()   # -topn_beam per stream (s2_semi only)

    def to_c(self) -> MgauCfg:
        c = MgauCfg()
        c.n_mgau, c.n_feat, c.n_density, c.n_sen = self.n_mgau, self.n_feat, self.n_density, self.n_sen
        for i, l in enumerate(self.featlen):
            c.featlen[i] = int(l)
        c.topn, c.aw, c.ds_ratio, c.logbase, c.device = self.topn, self.aw, self.ds_ratio, self.logbase, self.device
        for i, v in enumerate(self.topn_beam):
            c.topn_beam[i] = int(v)
        return c


class Mgau:
    """A ``ps_mgau_t`` living on the GPU (ms / ptm / s2_semi flavour)."""

    def __init__(self, handle, cfg: Optional[MgauConfig] = None):
        if not handle:
            raise B200Error(f"mgau init failed: {_lib.last_error()}")
        self._h = C.c_void_p(handle)
        self.cfg = cfg
        self.n_sen = lib.b200_mgau_n_sen(self._h)
        self.featdim = lib.b200_mgau_featdim(self._h)
        self.frame_idx = 0   # ps_mgau_t.frame_idx (PS/acmod.h:115)

    # vt->name
    @property
    def name(self) -> str:
        return lib.b200_mgau_name(self._h).decode()

    @property
    def path(self) -> int:
        return lib.b200_mgau_get_path(self._h)

    def set_path(self, path: int):
        check(lib.b200_mgau_set_path(self._h, path), "set_path")

    def tc_last_format(self) -> int:
        """1 = the last tensor-core call ran every tile on fp16 operands, 0 = on TF32, 2 = mixed, -1 = no plan."""
        return lib.b200_mgau_tc_last_format(self._h)

    def tied_stats(self):
        """(lists produced, lists re-done by the exact-scan fallback) of the last
        tensor-core scoring call of a ptm / s2_semi back-end."""
        out = (C.c_longlong * 3)()
        check(lib.b200_mgau_tied_stats(self._h, out), "tied_stats")
        self.tied_max_err = int(out[2])
        return int(out[0]), int(out[1])

    def cont_stats(self) -> dict:
        """Last tensor-core scoring call of a fully continuous ms back-end
        (include/b200sphinx.h: b200_mgau_cont_stats)."""
        out = (C.c_longlong * 7)()
        check(lib.b200_mgau_cont_stats(self._h, out), "cont_stats")
        k = ("pairs", "hard_pairs", "rescored_pairs", "rescanned_pairs", "overflow", "max_gemm_err", "hard_pairs_in_place")
        return dict(zip(k, (int(v) for v in out)))

    # vt->free
    def free(self):
        if self._h:
            lib.b200_mgau_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # vt->transform: the caller supplies already transformed + precomputed arrays
    def transform(self, mean, var, det):
        mean, var, det = _c(mean, np.float32), _c(var, np.float32), _c(det, np.float32)
        check(lib.b200_mgau_update_params(self._h, _p(mean, C.c_float), _p(var, C.c_float), _p(det, C.c_float)),
              "transform")

    # vt->frame_eval (PS/acmod.h:99-105)
    def frame_eval(self, senscr: np.ndarray, senone_active: Optional[np.ndarray], n_senone_active: int,
                   feat: Sequence[np.ndarray], frame: int, compallsen: bool) -> int:
        assert senscr.dtype == np.int16 and senscr.flags.c_contiguous and senscr.size >= self.n_sen
        streams = [_c(f, np.float32) for f in feat]
        ptrs = (C.POINTER(C.c_float) * len(streams))(*[_p(s, C.c_float) for s in streams])
        act = None
        if senone_active is not None:
            senone_active = _c(senone_active, np.uint8)
            act = _p(senone_active, C.c_uint8)
        check(lib.b200_mgau_frame_eval(self._h, _p(senscr, C.c_int16), act, int(n_senone_active), ptrs,
                                       int(frame), 1 if compallsen else 0), "frame_eval")
        return 0

    # batched compallsen scoring, host buffers (the e2e path)
    def score(self, feat: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        feat = _c(feat, np.float32)
        T = feat.shape[0] if feat.ndim == 2 else feat.size // max(self.featdim, 1)
        if out is None:
            out = np.empty((T, self.n_sen), np.int16)
        assert out.dtype == np.int16 and out.flags.c_contiguous and out.size >= T * self.n_sen
        check(lib.b200_mgau_score_host(self._h, feat.ctypes.data, T, out.ctypes.data), "score_host")
        return out

    def score_ptr(self, feat_ptr: int, T: int, out_ptr: int):
        """Host pointers (e.g. pinned torch tensors)."""
        check(lib.b200_mgau_score_host(self._h, feat_ptr, T, out_ptr), "score_host")

    def score_dev(self, d_feat: int, T: int, d_out: int, stream: int = 0):
        """Device pointers; synchronous and timed when stream == 0."""
        check(lib.b200_mgau_score_dev(self._h, d_feat, T, d_out, stream or None), "score_dev")

    def last_ms(self, which: int = 0) -> float:
        return lib.b200_mgau_last_ms(self._h, which)

    def timing_avg(self, n_calls: int, which: int = 2) -> float:
        """Mean device time (ms) of section `which` (0 total, 1 operand prep, 2
        scoring kernels, 3 normalise, 4 the exact fix-up kernels inside 2) over the last
        n_calls score_dev calls."""
        return lib.b200_mgau_timing_avg(self._h, n_calls, which)

    # utterance-batched serving (what the plug-in does)
    def utt_begin(self, feat: np.ndarray, frame0: int = 0):
        """Score a block of frames that starts at utterance frame `frame0` (only s2_semi -ds looks at it)."""
        feat = _c(feat, np.float32)
        T = feat.shape[0]
        check(lib.b200_mgau_utt_begin_at(self._h, feat.ctypes.data, T, int(frame0)), "utt_begin")

    def utt_frame(self, senscr: np.ndarray, senone_active: Optional[np.ndarray], n_senone_active: int,
                  frame: int, compallsen: bool):
        act = None
        if senone_active is not None:
            senone_active = _c(senone_active, np.uint8)
            act = _p(senone_active, C.c_uint8)
        check(lib.b200_mgau_utt_frame(self._h, _p(senscr, C.c_int16), act, int(n_senone_active), int(frame),
                                      1 if compallsen else 0), "utt_frame")


def ms_from_arrays(cfg: MgauConfig, mean, var_pre, det, mixw_q, sen2mgau) -> Mgau:
    """ms_mgau_init from already precomputed arrays (see b200_ms_create)."""
    mean, var_pre, det = _c(mean, np.float32), _c(var_pre, np.float32), _c(det, np.float32)
    mixw_q, sen2mgau = _c(mixw_q, np.uint8), _c(sen2mgau, np.uint32)
    c = cfg.to_c()
    h = lib.b200_ms_create(C.byref(c), _p(mean, C.c_float), _p(var_pre, C.c_float), _p(det, C.c_float),
                           _p(mixw_q, C.c_uint8), _p(sen2mgau, C.c_uint32))
    return Mgau(h, cfg)


def ms_from_files(mean: str, var: str, mixw: str, senmgau: str = ".cont.", sen2cb=None, varfloor=1e-4,
                  mixwfloor=1e-7, topn=4, aw=1, logbase=LOGBASE, device=0) -> Mgau:
    """ms_mgau_init(config) (PS/ms_mgau.c:79-141) on S3 parameter files."""
    s2c = None
    if sen2cb is not None:
        sen2cb = _c(sen2cb, np.uint8)
        s2c = _p(sen2cb, C.c_uint8)
    h = lib.b200_ms_load(mean.encode(), var.encode(), mixw.encode(), senmgau.encode(), s2c, varfloor,
                         mixwfloor, topn, aw, logbase, device)
    return Mgau(h, None)


def _tied(kind, cfg, mean, var_pre, det, mixw_rows, n_clust, mixw_cb, sen2cb):
    mean, var_pre, det = _c(mean, np.float32), _c(var_pre, np.float32), _c(det, np.float32)
    mixw_rows = _c(mixw_rows, np.uint8)
    cbp = None
    if n_clust:
        mixw_cb = _c(mixw_cb, np.uint8)
        cbp = _p(mixw_cb, C.c_uint8)
    c = cfg.to_c()
    if kind == 1:
        sen2cb = _c(sen2cb, np.uint8)
        h = lib.b200_ptm_create(C.byref(c), _p(mean, C.c_float), _p(var_pre, C.c_float), _p(det, C.c_float),
                                _p(mixw_rows, C.c_uint8), int(n_clust), cbp, _p(sen2cb, C.c_uint8))
    else:
        h = lib.b200_semi_create(C.byref(c), _p(mean, C.c_float), _p(var_pre, C.c_float), _p(det, C.c_float),
                                 _p(mixw_rows, C.c_uint8), int(n_clust), cbp)
    return Mgau(h, cfg)


def ptm_from_arrays(cfg, mean, var_pre, det, mixw_rows, sen2cb, n_clust=0, mixw_cb=None) -> Mgau:
    """ptm_mgau_init (PS/ptm_mgau.c:775-872) from precomputed arrays."""
    return _tied(1, cfg, mean, var_pre, det, mixw_rows, n_clust, mixw_cb, sen2cb)


def semi_from_arrays(cfg, mean, var_pre, det, mixw_rows, n_clust=0, mixw_cb=None) -> Mgau:
    """s2_semi_mgau_init (PS/s2_semi_mgau.c:1240-1330) from precomputed arrays."""
    return _tied(2, cfg, mean, var_pre, det, mixw_rows, n_clust, mixw_cb, None)


def tied_from_model_dir(hmmdir: str, n_sen: int, sen2cb=None, topn=4, varfloor=1e-4, mixwfloor=1e-7,
                        logbase=LOGBASE, device=0, topn_beam=(), ds_ratio=1) -> Mgau:
    """Back-end selection of acmod_init_am (PS/acmod.c:110-127) for a model
    directory holding means / variances / sendump|mixture_weights: one codebook
    -> s2_semi, otherwise ptm (sen2cb = bin_mdef sen2cimap must be given)."""
    import os
    g = read_gauden(os.path.join(hmmdir, "means"))
    v = read_gauden(os.path.join(hmmdir, "variances"))
    n_mgau, n_feat, n_density, veclen = g["n_mgau"], g["n_feat"], g["n_density"], g["veclen"]
    tot = sum(veclen)
    mean = g["data"].reshape(n_mgau, -1)
    var = v["data"].reshape(n_mgau, -1).copy()
    det = np.zeros((n_mgau, n_feat, n_density), np.float32)
    for mg in range(n_mgau):
        off = 0
        for f, l in enumerate(veclen):
            blk = var[mg, n_density * off:n_density * (off + l)].reshape(n_density, l)
            pv, pd = gauden_precompute(blk, l, varfloor, logbase)
            var[mg, n_density * off:n_density * (off + l)] = pv.reshape(-1)
            det[mg, f] = pd
            off += l
    assert mean.shape[1] == n_density * tot
    sd = os.path.join(hmmdir, "sendump")
    if os.path.exists(sd):
        d = read_sendump(sd, n_feat, n_density, n_sen)
        rows, n_clust, cb = d["mixw"], d["n_clust"], d["mixw_cb"]
    else:
        mw = read_mixw(os.path.join(hmmdir, "mixture_weights"))
        rows, n_clust, cb = mixw_quantize_tied(mw, mixwfloor, logbase), 0, None
    cfg = MgauConfig(n_mgau, n_feat, n_density, n_sen, veclen, topn=topn, logbase=logbase, device=device,
                     topn_beam=tuple(topn_beam), ds_ratio=ds_ratio)
    if n_mgau == 1:
        return semi_from_arrays(cfg, mean, var, det, rows, n_clust, cb)
    return ptm_from_arrays(cfg, mean, var, det, rows, sen2cb, n_clust, cb)


# --------------------------------------------------------------------- HMM
class HmmPopulation:
    """Structure-of-arrays batch of hmm_t, state-major ([state][hmm])."""

    def __init__(self, n_hmm: int, n_emit: int):
        self.n_hmm, self.n_emit = n_hmm, n_emit
        self.score = np.full((n_emit, n_hmm), WORST_SCORE, np.int32)
        self.history = np.full((n_emit, n_hmm), -1, np.int32)
        self.out_score = np.full(n_hmm, WORST_SCORE, np.int32)
        self.out_history = np.full(n_hmm, -1, np.int32)
        self.senid = np.zeros((n_emit, n_hmm), np.uint16)
        self.tmatid = np.zeros(n_hmm, np.int16)
        self.mpx = np.zeros(n_hmm, np.uint8)
        self.bestscore = np.full(n_hmm, WORST_SCORE, np.int32)

    def to_c(self) -> HmmSoa:
        for k in ("score", "history", "out_score", "out_history", "senid", "tmatid", "mpx", "bestscore"):
            a = getattr(self, k)
            assert a.flags.c_contiguous
        s = HmmSoa()
        s.n_hmm = self.n_hmm
        s.score, s.history = _p(self.score, C.c_int32), _p(self.history, C.c_int32)
        s.out_score, s.out_history = _p(self.out_score, C.c_int32), _p(self.out_history, C.c_int32)
        s.senid, s.tmatid = _p(self.senid, C.c_uint16), _p(self.tmatid, C.c_int16)
        s.mpx, s.bestscore = _p(self.mpx, C.c_uint8), _p(self.bestscore, C.c_int32)
        return s


class ChanTree:
    """The lexical tree of the forward tree search (root_chan_t / chan_t, ngram_search.h:64-104) on the
    GPU, and its prune / phone-transition stage: prune_root_chan + prune_nonroot_chan
    (ngram_search_fwdtree.c:714-869).  Arrays as include/b200sphinx.h describes them."""

    PAR = ("frame", "best_score", "beam", "pbeam", "lpbeam", "pip", "nwpen", "has_pls")

    def __init__(self, n_root, n_chan, child_off, child, ciphone, pw_off, pw_wid, pw_lastphone, n_ci, n_emit=3, device=0):
        a = [_c(x, np.int32) for x in (child_off, child, ciphone, pw_off, pw_wid, pw_lastphone)]
        h = lib.b200_chantree_create(int(n_root), int(n_chan), *[_p(x, C.c_int32) for x in a], int(n_ci), int(n_emit), device)
        if not h:
            raise B200Error(f"b200_chantree_create failed: {_lib.last_error()}")
        self._h = C.c_void_p(h)
        self.n_root, self.n_chan, self.n_ci, self.n_emit = int(n_root), int(n_chan), int(n_ci), int(n_emit)
        self.cand_cap = lib.b200_chantree_cand_cap(self._h)

    def free(self):
        if self._h:
            lib.b200_chantree_free(self._h)
            self._h = None

    def prune(self, par, pls_pen, acl_lists, score, history, out_score, out_history, bestscore, frame, list_cap=None):
        """One frame for n_utt utterances.  par: [n_utt][8] (ChanTree.PAR order); pls_pen [n_utt][n_ci] or
        None; acl_lists: one int array per utterance; state arrays state-major over n_utt * n_chan channels,
        updated in place.  Returns (next active lists, candidate arrays [n][3]) per utterance."""
        par = _c(par, np.int32).reshape(-1, 8)
        n_utt = par.shape[0]
        cap = max(1, self.n_chan - self.n_root) if list_cap is None else int(list_cap)
        acl = np.zeros((n_utt, cap), np.int32)
        n_act = np.zeros(n_utt, np.int32)
        for u, l in enumerate(acl_lists):
            l = np.asarray(l, np.int32)
            if len(l) > cap:
                raise B200Error("active list longer than list_cap")
            acl[u, :len(l)] = l
            n_act[u] = len(l)
        pen = None if pls_pen is None else _c(pls_pen, np.int32).reshape(n_utt, self.n_ci)
        for a in (score, history, out_score, out_history, bestscore, frame):
            assert a.dtype == np.int32 and a.flags.c_contiguous
        assert score.size == self.n_emit * n_utt * self.n_chan and frame.size == n_utt * self.n_chan
        ccap = max(1, self.cand_cap)
        nacl = np.zeros((n_utt, cap), np.int32)
        cand = np.zeros((n_utt, ccap, 3), np.int32)
        n_nacl, n_cand = np.zeros(n_utt, np.int32), np.zeros(n_utt, np.int32)
        check(lib.b200_fwdtree_prune_host(self._h, n_utt, _p(par, C.c_int32), None if pen is None else _p(pen, C.c_int32),
                                          _p(acl, C.c_int32), _p(n_act, C.c_int32), cap, _p(score, C.c_int32),
                                          _p(history, C.c_int32), _p(out_score, C.c_int32), _p(out_history, C.c_int32),
                                          _p(bestscore, C.c_int32), _p(frame, C.c_int32), _p(nacl, C.c_int32),
                                          _p(n_nacl, C.c_int32), _p(cand, C.c_int32), _p(n_cand, C.c_int32), ccap),
              "fwdtree_prune_host")
        return [nacl[u, :n_nacl[u]].copy() for u in range(n_utt)], [cand[u, :n_cand[u]].copy() for u in range(n_utt)]

    def prune_resident(self, ctx, frame, par, pls_pen, acl_lists):
        """The same stage on the RESIDENT population of an HmmContext (the tree's channels of n_utt utterances,
        as left by ctx.step / run_dev): b200_fwdtree_prune_dev on the context's stream, the channel states never
        leave HBM.  frame: host int32 [n_utt * n_chan] (hmm_frame; updated in place).  Only the lists come back."""
        par = _c(par, np.int32).reshape(-1, 8)
        n_utt, nc = par.shape[0], self.n_chan
        soa, st = ctx.device_arrays()
        assert soa.n_hmm == n_utt * nc and frame.dtype == np.int32 and frame.size == n_utt * nc
        cap, ccap = max(1, nc - self.n_root), max(1, self.cand_cap)
        acl = np.zeros((n_utt, cap), np.int32)
        n_act = np.array([len(l) for l in acl_lists], np.int32)
        for u, l in enumerate(acl_lists):
            acl[u, :len(l)] = l
        pen = np.zeros((n_utt, self.n_ci), np.int32) if pls_pen is None else _c(pls_pen, np.int32).reshape(n_utt, self.n_ci)
        host = [frame, par, pen, acl, n_act]
        sizes = [a.nbytes for a in host] + [n_utt * cap * 4, n_utt * 4, n_utt * ccap * 12, n_utt * 4]
        dev = [lib.b200_dev_alloc(max(1, n), 0) for n in sizes]
        try:
            if not all(dev):
                raise B200Error("device allocation failed")
            for a, d in zip(host, dev):
                check(lib.b200_dev_upload(d, a.ctypes.data, a.nbytes), "upload")
            p = _lib.PruneDev()
            as_vp = lambda x: C.cast(x, C.c_void_p)
            p.score, p.history, p.out_score = as_vp(soa.score), as_vp(soa.history), as_vp(soa.out_score)
            p.out_history, p.bestscore = as_vp(soa.out_history), as_vp(soa.bestscore)
            p.frame, p.state_stride = dev[0], soa.n_hmm
            p.par, p.pls_pen, p.acl, p.n_act, p.list_cap = dev[1], dev[2], dev[3], dev[4], cap
            p.nacl, p.n_nacl, p.cand, p.n_cand, p.cand_cap = dev[5], dev[6], dev[7], dev[8], ccap
            check(lib.b200_fwdtree_prune_dev(self._h, n_utt, C.byref(p), st), "fwdtree_prune_dev")
            check(lib.b200_dev_sync(0), "sync")
            nacl, cand = np.zeros((n_utt, cap), np.int32), np.zeros((n_utt, ccap, 3), np.int32)
            n_nacl, n_cand = np.zeros(n_utt, np.int32), np.zeros(n_utt, np.int32)
            for a, d in ((frame, dev[0]), (nacl, dev[5]), (n_nacl, dev[6]), (cand, dev[7]), (n_cand, dev[8])):
                check(lib.b200_dev_download(a.ctypes.data, d, a.nbytes), "download")
        finally:
            for d in dev:
                if d:
                    lib.b200_dev_free(d)
        return [nacl[u, :n_nacl[u]].copy() for u in range(n_utt)], [cand[u, :n_cand[u]].copy() for u in range(n_utt)]


class FwdtreeDevice:
    """The tree-internal part of one forward-tree frame, resident on the device for a batch of utterances:
    evaluate_channels' eval_root_chan + eval_nonroot_chan (b200_hmm_eval_list_dev) and prune_channels'
    prune_root_chan + prune_nonroot_chan (b200_fwdtree_prune_dev) alternate on the channel states of an
    HmmContext's resident population; the active list one stage writes is the list the next one reads.  What
    crosses the host link per frame: senone scores in, per-utterance best / list lengths / last-phone
    candidates out (the LM-dependent last_phone_transition and word_transition stay on the host)."""

    def __init__(self, tree: "ChanTree", ctx: "HmmContext", n_utt: int, frame: np.ndarray):
        self.tree, self.ctx, self.n_utt = tree, ctx, n_utt
        nc = tree.n_chan
        self.cap, self.ccap = max(1, nc - tree.n_root), max(1, tree.cand_cap)
        sizes = dict(frame=n_utt * nc * 4, par=n_utt * 32, pen=n_utt * tree.n_ci * 4, acl0=n_utt * self.cap * 4,
                     acl1=n_utt * self.cap * 4, n0=n_utt * 4, n1=n_utt * 4, cand=n_utt * self.ccap * 12, n_cand=n_utt * 4,
                     best=n_utt * 4, senscr=n_utt * ctx.n_sen * 2)
        self.d = {k: lib.b200_dev_alloc(v, 0) for k, v in sizes.items()}
        if not all(self.d.values()):
            raise B200Error("device allocation failed")
        self.cur = 0
        frame = _c(frame, np.int32)
        assert frame.size == n_utt * nc
        check(lib.b200_dev_upload(self.d["frame"], frame.ctypes.data, frame.nbytes), "upload")
        self.set_lists([np.zeros(0, np.int32)] * n_utt)

    def free(self):
        for v in self.d.values():
            lib.b200_dev_free(v)
        self.d = {}

    def _up(self, key, a):
        check(lib.b200_dev_upload(self.d[key], a.ctypes.data, a.nbytes), "upload")

    def _down(self, key, a):
        check(lib.b200_dev_download(a.ctypes.data, self.d[key], a.nbytes), "download")
        return a

    def set_lists(self, lists):
        acl = np.zeros((self.n_utt, self.cap), np.int32)
        n = np.array([len(l) for l in lists], np.int32)
        for u, l in enumerate(lists):
            acl[u, :len(l)] = l
        self._up(f"acl{self.cur}", acl)
        self._up(f"n{self.cur}", n)

    def lists(self):
        acl = self._down(f"acl{self.cur}", np.zeros((self.n_utt, self.cap), np.int32))
        n = self._down(f"n{self.cur}", np.zeros(self.n_utt, np.int32))
        return [acl[u, :n[u]].copy() for u in range(self.n_utt)]

    def frame_stamps(self):
        return self._down("frame", np.zeros(self.n_utt * self.tree.n_chan, np.int32))

    def set_frame_stamps(self, frame):
        self._up("frame", _c(frame, np.int32))

    def evaluate(self, senscr, frame_idx):
        """-> best[n_utt] over the evaluated tree channels (the caller folds in its word channels)."""
        par = np.zeros((self.n_utt, 8), np.int32)
        par[:, 0] = frame_idx
        self._up("par", par)
        self._up("senscr", _c(senscr, np.int16).reshape(self.n_utt, self.ctx.n_sen))
        t = self.tree
        check(lib.b200_hmm_eval_list_dev(self.ctx._h, t.n_root, t.n_chan, self.d["frame"], self.d["par"], self.d[f"acl{self.cur}"],
                                         self.d[f"n{self.cur}"], self.cap, self.d["senscr"], self.d["best"], None), "hmm_eval_list_dev")
        check(lib.b200_dev_sync(0), "sync")
        return self._down("best", np.zeros(self.n_utt, np.int32))

    def _args(self, with_pen=False):
        soa, st = self.ctx.device_arrays()
        p = _lib.PruneDev()
        as_vp = lambda x: C.cast(x, C.c_void_p)
        p.score, p.history, p.out_score = as_vp(soa.score), as_vp(soa.history), as_vp(soa.out_score)
        p.out_history, p.bestscore = as_vp(soa.out_history), as_vp(soa.bestscore)
        p.frame, p.state_stride = self.d["frame"], soa.n_hmm
        p.par, p.pls_pen = self.d["par"], (self.d["pen"] if with_pen else None)
        p.acl, p.n_act, p.list_cap = self.d[f"acl{self.cur}"], self.d[f"n{self.cur}"], self.cap
        return p, st

    def renormalize(self, frame_idx, norm):
        """renormalize_scores' tree part (ngram_search_fwdtree.c:557-576) by norm[n_utt]."""
        par = np.zeros((self.n_utt, 8), np.int32)
        par[:, 0] = frame_idx
        self._up("par", par)
        self._up("best", _c(norm, np.int32))
        p, st = self._args()
        check(lib.b200_fwdtree_renorm_dev(self.tree._h, self.n_utt, C.byref(p), self.d["best"], st), "fwdtree_renorm_dev")
        check(lib.b200_dev_sync(0), "sync")

    def deactivate(self, frame_idx):
        """deactivate_channels' root loop (ngram_search_fwdtree.c:1418-1431)."""
        par = np.zeros((self.n_utt, 8), np.int32)
        par[:, 0] = frame_idx
        self._up("par", par)
        p, st = self._args()
        check(lib.b200_fwdtree_deactivate_dev(self.tree._h, self.n_utt, C.byref(p), st), "fwdtree_deactivate_dev")
        check(lib.b200_dev_sync(0), "sync")

    def prune(self, par, pls_pen=None):
        """par [n_utt][8] (ChanTree.PAR order).  -> list of candidate arrays; the next list stays on the device."""
        par = _c(par, np.int32).reshape(self.n_utt, 8)
        self._up("par", par)
        if pls_pen is not None:
            self._up("pen", _c(pls_pen, np.int32).reshape(self.n_utt, self.tree.n_ci))
        p, st = self._args(pls_pen is not None)
        nxt = 1 - self.cur
        p.nacl, p.n_nacl, p.cand, p.n_cand, p.cand_cap = self.d[f"acl{nxt}"], self.d[f"n{nxt}"], self.d["cand"], self.d["n_cand"], self.ccap
        check(lib.b200_fwdtree_prune_dev(self.tree._h, self.n_utt, C.byref(p), st), "fwdtree_prune_dev")
        check(lib.b200_dev_sync(0), "sync")
        self.cur = nxt
        n_cand = self._down("n_cand", np.zeros(self.n_utt, np.int32))
        cand = self._down("cand", np.zeros((self.n_utt, self.ccap, 3), np.int32))
        return [cand[u, :n_cand[u]].copy() for u in range(self.n_utt)]


class PhoneLoop:
    """The phone-loop look-ahead search (phone_loop_search.c) for n_utt utterances in lock step, its phone HMMs
    resident on the device: start() = phone_loop_search_start, step() = phone_loop_search_step minus the acmod
    calls; step returns (best_score [n_utt], pls_pen [n_utt][n_phones]) -- the penalties
    phone_loop_search_score hands to prune_root_chan / prune_nonroot_chan."""

    def __init__(self, tp, senid, tmatid, n_sen, beam, pbeam, pip, n_utt=1, device=0):
        tp, senid, tmatid = _c(tp, np.uint8), _c(senid, np.uint16), _c(tmatid, np.int16)
        self.n_emit, self.n_phones = senid.shape
        assert tp.shape[1] == self.n_emit and tp.shape[2] == self.n_emit + 1
        h = lib.b200_phone_loop_create(self.n_phones, self.n_emit, _p(tp, C.c_uint8), tp.shape[0], _p(senid, C.c_uint16),
                                       _p(tmatid, C.c_int16), int(n_sen), int(beam), int(pbeam), int(pip), int(n_utt), device)
        if not h:
            raise B200Error(f"b200_phone_loop_create failed: {_lib.last_error()}")
        self._h, self.n_utt, self.n_sen = C.c_void_p(h), int(n_utt), int(n_sen)

    def free(self):
        if self._h:
            lib.b200_phone_loop_free(self._h)
            self._h = None

    def start(self):
        check(lib.b200_phone_loop_start(self._h), "phone_loop_start")

    def set_state(self, score, history, out_score, out_history, bestscore, frame, best):
        a = [_c(x, np.int32) for x in (score, history, out_score, out_history, bestscore, frame, best)]
        N = self.n_utt * self.n_phones
        assert a[0].size == self.n_emit * N and a[2].size == N and a[6].size == self.n_utt
        check(lib.b200_phone_loop_set_state(self._h, *[_p(x, C.c_int32) for x in a]), "phone_loop_set_state")

    def state(self):
        N = self.n_utt * self.n_phones
        out = dict(score=np.zeros((self.n_emit, N), np.int32), history=np.zeros((self.n_emit, N), np.int32),
                   out_score=np.zeros(N, np.int32), out_history=np.zeros(N, np.int32), bestscore=np.zeros(N, np.int32),
                   frame=np.zeros(N, np.int32), best=np.zeros(self.n_utt, np.int32), renorm=np.zeros(self.n_utt, np.int32))
        check(lib.b200_phone_loop_get_state(self._h, *[_p(out[k], C.c_int32) for k in
                                                       ("score", "history", "out_score", "out_history", "bestscore", "frame", "best", "renorm")]),
              "phone_loop_get_state")
        return out

    def step(self, senscr, frame_idx):
        s = _c(senscr, np.int16).reshape(self.n_utt, self.n_sen)
        pen = np.zeros((self.n_utt, self.n_phones), np.int32)
        best = np.zeros(self.n_utt, np.int32)
        check(lib.b200_phone_loop_step_host(self._h, _p(s, C.c_int16), int(frame_idx), _p(pen, C.c_int32), _p(best, C.c_int32)),
              "phone_loop_step_host")
        return best, pen


def s3hmm_vit_eval(n_emit: int, tp, sseq, n_sen: int, senscr, score, history, out_score, out_history, ssid, tmatid, mpx,
                   bestscore, device: int = 0):
    """sphinx3's hmm_vit_eval (libs3decoder/libam/hmm.c:852-873) for every HMM, once per row of
    senscr ([n_frames][n_sen] int32).  State-major int32 arrays [n_emit][n_hmm] (score, history,
    ssid), updated in place; tp int32 [n_tmat][n_emit][n_emit + 1]; sseq int16 [n_sseq][n_emit].
    Returns the per-frame best scores."""
    tp, sseq = _c(tp, np.int32), _c(sseq, np.int16)
    senscr = _c(senscr, np.int32).reshape(-1, n_sen)
    for a in (score, history, ssid, out_score, out_history, bestscore):
        assert a.dtype == np.int32 and a.flags.c_contiguous
    tm, mp = _c(tmatid, np.int32), _c(mpx, np.uint8)
    soa = _lib.S3HmmSoa()
    soa.n_hmm = out_score.shape[0]
    soa.score, soa.history, soa.ssid = _p(score, C.c_int32), _p(history, C.c_int32), _p(ssid, C.c_int32)
    soa.out_score, soa.out_history, soa.bestscore = _p(out_score, C.c_int32), _p(out_history, C.c_int32), _p(bestscore, C.c_int32)
    soa.tmatid, soa.mpx = _p(tm, C.c_int32), _p(mp, C.c_uint8)
    best = np.zeros(senscr.shape[0], np.int32)
    check(lib.b200_s3hmm_eval_host(n_emit, _p(tp, C.c_int32), tp.shape[0], _p(sseq, C.c_int16), sseq.shape[0], n_sen,
                                   C.byref(soa), _p(senscr, C.c_int32), senscr.shape[0], _p(best, C.c_int32), device),
          "s3hmm_eval_host")
    return best


class HmmContext:
    """hmm_context_t on the GPU: hmm_context_init(n_emit, tp, senscore, sseq)."""

    def __init__(self, n_emit: int, tp: np.ndarray, sseq: Optional[np.ndarray], n_sen: int, device: int = 0):
        tp = _c(tp, np.uint8)
        assert tp.ndim == 3 and tp.shape[1] == n_emit and tp.shape[2] == n_emit + 1
        n_sseq = 0 if sseq is None else sseq.shape[0]
        ss = None if sseq is None else _c(sseq, np.uint16)
        h = lib.b200_hmm_ctx_create(n_emit, _p(tp, C.c_uint8), tp.shape[0],
                                    None if ss is None else _p(ss, C.c_uint16), n_sseq, n_sen, device)
        if not h:
            raise B200Error(f"hmm_context_init failed: {_lib.last_error()}")
        self._h = C.c_void_p(h)
        self.n_emit, self.n_sen = n_emit, n_sen

    def free(self):
        if self._h:
            lib.b200_hmm_ctx_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def vit_eval(self, pop: HmmPopulation, senscr: np.ndarray) -> np.ndarray:
        """hmm_vit_eval for every HMM, once per row of senscr ([n_frames][n_sen]
        int16); pop is updated in place.  Returns the per-frame best score."""
        senscr = _c(senscr, np.int16).reshape(-1, self.n_sen)
        best = np.zeros(senscr.shape[0], np.int32)
        soa = pop.to_c()
        check(lib.b200_hmm_eval_host(self._h, C.byref(soa), _p(senscr, C.c_int16), senscr.shape[0],
                                     _p(best, C.c_int32)), "hmm_eval_host")
        return best

    def upload(self, pop: HmmPopulation):
        self.n_utt = 1
        soa = pop.to_c()
        check(lib.b200_hmm_pop_upload(self._h, C.byref(soa)), "pop_upload")

    def download(self, pop: HmmPopulation):
        soa = pop.to_c()
        check(lib.b200_hmm_pop_download(self._h, C.byref(soa)), "pop_download")

    def device_arrays(self):
        """(HmmSoa of DEVICE pointers, stream) of the resident population (b200_hmm_pop_device)."""
        soa, st = HmmSoa(), C.c_void_p()
        check(lib.b200_hmm_pop_device(self._h, C.byref(soa), C.byref(st)), "pop_device")
        return soa, st

    def set_utts(self, utt_off):
        """Split the resident population into utterances (see b200_hmm_pop_set_utts)."""
        off = _c(utt_off, np.int32)
        self.n_utt = off.size - 1
        check(lib.b200_hmm_pop_set_utts(self._h, self.n_utt, _p(off, C.c_int32)), "pop_set_utts")

    def step(self, senscr, beam: int, n_hmm: int, want_idx=True):
        """One search frame on the resident population: eval + beam + compaction +
        active-senone gather.  senscr: host int16 [n_utt][n_sen] or a device
        pointer.  Returns (best, survivors, mask); per-utterance arrays when the
        population is batched."""
        nu = getattr(self, "n_utt", 1)
        if isinstance(senscr, np.ndarray):
            s = _c(senscr, np.int16)
            check(lib.b200_hmm_step_host(self._h, _p(s, C.c_int16), int(beam)), "hmm_step_host")
        else:
            check(lib.b200_hmm_step_dev(self._h, senscr, int(beam), None), "hmm_step_dev")
        return self.step_results(n_hmm, want_idx)

    def step_results(self, n_hmm: int, want_idx=True):
        """(best, survivors, mask) of the last step / the last frame of run_dev."""
        nu = getattr(self, "n_utt", 1)
        best, nk = np.zeros(nu, np.int32), np.zeros(nu, np.int32)
        idx = np.zeros(n_hmm, np.int32) if want_idx else None
        mask = np.zeros((nu, (self.n_sen + 31) // 32), np.uint32)
        check(lib.b200_hmm_step_results(self._h, _p(best, C.c_int32), _p(nk, C.c_int32),
                                        _p(idx, C.c_int32) if want_idx else None, _p(mask, C.c_uint32)),
              "hmm_step_results")
        if nu == 1:
            return int(best[0]), (idx[:nk[0]] if want_idx else int(nk[0])), mask[0]
        return best, (idx[:nk.sum()] if want_idx else nk), mask

    def normalize(self, best_per_utt=None):
        """hmm_normalize over the resident population (PS/hmm.c:207-218); None = the
        per-utterance best of the last step()."""
        if best_per_utt is None:
            check(lib.b200_hmm_normalize_dev(self._h, None, None), "hmm_normalize")
            check(lib.b200_dev_sync(0), "sync")
            return
        b = _c(best_per_utt, np.int32)
        d = lib.b200_dev_alloc(b.nbytes, 0)
        try:
            check(lib.b200_dev_upload(d, b.ctypes.data, b.nbytes), "upload")
            check(lib.b200_hmm_normalize_dev(self._h, d, None), "hmm_normalize")
            check(lib.b200_dev_sync(0), "sync")
        finally:
            lib.b200_dev_free(d)

    def clear_pruned(self):
        """hmm_clear_scores for every HMM the last step()'s beam dropped."""
        check(lib.b200_hmm_clear_pruned_dev(self._h, None), "hmm_clear_pruned")
        check(lib.b200_dev_sync(0), "sync")

    def enter(self, idx, score, hist):
        """Batched hmm_enter with the callers' `only if better` test, list-order semantics."""
        i, s, h = _c(idx, np.int32), _c(score, np.int32), _c(hist, np.int32)
        check(lib.b200_hmm_enter_host(self._h, _p(i, C.c_int32), _p(s, C.c_int32), _p(h, C.c_int32), i.size), "hmm_enter")

    def run_dev(self, d_senscr: int, frame_stride: int, n_cycle: int, n_frames: int, beam: int, stream=None):
        """n_frames steps on device-resident senone scores (frame f at d_senscr + (f % n_cycle) * frame_stride
        int16 elements) in one persistent launch.  Read the last frame with step_results()."""
        check(lib.b200_hmm_run_dev(self._h, d_senscr, int(frame_stride), int(n_cycle), int(n_frames), int(beam), stream),
              "hmm_run_dev")

    def step_dev_async(self, d_senscr: int, beam: int):
        check(lib.b200_hmm_step_dev(self._h, d_senscr, int(beam), None), "hmm_step_dev")

    def last_ms(self) -> float:
        return lib.b200_hmm_last_ms(self._h)


# ------------------------------------------------------------ feature stage
def feat_1s_c_d_dd(cep, utt_off=None, cmn: bool = True, device: int = 0) -> np.ndarray:
    """Cepstra [T][13] of one or several utterances (utt_off: frame offsets,
    len n_utt+1) -> [T][39] features: feat_s2mfc2feat_block_utt + cmn current +
    feat_1s_c_d_dd_cep2feat (SB/feat/feat.c:726-769,1241-1265; cmn.c:150-186)."""
    cep = _c(cep, np.float32)
    T, cs = cep.shape
    off = np.array([0, T], np.int32) if utt_off is None else _c(utt_off, np.int32)
    out = np.zeros((T, 3 * cs), np.float32)
    check(lib.b200_feat_1s_c_d_dd_host(_p(cep, C.c_float), _p(off, C.c_int32), off.size - 1, cs, 1 if cmn else 0,
                                       _p(out, C.c_float), device), "feat_1s_c_d_dd")
    return out


FEAT_TYPES = ["1s_c_d_dd", "s3_1x39", "s2_4x", "1s_c_d_ld_dd", "1s_c", "1s_c_d"]


def parse_feat_type(ftype: str):
    """-feat string -> (type id, copy window, copy stream lengths): the named types of feat_init
    (SB/feat/feat.c:866-951) or its numeric form "n[,n..][:w]" (feat.c:952-1000)."""
    if ftype in FEAT_TYPES:
        return FEAT_TYPES.index(ftype), 0, []
    if ftype in ("1s_3c", "1s_4c"):
        # feat_s3_cepwin (feat.c:687-696) copies (2w+1)*cepsize CONTIGUOUS floats starting at frame t-w,
        # but the padded utterance is not contiguous (feat.c:1248-1259): its first and last w output
        # frames read stale / out-of-bounds memory in the reference.  Nothing to be identical to.
        raise ValueError(f"-feat {ftype}: undefined at the utterance edges in the reference (feat_s3_cepwin); "
                         f"use the numeric form '13:{ftype[3]}'")
    if ftype and ftype[0].isdigit():
        body, _, w = ftype.partition(":")
        return 6, int(w) if w else 0, [int(x) for x in body.split(",")]
    raise ValueError(f"unknown -feat type {ftype}")


class _FeatCfg(C.Structure):
    _fields_ = [("type", C.c_int32), ("cepsize", C.c_int32), ("cmn", C.c_int32), ("varnorm", C.c_int32),
                ("agc", C.c_int32), ("lda_rows", C.c_int32), ("lda_cols", C.c_int32), ("lda_dim", C.c_int32),
                ("lda", C.POINTER(C.c_float)), ("n_subvec", C.c_int32), ("subvec", C.POINTER(C.c_int32)),
                ("copy_window", C.c_int32), ("copy_streams", C.c_int32), ("copy_len", C.c_int32 * 4)]


lib.b200_feat_dims.argtypes = [C.POINTER(_FeatCfg), C.POINTER(C.c_int32)]
lib.b200_feat_compute_host.argtypes = [C.POINTER(_FeatCfg), C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int,
                                       C.POINTER(C.c_float), C.c_int]


def feat_compute(cep, utt_off=None, ftype: str = "1s_c_d_dd", cmn: bool = True, varnorm: bool = False,
                 agc: bool = False, lda=None, lda_dim: int = 0, subvec=None, device: int = 0) -> np.ndarray:
    """feat_s2mfc2feat_block_utt / feat_compute_utt for a batch of utterances
    (SB/feat/feat.c:1110-1135, 1241-1265): `-feat ftype`, `-cmn current|none`,
    `-varnorm`, `-agc max|none`, `-lda` ([rows][stream_len] matrix, `-ldadim`),
    `-svspec` (flat index list).  Returns [T][out_len] float32."""
    cep = _c(cep, np.float32)
    T, cs = cep.shape
    off = np.array([0, T], np.int32) if utt_off is None else _c(utt_off, np.int32)
    cfg = _FeatCfg()
    tid, cw, clen = parse_feat_type(ftype)
    cfg.type, cfg.cepsize, cfg.cmn, cfg.varnorm, cfg.agc = tid, cs, int(cmn), int(varnorm), int(agc)
    cfg.copy_window, cfg.copy_streams = cw, len(clen)
    for j, l in enumerate(clen):
        cfg.copy_len[j] = l
    keep = []
    if lda is not None:
        lda = _c(lda, np.float32)
        keep.append(lda)
        cfg.lda_rows, cfg.lda_cols, cfg.lda_dim, cfg.lda = lda.shape[0], lda.shape[1], lda_dim, _p(lda, C.c_float)
    if subvec is not None and len(subvec):
        sv = _c(np.asarray(subvec), np.int32)
        keep.append(sv)
        cfg.n_subvec, cfg.subvec = sv.size, _p(sv, C.c_int32)
    dims = (C.c_int32 * 3)()
    check(lib.b200_feat_dims(C.byref(cfg), dims), "feat_dims")
    out = np.zeros((T, dims[2]), np.float32)
    check(lib.b200_feat_compute_host(C.byref(cfg), _p(cep, C.c_float), _p(off, C.c_int32), off.size - 1,
                                     _p(out, C.c_float), device), "feat_compute")
    return out


# ------------------------------------------------------------ sphinx3 GMM
S3_LOGBASE = float(np.float32(1.0003))   # sphinx3 -logbase default (float32 option, cmdln_macro.h:246)


class S3Mgau:
    """sphinx3's mgau_model_t + fast_gmm_t behind the C ABI (S3/libam/cont_mgau.c,
    approx_cont_mgau.c).  Scores are int32, higher = better."""

    def __init__(self, handle):
        if not handle:
            raise B200Error(f"sphinx3 scorer: {_lib.last_error()}")
        self.h = handle
        d = (C.c_int32 * 5)()
        check(lib.b200_s3_dims(self.h, d), "s3_dims")
        self.n_sen, self.max_comp, self.veclen, self.n_ci_sen, self.ci_pbeam = (int(v) for v in d)

    @classmethod
    def from_arrays(cls, mean, var, mixw, cd2cisen, n_ci_sen, varfloor=1e-4, mixwfloor=1e-7, logbase=S3_LOGBASE,
                    device=0):
        mean, var, mixw = _c(mean, np.float32), _c(var, np.float32), _c(mixw, np.float32)
        S, M, D = mean.shape
        cd = _c(cd2cisen, np.int32)
        return cls(lib.b200_s3_create(S, M, D, _p(mean, C.c_float), _p(var, C.c_float), _p(mixw, C.c_float),
                                      varfloor, mixwfloor, logbase, _p(cd, C.c_int32), n_ci_sen, device))

    @classmethod
    def from_files(cls, meanfile, varfile, mixwfile, cd2cisen, n_ci_sen, varfloor=1e-4, mixwfloor=1e-7,
                   logbase=S3_LOGBASE, device=0):
        cd = _c(cd2cisen, np.int32)
        return cls(lib.b200_s3_load(meanfile.encode(), varfile.encode(), mixwfile.encode(), varfloor, mixwfloor,
                                    logbase, _p(cd, C.c_int32), n_ci_sen, device))

    def set_fast(self, ci_pbeam=1e-80, max_cd=100000, ds_ratio=1, tighten=0.5):
        check(lib.b200_s3_set_fast(self.h, ci_pbeam, max_cd, ds_ratio, tighten), "s3_set_fast")
        d = (C.c_int32 * 5)()
        lib.b200_s3_dims(self.h, d)
        self.ci_pbeam = int(d[4])

    def set_subvq(self, path, varfloor=1e-4, max_sv=-1, vqeval=3, subvqbeam=1e-3):
        """-subvq / -svmax / -vqeval / -subvqbeam (S3/libam/subvq.c); path None removes the layer."""
        check(lib.b200_s3_set_subvq(self.h, None if path is None else path.encode(), varfloor, max_sv, vqeval, subvqbeam),
              "s3_set_subvq")

    def set_gs(self, path):
        """-gs (S3/libam/gs.c); path None removes the layer."""
        check(lib.b200_s3_set_gs(self.h, None if path is None else path.encode()), "s3_set_gs")

    def utt_reset(self):
        check(lib.b200_s3_utt_reset(self.h), "s3_utt_reset")

    def params(self):
        S, M, D = self.n_sen, self.max_comp, self.veclen
        nc = np.zeros(S, np.int32)
        mean = np.zeros((S, M, D), np.float32); var = np.zeros((S, M, D), np.float32)
        lrd = np.zeros((S, M), np.float32); mixw = np.zeros((S, M), np.int32); scal = np.zeros(2, np.float64)
        check(lib.b200_s3_params(self.h, _p(nc, C.c_int32), _p(mean, C.c_float), _p(var, C.c_float),
                                 _p(lrd, C.c_float), _p(mixw, C.c_int32), _p(scal, C.c_double)), "s3_params")
        return nc, mean, var, lrd, mixw, scal

    def state(self):
        b = np.zeros(self.n_sen, np.int32); u = np.zeros(self.n_sen, np.int32)
        check(lib.b200_s3_state(self.h, _p(b, C.c_int32), _p(u, C.c_int32)), "s3_state")
        return b, u

    def eval_dense(self, feat):
        """mgau_eval(g, s, NULL, x, t, 1) for every frame and senone -> int32 [T][n_sen]."""
        feat = _c(feat, np.float32)
        out = np.zeros((feat.shape[0], self.n_sen), np.int32)
        check(lib.b200_s3_dense_host(self.h, _p(feat, C.c_float), feat.shape[0], _p(out, C.c_int32)), "s3_dense")
        return out

    def eval_utt(self, feat, sen_active=None, frame0=0, senscr0=None):
        """Per frame: CI pass + approx_cont_mgau_frame_eval.  -> (senscr [T][S], best [T], sen_active after)."""
        feat = _c(feat, np.float32)
        T = feat.shape[0]
        act = None if sen_active is None else np.ascontiguousarray(sen_active, np.uint8).copy()
        io = np.zeros(self.n_sen, np.int32) if senscr0 is None else _c(senscr0, np.int32).copy()
        out = np.zeros((T, self.n_sen), np.int32)
        best = np.zeros(T, np.int32)
        check(lib.b200_s3_score_utt_host(self.h, _p(feat, C.c_float), T, frame0,
                                         None if act is None else _p(act, C.c_uint8), _p(io, C.c_int32),
                                         _p(out, C.c_int32), _p(best, C.c_int32)), "s3_score_utt")
        self.last_row = io
        return out, best, act

    def frame_eval(self, x, frame, sen_active, senscr):
        """In-place single-frame drop-in (gmm_compute_lv1 + lv2); returns best."""
        x = _c(x, np.float32)
        best = C.c_int32(0)
        check(lib.b200_s3_frame_eval(self.h, _p(x, C.c_float), frame, _p(sen_active, C.c_uint8),
                                     _p(senscr, C.c_int32), C.byref(best)), "s3_frame_eval")
        return best.value

    def last_ms(self):
        return lib.b200_s3_last_ms(self.h)

    def free(self):
        if self.h:
            lib.b200_s3_free(self.h)
            self.h = None
