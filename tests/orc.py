"""Test-only access to the oracle (oracle/liboracle.so: our C restatement) and,
when present, the reference itself (oracle/_ref/libref_shim.so: the unmodified
reference sources compiled by oracle/Makefile).  Never imported by the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
DATA_DIR = os.path.join(REF_DIR, "data")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# -logbase is a float32 option (cmd_ln_float32_r, pocketsphinx.c:223): the
# reference's effective base is (double)(float)1.0001.
LOGBASE = float(np.float32(1.0001))

f32p, i32p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
i16p, u16p, u8p = C.POINTER(C.c_int16), C.POINTER(C.c_uint16), C.POINTER(C.c_uint8)
vp = C.c_void_p


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _load_port():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "sphinx_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
    L = C.CDLL(so)
    L.orc_logmath_init.restype = vp
    L.orc_logmath_init.argtypes = [C.c_double, C.c_int, C.c_int]
    L.orc_logmath_free.argtypes = [vp]
    L.orc_logmath_table.argtypes = [vp, i32p, C.c_int]
    L.orc_logmath_log.argtypes = [vp, C.c_double]
    L.orc_logmath_add.argtypes = [vp, C.c_int32, C.c_int32]
    L.orc_gauden_precompute.argtypes = [f32p, f32p, C.c_long, C.c_int, C.c_float, C.c_double]
    L.orc_mixw_quantize.argtypes = [f32p, u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_double]
    L.orc_mixw_quantize_tied.argtypes = [f32p, u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_double]
    L.orc_tmat_quantize.argtypes = [f32p, u8p, C.c_int, C.c_int, C.c_double, C.c_double]
    L.orc_flags2list.argtypes = [u32p, C.c_int, u8p]
    L.orc_ms_model_new.restype = vp
    L.orc_ms_model_new.argtypes = [C.c_int, C.c_int, i32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, f32p, f32p,
                                   u8p, u32p, C.c_double]
    L.orc_ms_model_free.argtypes = [vp]
    L.orc_ms_frame_eval.argtypes = [vp, f32p, u8p, C.c_int, C.c_int, i16p]
    L.orc_ms_eval_all.argtypes = [vp, f32p, C.c_int, i16p]
    L.orc_tied_new.restype = vp
    L.orc_tied_new.argtypes = [C.c_int, C.c_int, C.c_int, i32p, C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, u8p,
                               C.c_int, C.c_int, u8p, u8p, C.c_double]
    L.orc_tied_free.argtypes = [vp]
    L.orc_tied_reset.argtypes = [vp]
    L.orc_tied_set_topn_beam.argtypes = [vp, i32p]
    L.orc_tied_set_ds.argtypes = [vp, C.c_int]
    L.orc_tied_frame_eval.argtypes = [vp, f32p, u8p, C.c_int, C.c_int, C.c_int, i16p]
    L.orc_tied_eval_all.argtypes = [vp, f32p, C.c_int, i16p]
    L.orc_tied_lists.argtypes = [vp, i32p, i32p]
    L.orc_hmm_eval_batch.restype = C.c_int32
    L.orc_hmm_eval_batch.argtypes = [C.c_int, C.c_int, u8p, C.c_int, u16p, C.c_int, i16p, i32p, i32p, i32p, i32p,
                                     u16p, u16p, i16p, u8p, i32p, C.c_int]
    return L


port = _load_port()


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref_shim.so"))


_ref = None


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(REF_DIR, "libref_shim.so"))
        L.ref_logadd_table.argtypes = [C.c_double, C.c_int, i32p, C.c_int]
        L.ref_logmath_zero.argtypes = [C.c_double, C.c_int]
        L.ref_logmath_log.argtypes = [C.c_double, C.c_int, C.POINTER(C.c_double), C.c_int, i32p]
        L.ref_logmath_add.argtypes = [C.c_double, C.c_int, i32p, i32p, C.c_int, i32p]
        L.ref_ms_init.restype = vp
        L.ref_ms_init.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_double, C.c_double, C.c_int,
                                  C.c_int, C.c_double]
        L.ref_ms_free.argtypes = [vp]
        L.ref_ms_dims.argtypes = [vp, i32p, i32p]
        L.ref_ms_params.argtypes = [vp, f32p, f32p, f32p, u8p]
        L.ref_ms_eval_all.argtypes = [vp, f32p, C.c_int, i16p]
        L.ref_ms_eval_active.argtypes = [vp, f32p, u8p, C.c_int, C.c_int, i16p]
        L.ref_acmod_open_ex.restype = vp
        L.ref_acmod_open_ex.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_char_p]
        L.ref_acmod_open.restype = vp
        L.ref_acmod_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_double]
        L.ref_acmod_close.argtypes = [vp]
        L.ref_acmod_backend.restype = C.c_char_p
        L.ref_acmod_backend.argtypes = [vp]
        L.ref_acmod_info.argtypes = [vp, i32p, i32p]
        L.ref_acmod_score_feats.argtypes = [vp, f32p, C.c_int, i16p]
        L.ref_acmod_frame_eval.argtypes = [vp, f32p, u8p, C.c_int, C.c_int, C.c_int, i16p]
        L.ref_acmod_sen2cimap.argtypes = [vp, u8p]
        L.ref_acmod_cep2feat.argtypes = [vp, f32p, C.c_int, C.c_int, f32p, C.c_int]
        L.ref_feat_compute.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, f32p, C.c_int, C.c_int,
                                       C.c_int, C.c_char_p, f32p, C.c_int, f32p, i32p, i32p]
        L.ref_acmod_tables.argtypes = [vp, i32p, u8p, u16p]
        L.ref_tmat_load.argtypes = [C.c_char_p, C.c_double, C.c_double, u8p, C.c_int, i32p]
        L.ref_hmm_eval_batch.restype = C.c_int32
        L.ref_hmm_eval_batch.argtypes = [C.c_int, C.c_int, u8p, C.c_int, u16p, C.c_int, i16p, i32p, i32p, i32p,
                                         i32p, u16p, u16p, i16p, u8p, i32p, C.c_int]
        L.ref_hmm_maint.argtypes = [C.c_int, C.c_int, C.c_int, i32p, i32p, i32p, i32p, i32p, u8p, i32p, C.c_int, i32p,
                                    i32p, i32p]
        _ref = L
    return _ref


def ref_hmm_maint(op, score, history, out_score, out_history, bestscore, sel=None, arg=None, lidx=None, lscore=None,
                  lhist=None):
    """The reference's hmm_clear_scores (op 0, on sel) / hmm_normalize (op 1, by arg) / entry loop with
    hmm_enter (op 2, list order) on HMM-major arrays [n_hmm][n_emit]; returns the updated copies."""
    sc, hi = _c(score, np.int32).copy(), _c(history, np.int32).copy()
    os_, oh, bs = _c(out_score, np.int32).copy(), _c(out_history, np.int32).copy(), _c(bestscore, np.int32).copy()
    n_hmm, n_emit = sc.shape
    z8, z32 = np.zeros(1, np.uint8), np.zeros(1, np.int32)
    sel = z8 if sel is None else _c(sel, np.uint8)
    arg = z32 if arg is None else _c(arg, np.int32)
    n_list = 0 if lidx is None else len(lidx)
    li = z32 if lidx is None else _c(lidx, np.int32)
    ls = z32 if lscore is None else _c(lscore, np.int32)
    lh = z32 if lhist is None else _c(lhist, np.int32)
    ref().ref_hmm_maint(op, n_emit, n_hmm, _p(sc, C.c_int32), _p(hi, C.c_int32), _p(os_, C.c_int32), _p(oh, C.c_int32),
                        _p(bs, C.c_int32), _p(sel, C.c_uint8), _p(arg, C.c_int32), n_list, _p(li, C.c_int32),
                        _p(ls, C.c_int32), _p(lh, C.c_int32))
    return sc, hi, os_, oh, bs


# ------------------------------------------------------------------ wrappers
def port_logadd_table(base=1.0001, shift=10):
    lm = port.orc_logmath_init(base, shift, 1)
    n = port.orc_logmath_table(lm, None, 0)
    out = np.zeros(n, np.int32)
    port.orc_logmath_table(lm, _p(out, C.c_int32), n)
    port.orc_logmath_free(lm)
    return out


def port_precompute(var, length, varfloor=1e-4, logbase=LOGBASE):
    v = _c(var, np.float32).copy().reshape(-1, length)
    det = np.zeros(v.shape[0], np.float32)
    port.orc_gauden_precompute(_p(v, C.c_float), _p(det, C.c_float), v.shape[0], length, varfloor, logbase)
    return v.reshape(var.shape), det.reshape(var.shape[:-1])


def port_mixw_quantize(mixw, floor=1e-7, logbase=LOGBASE):
    m = _c(mixw, np.float32).copy()
    out = np.zeros(m.shape, np.uint8)
    port.orc_mixw_quantize(_p(m, C.c_float), _p(out, C.c_uint8), m.shape[0], m.shape[1], m.shape[2], floor, logbase)
    return out


def port_mixw_quantize_tied(mixw, floor=1e-7, logbase=LOGBASE):
    m = _c(mixw, np.float32).copy()
    out = np.zeros((m.shape[1], m.shape[2], m.shape[0]), np.uint8)
    port.orc_mixw_quantize_tied(_p(m, C.c_float), _p(out, C.c_uint8), m.shape[0], m.shape[1], m.shape[2], floor,
                                logbase)
    return out


def port_tmat_quantize(tp, floor=1e-4, logbase=LOGBASE):
    t = _c(tp, np.float32).copy()
    out = np.zeros(t.shape, np.uint8)
    port.orc_tmat_quantize(_p(t, C.c_float), _p(out, C.c_uint8), t.shape[0], t.shape[1], floor, logbase)
    return out


def port_flags2list(mask, n_sen):
    m = _c(mask, np.uint32)
    out = np.zeros(2 * n_sen + 8, np.uint8)
    n = port.orc_flags2list(_p(m, C.c_uint32), n_sen, _p(out, C.c_uint8))
    return out[:n].copy()


class PortMs:
    """orc_ms_model_t: arrays are PRECOMPUTED mean/var/det [mgau][feat][density][len], mixw [sen][feat][cw]."""

    def __init__(self, n_mgau, n_feat, featlen, n_density, n_sen, topn, aw, mean, var, det, mixw, sen2mgau,
                 logbase=LOGBASE):
        self.keep = [_c(mean, np.float32), _c(var, np.float32), _c(det, np.float32), _c(mixw, np.uint8),
                     _c(sen2mgau, np.uint32), _c(featlen, np.int32)]
        k = self.keep
        self.h = port.orc_ms_model_new(n_mgau, n_feat, _p(k[5], C.c_int32), n_density, n_sen, topn, aw,
                                       _p(k[0], C.c_float), _p(k[1], C.c_float), _p(k[2], C.c_float),
                                       _p(k[3], C.c_uint8), _p(k[4], C.c_uint32), logbase)
        self.n_sen = n_sen
        self.veclen = int(sum(featlen))

    def eval_all(self, feat):
        feat = _c(feat, np.float32).reshape(-1, self.veclen)
        out = np.zeros((feat.shape[0], self.n_sen), np.int16)
        port.orc_ms_eval_all(self.h, _p(feat, C.c_float), feat.shape[0], _p(out, C.c_int16))
        return out

    def frame_eval(self, feat, deltas, compallsen, senscr=None):
        feat = _c(feat, np.float32)
        if senscr is None:
            senscr = np.zeros(self.n_sen, np.int16)
        d = _c(deltas if deltas is not None else [], np.uint8)
        port.orc_ms_frame_eval(self.h, _p(feat, C.c_float), _p(d, C.c_uint8), d.size, 1 if compallsen else 0,
                               _p(senscr, C.c_int16))
        return senscr

    def __del__(self):
        port.orc_ms_model_free(self.h)


class PortTied:
    def __init__(self, kind, n_mgau, n_feat, featlen, n_density, n_sen, topn, mean, var, det, mixw_rows, n_clust,
                 mixw_cb, sen2cb, logbase=LOGBASE):
        mixw_rows = _c(mixw_rows, np.uint8)
        self.keep = [_c(mean, np.float32), _c(var, np.float32), _c(det, np.float32), mixw_rows,
                     _c(mixw_cb if mixw_cb is not None else np.zeros(16), np.uint8),
                     _c(sen2cb if sen2cb is not None else np.zeros(n_sen), np.uint8), _c(featlen, np.int32)]
        k = self.keep
        self.h = port.orc_tied_new(kind, n_mgau, n_feat, _p(k[6], C.c_int32), n_density, n_sen, topn,
                                   _p(k[0], C.c_float), _p(k[1], C.c_float), _p(k[2], C.c_float),
                                   _p(k[3], C.c_uint8), mixw_rows.shape[-1], int(n_clust), _p(k[4], C.c_uint8),
                                   _p(k[5], C.c_uint8), logbase)
        self.n_sen, self.veclen = n_sen, int(sum(featlen))
        self.shape = (n_mgau, n_feat, topn)

    def reset(self):
        port.orc_tied_reset(self.h)

    def set_ds(self, ds):
        port.orc_tied_set_ds(self.h, int(ds))

    def set_topn_beam(self, beam):
        bm = _c(list(beam) + [0] * 8, np.int32)
        port.orc_tied_set_topn_beam(self.h, _p(bm, C.c_int32))

    def eval_all(self, feat):
        feat = _c(feat, np.float32).reshape(-1, self.veclen)
        out = np.zeros((feat.shape[0], self.n_sen), np.int16)
        port.orc_tied_eval_all(self.h, _p(feat, C.c_float), feat.shape[0], _p(out, C.c_int16))
        return out

    def frame_eval(self, feat, deltas, compallsen, frame):
        feat = _c(feat, np.float32)
        out = np.zeros(self.n_sen, np.int16)
        d = _c(deltas if deltas is not None else [], np.uint8)
        port.orc_tied_frame_eval(self.h, _p(feat, C.c_float), _p(d, C.c_uint8), d.size, 1 if compallsen else 0,
                                 frame, _p(out, C.c_int16))
        return out

    def lists(self):
        cw = np.zeros(self.shape, np.int32)
        sc = np.zeros(self.shape, np.int32)
        port.orc_tied_lists(self.h, _p(cw, C.c_int32), _p(sc, C.c_int32))
        return cw, sc

    def __del__(self):
        port.orc_tied_free(self.h)


def hmm_eval(fn, n_emit, tp, sseq, senscr, score, history, out_score, out_history, senid, tmatid, mpx, bestscore,
             repeat=1):
    """fn = port.orc_hmm_eval_batch or ref().ref_hmm_eval_batch.  HMM-major arrays [hmm][state], in place."""
    n_hmm = out_score.shape[0]
    tp = _c(tp, np.uint8)
    sseq = _c(sseq, np.uint16)
    ssid = np.zeros(n_hmm, np.uint16)
    return fn(n_emit, n_hmm, _p(tp, C.c_uint8), tp.shape[0], _p(sseq, C.c_uint16), sseq.shape[0],
              _p(senscr, C.c_int16), _p(score, C.c_int32), _p(history, C.c_int32), _p(out_score, C.c_int32),
              _p(out_history, C.c_int32), _p(senid, C.c_uint16), _p(ssid, C.c_uint16), _p(tmatid, C.c_int16),
              _p(mpx, C.c_uint8), _p(bestscore, C.c_int32), repeat)


class RefAcmod:
    """The reference's acmod + whichever back-end it selects, on a model directory."""

    def __init__(self, hmmdir, senmgau="", topn=4, ds=1, logbase=LOGBASE, topn_beam=""):
        self.h = ref().ref_acmod_open_ex(hmmdir.encode(), senmgau.encode(), topn, ds, logbase, topn_beam.encode())
        if not self.h:
            raise RuntimeError(f"reference acmod_init failed for {hmmdir}")
        info = (C.c_int32 * 4)()
        sl = (C.c_int32 * 8)()
        ref().ref_acmod_info(self.h, info, sl)
        self.n_sen, self.n_feat, self.featdim, self.n_emit = info[0], info[1], info[2], info[3]
        self.streamlen = [sl[i] for i in range(self.n_feat)]
        self.backend = ref().ref_acmod_backend(self.h).decode()

    def cep2feat(self, cep):
        cep = _c(cep, np.float32).copy()   # the reference normalises its input IN PLACE (feat.c:1308)
        out = np.zeros((cep.shape[0] + 16, self.featdim), np.float32)
        n = ref().ref_acmod_cep2feat(self.h, _p(cep, C.c_float), cep.shape[0], cep.shape[1], _p(out, C.c_float),
                                     out.shape[0])
        return out[:n].copy()

    def score(self, feat):
        feat = _c(feat, np.float32)
        out = np.zeros((feat.shape[0], self.n_sen), np.int16)
        ref().ref_acmod_score_feats(self.h, _p(feat, C.c_float), feat.shape[0], _p(out, C.c_int16))
        return out

    def frame_eval(self, feat, deltas, frame, compallsen, out=None):
        feat = _c(feat, np.float32)
        if out is None:
            out = np.zeros(self.n_sen, np.int16)
        d = _c(deltas if deltas is not None else [], np.uint8)
        ref().ref_acmod_frame_eval(self.h, _p(feat, C.c_float), _p(d, C.c_uint8), d.size, frame,
                                   1 if compallsen else 0, _p(out, C.c_int16))
        return out

    def sen2cimap(self):
        out = np.zeros(self.n_sen, np.uint8)
        ref().ref_acmod_sen2cimap(self.h, _p(out, C.c_uint8))
        return out

    def tables(self):
        sizes = (C.c_int32 * 3)()
        ref().ref_acmod_tables(self.h, sizes, None, None)
        tp = np.zeros((sizes[0], sizes[1], sizes[1] + 1), np.uint8)
        sseq = np.zeros((sizes[2], self.n_emit), np.uint16)
        ref().ref_acmod_tables(self.h, sizes, _p(tp, C.c_uint8), _p(sseq, C.c_uint16))
        return tp, sseq

    def close(self):
        if self.h:
            ref().ref_acmod_close(self.h)
            self.h = None


def read_mfc(path):
    raw = np.fromfile(path, dtype="<i4", count=1)
    n = int(raw[0])
    data = np.fromfile(path, dtype="<f4", offset=4)
    if data.size != n:
        n = int(np.frombuffer(raw.tobytes(), dtype=">i4")[0])
        data = np.fromfile(path, dtype=">f4", offset=4).astype(np.float32)
    return data.reshape(-1, 13).astype(np.float32)


# ------------------------------------------------------- sphinx3 flavour
# sphinx3's -logbase default is the float32 option 1.0003 (cmdln_macro.h:246).
S3_LOGBASE = float(np.float32(1.0003))
S3_ZERO = np.int32(-939524096)   # 0xc8000000, s3types.h:192
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)

port.orc_s3_new.restype = vp
port.orc_s3_new.argtypes = [C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, C.c_double, C.c_double, C.c_double, i32p, C.c_int]
port.orc_s3_free.argtypes = [vp]
port.orc_s3_set_fast.argtypes = [vp, C.c_double, C.c_int, C.c_int, C.c_float]
port.orc_s3_ci_pbeam.restype = C.c_int32
port.orc_s3_ci_pbeam.argtypes = [vp]
port.orc_s3_utt_reset.argtypes = [vp]
port.orc_s3_params.argtypes = [vp, i32p, f32p, f32p, f32p, i32p, f64p]
port.orc_s3_state.argtypes = [vp, i32p, i32p]
port.orc_s3_mgau_eval.restype = C.c_int32
port.orc_s3_mgau_eval.argtypes = [vp, C.c_int, i32p, f32p, C.c_int, C.c_int]
port.orc_s3_eval_utt.argtypes = [vp, f32p, C.c_int, C.c_int, u8p, i32p, i32p, i32p]
port.orc_s3_counts.argtypes = [vp, i64p]


def have_ref_s3():
    return os.path.exists(os.path.join(REF_DIR, "libref_shim_s3.so"))


_ref3 = None


def ref_s3():
    global _ref3
    if _ref3 is None:
        L = C.CDLL(os.path.join(REF_DIR, "libref_shim_s3.so"))
        L.ref_s3_open.restype = vp
        L.ref_s3_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, i32p, C.c_int, C.c_int, C.c_double,
                                  C.c_double, C.c_double]
        L.ref_s3_set_fast.argtypes = [vp, C.c_double, C.c_int, C.c_int, C.c_float]
        L.ref_s3_dims.argtypes = [vp, i32p]
        L.ref_s3_cd2cisen.argtypes = [vp, i32p]
        L.ref_s3_params.argtypes = [vp, i32p, f32p, f32p, f32p, i32p, f64p]
        L.ref_s3_utt_reset.argtypes = [vp]
        L.ref_s3_state.argtypes = [vp, i32p, i32p]
        L.ref_s3_eval_dense.argtypes = [vp, f32p, C.c_int, i32p]
        L.ref_s3_eval_utt.argtypes = [vp, f32p, C.c_int, C.c_int, u8p, i32p, i32p, i32p]
        L.ref_s3_close.argtypes = [vp]
        _ref3 = L
    return _ref3


class _S3Common:
    """Same surface for the port (PortS3) and the reference shim (RefS3)."""

    def eval_utt(self, feat, sen_active=None, frame0=0, senscr0=None):
        """-> (out [T][S] int32, best [T], sen_active after (or None))."""
        feat = _c(feat, np.float32)
        T = feat.shape[0]
        act = None if sen_active is None else np.ascontiguousarray(sen_active, np.uint8).copy()
        io = np.zeros(self.n_sen, np.int32) if senscr0 is None else _c(senscr0, np.int32).copy()
        out = np.zeros((T, self.n_sen), np.int32)
        best = np.zeros(T, np.int32)
        self._eval_utt(self.h, _p(feat, C.c_float), T, frame0, None if act is None else _p(act, C.c_uint8),
                       _p(io, C.c_int32), _p(out, C.c_int32), _p(best, C.c_int32))
        self.last_row = io
        return out, best, act

    def params(self):
        S, M, D = self.n_sen, self.max_comp, self.veclen
        nc = np.zeros(S, np.int32)
        mean = np.zeros((S, M, D), np.float32); var = np.zeros((S, M, D), np.float32)
        lrd = np.zeros((S, M), np.float32); mixw = np.zeros((S, M), np.int32); scal = np.zeros(2, np.float64)
        self._params(self.h, _p(nc, C.c_int32), _p(mean, C.c_float), _p(var, C.c_float), _p(lrd, C.c_float),
                     _p(mixw, C.c_int32), _p(scal, C.c_double))
        return nc, mean, var, lrd, mixw, scal

    def state(self):
        b = np.zeros(self.n_sen, np.int32); u = np.zeros(self.n_sen, np.int32)
        self._state(self.h, _p(b, C.c_int32), _p(u, C.c_int32))
        return b, u

    def set_fast(self, ci_pbeam=1e-80, max_cd=100000, ds_ratio=1, tighten=0.5):
        self._set_fast(self.h, ci_pbeam, max_cd, ds_ratio, tighten)

    def utt_reset(self):
        self._reset(self.h)


class PortS3(_S3Common):
    def __init__(self, mean, var, mixw, cd2cisen, n_ci_sen, varfloor=1e-4, mixwfloor=1e-7, logbase=S3_LOGBASE):
        mean, var, mixw = _c(mean, np.float32), _c(var, np.float32), _c(mixw, np.float32)
        self.n_sen, self.max_comp, self.veclen = mean.shape
        cd = _c(cd2cisen, np.int32)
        self.h = port.orc_s3_new(self.n_sen, self.max_comp, self.veclen, _p(mean, C.c_float), _p(var, C.c_float),
                                 _p(mixw, C.c_float), varfloor, mixwfloor, logbase, _p(cd, C.c_int32), n_ci_sen)
        self._eval_utt, self._params, self._state = port.orc_s3_eval_utt, port.orc_s3_params, port.orc_s3_state
        self._set_fast, self._reset = port.orc_s3_set_fast, port.orc_s3_utt_reset

    def free(self):
        port.orc_s3_free(self.h)


class RefS3(_S3Common):
    def __init__(self, meanfile, varfile, mixwfile, mdef_file=None, cd2cisen=None, n_ci_sen=0, varfloor=1e-4,
                 mixwfloor=1e-7, logbase=S3_LOGBASE):
        L = ref_s3()
        cd = None if cd2cisen is None else _c(cd2cisen, np.int32)
        self.h = L.ref_s3_open(meanfile.encode(), varfile.encode(), mixwfile.encode(),
                               None if mdef_file is None else mdef_file.encode(),
                               None if cd is None else _p(cd, C.c_int32), 0 if cd is None else len(cd), n_ci_sen,
                               varfloor, mixwfloor, logbase)
        d = np.zeros(5, np.int32)
        L.ref_s3_dims(self.h, _p(d, C.c_int32))
        self.n_sen, self.max_comp, self.veclen, self.n_ci_sen, self.ci_pbeam = (int(v) for v in d)
        self._eval_utt, self._params, self._state = L.ref_s3_eval_utt, L.ref_s3_params, L.ref_s3_state
        self._set_fast, self._reset = L.ref_s3_set_fast, L.ref_s3_utt_reset

    def cd2cisen(self):
        out = np.zeros(self.n_sen, np.int32)
        ref_s3().ref_s3_cd2cisen(self.h, _p(out, C.c_int32))
        return out

    def eval_dense(self, feat):
        feat = _c(feat, np.float32)
        out = np.zeros((feat.shape[0], self.n_sen), np.int32)
        ref_s3().ref_s3_eval_dense(self.h, _p(feat, C.c_float), feat.shape[0], _p(out, C.c_int32))
        return out

    def free(self):
        ref_s3().ref_s3_close(self.h)


# ------------------------------------------------------- sphinx3 sub-vector quantised shortlists
port.orc_s3_set_svq.restype = C.c_int
port.orc_s3_set_svq.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, i32p, i32p, f32p, f32p, i32p, C.c_double, C.c_double]
port.orc_s3_svq_tables.argtypes = [vp, C.c_int, f32p, f32p, f32p, f64p]
port.orc_s3_svq_map.argtypes = [vp, i32p]
port.orc_s3_svq_beam.restype = C.c_int32
port.orc_s3_svq_beam.argtypes = [vp]
port.orc_s3_svq_eval.argtypes = [vp, f32p]
port.orc_s3_svq_dist.argtypes = [vp, i32p]
port.orc_s3_svq_shortlist.restype = C.c_int
port.orc_s3_svq_shortlist.argtypes = [vp, C.c_int]


def read_subvq(path):
    """Independent reader of the text format subvq_init parses (S3/libam/subvq.c:206-340):
    -> dict(r, c, n_sv, vqsize, veclen [n_sv], featdim [list of arrays], mean / var [list of [vqsize][veclen]],
    map [r][c][n_sv] int32 (entries < 0 = component not present))."""
    lines = open(path).read().split("\n")
    i = 0
    while not lines[i].startswith("VQParam"):
        i += 1
    t = lines[i].split()
    r, c, n_sv, vqsize = int(t[1]), int(t[2]), int(t[4]), int(t[5])
    i += 1
    veclen, featdim = [], []
    for s in range(n_sv):
        t = lines[i].split(); i += 1
        assert t[0] == "Subvector" and int(t[1]) == s
        L = int(t[3]); veclen.append(L); featdim.append(np.array([int(x) for x in t[4:4 + L]], np.int32))
    mean, var = [], []
    mp = np.zeros((r, c, n_sv), np.int32)
    for s in range(n_sv):
        assert lines[i].split()[:2] == ["Codebook", str(s)]; i += 1
        mv = np.array([[float(x) for x in lines[i + k].split()] for k in range(vqsize)], np.float64); i += vqsize
        mean.append(np.ascontiguousarray(mv[:, 0::2], np.float32)); var.append(np.ascontiguousarray(mv[:, 1::2], np.float32))
        assert lines[i].split()[:2] == ["Map", str(s)]; i += 1
        for k in range(r):
            mp[k, :, s] = [int(x) for x in lines[i + k].split()]
        i += r
    assert lines[i].strip() == "End"
    return dict(r=r, c=c, n_sv=n_sv, vqsize=vqsize, veclen=np.array(veclen, np.int32), featdim=featdim, mean=mean, var=var, map=mp)


def write_subvq(path, q):
    """The same format from arrays (synthetic sub-VQ models for the tests)."""
    with open(path, "w") as f:
        f.write("VQParam %d %d -> %d %d\n" % (q["r"], q["c"], q["n_sv"], q["vqsize"]))
        for s in range(q["n_sv"]):
            f.write("Subvector %d length %d " % (s, q["veclen"][s]) + " ".join("%2d" % d for d in q["featdim"][s]) + "\n")
        for s in range(q["n_sv"]):
            f.write("Codebook %d Sqerr 0.0\n" % s)
            for k in range(q["vqsize"]):
                f.write(" ".join("%.8e %.8e" % (q["mean"][s][k, i], q["var"][s][k, i]) for i in range(q["veclen"][s])) + "\n")
            f.write("Map %d\n" % s)
            for k in range(q["r"]):
                f.write(" ".join("%d" % v for v in q["map"][k, :, s]) + "\n")
        f.write("End\n")


def synthetic_subvq(mean, var, n_comp_valid, n_sv, vqsize, seed=5):
    """A sub-VQ model for a synthetic acoustic model: sub-vectors = contiguous slices of the
    feature vector, codewords = randomly chosen (mean, variance) sub-vectors of the model's own
    Gaussians, map = nearest codeword by Euclidean distance of the means; components the
    acoustic model drops (mgau_uninit_compact) are marked -1 as gausubvq writes them."""
    rng = np.random.default_rng(seed)
    S, M, D = mean.shape
    edges = np.linspace(0, D, n_sv + 1).astype(int)
    q = dict(r=S, c=M, n_sv=n_sv, vqsize=vqsize, veclen=np.diff(edges).astype(np.int32), featdim=[], mean=[], var=[],
             map=np.zeros((S, M, n_sv), np.int32))
    flat_m, flat_v = mean.reshape(S * M, D), var.reshape(S * M, D)
    for s in range(n_sv):
        dims = np.arange(edges[s], edges[s + 1], dtype=np.int32)
        pick = rng.choice(np.flatnonzero(n_comp_valid.ravel()), vqsize, replace=False)
        cm, cv = flat_m[pick][:, dims].copy(), np.maximum(flat_v[pick][:, dims], 1e-3).copy()
        q["featdim"].append(dims); q["mean"].append(cm.astype(np.float32)); q["var"].append(cv.astype(np.float32))
        d2 = ((flat_m[:, None, dims] - cm[None]) ** 2).sum(-1)
        q["map"][:, :, s] = d2.argmin(1).reshape(S, M)
    q["map"][~n_comp_valid] = -1
    return q


def port_set_svq(p3, q, varfloor=1e-4, max_sv=-1, vqeval=3, subvqbeam=1e-3):
    """Attach the sub-VQ model `q` (read_subvq) to a PortS3."""
    veclen = _c(q["veclen"], np.int32)
    fd = _c(np.concatenate(q["featdim"]), np.int32)
    mean = _c(np.concatenate([m.ravel() for m in q["mean"]]), np.float32)
    var = _c(np.concatenate([v.ravel() for v in q["var"]]), np.float32)
    mp = _c(q["map"], np.int32)
    rc = port.orc_s3_set_svq(p3.h, q["n_sv"], max_sv, q["vqsize"], vqeval, _p(veclen, C.c_int32), _p(fd, C.c_int32),
                             _p(mean, C.c_float), _p(var, C.c_float), _p(mp, C.c_int32), varfloor, subvqbeam)
    assert rc == 0, "sub-VQ map does not match the model's components"
    p3.svq = q
    p3.n_sv_use = q["n_sv"] if max_sv < 0 else min(max_sv, q["n_sv"])


def ref_set_svq(r3, path, varfloor=1e-4, max_sv=-1, vqeval=3, subvqbeam=1e-3):
    L = ref_s3()
    L.ref_s3_open_svq.restype = C.c_int
    L.ref_s3_open_svq.argtypes = [vp, C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_double]
    L.ref_s3_svq_dims.argtypes = [vp, i32p]
    L.ref_s3_svq_tables.restype = C.c_int
    L.ref_s3_svq_tables.argtypes = [vp, C.c_int, i32p, f32p, f32p, f32p, f64p]
    L.ref_s3_svq_map.argtypes = [vp, i32p]
    L.ref_s3_svq_vqdist.argtypes = [vp, f32p, C.c_int, i32p]
    L.ref_s3_svq_shortlist.restype = C.c_int
    L.ref_s3_svq_shortlist.argtypes = [vp, C.c_int, u8p]
    assert L.ref_s3_open_svq(r3.h, path.encode(), varfloor, max_sv, vqeval, subvqbeam) == 0
    d = np.zeros(6, np.int32)
    L.ref_s3_svq_dims(r3.h, _p(d, C.c_int32))
    r3.svq_dims = [int(v) for v in d]       # n_sv, vqsize, r, c, VQ_EVAL, (beam valid after set_fast)


# ------------------------------------------------------- sphinx3 Gaussian selector (S3/libam/gs.c)
port.orc_s3_set_gs.argtypes = [vp, C.c_int, C.c_int, f32p, C.POINTER(C.c_uint32)]
port.orc_s3_gs_closest.restype = C.c_int
port.orc_s3_gs_closest.argtypes = [vp, f32p]


def write_gs(path, codeword, bits, n_density):
    """The binary map gs_read parses (gs.c:156-218): five int32 (n_mgau, n_feat = 1, n_density, n_code, featlen),
    then per codeword its featlen float32 followed by one bit vector (bitvec_size(n_density) uint32 words, native
    byte order) per mixture.  bits [n_mgau][n_code] uint32 = the first word (the only one the reference keeps)."""
    codeword = _c(codeword, np.float32); bits = _c(bits, np.uint32)
    n_code, featlen = codeword.shape
    n_mgau = bits.shape[0]
    words = (n_density + 31) // 32
    with open(path, "wb") as f:
        f.write(np.array([n_mgau, 1, n_density, n_code, featlen], np.int32).tobytes())
        for k in range(n_code):
            f.write(codeword[k].tobytes())
            row = np.zeros((n_mgau, words), np.uint32); row[:, 0] = bits[:, k]
            f.write(row.tobytes())


def synthetic_gs(mean, n_code, seed=9):
    """Codewords = means of randomly picked Gaussians (codeword 0 far away from everything: the reference asserts
    best_cid > 0, approx_cont_mgau.c:209); map bit c of (senone, codeword) set when component c's mean is among the
    senone's nearest half to the codeword -- so shortlists really drop components; a few maps are left empty to
    exercise the all-components fall-back of gs_mgau_shortlist."""
    rng = np.random.default_rng(seed)
    S, M, D = mean.shape
    cw = mean.reshape(S * M, D)[rng.choice(S * M, n_code, replace=False)].copy()
    cw[0] = 1e4
    d2 = ((mean[:, None, :, :] - cw[None, :, None, :]) ** 2).sum(-1)            # [S][n_code][M]
    keep = d2 <= np.median(d2, axis=2, keepdims=True)
    bits = (keep * (np.uint64(1) << np.arange(M, dtype=np.uint64))[None, None, :]).sum(-1).astype(np.uint32)
    bits[rng.random(bits.shape) < 0.02] = 0
    return cw.astype(np.float32), bits


def port_set_gs(p3, codeword, bits):
    codeword = _c(codeword, np.float32); bits = _c(bits, np.uint32)
    port.orc_s3_set_gs(p3.h, codeword.shape[0], codeword.shape[1], _p(codeword, C.c_float), bits.ctypes.data_as(C.POINTER(C.c_uint32)))


def ref_set_gs(r3, path):
    L = ref_s3()
    L.ref_s3_open_gs.restype = C.c_int
    L.ref_s3_open_gs.argtypes = [vp, C.c_char_p]
    L.ref_s3_gs_closest.argtypes = [vp, f32p, C.c_int, i32p]
    assert L.ref_s3_open_gs(r3.h, path.encode()) == 0


# ------------------------------------------------------- sphinx3 hmm_vit_eval
port.orc_s3hmm_eval_batch.restype = C.c_int32
port.orc_s3hmm_eval_batch.argtypes = [C.c_int, C.c_int, i32p, C.c_int, i16p, C.c_int, i32p, i32p, i32p, i32p, i32p, i32p,
                                      i32p, u8p, i32p, C.c_int]


def s3hmm_eval(which, n_emit, tp, sseq, senscr, score, history, out_score, out_history, ssid, tmatid, mpx, bestscore,
               repeat=1):
    """which = "port" (oracle/sphinx_oracle.c) or "ref" (sphinx3's own libam/hmm.c via oracle/_ref).
    HMM-major int32 arrays [n_hmm][n_emit], updated in place; returns the best score of the last pass."""
    if which == "ref":
        fn = ref_s3().ref_s3hmm_eval_batch
        fn.restype = C.c_int32
        fn.argtypes = port.orc_s3hmm_eval_batch.argtypes
    else:
        fn = port.orc_s3hmm_eval_batch
    tp, sseq, senscr = _c(tp, np.int32), _c(sseq, np.int16), _c(senscr, np.int32)
    return fn(n_emit, out_score.shape[0], _p(tp, C.c_int32), tp.shape[0], _p(sseq, C.c_int16), sseq.shape[0],
              _p(senscr, C.c_int32), _p(score, C.c_int32), _p(history, C.c_int32), _p(out_score, C.c_int32),
              _p(out_history, C.c_int32), _p(ssid, C.c_int32), _p(_c(tmatid, np.int32), C.c_int32),
              _p(_c(mpx, np.uint8), C.c_uint8), _p(bestscore, C.c_int32), repeat)


def have_ref_s3():
    return os.path.exists(os.path.join(REF_DIR, "libref_shim_s3.so"))


# ------------------------------------------------------- feature stage
port.orc_feat_1s_c_d_dd.argtypes = [f32p, C.c_int, C.c_int, C.c_int, f32p]


def port_feat_1s_c_d_dd(cep, cmn=True):
    cep = _c(cep, np.float32)
    out = np.zeros((cep.shape[0], 3 * cep.shape[1]), np.float32)
    port.orc_feat_1s_c_d_dd(_p(cep, C.c_float), cep.shape[0], cep.shape[1], 1 if cmn else 0, _p(out, C.c_float))
    return out


FEAT_TYPES = ["1s_c_d_dd", "s3_1x39", "s2_4x", "1s_c_d_ld_dd", "1s_c", "1s_c_d"]
port.orc_feat_set_copy.argtypes = [C.c_int, C.c_int, i32p]
port.orc_feat_compute.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, C.c_int, i32p, C.c_int,
                                  f32p, C.c_int, f32p]


def parse_svspec(spec):
    """'0-12/13-25/26-38' or '0,2,4/1,3' -> flat index list (sphinxbase parse_subvecs; the
    stream boundaries do not change the memory layout)."""
    idx = []
    for part in spec.split("/"):
        for piece in part.split(","):
            if "-" in piece:
                a, b = piece.split("-")
                idx.extend(range(int(a), int(b) + 1))
            else:
                idx.append(int(piece))
    return idx


def port_feat_compute(cep, ftype="1s_c_d_dd", cmn=True, varnorm=False, agc=False, lda=None, lda_dim=0, svspec=None):
    cep = _c(cep, np.float32)
    T, cs = cep.shape
    sv = np.array(parse_svspec(svspec) if svspec else [], np.int32)
    if lda is not None:
        lda = _c(lda, np.float32)
        if lda_dim <= 0 or lda_dim > lda.shape[0]:
            lda_dim = lda.shape[0]
        lda = np.ascontiguousarray(lda[:lda_dim])
    from cmusphinx_b200.engine import parse_feat_type
    tid, cw, clen = parse_feat_type(ftype)
    cl = np.array(clen + [0] * 8, np.int32)
    port.orc_feat_set_copy(cw, len(clen), _p(cl, C.c_int32))
    out = np.zeros((T, max(4 * max(cs, 13), 15 * cs)), np.float32)
    n = port.orc_feat_compute(tid, cs, int(cmn), int(varnorm), int(agc),
                              _p(lda, C.c_float) if lda is not None else None, lda_dim,
                              _p(sv, C.c_int32) if sv.size else None, sv.size, _p(cep, C.c_float), T,
                              _p(out, C.c_float))
    if n < 0:
        raise ValueError("bad feature configuration")
    flat = out.reshape(-1)[:T * n]
    return flat.reshape(T, n).copy()


def ref_feat_compute(cep, ftype="1s_c_d_dd", cmn=True, varnorm=False, agc=False, lda=None, lda_dim=0, svspec=None):
    """The reference's feat_t on one utterance -> [T][out_dim] (the valid prefix of its rows)."""
    cep = _c(cep, np.float32)
    T, cs = cep.shape
    out = np.zeros((T + 16, max(4 * max(cs, 13), 15 * cs)), np.float32)
    row = (C.c_int32 * 1)()
    od = (C.c_int32 * 1)()
    if lda is not None:
        lda = _c(lda, np.float32)
    n = ref().ref_feat_compute(ftype.encode(), b"current" if cmn else b"none", int(varnorm),
                               b"max" if agc else b"none", cs,
                               _p(lda, C.c_float) if lda is not None else None,
                               lda.shape[0] if lda is not None else 0, lda.shape[1] if lda is not None else 0, lda_dim,
                               svspec.encode() if svspec else None, _p(cep, C.c_float), T, _p(out, C.c_float), row, od)
    if n < 0:
        raise ValueError("reference rejected the feature configuration")
    rows = out.reshape(-1)[:n * row[0]].reshape(n, row[0])
    return rows[:, :od[0]].copy()


# ---- prune / phone-transition stage (ngram_search_fwdtree.c:714-869)
PRUNE_PAR = ("frame", "best_score", "beam", "pbeam", "lpbeam", "pip", "nwpen", "has_pls")


def prune_rows_to_soa(rows):
    """Trace / golden rows [n_chan][10] (score 0..2, history 0..2, out_score, out_history, bestscore, frame)
    -> state-major arrays."""
    rows = np.asarray(rows, np.int32)
    return dict(score=_c(rows[:, 0:3].T, np.int32), history=_c(rows[:, 3:6].T, np.int32), out_score=_c(rows[:, 6], np.int32),
                out_history=_c(rows[:, 7], np.int32), bestscore=_c(rows[:, 8], np.int32), frame=_c(rows[:, 9], np.int32))


def prune_soa_to_rows(s):
    return np.concatenate([s["score"].T, s["history"].T, s["out_score"][:, None], s["out_history"][:, None],
                           s["bestscore"][:, None], s["frame"][:, None]], axis=1).astype(np.int32)


def port_fwdtree_prune(topo, par, pls_pen, acl, soa):
    """oracle/sphinx_oracle.c orc_fwdtree_prune on copies; -> (soa_after, nacl, cand[n][3])."""
    L = _load_port()
    s = {k: _c(v, np.int32).copy() for k, v in soa.items()}
    n_chan, ne = int(topo["n_chan"]), s["score"].shape[0]
    t = {k: _c(topo[k], np.int32) for k in ("child_off", "child", "ciphone", "pw_off", "pw_wid", "pw_lastphone")}
    parv = _c([int(par[k]) for k in PRUNE_PAR], np.int32)
    pen = _c(pls_pen, np.int32)
    acl = _c(acl, np.int32)
    nacl = np.zeros(max(1, n_chan), np.int32)
    cand = np.zeros((max(1, len(t["pw_wid"])), 3), np.int32)
    n1, n2 = C.c_int32(0), C.c_int32(0)
    L.orc_fwdtree_prune.restype = None
    L.orc_fwdtree_prune(C.c_int(int(topo["n_root"])), C.c_int(n_chan), C.c_int(ne), _p(t["child_off"], C.c_int32),
                        _p(t["child"], C.c_int32), _p(t["ciphone"], C.c_int32), _p(t["pw_off"], C.c_int32),
                        _p(t["pw_wid"], C.c_int32), _p(t["pw_lastphone"], C.c_int32), _p(parv, C.c_int32), _p(pen, C.c_int32),
                        _p(acl, C.c_int32), C.c_int(len(acl)), _p(s["score"], C.c_int32), _p(s["history"], C.c_int32),
                        _p(s["out_score"], C.c_int32), _p(s["out_history"], C.c_int32), _p(s["bestscore"], C.c_int32),
                        _p(s["frame"], C.c_int32), _p(nacl, C.c_int32), C.byref(n1), _p(cand, C.c_int32), C.byref(n2))
    return s, nacl[:n1.value].copy(), cand[:n2.value].copy()


# ---- phone-loop look-ahead search (phone_loop_search.c:253-291)
def port_phone_loop_step(tp, rec):
    """One frame through oracle/sphinx_oracle.c orc_phone_loop_step.  rec: a frame dict of
    fwdtree_trace.read_pls_trace (state before the step, senscr [n_phones][ne] = the frame's score of every
    state's senone).  -> dict of the state after the step (+ best_score, renorm)."""
    L = _load_port()
    n, ne = rec["score"].shape
    tp = _c(tp, np.uint8)
    sc, hi = _c(rec["score"], np.int32).copy(), _c(rec["history"], np.int32).copy()
    os_, oh, bs, fr = (_c(rec[k], np.int32).copy() for k in ("out_score", "out_history", "bestscore", "frame_of"))
    senscr = _c(rec["senscr"], np.int16).reshape(-1)                 # compact: senone of (phone i, state s) = i * ne + s
    senid = np.arange(n * ne, dtype=np.uint16)
    tm = _c(rec["tmatid"], np.int16)
    par = np.array([rec["frame"], rec["best_score"], rec["beam"], rec["pbeam"], rec["pip"]], np.int32)
    rn = C.c_int32(0)
    L.orc_phone_loop_step.restype = C.c_int32
    best = L.orc_phone_loop_step(C.c_int(n), C.c_int(ne), _p(tp, C.c_uint8), _p(senscr, C.c_int16), _p(par, C.c_int32), _p(sc, C.c_int32),
                                 _p(hi, C.c_int32), _p(os_, C.c_int32), _p(oh, C.c_int32), _p(bs, C.c_int32), _p(fr, C.c_int32),
                                 _p(senid, C.c_uint16), _p(tm, C.c_int16), C.byref(rn))
    return dict(score=sc, history=hi, out_score=os_, out_history=oh, bestscore=bs, frame_of=fr, best_score=int(best), renorm=rn.value)
