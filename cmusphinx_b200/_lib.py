"""ctypes binding of libb200sphinx.so (the drop-in boundary, include/b200sphinx.h)."""
import ctypes as C
import os

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb200sphinx.so")


class B200Error(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C cmusphinx_b200/csrc`. There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

c_i32p = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)
c_i16p = C.POINTER(C.c_int16)
c_u16p = C.POINTER(C.c_uint16)
c_u8p = C.POINTER(C.c_uint8)
c_f32p = C.POINTER(C.c_float)


class MgauCfg(C.Structure):
    _fields_ = [("n_mgau", C.c_int32), ("n_feat", C.c_int32), ("n_density", C.c_int32),
                ("n_sen", C.c_int32), ("featlen", C.c_int32 * 4), ("topn", C.c_int32),
                ("aw", C.c_int32), ("ds_ratio", C.c_int32), ("logbase", C.c_double),
                ("device", C.c_int32), ("topn_beam", C.c_int32 * 4)]


class HmmSoa(C.Structure):
    _fields_ = [("n_hmm", C.c_int32), ("score", c_i32p), ("history", c_i32p),
                ("out_score", c_i32p), ("out_history", c_i32p), ("senid", c_u16p),
                ("tmatid", c_i16p), ("mpx", c_u8p), ("bestscore", c_i32p)]


class S3HmmSoa(C.Structure):
    _fields_ = [("n_hmm", C.c_int32), ("score", c_i32p), ("history", c_i32p), ("ssid", c_i32p),
                ("out_score", c_i32p), ("out_history", c_i32p), ("bestscore", c_i32p),
                ("tmatid", c_i32p), ("mpx", c_u8p)]


def _sig(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


vp = C.c_void_p
_sig("b200_last_error", C.c_char_p)
_sig("b200_abi_version", C.c_int)
_sig("b200_device_count", C.c_int)
_sig("b200_launch_count", C.c_longlong)
_sig("b200_logadd_table", C.c_int, C.c_double, C.c_int, c_u32p, C.c_int)
_sig("b200_logmath_log", C.c_int32, C.c_double, C.c_int, C.c_double)
_sig("b200_logmath_add", C.c_int32, C.c_double, C.c_int, C.c_int32, C.c_int32)
_sig("b200_gauden_precompute", C.c_int, c_f32p, c_f32p, C.c_long, C.c_int, C.c_float, C.c_double)
_sig("b200_mixw_quantize_ms", C.c_int, c_f32p, c_u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_double)
_sig("b200_mixw_quantize_tied", C.c_int, c_f32p, c_u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_double)
_sig("b200_tmat_quantize", C.c_int, c_f32p, c_u8p, C.c_int, C.c_int, C.c_double, C.c_double)
_sig("b200_s3_read_gauden", C.c_int, C.c_char_p, c_i32p, c_i32p, c_f32p)
_sig("b200_s3_read_mixw", C.c_int, C.c_char_p, c_i32p, c_f32p)
_sig("b200_s3_read_tmat", C.c_int, C.c_char_p, c_i32p, c_f32p)
_sig("b200_s3_read_sendump", C.c_int, C.c_char_p, c_i32p, c_u8p, c_u8p)
_sig("b200_mdef_read_maps", C.c_int, C.c_char_p, c_i32p, c_i16p, c_i16p)
_sig("b200_sen_write", C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_double, c_i16p, C.c_int, vp, c_i32p)
_sig("b200_sen_read", C.c_int, C.c_char_p, c_i32p, C.POINTER(C.c_double), c_i16p, c_i32p)
_sig("b200_ms_create", vp, C.POINTER(MgauCfg), c_f32p, c_f32p, c_f32p, c_u8p, c_u32p)
_sig("b200_ms_load", vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, c_u8p, C.c_double,
     C.c_double, C.c_int, C.c_int, C.c_double, C.c_int)
_sig("b200_ptm_create", vp, C.POINTER(MgauCfg), c_f32p, c_f32p, c_f32p, c_u8p, C.c_int, c_u8p, c_u8p)
_sig("b200_semi_create", vp, C.POINTER(MgauCfg), c_f32p, c_f32p, c_f32p, c_u8p, C.c_int, c_u8p)
_sig("b200_mgau_free", None, vp)
_sig("b200_mgau_name", C.c_char_p, vp)
_sig("b200_mgau_n_sen", C.c_int, vp)
_sig("b200_mgau_featdim", C.c_int, vp)
_sig("b200_mgau_update_params", C.c_int, vp, c_f32p, c_f32p, c_f32p)
_sig("b200_mgau_set_path", C.c_int, vp, C.c_int)
_sig("b200_mgau_get_path", C.c_int, vp)
_sig("b200_mgau_tc_last_format", C.c_int, vp)
_sig("b200_mgau_tied_stats", C.c_int, vp, C.POINTER(C.c_longlong))
_sig("b200_mgau_cont_stats", C.c_int, vp, C.POINTER(C.c_longlong))
_sig("b200_mgau_score_host", C.c_int, vp, vp, C.c_int, vp)
_sig("b200_mgau_score_dev", C.c_int, vp, vp, C.c_int, vp, vp)
_sig("b200_mgau_frame_eval", C.c_int, vp, c_i16p, c_u8p, C.c_int32, C.POINTER(c_f32p), C.c_int32, C.c_int32)
_sig("b200_mgau_utt_begin", C.c_int, vp, vp, C.c_int)
_sig("b200_mgau_utt_begin_at", C.c_int, vp, vp, C.c_int, C.c_int)
_sig("b200_mgau_utt_frame", C.c_int, vp, c_i16p, c_u8p, C.c_int32, C.c_int32, C.c_int32)
_sig("b200_mgau_last_ms", C.c_float, vp, C.c_int)
_sig("b200_mgau_timing_avg", C.c_float, vp, C.c_int, C.c_int)
_sig("b200_hmm_ctx_create", vp, C.c_int, c_u8p, C.c_int, c_u16p, C.c_int, C.c_int, C.c_int)
_sig("b200_hmm_ctx_free", None, vp)
_sig("b200_hmm_eval_host", C.c_int, vp, C.POINTER(HmmSoa), c_i16p, C.c_int, c_i32p)
_sig("b200_s3hmm_eval_host", C.c_int, C.c_int, c_i32p, C.c_int, c_i16p, C.c_int, C.c_int, C.POINTER(S3HmmSoa), c_i32p,
     C.c_int, c_i32p, C.c_int)
_sig("b200_hmm_pop_upload", C.c_int, vp, C.POINTER(HmmSoa))
_sig("b200_hmm_pop_download", C.c_int, vp, C.POINTER(HmmSoa))
_sig("b200_hmm_pop_set_utts", C.c_int, vp, C.c_int, c_i32p)
_sig("b200_hmm_step_dev", C.c_int, vp, vp, C.c_int32, vp)
_sig("b200_hmm_run_dev", C.c_int, vp, vp, C.c_long, C.c_int, C.c_int, C.c_int32, vp)
_sig("b200_hmm_step_results", C.c_int, vp, c_i32p, c_i32p, c_i32p, c_u32p)
_sig("b200_hmm_step_host", C.c_int, vp, c_i16p, C.c_int32)
_sig("b200_hmm_last_ms", C.c_float, vp)
_sig("b200_hmm_normalize_dev", C.c_int, vp, vp, vp)
_sig("b200_hmm_clear_pruned_dev", C.c_int, vp, vp)
_sig("b200_hmm_enter_dev", C.c_int, vp, vp, vp, vp, C.c_int, vp)
_sig("b200_hmm_enter_host", C.c_int, vp, c_i32p, c_i32p, c_i32p, C.c_int)
class PruneDev(C.Structure):
    _fields_ = [("score", vp), ("history", vp), ("out_score", vp), ("out_history", vp), ("bestscore", vp), ("frame", vp),
                ("state_stride", C.c_long), ("par", vp), ("pls_pen", vp), ("acl", vp), ("n_act", vp), ("list_cap", C.c_int32),
                ("nacl", vp), ("n_nacl", vp), ("cand", vp), ("n_cand", vp), ("cand_cap", C.c_int32)]


_sig("b200_hmm_pop_device", C.c_int, vp, C.POINTER(HmmSoa), C.POINTER(vp))
_sig("b200_hmm_eval_list_dev", C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp)
_sig("b200_phone_loop_create", vp, C.c_int, C.c_int, c_u8p, C.c_int, c_u16p, c_i16p, C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int, C.c_int)
_sig("b200_phone_loop_free", None, vp)
_sig("b200_phone_loop_start", C.c_int, vp)
_sig("b200_phone_loop_set_state", C.c_int, vp, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p)
_sig("b200_phone_loop_get_state", C.c_int, vp, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p)
_sig("b200_phone_loop_step_dev", C.c_int, vp, vp, C.c_int, vp, vp)
_sig("b200_phone_loop_step_host", C.c_int, vp, c_i16p, C.c_int, c_i32p, c_i32p)
_sig("b200_chantree_create", vp, C.c_int, C.c_int, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, C.c_int, C.c_int, C.c_int)
_sig("b200_chantree_free", None, vp)
_sig("b200_chantree_cand_cap", C.c_int, vp)
_sig("b200_fwdtree_prune_dev", C.c_int, vp, C.c_int, C.POINTER(PruneDev), vp)
_sig("b200_fwdtree_renorm_dev", C.c_int, vp, C.c_int, C.POINTER(PruneDev), vp, vp)
_sig("b200_fwdtree_deactivate_dev", C.c_int, vp, C.c_int, C.POINTER(PruneDev), vp)
_sig("b200_fwdtree_prune_host", C.c_int, vp, C.c_int, c_i32p, c_i32p, c_i32p, c_i32p, C.c_int, c_i32p, c_i32p, c_i32p, c_i32p,
     c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, C.c_int)
_sig("b200_s3_create", vp, C.c_int, C.c_int, C.c_int, c_f32p, c_f32p, c_f32p, C.c_double, C.c_double, C.c_double,
     c_i32p, C.c_int, C.c_int)
_sig("b200_s3_load", vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_double, C.c_double, C.c_double, c_i32p, C.c_int,
     C.c_int)
_sig("b200_s3_free", None, vp)
_sig("b200_s3_dims", C.c_int, vp, c_i32p)
_sig("b200_s3_set_fast", C.c_int, vp, C.c_double, C.c_int, C.c_int, C.c_float)
_sig("b200_s3_set_subvq", C.c_int, vp, C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_double)
_sig("b200_s3_set_gs", C.c_int, vp, C.c_char_p)
_sig("b200_s3_set_fast_log", C.c_int, vp, C.c_int32, C.c_int, C.c_int, C.c_float, C.c_int32)
_sig("b200_s3_utt_reset", C.c_int, vp)
_sig("b200_s3_params", C.c_int, vp, c_i32p, c_f32p, c_f32p, c_f32p, c_i32p, C.POINTER(C.c_double))
_sig("b200_s3_state", C.c_int, vp, c_i32p, c_i32p)
_sig("b200_s3_dense_host", C.c_int, vp, c_f32p, C.c_int, c_i32p)
_sig("b200_s3_dense_dev", C.c_int, vp, vp, C.c_int, vp, vp)
_sig("b200_s3_score_utt_host", C.c_int, vp, c_f32p, C.c_int, C.c_int, c_u8p, c_i32p, c_i32p, c_i32p)
_sig("b200_s3_score_utt_dev", C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp)
_sig("b200_s3_frame_eval", C.c_int, vp, c_f32p, C.c_int32, c_u8p, c_i32p, c_i32p)
_sig("b200_s3_last_ms", C.c_float, vp)
_sig("b200_fp64_issue_rate", C.c_double, C.c_int)
_sig("b200_feat_1s_c_d_dd_dev", C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp)
_sig("b200_feat_1s_c_d_dd_host", C.c_int, c_f32p, c_i32p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_int)
_sig("b200_flags2list", C.c_int, c_u32p, C.c_int, c_u8p, C.c_int)
_sig("b200_dev_alloc", vp, C.c_size_t, C.c_int)
_sig("b200_dev_free", None, vp)
_sig("b200_dev_upload", C.c_int, vp, vp, C.c_size_t)
_sig("b200_dev_download", C.c_int, vp, vp, C.c_size_t)
_sig("b200_host_alloc_pinned", vp, C.c_size_t)
_sig("b200_host_free_pinned", None, vp)
_sig("b200_dev_sync", C.c_int, C.c_int)


def last_error():
    return lib.b200_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc is None or (isinstance(rc, int) and rc < 0):
        raise B200Error(f"{what}: {last_error()} (rc={rc})")
    return rc
