"""cmusphinx_b200 -- B200-native acoustic scoring / HMM evaluation for the
pocketsphinx decode path (acmod_score -> ps_mgau_frame_eval, hmm_vit_eval).

The product is the C-ABI shared library ``libb200sphinx.so`` (CUDA, sm_100a
only; declared in ``include/b200sphinx.h``).  This package is the thin host-side
mirror of the reference's back-end interface on top of it; it never computes
scores itself and has no CPU fallback: importing works without a GPU (so the
ABI can be inspected), constructing a scorer does not.
"""
from ._lib import lib, LIB_PATH, B200Error, last_error  # noqa: F401
from .engine import (  # noqa: F401
    MgauConfig, Mgau, ms_from_files, ms_from_arrays, ptm_from_arrays,
    semi_from_arrays, tied_from_model_dir, HmmContext, HmmPopulation,
    logadd_table, gauden_precompute, mixw_quantize_ms, mixw_quantize_tied,
    tmat_quantize, flags2list, read_gauden, read_mixw, read_tmat, read_sendump,
    device_count, launch_count, S3Mgau, s3hmm_vit_eval, read_s3_cont_arrays, feat_1s_c_d_dd, feat_compute, FEAT_TYPES, sen_write, sen_read, mdef_maps, ChanTree, FwdtreeDevice, PhoneLoop,
)
from . import s3io, synth, shard  # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
