"""Writers for the reference's S3 binary parameter files (test/bench fixtures).

Formats: SURVEY.md Appendix A -- sphinxbase bio.c:137-262 header, then
  means/variances  : int32 n_mgau,n_feat,n_density, int32 veclen[n_feat], int32 n, float32 data
  mixture_weights  : int32 n_sen,n_feat,n_cw,n, float32 [sen][feat][cw]
  transition_matrices: int32 n_tmat,n_src,n_dst,n, float32 [tmat][src][dst]
No `chksum0` header line is written, so readers skip checksum verification
(PS/ms_gauden.c:208-217).
"""
import struct

import numpy as np


def _hdr(fp, version="1.0"):
    fp.write(b"s3\n")
    fp.write(f"version {version}\n".encode())
    fp.write(b"endhdr\n")
    fp.write(struct.pack("<I", 0x11223344))


def write_gauden(path, arr, veclen):
    """arr: [n_mgau][n_feat-concatenated...] given as list per stream or a
    float32 array [n_mgau, n_density, sum(veclen)] for 1 stream / per-stream list."""
    if isinstance(arr, (list, tuple)):
        streams = [np.ascontiguousarray(a, np.float32) for a in arr]   # each [n_mgau][n_density][len]
    else:
        a = np.ascontiguousarray(arr, np.float32)
        assert len(veclen) == 1
        streams = [a]
    n_mgau, n_density = streams[0].shape[:2]
    n = sum(s.size for s in streams)
    with open(path, "wb") as fp:
        _hdr(fp)
        fp.write(struct.pack("<3i", n_mgau, len(streams), n_density))
        fp.write(struct.pack(f"<{len(streams)}i", *[int(v) for v in veclen]))
        fp.write(struct.pack("<i", n))
        for m in range(n_mgau):
            for s in streams:
                fp.write(s[m].astype("<f4").tobytes())


def write_mixw(path, mixw):
    mixw = np.ascontiguousarray(mixw, np.float32)
    n_sen, n_feat, n_cw = mixw.shape
    with open(path, "wb") as fp:
        _hdr(fp)
        fp.write(struct.pack("<4i", n_sen, n_feat, n_cw, mixw.size))
        fp.write(mixw.astype("<f4").tobytes())


def write_tmat(path, tp):
    tp = np.ascontiguousarray(tp, np.float32)
    n_tmat, n_src, n_dst = tp.shape
    with open(path, "wb") as fp:
        _hdr(fp)
        fp.write(struct.pack("<4i", n_tmat, n_src, n_dst, tp.size))
        fp.write(tp.astype("<f4").tobytes())
