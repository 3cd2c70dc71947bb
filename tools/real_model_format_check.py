import os, sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import orc, cmusphinx_b200 as b
d = os.path.join(orc.DATA_DIR, "hmm", "cont")
m = b.ms_from_files(*(os.path.join(d, n) for n in ("means", "variances", "mixture_weights")), ".cont.", topn=4)
r = orc.RefAcmod(d); cep = orc.read_mfc(os.path.join(orc.DATA_DIR, "test", "pittsburgh.littleendian.mfc")); feat = r.cep2feat(cep); r.close()
got = m.score(feat)
fmt = m.tc_last_format()
import ctypes as C
m.set_path(0); want = m.score(feat)
dd = np.abs(got.astype(np.int32) - want)
print("path1 format", fmt, "frames", feat.shape[0], "max|d|", dd.max(), "mismatch", float((dd != 0).mean()))
