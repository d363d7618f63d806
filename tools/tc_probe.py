"""First-light check of the tcgen05 scoring kernel on a small byte-valued index: results against the mma.sync ring
kernel and the row-major kernel, plus the kernel's watchdog record."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from similaritysearchbyrdf_b200 import synth, _lib as B
from tests import util as U

n, nq, d = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000, 512, 128
X, Q = synth.config2(n=n, nq=nq, d=d)
A, chain = synth.angle_family(d, 128, 10, 3, 32, 88389)
Ap = synth.partitioner_family(30, 3, 88390)
ix = U.make_index(d, A, chain, Ap, bucket_overflow=100)
ix.fit_dense(X)
ref = ix.query_topk_dense(Q, None, 0, 10, B.METRIC_DOT)
print("ring ok", ix.stats()["bm_survivors"], flush=True)
with ix.debug_options(u8i_kernel=3):
    got = ix.query_topk_dense(Q, None, 0, 10, B.METRIC_DOT)
print("tc diag", ix.tc_diag()[:8].tolist(), "stats", {k: v for k, v in ix.stats().items() if k.startswith("bm_")}, flush=True)
print("ids equal", float((ref[0] == got[0]).mean()), "scores equal", float((ref[1] == got[1]).mean()))
bad = np.nonzero((ref[0] != got[0]).any(axis=1))[0]
print("bad queries", len(bad), bad[:10].tolist())
if len(bad):
    q = bad[0]
    print(ref[0][q], got[0][q]); print(ref[1][q], got[1][q])
