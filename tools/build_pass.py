"""Index build of BASELINE configs[1] through the host API and the device API, cold and warm, with DPF_TRACE phases."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DPF_TRACE"] = "1"
import numpy as np
import torch
from similaritysearchbyrdf_b200 import DPFIndex, synth

n, d = 1_000_000, 128
X, Q = synth.config2(n, 16, d)
A, chain = synth.angle_family(d, max(100, d), 10, 3, 32, 88387 + 2)
Ap = synth.partitioner_family(30, 3, 88387 + 3)
Xh = torch.from_numpy(X).pin_memory()
Xd = torch.from_numpy(X).cuda()


def mk():
    ix = DPFIndex(d=d, L=30, k=32, pb=3)
    ix.set_family(A, chain)
    ix.set_partitioners(Ap)
    ix.set_profiling(True)
    return ix


for rep in range(3):
    ix = mk()
    torch.cuda.synchronize()
    t0 = time.time()
    ix.fit_dense(Xh.numpy())
    print(f"host-API fit #{rep}: {1e3 * (time.time() - t0):.1f} ms", {k: round(v, 3) for k, v in ix.stage_times_ms().items() if v}, flush=True)
    ix.close()
for rep in range(3):
    ix = mk()
    torch.cuda.synchronize()
    t0 = time.time()
    ix.fit_dense_dev(Xd.data_ptr(), n)
    print(f"device-API fit #{rep}: {1e3 * (time.time() - t0):.1f} ms", {k: round(v, 3) for k, v in ix.stage_times_ms().items() if v}, flush=True)
    ix.close()
