"""Sequential rotations: per-step device time of rotate and cleanup, and wall time (allocator stalls show up there)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from symmer_b200 import PauliwordOp, ops
import math

ops.device()
np.random.seed(4)
P = PauliwordOp.random(1000, 100000)
gens = []
for k in range(14):
    G = PauliwordOp.random(1000, 1); G.coeff_vec[0] = 1
    gens.append((G, 0.1 + 1.3 * np.random.rand()))
for rep in range(2):
    xz, c = P._xz, P._coeff_dev()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    tot_r = tot_c = 0.0
    for G, th in gens:
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t0 = time.perf_counter()
        e[0].record()
        rxz, rc = ops.rotate(xz, c, G._xz, math.cos(th), math.sin(th), 0)
        e[1].record()
        nxz, nc = ops.cleanup(rxz, rc)
        e[2].record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        tr, tc = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        tot_r += tr; tot_c += tc
        print(f"rep {rep} rows {xz.shape[0]:>9} -> {rxz.shape[0]:>9} -> {nxz.shape[0]:>9}: rotate {tr:7.3f} ms  cleanup {tc:7.3f} ms  wall {wall:7.3f} ms", flush=True)
        xz, c = nxz, nc
        del rxz, rc
    print(f"rep {rep}: rotate {tot_r:.2f} ms, cleanup {tot_c:.2f} ms, wall {(time.perf_counter() - w0) * 1e3:.2f} ms")
