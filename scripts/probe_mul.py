"""Quick timing probe of the fused multiply+cleanup at several sizes (not the bench)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import symmer_b200.ops as ops
from oracle import pauli_oracle as po

def run(n, M, N, reps=3, square=False):
    a_s, a_c = po.random_operator(n, M, seed=1)
    if square:
        b_s, b_c = a_s, a_c
    else:
        b_s, b_c = po.random_operator(n, N, seed=2)
    a = ops.pack(torch.from_numpy(a_s), n); ac = torch.from_numpy(a_c).cuda()
    b = ops.pack(torch.from_numpy(b_s), n); bc = torch.from_numpy(b_c).cuda()
    T = a.shape[0] * b.shape[0]
    for _ in range(2):
        oxz, oc = ops.mul_cleanup(a, ac, b, bc)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); oxz, oc = ops.mul_cleanup(a, ac, b, bc); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    U = oxz.shape[0]
    ms = min(ts)
    R = 16 * a.shape[1] // 2 + 16
    alg = T * (2 * R + (U / T) * R)
    print(f"n={n} M={M} N={b.shape[0]} T={T:.3e} U={U:.3e} ms={ms:.3f} ct/s={T/ms*1e3:.3e} alg_GB/s={alg/ms/1e6:.1f} "
          f"mem={torch.cuda.max_memory_allocated()/1e9:.1f}GB", flush=True)
    del oxz, oc

if __name__ == "__main__":
    ops.device()
    run(1000, 500, 500, square=True)
    run(1000, 2000, 2000)
    run(1000, 12500, 1000)
    run(1000, 12500, 10000, reps=2)
