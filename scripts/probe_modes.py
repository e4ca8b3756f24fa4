"""A/B of the product modes by size: scatter-by-t (small products) vs ordered tiles vs sorted-hash order."""
import sys
import torch
sys.path.insert(0, ".")
import symmer_b200.ops as ops
from oracle import pauli_oracle as po


def time_mul(a, ac, b, bc, reps=5):
    for _ in range(2):
        ops.mul_cleanup(a, ac, b, bc)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); oxz, oc = ops.mul_cleanup(a, ac, b, bc); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), oxz.shape[0]


ops.device()
for n, M, N, square in [(1000, 500, 500, True), (1000, 1000, 1000, True), (1000, 1024, 1024, False), (1000, 2048, 2048, False),
                        (1000, 4000, 1000, False), (64, 2048, 2048, False), (24, 3000, 3000, True)]:
    a_s, a_c = po.random_operator(n, M, seed=1)
    b_s, b_c = (a_s, a_c) if square else po.random_operator(n, N, seed=2)
    a = ops.pack(torch.from_numpy(a_s), n); ac = torch.from_numpy(a_c).cuda()
    b = ops.pack(torch.from_numpy(b_s), n); bc = torch.from_numpy(b_c).cuda()
    out = []
    for name, limit, knob6 in [("by_t", 1 << 40, 1), ("tiles", 0, 1), ("sorted", 0, 0)]:
        ops.set_tuning(0, limit); ops.set_tuning(6, knob6)
        ms, U = time_mul(a, ac, b, bc)
        out.append(f"{name} {ms:.3f} ms")
    print(f"n={n} {M}x{b.shape[0]}{' square' if square else ''} T={M * b.shape[0]:.2e} U={U}: " + ", ".join(out), flush=True)
ops.set_tuning(0, 1 << 22); ops.set_tuning(6, 1)
