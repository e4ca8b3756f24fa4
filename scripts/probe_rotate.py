import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import symmer_b200.ops as ops
from symmer_b200 import PauliwordOp
from oracle import pauli_oracle as po
ops.device()
n = 1000
def t_ms(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); o = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)); del o
    return min(ts)
for M in [100_000, 1_000_000, 10_000_000]:
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    xz = torch.randint(-2**63, 2**63 - 1, (M, 32), dtype=torch.int64, device="cuda", generator=g)
    xz[:, 15] &= (1 << 40) - 1; xz[:, 31] &= (1 << 40) - 1      # 1000 qubits: top 24 bits of the last word are padding
    c = torch.randn(M, dtype=torch.complex128, device="cuda")
    P = PauliwordOp._from_device(xz, c, n)
    q_s, _ = po.random_operator(n, 1, seed=5)
    Q = PauliwordOp(q_s, [1])
    R = 272
    ms_rot = t_ms(lambda: ops.rotate(P._xz, P._c, Q._xz, 0.7, 0.6, 0))
    ms_all = t_ms(lambda: P.perform_rotations([(Q, 0.731)]))
    ms_cl = t_ms(lambda: P.perform_rotations([(Q, np.pi / 2)]))
    ms_clean = t_ms(lambda: ops.cleanup(P._xz, P._c))
    out = P.perform_rotations([(Q, 0.731)])
    f = (out.n_terms - M) / M
    print(f"M={M:.0e}: rotate-only {ms_rot:.3f} ms ({M*(R+(1+f)*R)/ms_rot/1e6:.0f} GB/s r+w) | general rotation incl dedup {ms_all:.3f} ms "
          f"({M/ms_all*1e3:.3e} rows/s, model 5.5R -> {M*5.5*R/ms_all/1e6:.0f} GB/s) | clifford {ms_cl:.3f} ms | cleanup of M unique rows {ms_clean:.3f} ms "
          f"({M*2*R/ms_clean/1e6:.0f} GB/s r+w) rows_out={out.n_terms}", flush=True)
    del P, xz, c, out
