import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from symmer_b200 import PauliwordOp, ops
ops.device()
np.random.seed(2)
P = PauliwordOp.random(1000, 100000)
Q = PauliwordOp.random(1000, 1)
Q.coeff_vec[0] = 1
def t(fn, n=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("rotation, no host view: %.3f ms" % t(lambda: P.perform_rotations([(Q, 0.731)])))
_ = P.coeff_vec[:10]
print("rotation, coeff host view handed out: %.3f ms" % t(lambda: P.perform_rotations([(Q, 0.731)])))
print("  _coeff_dev alone: %.3f ms" % t(lambda: P._coeff_dev()))
_ = P.symp_matrix[:10]
print("rotation, symp host view too: %.3f ms" % t(lambda: P.perform_rotations([(Q, 0.731)])))
print("100 different angles: %.3f ms each" % (t(lambda: [P.perform_rotations([(Q, 0.1 + 0.013 * k)]) for k in range(100)], n=1) / 100))
