import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import symmer_b200.ops as ops
from symmer_b200 import PauliwordOp
from oracle import pauli_oracle as po
ops.device()
n, M = 1000, 10_000_000
g = torch.Generator(device="cuda"); g.manual_seed(1)
xz = torch.randint(-2**63, 2**63 - 1, (M, 32), dtype=torch.int64, device="cuda", generator=g)
xz[:, 15] &= (1 << 40) - 1; xz[:, 31] &= (1 << 40) - 1
c = torch.randn(M, dtype=torch.complex128, device="cuda")
P = PauliwordOp._from_device(xz, c, n)
q_s, _ = po.random_operator(n, 1, seed=5)
Q = PauliwordOp(q_s, [1])
out = P.perform_rotations([(Q, 0.731)])
torch.cuda.synchronize()
