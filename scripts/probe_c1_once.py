"""Three C1 products (1000 q, 500 x 500 terms, A*A) for ncu captures of the small-product kernels."""
import os, sys
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import pauli_oracle as po
from symmer_b200 import PauliwordOp, ops
ops.device()
s, c = po.random_operator(1000, 500, seed=1)
A = PauliwordOp(s, c); ac = A._coeff_dev(); axz = A._xz
for _ in range(3):
    out = ops.mul_cleanup(axz, ac, axz, ac)
torch.cuda.synchronize()
print(out[0].shape)
