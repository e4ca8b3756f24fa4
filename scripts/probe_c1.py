"""Config C1 (square of random(1000 q, 500 terms)) through the API, for an ncu launch list."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from symmer_b200 import PauliwordOp, ops
ops.device()
np.random.seed(1)
P = PauliwordOp.random(1000, 500)
for _ in range(3):
    S = P * P
torch.cuda.synchronize()
print(S.n_terms)
