import sys
import torch
sys.path.insert(0, ".")
import symmer_b200.ops as ops
from oracle import pauli_oracle as po
ops.device()
n, M = 1000, 40000
a_s, _ = po.random_operator(n, M, seed=3)
a = ops.pack(torch.from_numpy(a_s), n)
blk = a[:10000].contiguous()
o = ops.commute_mma(blk, a); torch.cuda.synchronize()
o = ops.commute(blk, a); torch.cuda.synchronize()
