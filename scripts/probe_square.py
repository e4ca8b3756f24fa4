"""One square product (heavy duplicates) in a chosen mode, for an ncu launch list."""
import sys
import torch
sys.path.insert(0, ".")
import symmer_b200.ops as ops
from oracle import pauli_oracle as po
mode = sys.argv[1] if len(sys.argv) > 1 else "tiles"
ops.device()
limit, knob6 = {"by_t": (1 << 40, 1), "tiles": (0, 1), "sorted": (0, 0)}[mode]
ops.set_tuning(0, limit); ops.set_tuning(6, knob6)
a_s, a_c = po.random_operator(1000, 1000, seed=1)
a = ops.pack(torch.from_numpy(a_s), 1000); ac = torch.from_numpy(a_c).cuda()
for _ in range(2):
    xz, c = ops.mul_cleanup(a, ac, a, ac)
torch.cuda.synchronize()
print(mode, xz.shape)
