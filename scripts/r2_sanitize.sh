#!/bin/bash
# compute-sanitizer over the class-dedup / group / join / mirror kernels on small inputs
set -u
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import pauli_oracle as po
from symmer_b200 import ops, PauliwordOp, QuantumState
ops.device()
def dev_op(s, c):
    n = s.shape[1] // 2
    return ops.pack(torch.from_numpy(np.ascontiguousarray(s)), n), torch.from_numpy(np.asarray(c, dtype=complex)).cuda()
ops.set_tuning(0, 0)
rng = np.random.default_rng(1)
for variant in (2, 1, 0):
    ops.set_tuning(11, variant)
    for n, m1, m2 in [(64, 700, 300), (1000, 300, 120), (128, 1500, 40)]:
        a_s, a_c = po.random_operator(n, m1, seed=n + m1)
        b_s, b_c = po.random_operator(n, m2, seed=n + m2 + 1)
        b_s[: m2 // 3] = a_s[: m2 // 3]
        a_s[m1 // 2:] = a_s[: m1 - m1 // 2]
        for thr in (1e-15, 0.7):
            xz, c = ops.mul_cleanup(*dev_op(a_s, a_c), *dev_op(b_s, b_c), thr)
    # overflow: identical rows
    a_s, a_c = po.random_operator(128, 20000, seed=3); a_s[:19500] = a_s[0]
    b_s, b_c = po.random_operator(128, 20, seed=4)
    xz, c = ops.mul_cleanup(*dev_op(a_s, a_c), *dev_op(b_s, b_c))
    # block list
    a_s, a_c = po.random_operator(1000, 600, seed=31); b_s, b_c = po.random_operator(1000, 90, seed=32); b_s[:30] = a_s[:30]
    a, ac = dev_op(a_s, a_c); b, bc = dev_op(b_s, b_c)
    ops.mul_blocks_cleanup(a, ac, b, bc, [(0, 300, 40, 90), (300, 301, 0, 40), (301, 600, 0, 35), (0, 200, 0, 40)])
    print("variant", variant, "ok", flush=True)
ops.set_tuning(11, 2); ops.set_tuning(0, 1 << 22)
# fused rotation, lex order, join, mirror
s, c = po.random_operator(128, 3000, seed=7); s[1500:] = s[:1500]
xz, cc = dev_op(s, c); q = ops.pack(torch.from_numpy(po.random_operator(128, 1, seed=8)[0]), 128)
ops.rotate_dedup(xz, cc, q, 0.8, 0.6)
ops.lex_order(xz)
P = PauliwordOp(s, c); assert P == PauliwordOp(s[::-1].copy(), c[::-1].copy())
bra = QuantumState(rng.integers(0, 2, (50, 70)), rng.standard_normal(50), vec_type='bra'); ket = QuantumState(rng.integers(0, 2, (60, 70)), rng.standard_normal(60))
bra * ket
big, _ = po.random_operator(130, 8300, seed=9)
ops.commute_self(ops.pack(torch.from_numpy(big), 130), block_rows=2048)
print("all ok", flush=True)
PY
TOOL=${SAN_TOOL:-memcheck}
timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 python /tmp/san_small.py > gpurun_out/r2_san_$TOOL.log 2>&1; echo "$TOOL exit $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|variant|all ok|Error" gpurun_out/r2_san_$TOOL.log | head -20
