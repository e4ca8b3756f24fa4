import sys, ctypes
import torch
sys.path.insert(0, ".")
import symmer_b200.ops as ops
from symmer_b200 import _cabi
from oracle import pauli_oracle as po

ops.device()
n, M, N = 1000, 12500, 2000
a_s, a_c = po.random_operator(n, M, seed=1); b_s, b_c = po.random_operator(n, N, seed=2)
a = ops.pack(torch.from_numpy(a_s), n); ac = torch.from_numpy(a_c).cuda()
b = ops.pack(torch.from_numpy(b_s), n); bc = torch.from_numpy(b_c).cuda()
L = ops.lib(); W = 16
for by_t in [0, 1 << 40]:
    ops.set_tuning(0, by_t)
    ws = ops.workspace(L.sym_mul_cleanup_ws_bytes(M, N, W))
    U = ctypes.c_int64(0)
    _cabi.check(L.sym_mul_cleanup_count(ops._p(a), ops._p(ac), M, ops._p(b), ops._p(bc), N, W, 1e-15, None, ctypes.byref(U), ops._p(ws), ws.numel(), ops._stream()))
    U = U.value
    out_xz = torch.empty((U, 2 * W), dtype=torch.int64, device="cuda"); out_c = torch.empty(U, dtype=torch.complex128, device="cuda")
    for variant in [1, 2, 3]:
        ops.set_tuning(1, variant)
        ts = []
        for _ in range(4):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            _cabi.check(L.sym_mul_cleanup_emit(ops._p(a), ops._p(ac), M, ops._p(b), ops._p(bc), N, W, U, ops._p(out_xz), ops._p(out_c), ops._p(ws), ws.numel(), ops._stream()))
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(f"by_t={by_t>0} variant={variant} emit ms={min(ts):.3f} GB/s={U*272/min(ts)/1e6:.0f}", flush=True)
