import sys, ctypes
import torch
sys.path.insert(0, ".")
import symmer_b200.ops as ops
from symmer_b200 import _cabi
ops.device()
L = ops.lib()
T = 125_000_000
g = torch.Generator(device="cuda"); g.manual_seed(0)
keys0 = torch.randint(-2**63, 2**63 - 1, (T,), dtype=torch.int64, device="cuda", generator=g)
out = torch.empty_like(keys0)
counts = torch.zeros(256, dtype=torch.int64, device="cuda")
ws = ops.workspace(L.sym_partition_ws_bytes(T))
for variant in [0, 1, 2, 3, 5]:
    ops.set_tuning(2, variant)
    ts = []
    for _ in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _cabi.check(L.sym_partition_records(ops._p(keys0), T, 8, ops._p(out), ops._p(counts), ops._p(ws), ws.numel(), ops._stream()))
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ok = bool((counts.sum() == T).item())
    srt = bool(((out[1:].view(torch.int64) >> 56) >= (out[:-1] >> 56)).all().item())
    print(f"scatter variant={variant} one 8-bit pass (hist+scan+scatter) ms={min(ts):.3f}  GB/s(24B/elem)={T*24/min(ts)/1e6:.0f} counts_ok={ok} sorted={srt}", flush=True)
