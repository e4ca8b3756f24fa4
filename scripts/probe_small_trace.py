"""Kernel timeline (torch.profiler / CUPTI) of one C1 product + cleanup: per-kernel device time and the gaps between them."""
import os, sys, json
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import pauli_oracle as po
from symmer_b200 import PauliwordOp, ops
dev = ops.device()
s, c = po.random_operator(1000, 500, seed=1)
A = PauliwordOp(s, c); ac = A._coeff_dev(); axz = A._xz
bxz, bc = axz, ac
if os.environ.get("AB"):
    s2, c2 = po.random_operator(1000, 500, seed=2)
    B = PauliwordOp(s2, c2); bc = B._coeff_dev(); bxz = B._xz
if os.environ.get("TILE"): ops.set_tuning(0, 0)
fn = lambda: ops.mul_cleanup(axz, ac, bxz, bc)
if os.environ.get("ROT"):
    s3, c3 = po.random_operator(1000, 100000, seed=3)
    C = PauliwordOp(s3, c3); C._coeff_dev()
    q = PauliwordOp(po.random_operator(1000, 1, seed=4)[0], [1.0])
    fn = lambda: C._rotate_by_single_Pword(q, 0.3)
if os.environ.get("C2"):
    from symmer_b200 import IndependentOp
    d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "hamiltonians", "H2O_STO3G.npz"))
    n = int(d["n_qubits"][0])
    H = PauliwordOp(np.unpackbits(d["symp"], axis=1)[:, :2 * n].astype(bool), d["coeff"]); H._coeff_dev()
    fn = lambda: IndependentOp.symmetry_generators(H)
for _ in range(30): fn()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(10): fn()
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/c1_trace.json")
ev = json.load(open("/tmp/c1_trace.json"))["traceEvents"]
k = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")], key=lambda e: e["ts"])
n = len(k) // 10
print("device events per call:", n)
last = k[-n:]
t0 = last[0]["ts"]
for e in last:
    print(f"{e['ts'] - t0:9.1f} us  +{e['dur']:7.1f} us  {e['name'][:90]}")
print("busy", sum(e["dur"] for e in last), "span", last[-1]["ts"] + last[-1]["dur"] - t0)
