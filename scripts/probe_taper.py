"""Where taper_it(H2O STO-3G) spends its time: cProfile of the warm call + CUDA launch count."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from symmer_b200 import PauliwordOp, QuantumState, ops
from symmer_b200.projection import QubitTapering
ops.device()
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "hamiltonians", "H2O_STO3G.npz"))
n = int(d["n_qubits"][0])
H = PauliwordOp(np.unpackbits(d["symp"], axis=1)[:, :2 * n].astype(bool), d["coeff"])
hf = np.array([1] * 10 + [0] * (n - 10))
def run():
    qt = QubitTapering(H)
    return qt.taper_it(ref_state=hf)
for _ in range(5): run()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20): out = run()
torch.cuda.synchronize()
print("taper_it(H2O) incl. symmetry generators: %.3f ms/call" % ((time.perf_counter() - t) / 20 * 1e3))
qt = QubitTapering(H); qt.taper_it(ref_state=hf)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20): out = qt.taper_it(ref_state=hf)
torch.cuda.synchronize()
print("taper_it(H2O) only: %.3f ms/call, %d -> %d qubits" % ((time.perf_counter() - t) / 20 * 1e3, n, out.n_qubits))
l0 = ops.launch_count() if hasattr(ops, "launch_count") else 0
qt.taper_it(ref_state=hf)
print("launches per taper_it:", (ops.launch_count() - l0) if hasattr(ops, "launch_count") else "n/a")
pr = cProfile.Profile(); pr.enable()
for _ in range(20): qt.taper_it(ref_state=hf)
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)
