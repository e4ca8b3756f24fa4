"""High-collision variant of the multiply+cleanup measurement: both operands drawn from the span of
g = 28 fixed generators, so the cross terms collide heavily (U <= 2^28 distinct rows) and the dedup does real
merging — against the collision-free C5 operands of bench.py where U = T. Prints one JSON line per size.
Small sizes are checked against the oracle first.

    python scripts/probe_collisions.py [rows_a rows_b]      # default 12500 x 10000 at 1000 qubits (config C5/8)
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import pauli_oracle as po  # noqa: E402
from symmer_b200 import ops  # noqa: E402

N_QUBITS, N_GEN = 1000, 28


def span_operator(gens, n_rows, rng):
    """n_rows random elements of the span of the generator rows (GF(2) combinations), random complex coefficients."""
    pick = rng.random((n_rows, gens.shape[0])) < 0.5
    symp = (pick.astype(np.uint8) @ gens.astype(np.uint8)) % 2
    coeff = rng.standard_normal(n_rows) + 1j * rng.standard_normal(n_rows)
    return symp.astype(bool), coeff


def main():
    dev = ops.device()
    rows_a, rows_b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (12500, 10000)
    rng = np.random.default_rng(11)
    gens = rng.random((N_GEN, 2 * N_QUBITS)) < 0.3
    # parity at a size the oracle finishes in seconds
    a_s, a_c = span_operator(gens[:8], 300, rng)
    b_s, b_c = span_operator(gens[:8], 200, rng)
    xz, c = ops.mul_cleanup(ops.pack(torch.from_numpy(a_s), N_QUBITS), torch.from_numpy(a_c).to(dev),
                            ops.pack(torch.from_numpy(b_s), N_QUBITS), torch.from_numpy(b_c).to(dev), 1e-15)
    ref_s, ref_c = po.multiply_by_operator(a_s, a_c, b_s, b_c)
    ok, why = po.compare_term_sets(ops.unpack(xz, N_QUBITS).cpu().numpy(), c.cpu().numpy(), ref_s, ref_c,
                                   scale=float(np.abs(a_c).max() * np.abs(b_c).max()) * 300)
    assert ok, why
    a_s, a_c = span_operator(gens, rows_a, rng)
    b_s, b_c = span_operator(gens, rows_b, rng)
    a, ac = ops.pack(torch.from_numpy(a_s), N_QUBITS), torch.from_numpy(a_c).to(dev)
    b, bc = ops.pack(torch.from_numpy(b_s), N_QUBITS), torch.from_numpy(b_c).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times, U = [], 0
    for it in range(6):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out_xz, out_c = ops.mul_cleanup(a, ac, b, bc, 1e-15)
        e1.record()
        torch.cuda.synchronize()
        U = out_xz.shape[0]
        if it >= 2:
            times.append(e0.elapsed_time(e1))
        del out_xz, out_c
    T = rows_a * rows_b
    ms = float(np.mean(times))
    print(json.dumps({"path": f"high-collision product: {rows_a} x {rows_b} terms from the span of {N_GEN} generators, 1000 q",
                      "cross_terms": T, "unique_terms": U, "ms": ms, "cross_terms_per_s": T / (ms * 1e-3),
                      "model_bytes": (2 + U / T) * 272 * T, "model_gbs": (2 + U / T) * 272 * T / (ms * 1e-3) / 1e9}), flush=True)


if __name__ == "__main__":
    main()
