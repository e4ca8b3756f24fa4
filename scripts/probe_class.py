"""Step time of the C5/8 product (12500 x 10000 terms, 1000 q) under the dedup variants: record sort (knob 10 = 0),
class-local dedup with 512-thread CTAs (11 = 0) and 1024-thread CTAs (11 = 1); collision-free and span operands."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import pauli_oracle as po  # noqa: E402
from symmer_b200 import ops  # noqa: E402

N_QUBITS = 1000


def span_operator(gens, n_rows, rng):
    pick = rng.random((n_rows, gens.shape[0])) < 0.5
    symp = (pick.astype(np.uint8) @ gens.astype(np.uint8)) % 2
    return symp.astype(bool), rng.standard_normal(n_rows) + 1j * rng.standard_normal(n_rows)


def time_product(a, ac, b, bc, reps=6):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=a.device)
    ts, U = [], 0
    for it in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        xz, c = ops.mul_cleanup(a, ac, b, bc, 1e-15)
        e1.record()
        torch.cuda.synchronize()
        U = xz.shape[0]
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
        del xz, c
    return float(np.mean(ts)), U


def main():
    dev = ops.device()
    rows_a, rows_b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (12500, 10000)
    a_s, a_c = po.random_operator(N_QUBITS, rows_a, seed=100)
    b_s, b_c = po.random_operator(N_QUBITS, rows_b, seed=7)
    rng = np.random.default_rng(11)
    gens = rng.random((28, 2 * N_QUBITS)) < 0.3
    sa_s, sa_c = span_operator(gens, rows_a, rng)
    sb_s, sb_c = span_operator(gens, rows_b, rng)
    for name, (xs, xc, ys, yc) in (("collision-free", (a_s, a_c, b_s, b_c)), ("span-28", (sa_s, sa_c, sb_s, sb_c))):
        if os.environ.get("PROBE_SPAN_ONLY") and name != "span-28":
            continue
        a, ac = ops.pack(torch.from_numpy(xs), N_QUBITS), torch.from_numpy(xc).to(dev)
        b, bc = ops.pack(torch.from_numpy(ys), N_QUBITS), torch.from_numpy(yc).to(dev)
        for label, knobs in (("sort", {10: 0}), ("class1024", {10: 1, 11: 1}), ("class32-table", {10: 1, 11: 3}), ("class32-prefilter", {10: 1, 11: 2})):
            if os.environ.get("PROBE_ONLY") and os.environ["PROBE_ONLY"] not in label:
                continue
            for k, v in knobs.items():
                ops.set_tuning(k, v)
            ms, U = time_product(a, ac, b, bc)
            print(json.dumps({"operands": name, "dedup": label, "ms": ms, "U": U, "T": rows_a * rows_b}), flush=True)
    ops.set_tuning(10, 1)
    ops.set_tuning(11, 2)


if __name__ == "__main__":
    main()
