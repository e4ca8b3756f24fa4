"""Secondary measurements of the other §8 rows (commute, rotations, expval, GF(2), C1) with the
oracle port timed beside them on bounded samples. Prints one JSON object per path."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import pauli_oracle as po  # noqa: E402
from symmer_b200 import PauliwordOp, QuantumState, IndependentOp, ops  # noqa: E402

PEAK = 6540.2
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]


def gpu_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
        del out
    return min(ts)


def cpu_time(fn, reps=1):
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def load_ham(tag):
    d = np.load(os.path.join("tests", "golden", "hamiltonians", tag + ".npz"))
    n = int(d["n_qubits"][0])
    return np.unpackbits(d["symp"], axis=1)[:, :2 * n].astype(bool), d["coeff"], n


def emit(name, **kw):
    print(json.dumps({"path": name, **kw}), flush=True)


def main():
    ops.device()
    which = sys.argv[1:] or ["c1", "commute", "rotate", "expval", "gf2"]

    if "c1" in which:
        np.random.seed(1)
        P = PauliwordOp.random(1000, 500)
        t = gpu_time(lambda: P * P, reps=10)
        s, c = P.symp_matrix, P.coeff_vec
        tc = cpu_time(lambda: po.multiply(s, c, s, c))
        emit("C1 square 1000q x 500 terms (250k cross terms + cleanup)", gpu_s=t, cross_terms_per_s=250000 / t,
             cpu_port_s=tc, cpu_cross_terms_per_s=250000 / tc, speedup=tc / t)

    if "commute" in which:
        for n, M in [(36, 42599), (1000, 100000)]:
            a_s, _ = po.random_operator(n, M, seed=3)
            a = ops.pack(torch.from_numpy(a_s), n)
            rows = min(M, 20000)                       # tile of the self-adjacency: rows x M bytes of output
            blk = a[:rows].contiguous()
            t = gpu_time(lambda: ops.commute(blk, a), reps=5)
            pairs = rows * M
            tb = gpu_time(lambda: ops.commute_bits(blk, a), reps=5)
            sub = a_s[:min(M, 4000)]
            tc = cpu_time(lambda: po.commutes_termwise(sub[:2000], sub))
            W = (n + 63) // 64
            emit(f"commute {n}q: {rows}x{M} block of the adjacency matrix", gpu_s=t, pairs_per_s=pairs / t,
                 out_write_gbs=pairs / t / 1e9, frac_hbm_write=pairs / t / 1e9 / PEAK,
                 lop3_per_pair=4 * W, bits_variant_pairs_per_s=pairs / tb,
                 cpu_port_pairs_per_s=2000 * len(sub) / tc, cpu_threads="BLAS default")

    if "rotate" in which:
        np.random.seed(2)
        P = PauliwordOp.random(1000, 100000)
        Q = PauliwordOp.random(1000, 1)
        Q.coeff_vec[0] = 1
        t = gpu_time(lambda: P.perform_rotations([(Q, 0.731)]), reps=5)
        out = P.perform_rotations([(Q, 0.731)])
        tcl = gpu_time(lambda: P.perform_rotations([(Q, np.pi / 2)]), reps=5)
        sub_s, sub_c = P.symp_matrix[:10000], P.coeff_vec[:10000]
        tc = cpu_time(lambda: po.perform_rotations(sub_s, sub_c, [(Q.symp_matrix[0], 0.731)]))
        R = 272
        emit("C3 one non-Clifford rotation of 1000q x 100k terms (incl. dedup)", gpu_s=t, rows_per_s=100000 / t,
             rows_out=out.n_terms, model_gbs=100000 * 5.5 * R / t / 1e9, frac_hbm=100000 * 5.5 * R / t / 1e9 / PEAK,
             clifford_gpu_s=tcl, cpu_port_rows_per_s=10000 / tc, cpu_sample="10k rows")
        # 100 independent single rotations (README claim x100)
        t100 = gpu_time(lambda: [P.perform_rotations([(Q, 0.1 + 0.013 * k)]) for k in range(100)], reps=1, warm=1)   # warm allocator
        emit("C3(i) 100 independent non-Clifford rotations of the same 100k-term operator", gpu_s=t100,
             rows_per_s=100 * 100000 / t100)

    if "rotseq" in which or "rotate" in which:
        # C3(ii): the longest SEQUENTIAL prefix of a random non-Clifford rotation sequence that fits a row cap
        # (rows grow ~1.5x per rotation: 100 sequential random rotations of 1e5 rows are impossible)
        cap = 40_000_000                                  # rows: 10.9 GB of packed operator + dedup workspace
        np.random.seed(4)
        op = PauliwordOp.random(1000, 100000)
        gens = []
        for k in range(24):
            G = PauliwordOp.random(1000, 1)
            G.coeff_vec[0] = 1
            gens.append((G, 0.1 + 1.3 * np.random.rand()))
        start = op
        times = []
        for rep in range(2):      # the first pass pays cudaMalloc for every new buffer size (240 ms cold vs 34 ms warm)
            op = start
            torch.cuda.synchronize()
            rows_in, steps, sizes = 0, 0, [op.n_terms]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for G, th in gens:
                if op.n_terms * 2 > cap:
                    break
                rows_in += op.n_terms
                op = op.perform_rotations([(G, th)])
                steps += 1
                sizes.append(op.n_terms)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e-3)
        t = times[-1]
        emit("C3(ii) sequential non-Clifford rotations of a 1000q x 100k-term operator until the next one would exceed "
             f"{cap:.0e} rows", gpu_s=t, gpu_s_cold_allocator=times[0], rotations=steps, rows_in_total=rows_in,
             rows_per_s=rows_in / t, rows_after_each=sizes, model_gbs=rows_in * 5.5 * 272 / t / 1e9,
             frac_hbm=rows_in * 5.5 * 272 / t / 1e9 / PEAK)
        del op, start

    if "expval" in which:
        for tag in ["H2O_STO3G", "NH3_STO3G", "HOOH_STO3G"]:
            symp, coeff, n = load_ham(tag)
            H = PauliwordOp(symp, coeff)
            rng = np.random.default_rng(0)
            psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
            psi /= np.linalg.norm(psi)
            psi_d = torch.from_numpy(psi).cuda()
            H._terms_sorted()
            t = gpu_time(lambda: H.expval_dense(psi_d), reps=3, warm=1)
            e = complex(H.expval_dense(psi_d).cpu().numpy())
            rec = {"gpu_s": t, "sign_evals_per_s": (1 << n) * H.n_terms / t, "expval": [e.real, e.imag],
                   "n_terms": H.n_terms, "n_qubits": n}
            if n <= 16:
                tc = cpu_time(lambda: po.expval_dense(symp, coeff, psi))
                rec["cpu_port_s"] = tc
                rec["cpu_expval_real"] = po.expval_dense(symp, coeff, psi).real
            emit(f"C4 matrix-free expval {tag}", **rec)

    if "g" in which:
        # kernels of rows g1-g3: qubit-wise commutation, qubit gather, Walsh-Hadamard decomposition
        n, M = 1000, 50000
        a_s = np.random.default_rng(3).random((M, 2 * n)) < 0.3
        a = ops.pack(torch.from_numpy(a_s), n)
        blk = a[:20000].contiguous()
        t = gpu_time(lambda: ops.commute_qwc(blk, a), reps=3, warm=1)
        emit("qwc 1000q: 20000x50000 block", gpu_s=t, pairs_per_s=20000 * M / t, out_write_gbs=20000 * M / t / 1e9,
             lop3_per_pair=5 * 2 * 16)
        n36, M36 = 36, 42599
        b_s = np.random.default_rng(4).random((M36, 2 * n36)) < 0.3
        b = ops.pack(torch.from_numpy(b_s), n36)
        t = gpu_time(lambda: ops.commute_qwc(b[:20000].contiguous(), b), reps=3, warm=1)
        emit("qwc 36q: 20000x42599 block (NaCl size)", gpu_s=t, pairs_per_s=20000 * M36 / t,
             frac_hbm_write=20000 * M36 / t / 1e9 / PEAK)
        perm = np.random.default_rng(0).permutation(n)
        t = gpu_time(lambda: ops.gather_qubits(a, perm, n), reps=3, warm=1)
        emit("gather_qubits: random permutation of 1000 qubits, 50000 rows", gpu_s=t, rows_per_s=M / t,
             rw_gbs=2 * M * 256 / t / 1e9)
        for nq, K in [(12, 4096), (20, 64), (24, 4)]:
            table = torch.randn(K, 1 << nq, dtype=torch.complex128, device="cuda")
            t = gpu_time(lambda: ops.pauli_decompose_diagonals(table, nq), reps=3, warm=1)
            passes = 1 + max(0, nq - 12)
            emit(f"walsh-hadamard of {K} diagonals of 2^{nq} entries (in place)", gpu_s=t, entries_per_s=K * (1 << nq) / t,
                 passes=passes, rw_gbs=passes * 2 * K * (1 << nq) * 16 / t / 1e9,
                 frac_hbm=passes * 2 * K * (1 << nq) * 16 / t / 1e9 / PEAK)
            del table

    if "gf2" in which or "c2" in which:
        from symmer_b200 import QubitTapering
        symp, coeff, n = load_ham("H2O_STO3G")
        H = PauliwordOp(symp, coeff)
        t = gpu_time(lambda: IndependentOp.symmetry_generators(H), reps=20, warm=3)
        t_seam = gpu_time(lambda: IndependentOp._symmetry_generators_host(H), reps=20, warm=3)
        tc = cpu_time(lambda: po.symmetry_generator_rows(symp), reps=5)
        ta = gpu_time(lambda: ops.commute(H.device_rows, H.device_rows), reps=10)
        tca = cpu_time(lambda: po.commutes_termwise(symp, symp), reps=3)
        hf = np.asarray(np.load(os.path.join("tests", "golden", "hamiltonians", "H2O_STO3G.npz"))["hf_array"], dtype=int)
        t_taper = None
        if hf is not None:
            t_taper = gpu_time(lambda: QubitTapering(H).taper_it(ref_state=hf), reps=5, warm=2)
        emit("C2 symmetry generators H2O STO-3G (14q, 1086 terms)", gpu_s=t, gpu_s_array_seam_form=t_seam, cpu_port_s=tc,
             adjacency_gpu_s=ta, adjacency_pairs_per_s=1086 * 1086 / ta, adjacency_cpu_s=tca,
             taper_it_end_to_end_gpu_s=t_taper)
    if "gf2" in which:
        rng = np.random.default_rng(1)
        m = rng.random((2000, 100000)) < 0.3
        bits = ops.pack_matrix(torch.from_numpy(m))
        def run():
            b = bits.clone()
            ops.rref_packed(b, m.shape[1])
            return b
        t = gpu_time(run, reps=1, warm=0)
        emit("GF(2) rref 2000 x 100000 (generator_reconstruction scale)", gpu_s=t,
             model_gbs=2000 * 2 * 2000 * 1563 * 8 / t / 1e9 / 2)


if __name__ == "__main__":
    main()
