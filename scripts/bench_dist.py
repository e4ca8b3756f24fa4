"""Multi-GPU measurements of SURVEY.md section 8e (rows e1-e5) under torchrun, one JSON line per path on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_dist.py

e1  product + cleanup: exchange-free owner classes against the hash-partitioned all-to-all of records, weak workload
    (C5/8 per GPU) and STRONG workload (all of C5 = 1e9 cross terms over the N ranks), with the bytes that cross NVLink
e2  cleanup of a term-sharded operator (rows travel to their owner: variable-size all-to-all)
e3  commutation matrix in row blocks (no collective); self-adjacency computes the upper block triangle only
e4  general rotation of a term-sharded operator (rotation local, one exchange + dedup)
e5  config C4: matrix-free <psi|H|psi> of HOOH STO-3G (24 q), 2^24 basis rows sharded, one all-reduce
Every time is device time (CUDA events), the maximum over the ranks.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import pauli_oracle as po  # noqa: E402  (input generator only)
from symmer_b200 import PauliwordOp, ops  # noqa: E402
from symmer_b200 import dist as sdist  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = ops.device()
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ONLY = set(filter(None, os.environ.get("DIST_ONLY", "").split(",")))
N_Q = 1000


def want(tag):
    return not ONLY or tag in ONLY


def timed(fn, reps=4, warm=2):
    ms = []
    for it in range(warm + reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= warm:
            ms.append(e0.elapsed_time(e1))
        del out
    t = torch.tensor([float(np.mean(ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def total(x):
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    return float(t.item())


def emit(**kw):
    if rank == 0:
        print(json.dumps(dict(n_gpus=world, **kw)), flush=True)


def free_gb():
    return torch.cuda.mem_get_info()[0] / 2 ** 30


# ------------------------------------------------------------------ e1: product, weak and strong
b_s, b_c = po.random_operator(N_Q, 10000, seed=7)
b, bc = ops.pack(torch.from_numpy(b_s), N_Q), torch.from_numpy(b_c).to(dev)
if want("e1"):
    a_s, a_c = po.random_operator(N_Q, 12500, seed=100 + rank)
    a, ac = ops.pack(torch.from_numpy(a_s), N_Q), torch.from_numpy(a_c).to(dev)
    T = 12500 * 10000
    for method in ("owner", "alltoall") if world > 1 else ("owner",):
        cache = {}
        st = {}

        def step():
            xz, c, info = sdist.sharded_product(a, ac, b, bc, method=method, block_sizes=[12500] * world, cache=cache)
            st["U"] = xz.shape[0]
            return xz, c

        ms = timed(step)
        U = total(st["U"])
        nvlink = ((world - 1) * 12500 * 272 if method == "owner" else T * 8 * (world - 1) / world + (world - 1) * 12500 * 272)
        emit(path="e1 product+cleanup, weak (C5/8 per GPU)", method=method, ms=ms, cross_terms=T * world, unique_terms=U,
             cross_terms_per_s=T * world / (ms * 1e-3), nvlink_bytes_in_per_gpu=nvlink,
             collective="all_gather_into_tensor of A" + ("" if method == "owner" else " + all_to_all_single of 8-byte records"))
    if world > 1:
        # phases of the owner method
        def ph_gather():
            return sdist.all_gather_operator(a, ac, sizes=[12500] * world)
        ms_g = timed(ph_gather)
        a_full, a_c_full, _ = sdist.all_gather_operator(a, ac, sizes=[12500] * world)
        lg = sdist.log2_exact(world)
        ms_p = timed(lambda: sdist.partition_by_owner(a_full, a_c_full, lg))
        b_part = sdist.partition_by_owner(b, bc, lg)
        a_part = sdist.partition_by_owner(a_full, a_c_full, lg)
        ms_x = timed(lambda: sdist.owned_product(a_full, a_c_full, b, bc, lg, rank, a_part=a_part, b_part=b_part)[:2])
        emit(path="e1 owner method, phases", all_gather_ms=ms_g, partition_a_ms=ms_p, owned_product_ms=ms_x,
             note="owned_product = class-local dedup + tiled emission of this rank's 8 blocks")
        del a_full, a_c_full, a_part, b_part
    del a, ac
    ops.release_workspace()
    torch.cuda.empty_cache()
    # strong scaling: all of C5 (1e5 x 1e4 terms) over the ranks
    rows = 100000 // world
    need_gb = rows * 10000 * (272 + 60) / 2 ** 30
    if world > 1 and need_gb < free_gb() - 6:
        blocks = [po.random_operator(N_Q, 12500, seed=100 + r) for r in range(rank * 8 // world, (rank + 1) * 8 // world)]
        a_s2, a_c2 = np.vstack([x[0] for x in blocks]), np.hstack([x[1] for x in blocks])
        a2, ac2 = ops.pack(torch.from_numpy(a_s2), N_Q), torch.from_numpy(a_c2).to(dev)
        del blocks, a_s2
        cache = {}
        st = {}

        def step_strong():
            xz, c, info = sdist.sharded_product(a2, ac2, b, bc, block_sizes=[rows] * world, cache=cache)
            st["U"] = xz.shape[0]
            return xz, c

        ms = timed(step_strong, reps=3, warm=1)
        emit(path="e1 product+cleanup, STRONG: config C5 itself (1e5 x 1e4 terms, 1e9 cross terms)", method="owner", ms=ms,
             cross_terms=10 ** 9, unique_terms=total(st["U"]), cross_terms_per_s=10 ** 9 / (ms * 1e-3),
             output_gb_per_gpu=st["U"] * 272 / 1e9)
        del a2, ac2
        ops.release_workspace()
        torch.cuda.empty_cache()
    elif world > 1:
        emit(path="e1 STRONG skipped", reason=f"needs {need_gb:.0f} GB per GPU, {free_gb():.0f} GB free")

# ------------------------------------------------------------------ e2 / e4: term-sharded cleanup and rotation
if want("e2") or want("e4"):
    g = torch.Generator(device=dev)
    g.manual_seed(5)                                    # the pool is the same on every rank
    pool_rows = 8_000_000
    W2 = 2 * ((N_Q + 63) // 64)
    pool = torch.randint(-2 ** 63, 2 ** 63 - 1, (pool_rows, W2), dtype=torch.int64, device=dev, generator=g)
    pool[:, W2 // 2 - 1] &= (1 << (N_Q - 64 * (W2 // 2 - 1))) - 1       # zero padding bits above qubit 999
    pool[:, W2 - 1] &= (1 << (N_Q - 64 * (W2 // 2 - 1))) - 1
    g2 = torch.Generator(device=dev)
    g2.manual_seed(50 + rank)
    n_local = 10_000_000 // world
    idx = torch.randint(0, pool_rows, (n_local,), device=dev, generator=g2)
    xz_local = pool[idx].contiguous()
    c_local = torch.randn(n_local, dtype=torch.float64, device=dev, generator=g2).to(torch.complex128)
    del idx
    if want("e2"):
        st = {}

        def step_clean():
            xz, c = sdist.sharded_cleanup(xz_local, c_local)
            st["U"] = xz.shape[0]
            return xz, c

        ms = timed(step_clean)
        emit(path="e2 cleanup of a term-sharded operator (1e7 rows drawn from 8e6 distinct, 1000 q)", ms=ms, rows=n_local * world,
             unique_rows=total(st["U"]), rows_per_s=n_local * world / (ms * 1e-3),
             nvlink_bytes_out_per_gpu=n_local * 272 * (world - 1) / world,
             collective="all_to_all_single of rows (256 B) + of coefficients (16 B), variable splits")
    if want("e4"):
        cxz, cc = sdist.sharded_cleanup(xz_local, c_local)
        q_row = pool[:1].contiguous()
        st = {}

        def step_rot():
            xz, c = sdist.sharded_rotation(cxz, cc, q_row, 0.37)
            st["U"] = xz.shape[0]
            return xz, c

        ms = timed(step_rot)
        rows_in = total(cxz.shape[0])
        emit(path="e4 general rotation of a term-sharded operator (1000 q)", ms=ms, rows_in=rows_in, rows_out=total(st["U"]),
             rows_per_s=rows_in / (ms * 1e-3), collective="all_to_all_single of the rotated rows, then local dedup")
        ms = timed(lambda: sdist.sharded_rotation(cxz, cc, q_row, None))
        emit(path="e4 Clifford rotation of a term-sharded operator (1000 q)", ms=ms, rows_in=rows_in,
             rows_per_s=rows_in / (ms * 1e-3), collective="none (relabelling, dedup deferred)")
        del cxz, cc
    del pool, xz_local, c_local
    ops.release_workspace()
    torch.cuda.empty_cache()

# ------------------------------------------------------------------ e3: commutation matrix in row blocks
if want("e3"):
    g = torch.Generator(device=dev)
    g.manual_seed(9)
    M = 131072
    W2 = 2 * ((N_Q + 63) // 64)
    big = torch.randint(-2 ** 63, 2 ** 63 - 1, (M, W2), dtype=torch.int64, device=dev, generator=g)
    ms = timed(lambda: sdist.sharded_commute(big, big)[0], reps=3, warm=1)
    emit(path="e3 commutes_termwise(A, A), 131072 rows at 1000 q, full row blocks", ms=ms, pairs=M * M, pairs_per_s=M * M / (ms * 1e-3),
         collective="none")
    if hasattr(sdist, "sharded_adjacency"):
        ms = timed(lambda: sdist.sharded_adjacency(big)[0], reps=3, warm=1)
        emit(path="e3 adjacency_matrix(A): upper block triangle only, mirrored", ms=ms, pairs=M * M, pairs_per_s=M * M / (ms * 1e-3),
             collective="none")
    del big
    torch.cuda.empty_cache()

# ------------------------------------------------------------------ e5: config C4
if want("e5"):
    d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "hamiltonians", "HOOH_STO3G.npz"))
    n = int(d["n_qubits"][0])
    H = PauliwordOp(np.unpackbits(d["symp"], axis=1)[:, :2 * n].astype(bool), d["coeff"])
    rng = np.random.default_rng(0)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    psi_d = torch.from_numpy(psi).to(dev)
    xm, zm, cp = H._terms_sorted()
    res = {}

    def step_e():
        res["e"] = sdist.sharded_expval(xm, zm, cp, n, psi_d)
        return None

    ms = timed(step_e, reps=5, warm=2)
    emit(path="e5 config C4: matrix-free <psi|H|psi>, HOOH STO-3G (24 q, 14905 terms), dense 2^24 state, basis rows sharded", ms=ms,
         sign_evals_per_s=(1 << n) * H.n_terms / (ms * 1e-3), expval_real=res["e"].real, expval_imag=res["e"].imag,
         collective="all_reduce(SUM) of one complex128")

if world > 1:
    dist.barrier()
    dist.destroy_process_group()
