"""world_size-2 gloo worker for tests/test_dist_gloo.py (host-side logic of symmer_b200.dist on CPU tensors)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from symmer_b200 import dist as sdist  # noqa: E402
from oracle import pauli_oracle as po  # noqa: E402


def main():
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['MASTER_PORT']}",
                            rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    lg = sdist.log2_exact(world)

    # 1. variable-size row gather
    rows = torch.full((3 + rank, 4), rank, dtype=torch.int64)
    full, offs = sdist.all_gather_rows(rows)
    assert offs == [0, 3, 7] and full.shape == (7, 4)
    assert bool((full[:3] == 0).all()) and bool((full[3:] == 1).all())

    # 2. hash-partitioned record exchange: records carry (owner in the top bit, source rank, serial)
    rng = np.random.default_rng(100 + rank)
    n_rec = 1000 + 37 * rank
    owner = rng.integers(0, world, size=n_rec).astype(np.uint64)
    recs = (owner << np.uint64(64 - lg)) | (np.uint64(rank) << np.uint64(32)) | np.arange(n_rec, dtype=np.uint64)
    order = np.argsort(owner, kind="stable")                      # what sym_partition_records does on the GPU
    part = torch.from_numpy(recs[order].view(np.int64))
    counts = torch.from_numpy(np.bincount(owner.astype(np.int64), minlength=world).astype(np.int64))
    mine = sdist.exchange_records(part, counts).numpy().view(np.uint64)
    assert np.all((mine >> np.uint64(64 - lg)) == rank)           # I own everything I received
    total = torch.tensor([mine.size], dtype=torch.int64)
    dist.all_reduce(total)
    assert int(total) == 1000 + 1037                              # nothing lost or duplicated
    src = (mine >> np.uint64(32)) & np.uint64(0xFFFF)
    for s in range(world):                                        # grouped by source, order preserved
        serial = (mine[src == s] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        assert np.all(np.diff(serial) > 0)

    # 3. the whole sharded product with the oracle standing in for the CUDA kernels: block -> records
    #    -> owner exchange -> per-owner dedup; the union over ranks must equal the plain product.
    n, M, N = 9, 24, 10
    a_s, a_c = po.random_operator(n, M, seed=5)
    b_s, b_c = po.random_operator(n, N, seed=6)
    bounds = sdist.block_bounds(M, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    rows_blk, coeff_blk = po.cross_terms(a_s[lo:hi], a_c[lo:hi], b_s, b_c)
    packed = po.pack_bits(rows_blk)
    key = packed[:, 0] * np.uint64(0x9E3779B97F4A7C15) ^ (packed[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F))
    own = (key >> np.uint64(64 - lg)).astype(np.int64)
    order = np.argsort(own, kind="stable")
    counts = torch.from_numpy(np.bincount(own, minlength=world).astype(np.int64))
    # ship (x, z) words and the coefficient (re, im bit patterns) as four record streams
    streams = [packed[order, 0].view(np.int64), packed[order, 1].view(np.int64),
               np.ascontiguousarray(coeff_blk[order].real).view(np.int64),
               np.ascontiguousarray(coeff_blk[order].imag).view(np.int64)]
    got = [sdist.exchange_records(torch.from_numpy(s.copy()), counts).numpy() for s in streams]
    rx = po.unpack_bits(np.stack([got[0].view(np.uint64), got[1].view(np.uint64)], axis=1), n)
    rc = got[2].view(np.float64) + 1j * got[3].view(np.float64)
    loc_s, loc_c = po.symplectic_cleanup(rx, rc, 1e-15)
    gathered = [None] * world
    dist.all_gather_object(gathered, (loc_s, loc_c))
    all_s = np.vstack([g[0] for g in gathered])
    all_c = np.hstack([g[1] for g in gathered])
    assert len(np.unique(all_s, axis=0)) == len(all_s)            # owners are disjoint
    ref_s, ref_c = po.multiply_by_operator(a_s, a_c, b_s, b_c)
    ok, why = po.compare_term_sets(all_s, all_c, ref_s, ref_c, scale=np.abs(a_c).max() * np.abs(b_c).max())
    assert ok, why
    # 4. operator all-gather (equal blocks gather straight into the result; ragged blocks are padded)
    for sizes in ([5, 5], [4, 7]):
        xz_blk = torch.full((sizes[rank], 6), rank + 1, dtype=torch.int64)
        c_blk = torch.full((sizes[rank],), complex(rank, -rank), dtype=torch.complex128)
        xz_full, c_full, offs = sdist.all_gather_operator(xz_blk, c_blk)
        assert offs == [0, sizes[0], sizes[0] + sizes[1]]
        assert bool((xz_full[:sizes[0]] == 1).all()) and bool((xz_full[sizes[0]:] == 2).all())
        assert bool((c_full[:sizes[0]] == 0).all()) and bool((c_full[sizes[0]:] == complex(1, -1)).all())

    # 4b. row exchange (sharded_cleanup's collective): rows grouped by destination, order preserved
    n_rows = 6 + rank
    dest = (np.arange(n_rows) + rank) % world
    order_r = np.argsort(dest, kind="stable")
    xz_rows = torch.from_numpy((np.arange(n_rows, dtype=np.int64)[:, None] * 10 + rank + np.zeros((1, 4), np.int64))[order_r])
    c_rows = torch.from_numpy((np.arange(n_rows) + 1j * rank).astype(complex)[order_r])
    cnt = torch.from_numpy(np.bincount(dest, minlength=world).astype(np.int64))
    got_xz, got_c = sdist.exchange_rows(xz_rows, c_rows, cnt)
    assert got_xz.shape[0] == got_c.shape[0] and got_xz.shape[1] == 4
    src_rank = got_xz[:, 0].numpy() % 10
    serial = got_xz[:, 0].numpy() // 10
    assert np.all((serial + src_rank) % world == rank)           # everything I received was addressed to me
    assert np.array_equal(got_c.numpy().real, serial) and np.array_equal(got_c.numpy().imag, src_rank)
    tot = torch.tensor([got_xz.shape[0]], dtype=torch.int64)
    dist.all_reduce(tot)
    assert int(tot) == 6 + 7

    # 5. exchange-free owner partition (the default product path) with the oracle standing in for the
    #    kernels: class = a GF(2)-linear functional of the row, rank r multiplies class a of A with
    #    class a^r of B; the union over ranks must equal the plain product with disjoint owners.
    functional = np.random.default_rng(9).integers(0, 2, size=(2 * n, lg)).astype(np.int64)

    def classes(symp):
        bits = (symp.astype(np.int64) @ functional) & 1
        return (bits << np.arange(lg)).sum(axis=1)

    a_cls, b_cls = classes(a_s), classes(b_s)
    a_ord, b_ord = np.argsort(a_cls, kind="stable"), np.argsort(b_cls, kind="stable")
    a_p, a_cp, b_p, b_cp = a_s[a_ord], a_c[a_ord], b_s[b_ord], b_c[b_ord]
    blocks = sdist.owner_blocks(np.bincount(a_cls, minlength=world).tolist(), np.bincount(b_cls, minlength=world).tolist(),
                                rank)
    assert sum((p1 - p0) * (q1 - q0) for p0, p1, q0, q1 in blocks) > 0
    rows_l, coeff_l = [], []
    for p0, p1, q0, q1 in blocks:
        if p1 > p0 and q1 > q0:
            r_, c_ = po.cross_terms(a_p[p0:p1], a_cp[p0:p1], b_p[q0:q1], b_cp[q0:q1])
            assert np.all(classes(r_) == rank)                    # every generated term is mine
            rows_l.append(r_)
            coeff_l.append(c_)
    loc_s, loc_c = po.symplectic_cleanup(np.vstack(rows_l), np.hstack(coeff_l), 1e-15)
    gathered = [None] * world
    dist.all_gather_object(gathered, (loc_s, loc_c, sum(len(r_) for r_ in rows_l)))
    all_s = np.vstack([g[0] for g in gathered])
    all_c = np.hstack([g[1] for g in gathered])
    assert sum(g[2] for g in gathered) == M * N                   # every cross term generated exactly once
    assert len(np.unique(all_s, axis=0)) == len(all_s)
    ok, why = po.compare_term_sets(all_s, all_c, ref_s, ref_c, scale=np.abs(a_c).max() * np.abs(b_c).max())
    assert ok, why
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
