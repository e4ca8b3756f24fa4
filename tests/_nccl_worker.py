"""2+-GPU NCCL worker for tests/test_gpu_dist.py: sharded product / commute / expval against the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import pauli_oracle as po  # noqa: E402
from symmer_b200 import dist as sdist  # noqa: E402
from symmer_b200 import ops  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = ops.device()
    dist.init_process_group("nccl", device_id=dev)
    for n, M, N, dup in [(70, 96, 41, True), (1000, 300, 120, False), (5, 64, 64, True)]:
        a_s, a_c = po.random_operator(n, M, seed=31)
        b_s, b_c = po.random_operator(n, N, seed=32)
        if dup:
            b_s[:20] = a_s[:20]
        bounds = sdist.block_bounds(M, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        a_blk = ops.pack(torch.from_numpy(a_s[lo:hi].copy()), n)
        a_blk_c = torch.from_numpy(a_c[lo:hi].copy()).to(dev)
        b = ops.pack(torch.from_numpy(b_s), n)
        bc = torch.from_numpy(b_c).to(dev)
        for method in ("owner", "alltoall"):
            xz, c, info = sdist.sharded_product(a_blk, a_blk_c, b, bc, method=method)
            assert info["rows_total_a"] == M
            loc = (ops.unpack(xz, n).cpu().numpy(), c.cpu().numpy(), info["cross_terms_generated"])
            gathered = [None] * world
            dist.all_gather_object(gathered, loc)
            if rank == 0:
                s = np.vstack([g[0] for g in gathered])
                cc = np.hstack([g[1] for g in gathered])
                assert sum(g[2] for g in gathered) == M * N, method
                assert len(np.unique(s, axis=0)) == len(s), "owners overlap"
                ref_s, ref_c = po.multiply_by_operator(a_s, a_c, b_s, b_c)
                ok, why = po.compare_term_sets(s, cc, ref_s, ref_c, scale=np.abs(a_c).max() * np.abs(b_c).max())
                assert ok, (method, why)
        # term-sharded cleanup and rotations: every rank holds an arbitrary slice of a list with duplicates
        dup_s = np.vstack([a_s, a_s[::3], b_s[:N // 2]]) if a_s.shape[1] == b_s.shape[1] else a_s
        dup_c = np.hstack([a_c, -a_c[::3] * (np.arange(len(a_c[::3])) % 2), b_c[:N // 2]])
        sl = slice(rank, None, world)
        loc_xz = ops.pack(torch.from_numpy(dup_s[sl].copy()), n)
        loc_c = torch.from_numpy(dup_c[sl].copy()).to(dev)
        cxz, cc_ = sdist.sharded_cleanup(loc_xz, loc_c)
        q_row = b_s[0]
        q_xz = ops.pack(torch.from_numpy(q_row.reshape(1, -1).copy()), n)
        rxz, rc = sdist.sharded_rotation(cxz, cc_, q_xz, 0.37)
        kxz, kc = sdist.sharded_rotation(cxz, cc_, q_xz, None)
        parts = [None] * world
        dist.all_gather_object(parts, tuple(t.cpu().numpy() for t in (ops.unpack(cxz, n), cc_, ops.unpack(rxz, n), rc,
                                                                    ops.unpack(kxz, n), kc)))
        if rank == 0:
            cs, ccs = np.vstack([p_[0] for p_ in parts]), np.hstack([p_[1] for p_ in parts])
            assert len(np.unique(cs, axis=0)) == len(cs), "cleanup owners overlap"
            ref_s, ref_c = po.cleanup(dup_s, dup_c)
            ok, why = po.compare_term_sets(cs, ccs, ref_s, ref_c, scale=np.abs(dup_c).max())
            assert ok, ("sharded_cleanup", why)
            rs, rcs = np.vstack([p_[2] for p_ in parts]), np.hstack([p_[3] for p_ in parts])
            assert len(np.unique(rs, axis=0)) == len(rs), "rotation owners overlap"
            ref_rs, ref_rc = po.perform_rotations(ref_s, ref_c, [(q_row, 0.37)])
            ok, why = po.compare_term_sets(rs, rcs, ref_rs, ref_rc, scale=np.abs(dup_c).max())
            assert ok, ("sharded_rotation", why)
            ks, kcs = np.vstack([p_[4] for p_ in parts]), np.hstack([p_[5] for p_ in parts])
            ref_ks, ref_kc = po.perform_rotations(ref_s, ref_c, [(q_row, None)])
            ok, why = po.compare_term_sets(ks, kcs, ref_ks, ref_kc, scale=np.abs(dup_c).max())
            assert ok, ("sharded Clifford rotation", why)
        # commute row blocks
        a_full = ops.pack(torch.from_numpy(a_s), n)
        blk, row0 = sdist.sharded_commute(a_full, b)
        ref = po.commutes_termwise(a_s, b_s)
        assert np.array_equal(blk.cpu().numpy(), ref[row0:row0 + blk.shape[0]])
    # basis-sharded expval
    n = 10
    h_s, h_c = po.random_operator(n, 200, seed=33)
    rng = np.random.default_rng(0)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    xm, zm, cp = ops.term_masks_sorted(ops.pack(torch.from_numpy(h_s), n), torch.from_numpy(h_c).to(dev), n)
    e = sdist.sharded_expval(xm, zm, cp, n, torch.from_numpy(psi).to(dev))
    assert np.isclose(e, po.expval_dense(h_s, h_c, psi), rtol=1e-12)
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
