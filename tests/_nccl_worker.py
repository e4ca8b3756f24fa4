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
