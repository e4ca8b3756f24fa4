"""Multi-GPU paths (one process per GPU, torch.distributed over NCCL/NVLink) — SURVEY.md §8e.

* product + cleanup: the larger operand's rows are sharded by term blocks; every rank generates the
  8-byte dedup records of its block, routes each record to the rank that owns its hash range with
  one variable-size all-to-all, and the owner dedups locally and rebuilds rows from its replicas of
  the (small) operands. Only records cross NVLink (8 B per cross term instead of a 272 B row).
  The result stays hash-partitioned across ranks.
* commute / adjacency: row blocks of the output, no collective.
* expval: the 2^n basis is sharded by row range, one all-reduce of a complex scalar at the end.

The collective plumbing (`exchange_records`, `all_gather_rows`) only touches torch tensors and also
runs on CPU tensors with the gloo backend (covered by world_size-2 tests); the kernels are CUDA only.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


def _world(group=None) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def log2_exact(n: int) -> int:
    lg = n.bit_length() - 1
    if n < 1 or (1 << lg) != n:
        raise ValueError(f"the hash partition needs a power-of-two number of ranks, got {n}")
    return lg


def block_bounds(n_rows: int, world: int) -> List[int]:
    """Row block [bounds[r], bounds[r+1]) of rank r: contiguous, sizes differ by at most one."""
    base, rem = divmod(n_rows, world)
    out = [0]
    for r in range(world):
        out.append(out[-1] + base + (1 if r < rem else 0))
    return out


def all_gather_rows(local: torch.Tensor, group=None) -> Tuple[torch.Tensor, List[int]]:
    """Concatenate per-rank row blocks (variable row counts) on every rank; returns (full, offsets)."""
    rank, world = _world(group)
    if world == 1:
        return local, [0, local.shape[0]]
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    # equal-size gather of blocks padded to the largest one (uneven all_gather is not portable)
    pad = max(sizes)
    mine = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    mine[:local.shape[0]] = local
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    offsets = [0]
    for s in sizes:
        offsets.append(offsets[-1] + s)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0), offsets


def exchange_records(part: torch.Tensor, counts: torch.Tensor, group=None) -> torch.Tensor:
    """Variable-size all-to-all: `part` holds this rank's records grouped by destination rank,
    `counts[d]` of them for rank d. Returns the records this rank owns (grouped by source rank)."""
    rank, world = _world(group)
    if world == 1:
        return part
    counts = counts.to(torch.int64)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    send = [int(x) for x in counts.cpu().tolist()]
    recv = [int(x) for x in recv_counts.cpu().tolist()]
    out = torch.empty(sum(recv), dtype=part.dtype, device=part.device)
    dist.all_to_all_single(out, part.contiguous(), output_split_sizes=recv, input_split_sizes=send, group=group)
    return out


def sharded_product(a_block_xz: torch.Tensor, a_block_c: torch.Tensor, b_xz: torch.Tensor, b_c: torch.Tensor,
                    zero_threshold: Optional[float] = 1e-15, group=None):
    """(A * B).cleanup() with A's rows sharded over the ranks (this rank holds a_block) and B
    replicated. Returns this rank's hash partition of the result as (xz, c) device tensors plus a
    dict of sizes. Rows are unique across ranks."""
    rank, world = _world(group)
    lg = log2_exact(world)
    a_full, offsets = all_gather_rows(a_block_xz, group)
    a_c_full, _ = all_gather_rows(a_block_c, group)
    recs = ops.pair_records(a_full, offsets[rank], offsets[rank + 1], b_xz)
    if world > 1:
        part, counts = ops.partition_records(recs, lg)
        del recs
        mine = exchange_records(part, counts, group)
        del part
    else:
        mine = recs
    out_xz, out_c = ops.dedup_records(mine, a_full, a_c_full, b_xz, b_c, zero_threshold)
    info = {"cross_terms_generated": int((offsets[rank + 1] - offsets[rank]) * b_xz.shape[0]),
            "records_owned": int(mine.numel()), "rows_total_a": int(a_full.shape[0])}
    return out_xz, out_c, info


def sharded_commute(a_xz: torch.Tensor, b_xz: torch.Tensor, group=None):
    """Row block of commutes_termwise(A, B) owned by this rank: (bool[rows, N], row_begin). Inputs are
    replicated; no collective."""
    rank, world = _world(group)
    bounds = block_bounds(a_xz.shape[0], world)
    lo, hi = bounds[rank], bounds[rank + 1]
    return ops.commute(a_xz[lo:hi].contiguous(), b_xz), lo


def sharded_expval(xm: torch.Tensor, zm: torch.Tensor, cp: torch.Tensor, n_qubits: int, psi: torch.Tensor,
                   group=None) -> complex:
    """<psi|H|psi> with the 2^n basis rows sharded over the ranks (psi and the terms replicated) and
    one all-reduce of the complex partial sums."""
    rank, world = _world(group)
    bounds = block_bounds(1 << n_qubits, world)
    partial = ops.expval_dense(xm, zm, cp, n_qubits, psi, bounds[rank], bounds[rank + 1])
    if world > 1:
        buf = torch.view_as_real(partial.reshape(1)).clone()
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        partial = torch.view_as_complex(buf)[0]
    return complex(partial.cpu().numpy())
