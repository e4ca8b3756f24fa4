"""Multi-GPU paths (one process per GPU, torch.distributed over NCCL/NVLink) — SURVEY.md §8e.

* product + cleanup: the larger operand's rows are sharded by term blocks and all-gathered once
  (25.6 MB at C5). Default path ("owner"): the owner of a row is a GF(2)-linear function of it, so
  owner(A[p]^B[q]) = class(A[p]) ^ class(B[q]) and every rank generates exactly the cross terms it
  owns — duplicates meet on one rank and nothing but the operands ever crosses NVLink.
  Alternative ("alltoall"): every rank generates the 8-byte dedup records of its block, routes each
  record to the rank that owns its hash range with one variable-size all-to-all, and the owner
  dedups locally (8 B per cross term on NVLink instead of a 272 B row).
  Either way the owner rebuilds rows from its replicas of the operands and the result stays
  hash-partitioned across ranks.
* cleanup / rotations of a term-sharded operator: terms travel to the owner of their row with one
  variable-size all-to-all of rows (`sharded_cleanup`); Clifford rotations exchange nothing.
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


def all_gather_operator(xz: torch.Tensor, c: torch.Tensor, group=None, sizes: Optional[List[int]] = None):
    """All-gather an operator whose rows are sharded by blocks: (xz_full, c_full, offsets). One size
    exchange for both tensors (skipped when the caller already knows every rank's row count: `sizes`);
    equal blocks (the common case) gather straight into the result."""
    rank, world = _world(group)
    if world == 1:
        return xz, c, [0, xz.shape[0]]
    if sizes is None:
        n_local = torch.tensor([xz.shape[0]], dtype=torch.int64, device=xz.device)
        sizes_t = torch.empty(world, dtype=torch.int64, device=xz.device)
        dist.all_gather_into_tensor(sizes_t, n_local, group=group)
        sizes = [int(v) for v in sizes_t.cpu().tolist()]
    assert len(sizes) == world and sizes[rank] == xz.shape[0], "row counts of the ranks do not match the local block"
    offsets = [0]
    for sz in sizes:
        offsets.append(offsets[-1] + sz)
    if min(sizes) == max(sizes):
        xz_full = torch.empty((offsets[-1],) + tuple(xz.shape[1:]), dtype=xz.dtype, device=xz.device)
        c_real = torch.view_as_real(c.contiguous())
        c_full = torch.empty((offsets[-1], 2), dtype=c_real.dtype, device=c.device)
        dist.all_gather_into_tensor(xz_full, xz.contiguous(), group=group)
        dist.all_gather_into_tensor(c_full, c_real, group=group)
        return xz_full, torch.view_as_complex(c_full), offsets
    xz_full, _ = all_gather_rows(xz, group)
    c_full, _ = all_gather_rows(c, group)
    return xz_full, c_full, offsets


def owner_blocks(a_counts: List[int], b_counts: List[int], owner: int) -> List[Tuple[int, int, int, int]]:
    """Row blocks (p0, p1, q0, q1) of the class-grouped operands whose cross terms `owner` owns:
    class a of A against class a ^ owner of B (owner(A[p] ^ B[q]) = class(A[p]) ^ class(B[q]))."""
    parts = len(a_counts)
    assert len(b_counts) == parts and 0 <= owner < parts
    a_off, b_off = [0], [0]
    for n in a_counts:
        a_off.append(a_off[-1] + int(n))
    for n in b_counts:
        b_off.append(b_off[-1] + int(n))
    return [(a_off[a], a_off[a + 1], b_off[a ^ owner], b_off[(a ^ owner) + 1]) for a in range(parts)]


def partition_by_owner(xz: torch.Tensor, c: torch.Tensor, log2_parts: int):
    """Operator grouped by owner class: (xz', c', counts as a host list). One small device->host read."""
    p_xz, p_c, _, counts = ops.class_partition(xz, c, log2_parts)
    return p_xz, p_c, [int(v) for v in counts.cpu().tolist()]


def owned_product(a_xz: torch.Tensor, a_c: torch.Tensor, b_xz: torch.Tensor, b_c: torch.Tensor, log2_parts: int,
                  owner: int, zero_threshold: Optional[float] = 1e-15, a_part=None, b_part=None):
    """The part of (A * B).cleanup() owned by `owner` out of 2**log2_parts, computed locally from the
    full operands with no exchange. a_part / b_part: results of `partition_by_owner` when the caller
    already has them (a replicated operand that does not change between calls, or the other owners'
    parts of the same product). Returns (xz, c, info); info carries the class-grouped operands and
    the blocks, so that a caller can check output rows against their operands."""
    a_p, a_cp, a_counts = a_part if a_part is not None else partition_by_owner(a_xz, a_c, log2_parts)
    b_p, b_cp, b_counts = b_part if b_part is not None else partition_by_owner(b_xz, b_c, log2_parts)
    blocks = owner_blocks(a_counts, b_counts, owner)
    out_xz, out_c, n_recs = ops.mul_blocks_cleanup(a_p, a_cp, b_p, b_cp, blocks, zero_threshold)
    return out_xz, out_c, {"cross_terms_generated": n_recs, "records_owned": n_recs, "rows_total_a": int(a_xz.shape[0]),
                           "blocks": blocks, "a_part": (a_p, a_cp, a_counts), "b_part": (b_p, b_cp, b_counts)}


def streamed_product(a_xz: torch.Tensor, a_c: torch.Tensor, b_xz: torch.Tensor, b_c: torch.Tensor, log2_parts: int,
                     consumer, zero_threshold: Optional[float] = 1e-15):
    """(A * B).cleanup() of a product whose result does not fit the device (config C5 on one GPU:
    1e9 cross terms = 272 GB): the result is generated in 2**log2_parts hash partitions, one after
    the other, and each is handed to `consumer(part_index, xz, c)` and dropped. The partitions are
    the owner classes of the multi-GPU product, so equal rows always fall into the same partition:
    every partition is final (fully merged, thresholded) when it is handed over, and their union is
    the whole result. Returns the list of what the consumer returned."""
    a_part = partition_by_owner(a_xz, a_c, log2_parts)
    b_part = partition_by_owner(b_xz, b_c, log2_parts)
    out = []
    for part in range(1 << log2_parts):
        xz, c, _ = owned_product(a_xz, a_c, b_xz, b_c, log2_parts, part, zero_threshold, a_part=a_part, b_part=b_part)
        out.append(consumer(part, xz, c))
        del xz, c
    return out


def sharded_product(a_block_xz: torch.Tensor, a_block_c: torch.Tensor, b_xz: torch.Tensor, b_c: torch.Tensor,
                    zero_threshold: Optional[float] = 1e-15, group=None, method: str = "owner",
                    block_sizes: Optional[List[int]] = None, cache: Optional[dict] = None):
    """(A * B).cleanup() with A's rows sharded over the ranks (this rank holds a_block) and B
    replicated. Returns this rank's hash partition of the result as (xz, c) device tensors plus a
    dict of sizes. Rows are unique across ranks.

    method="owner" (default): exchange-free. Ownership is a GF(2)-linear function of the row, so
      after one all-gather of A (25.6 MB at config C5) every rank generates exactly the cross
      terms it owns (sym_class_partition + sym_pair_records_blocks) — no record crosses NVLink.
    method="alltoall": every rank generates the records of its own block of A, partitions them by
      owner (top hash bits) and routes them with one variable-size all-to-all (8 B per cross term).
    block_sizes: row count of every rank's block of A when the caller knows them (saves the size exchange
    and its host synchronisation). cache: a dict the caller keeps between calls; the owner-class grouping
    of the replicated operand B is stored there and reused while B is the same tensor."""
    rank, world = _world(group)
    lg = log2_exact(world)
    a_full, a_c_full, offsets = all_gather_operator(a_block_xz, a_block_c, group, sizes=block_sizes)
    if method == "owner":
        b_part = None
        if cache is not None:
            key = ("b_part", b_xz.data_ptr(), b_c.data_ptr(), tuple(b_xz.shape), b_xz._version, b_c._version, lg)
            if cache.get("b_key") != key:
                cache["b_key"], cache["b_part"] = key, partition_by_owner(b_xz, b_c, lg)
            b_part = cache["b_part"]
        return owned_product(a_full, a_c_full, b_xz, b_c, lg, rank, zero_threshold, b_part=b_part)
    if method != "alltoall":
        raise ValueError(f"unknown method {method!r}")
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


def exchange_rows(xz_by_owner: torch.Tensor, c_by_owner: torch.Tensor, counts: torch.Tensor, group=None):
    """Variable-size all-to-all of whole terms (packed rows + coefficients) grouped by destination
    rank: the general hash-partitioned exchange, used when the rows already exist (sums, rotations).
    272 B per term at 1000 qubits cross NVLink."""
    rank, world = _world(group)
    if world == 1:
        return xz_by_owner, c_by_owner
    counts = counts.to(torch.int64)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    send = [int(v) for v in counts.cpu().tolist()]
    recv = [int(v) for v in recv_counts.cpu().tolist()]
    out_xz = torch.empty((sum(recv),) + tuple(xz_by_owner.shape[1:]), dtype=xz_by_owner.dtype, device=xz_by_owner.device)
    dist.all_to_all_single(out_xz, xz_by_owner.contiguous(), output_split_sizes=recv, input_split_sizes=send, group=group)
    c_real = torch.view_as_real(c_by_owner.contiguous())
    out_c = torch.empty((sum(recv), 2), dtype=c_real.dtype, device=c_real.device)
    dist.all_to_all_single(out_c, c_real, output_split_sizes=recv, input_split_sizes=send, group=group)
    return out_xz, torch.view_as_complex(out_c)


def sharded_cleanup(xz_local: torch.Tensor, c_local: torch.Tensor, zero_threshold: Optional[float] = 1e-15, group=None):
    """cleanup() of an operator whose terms are spread over the ranks in any way: every term goes to
    the rank that owns its row (GF(2)-linear owner class, sym_class_partition), one variable-size
    all-to-all of rows, then the local dedup. The result is hash-partitioned: rows are unique across
    ranks, and an operator that already is hash-partitioned exchanges nothing but its new rows' owners."""
    rank, world = _world(group)
    if world == 1:
        return ops.cleanup(xz_local, c_local, zero_threshold)
    lg = log2_exact(world)
    by_owner_xz, by_owner_c, _, counts = ops.class_partition(xz_local, c_local, lg)
    mine_xz, mine_c = exchange_rows(by_owner_xz, by_owner_c, counts, group)
    return ops.cleanup(mine_xz.contiguous(), mine_c.contiguous(), zero_threshold)


def sharded_rotation(xz_local: torch.Tensor, c_local: torch.Tensor, q_xz: torch.Tensor, angle: Optional[float],
                     group=None):
    """One rotation R P R^dagger, R = exp(i*angle/2*Q), of a term-sharded operator (base.py:1090-1161):
    the rotation itself is local (the generator row is replicated, 256 B); a Clifford angle is a
    relabelling of rows and needs no exchange at all (its dedup is deferred like in
    PauliwordOp.perform_rotations); a general angle creates new rows P*Q whose owners differ, so it
    ends with one hash-partitioned exchange + local dedup."""
    import math
    angle = math.pi / 2 if angle is None else float(angle)
    multiple = angle * 2 / math.pi
    int_part = round(multiple)
    if abs(int_part - multiple) <= 1e-18:
        sign = -1.0 if int_part in [2, 3] else 1.0
        return ops.rotate(xz_local, c_local, q_xz, 0.0, 0.0, 1 if int_part % 2 else 2, sign)
    xz, c = ops.rotate(xz_local, c_local, q_xz, math.cos(angle), math.sin(angle), 0)
    return sharded_cleanup(xz.contiguous(), c.contiguous(), group=group)


def sharded_commute(a_xz: torch.Tensor, b_xz: torch.Tensor, group=None):
    """Row block of commutes_termwise(A, B) owned by this rank: (bool[rows, N], row_begin). Inputs are
    replicated; no collective."""
    rank, world = _world(group)
    bounds = block_bounds(a_xz.shape[0], world)
    lo, hi = bounds[rank], bounds[rank + 1]
    return ops.commute(a_xz[lo:hi].contiguous(), b_xz), lo


def sharded_adjacency(a_xz: torch.Tensor, group=None, block_rows: int = 4096):
    """This rank's share of adjacency_matrix(A) (base.py:1054-1062), upper block triangle only: the block rows
    i with i % world == rank (cyclic, so the triangle is balanced), each as (row_begin, bool[rows, M - row_begin])
    = commutes_termwise(A[i0:i1), A[i0:]). The matrix is symmetric: the lower triangle is the mirror image and is
    never computed. Inputs are replicated; no collective."""
    rank, world = _world(group)
    M = a_xz.shape[0]
    out = []
    for k, i0 in enumerate(range(0, M, block_rows)):
        if k % world != rank:
            continue
        i1 = min(M, i0 + block_rows)
        out.append((i0, ops.commute(a_xz[i0:i1], a_xz[i0:])))
    return out


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
