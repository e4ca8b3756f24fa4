"""CPU test double of `symmer_b200.ops` — TEST INFRASTRUCTURE ONLY.

The product has no CPU path: without a CUDA device every operator raises. To exercise the
HOST logic of `symmer_b200.base` / `utils` / `independent_op` / `projection` (argument validation, caching,
index bookkeeping, the reference's control flow) on the CPU-only build box, the `-m "not gpu"` tests swap
every `ops` entry point the host layer uses for a NumPy restatement of what that kernel computes, built on
the oracle (`oracle/pauli_oracle.py`) and operating on CPU tensors in the same packed layout. Nothing under
`symmer_b200/` imports this module; the `-m gpu` tests run the same assertions against the real kernels.

    with host_double():           # or the `host_ops` fixture of tests/test_host_logic.py
        P = symmer_b200.PauliwordOp.from_list(['XX', 'ZZ'])
"""
import contextlib
import os
import sys

import numpy as np
import scipy.sparse as sps
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pauli_oracle as po  # noqa: E402

CPU = torch.device("cpu")


def _np_rows(xz):
    return np.ascontiguousarray(xz.numpy()).view(np.uint64)


def _to_xz(packed_u64):
    return torch.from_numpy(np.ascontiguousarray(packed_u64).view(np.int64).copy())


def _wide(xz):
    """bool[M, 2*64W]: the packed rows unpacked over their full padded width."""
    W = xz.shape[1] // 2
    return po.unpack_bits(_np_rows(xz), 64 * W), W


def _c(c):
    return np.ascontiguousarray(c.numpy())


def _tc(c):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(c, dtype=complex)))


def _narrow(symp_wide, W):
    """Rows unpacked over 64W qubits -> packed [M, 2W]."""
    if symp_wide.shape[0] == 0:
        return torch.zeros((0, 2 * W), dtype=torch.int64)
    return _to_xz(po.pack_bits(symp_wide))


# ------------------------------------------------------------------------------------------ layout
def device():
    return CPU


def pack(symp, n_qubits):
    symp = symp.numpy().astype(bool)
    if symp.shape[0] == 0:
        return torch.zeros((0, 2 * max(1, (int(n_qubits) + 63) // 64)), dtype=torch.int64)
    if n_qubits == 0:
        return torch.zeros((symp.shape[0], 2), dtype=torch.int64)
    return _to_xz(po.pack_bits(symp.reshape(symp.shape[0], 2 * int(n_qubits))))


def unpack(xz, n_qubits):
    return torch.from_numpy(po.unpack_bits(_np_rows(xz), int(n_qubits)))


def ycount(xz):
    rows, _ = _wide(xz)
    return torch.from_numpy(po.y_count(rows).astype(np.int32))


def sketch(xz):
    rows = _np_rows(xz)
    mult = (np.arange(1, rows.shape[1] + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    with np.errstate(over="ignore"):
        key = np.bitwise_xor.reduce(rows * mult, axis=1) if rows.shape[0] else np.zeros(0, dtype=np.uint64)
    return torch.from_numpy(key.view(np.int64).copy())


def gather_qubits(xz, src, n_in):
    rows, W = _wide(xz)
    src = np.asarray(src, dtype=np.int64)
    n_out = src.size
    Wo = max(1, (n_out + 63) // 64)
    out = np.zeros((rows.shape[0], 2 * 64 * Wo), dtype=bool)
    live = np.flatnonzero(src >= 0)
    out[:, live] = rows[:, src[live]]
    out[:, 64 * Wo + live] = rows[:, 64 * W + src[live]]
    return _narrow(out, Wo)


# ------------------------------------------------------------------------- product, cleanup, commute
def _thr(zero_threshold):
    return None if zero_threshold is None else float(zero_threshold)


def cleanup(xz, c, zero_threshold=1e-15):
    rows, W = _wide(xz)
    s, cc = po.symplectic_cleanup(rows, _c(c), _thr(zero_threshold))
    return _narrow(s, W), _tc(cc)


def mul_cleanup(a_xz, a_c, b_xz, b_c, zero_threshold=1e-15):
    a, W = _wide(a_xz)
    b, _ = _wide(b_xz)
    s, cc = po.multiply_by_operator(a, _c(a_c), b, _c(b_c), _thr(zero_threshold))
    return _narrow(s, W), _tc(cc)


def cross_mul(a_xz, a_c, b_xz, b_c):
    a, W = _wide(a_xz)
    b, _ = _wide(b_xz)
    s, cc = po.cross_terms(a, _c(a_c), b, _c(b_c))
    return _narrow(s, W), _tc(cc)


def commute(a_xz, b_xz):
    a, _ = _wide(a_xz)
    b, _ = _wide(b_xz)
    return torch.from_numpy(po.commutes_termwise(a, b))


def commute_self(a_xz, row_begin=0, row_end=None, block_rows=None):
    return commute(a_xz, a_xz)


def gather_rows(xz, c, perm):
    idx = perm.to(torch.int64)
    return xz[idx].contiguous(), (c[idx].contiguous() if c is not None else None)


def lex_order(xz):
    """np.lexsort over the unpacked columns (last column = primary key), like base.py:469-470."""
    rows, _ = _wide(xz)
    order = np.lexsort(rows.T) if rows.shape[0] else np.zeros(0, dtype=np.int64)
    return torch.from_numpy(order.astype(np.int32))


def join_rows(keys_l, rows_l, keys_r, rows_r):
    table = {r.tobytes(): i for i, r in reversed(list(enumerate(rows_r.numpy())))}
    return torch.tensor([table.get(r.tobytes(), -1) for r in rows_l.numpy()], dtype=torch.int32)


def commute_qwc(a_xz, b_xz):
    a, _ = _wide(a_xz)
    b, _ = _wide(b_xz)
    return torch.from_numpy(po.qubitwise_commutes_termwise(a, b))


def rotate(xz, c, q_xz, cos_a, sin_a, mode, sign=1.0, padded_ok=False):
    """The contract of sym_rotate (include/symmer_b200.h): no dedup; mode 0 keeps row i in slot i and appends
    -i sin P Q for the anticommuting rows; modes 1 / 2 are the Clifford relabels. (padded_ok only allows the
    device library a different intermediate layout; the compact form is always a valid answer.)"""
    rows, W = _wide(xz)
    q, _ = _wide(q_xz.reshape(1, -1))
    cc = _c(c).copy()
    ac = ~po.commutes_termwise(rows, q).reshape(-1)
    pq_s, pq_c = po.cross_terms(rows[ac], cc[ac], q, np.ones(1, dtype=complex))
    if mode == 0:
        out_s = np.vstack([rows, pq_s])
        head = cc.copy()
        head[ac] *= cos_a
        return _narrow(out_s, W), _tc(np.hstack([head, pq_c * (-1j * sin_a)]))
    out_s = rows.copy()
    if mode == 1:
        out_s[ac] = pq_s
        cc[ac] = pq_c * (-1j) * sign
    else:
        cc[ac] = cc[ac] * sign
    return _narrow(out_s, W), _tc(cc)


def rotate_dedup(xz, c, q_xz, cos_a, sin_a, zero_threshold=1e-15):
    r_xz, r_c = rotate(xz, c, q_xz, cos_a, sin_a, 0)
    return cleanup(r_xz, r_c, zero_threshold)


def project(xz, c, n_qubits, stab_cols, stab_eigs, free_qubits):
    rows = po.unpack_bits(_np_rows(xz), int(n_qubits))
    n = int(n_qubits)
    cc = _c(c).copy()
    keep = np.ones(rows.shape[0], dtype=bool)
    for col, eig in zip(np.asarray(stab_cols).tolist(), np.asarray(stab_eigs).tolist()):
        q = col if col < n else col - n
        xb, zb = rows[:, q], rows[:, n + q]
        keep &= ~(zb if col < n else xb)
        cc = np.where(xb if col < n else zb, cc * eig, cc)
    free = np.asarray(free_qubits, dtype=np.int64)
    sub = np.hstack([rows[:, free], rows[:, n + free]])[keep]
    Wo = max(1, (free.size + 63) // 64)
    if sub.shape[0] == 0:
        return torch.zeros((0, 2 * Wo), dtype=torch.int64), _tc(np.zeros(0))
    if free.size == 0:
        return torch.zeros((sub.shape[0], 2), dtype=torch.int64), _tc(cc[keep])
    return _to_xz(po.pack_bits(sub)), _tc(cc[keep])


# ------------------------------------------------------------------------------------- matrix-free
def _masks(xz, c, n):
    rows = po.unpack_bits(_np_rows(xz), n)
    weights = (1 << np.arange(n - 1, -1, -1)).astype(np.int64)         # qubit 0 = most significant bit
    xm = rows[:, :n].astype(np.int64) @ weights
    zm = rows[:, n:].astype(np.int64) @ weights
    cp = _c(c) * (-1j) ** (po.y_count(rows) % 4)
    return xm, zm, cp


def unsorted_masks(op):
    xm, zm, cp = _masks(op._xz, op._coeff_dev(), op.n_qubits)
    return torch.from_numpy(xm), torch.from_numpy(zm), _tc(cp)


def term_masks_sorted(xz, c, n_qubits):
    xm, zm, cp = _masks(xz, c, int(n_qubits))
    order = np.argsort(xm, kind="stable")
    return torch.from_numpy(xm[order]), torch.from_numpy(zm[order]), _tc(cp[order])


def _parity(v):
    v = v.copy()
    for s in (32, 16, 8, 4, 2, 1):
        v ^= v >> s
    return v & 1


def _dense_rows(xm, zm, cp, n):
    side = 1 << int(n)
    r = np.arange(side, dtype=np.int64)
    xm, zm, cp = xm.numpy(), zm.numpy(), _c(cp)
    rows, cols, vals = [], [], []
    for x, z, cc in zip(xm, zm, cp):
        rows.append(r)
        cols.append(r ^ x)
        vals.append(cc * (1 - 2 * _parity(r & z)))
    return side, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)


def to_csr(xm, zm, cp, n_qubits):
    """Every row gets one entry per distinct x mask, sorted by column, explicit zeros kept (sym_to_csr)."""
    side, rows, cols, vals = _dense_rows(xm, zm, cp, n_qubits)
    groups = np.unique(xm.numpy())
    key = rows * side + cols
    order = np.argsort(key, kind="stable")
    key, vals = key[order], vals[order]
    first = np.flatnonzero(np.r_[True, key[1:] != key[:-1]])
    data = np.add.reduceat(vals, first)
    indices = key[first] % side
    indptr = np.arange(side + 1, dtype=np.int64) * len(groups)
    return _tc(data), torch.from_numpy(indices), torch.from_numpy(indptr)


def apply_dense(xm, zm, cp, n_qubits, psi, row_begin=0, row_end=None):
    side, rows, cols, vals = _dense_rows(xm, zm, cp, n_qubits)
    mat = sps.csr_matrix((vals, (rows, cols)), shape=(side, side))
    y = mat @ psi.numpy()
    return _tc(y[row_begin:side if row_end is None else row_end])


def expval_dense(xm, zm, cp, n_qubits, psi, row_begin=0, row_end=None):
    y = apply_dense(xm, zm, cp, n_qubits, psi).numpy()
    p = psi.numpy()
    side = p.size
    sl = slice(row_begin, side if row_end is None else row_end)
    return torch.tensor(complex(np.vdot(p[sl], y[sl])), dtype=torch.complex128)


def _wht(table):
    """Walsh-Hadamard transform along the last axis (natural order), scaled by 1/len."""
    t = np.array(table, dtype=complex)
    side = t.shape[-1]
    h = 1
    while h < side:
        t = t.reshape(t.shape[0], -1, 2, h)
        t = np.stack([t[:, :, 0] + t[:, :, 1], t[:, :, 0] - t[:, :, 1]], axis=2)
        h *= 2
    return t.reshape(t.shape[0], side) / side


def pauli_decompose_dense(matrix, n_qubits):
    m = matrix.numpy()
    side = 1 << int(n_qubits)
    r = np.arange(side)
    diag = np.stack([m[r, r ^ x] for x in range(side)])
    return _tc(_wht(diag))


def pauli_decompose_diagonals(diag, n_qubits):
    diag.copy_(_tc(_wht(diag.numpy())))
    return diag


def rows_from_masks(xm, zm, cp, n_qubits):
    n = int(n_qubits)
    shifts = np.arange(n - 1, -1, -1)
    xb = ((xm.numpy()[:, None] >> shifts) & 1).astype(bool)
    zb = ((zm.numpy()[:, None] >> shifts) & 1).astype(bool)
    c = _c(cp) * (1j) ** (np.sum(xb & zb, axis=1) % 4)
    if xb.shape[0] == 0:
        return torch.zeros((0, 2), dtype=torch.int64), _tc(c)
    return _to_xz(po.pack_bits(np.hstack([xb, zb]))), _tc(c)


# ------------------------------------------------------------------------------------------- GF(2)
def pack_matrix(m):
    m = m.numpy().astype(bool)
    R, C = m.shape
    Cw = max(1, (C + 63) // 64)
    pad = np.zeros((R, Cw * 64), dtype=np.uint8)
    pad[:, :C] = m
    return torch.from_numpy(np.packbits(pad, axis=1, bitorder="little").view("<u8").view(np.int64).copy())


def unpack_matrix(bits, C):
    by = _np_rows(bits).view(np.uint8)
    return torch.from_numpy(np.unpackbits(by, axis=1, bitorder="little")[:, :int(C)].astype(bool))


def rref_packed(bits, C):
    """In-place row reduction with the reference's row-driven pivot rule (utils.py:292-315); returns the pivot
    column of every row (-1 for a zero row), like sym_rref."""
    m = unpack_matrix(bits, bits.shape[1] * 64).numpy()
    R = m.shape[0]
    piv = np.full(R, -1, dtype=np.int32)
    for i in range(R):
        cols = np.flatnonzero(m[i, :int(C)])
        if cols.size == 0:
            continue
        piv[i] = cols[0]
        others = np.flatnonzero(m[:, cols[0]])
        others = others[others != i]
        m[others] ^= m[i]
    bits.copy_(pack_matrix(torch.from_numpy(m)))
    return torch.from_numpy(piv)


def bit_transpose(bits):
    R = bits.shape[0]
    m = unpack_matrix(bits, bits.shape[1] * 64).numpy()
    return pack_matrix(torch.from_numpy(np.ascontiguousarray(m.T)))


def or_rows(bits, rows=None):
    sel = _np_rows(bits) if rows is None else _np_rows(bits)[rows.numpy().astype(np.int64)]
    out = np.bitwise_or.reduce(sel, axis=0) if sel.shape[0] else np.zeros(bits.shape[1], dtype=np.uint64)
    return torch.from_numpy(out.view(np.int64).copy())


_SWAPPED = ["device", "pack", "unpack", "ycount", "sketch", "gather_qubits", "cleanup", "mul_cleanup", "cross_mul",
            "commute", "commute_self", "gather_rows", "lex_order", "join_rows", "commute_qwc", "rotate", "rotate_dedup", "project", "term_masks_sorted", "to_csr", "apply_dense",
            "expval_dense", "pauli_decompose_dense", "pauli_decompose_diagonals", "rows_from_masks", "pack_matrix", "unpack_matrix", "rref_packed", "bit_transpose", "or_rows"]


@contextlib.contextmanager
def host_double():
    """Swap the kernels behind the host layer for their NumPy restatements (CPU tensors) inside the block."""
    from symmer_b200 import base, ops
    saved = {name: getattr(ops, name) for name in _SWAPPED}
    saved_masks = base._unsorted_masks
    here = sys.modules[__name__]
    try:
        for name in _SWAPPED:
            setattr(ops, name, getattr(here, name))
        base._unsorted_masks = unsorted_masks
        yield
    finally:
        for name, fn in saved.items():
            setattr(ops, name, fn)
        base._unsorted_masks = saved_masks
