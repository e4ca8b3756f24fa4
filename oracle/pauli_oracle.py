"""CPU oracle for the symplectic Pauli-algebra hot path — TEST INFRASTRUCTURE, not product code.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import this module. The product path (`symmer_b200/`) never does: it fails loudly without its CUDA
library.

Every function restates, in plain NumPy on unpacked `bool[M, 2n]` symplectic matrices (the
reference's own data layout), the algorithm of one reference function and cites it. Paths are
relative to the reference tree (UCL-CCS/symmer).

Parity status: PINNED. `tests/golden/make_golden.py` imports the real reference (through
`oracle/shim`, which only replaces uninstalled third-party packages) in the build container, runs the
reference's own golden cases plus seeded random cases, and commits inputs+outputs under
`tests/golden/`; `tests/test_oracle_golden.py` checks this module against those vectors.

Third-party arithmetic restated here because the dependency is absent from the reference tree:
qiskit 1.2.4 (`poetry.lock`), module `qiskit._accelerate.sparse_pauli_op`:
  * `unordered_unique`  -> `unordered_unique` below (call site utils.py:271)
  * `to_matrix_sparse`  -> `zx_to_matrix_sparse` below (call site base.py:1500-1510)
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sps

_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB = None


def _clib():
    """Optional C helpers (oracle/oracle.c -> oracle/_build/liboracle.so). Still oracle code."""
    global _CLIB
    if _CLIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.orc_unordered_unique.restype = ctypes.c_int64
            lib.orc_unordered_unique.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                                 ctypes.c_void_p, ctypes.c_void_p]
            lib.orc_to_matrix_sparse.restype = ctypes.c_int64
            lib.orc_to_matrix_sparse.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
            lib.orc_add_at.restype = None
            lib.orc_add_at.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
            _CLIB = lib
        else:
            _CLIB = False
    return _CLIB or None


# ----------------------------------------------------------------------------------------------
# helpers shared with the tests (canonical order, packing used only for comparisons)
# ----------------------------------------------------------------------------------------------

def pack_bits(symp: np.ndarray) -> np.ndarray:
    """bool[M, 2n] -> uint64[M, 2W] with W = ceil(n/64); bit q of a block sits in word q//64 at
    position q%64. This is the product's device layout, restated independently so
    tests can compare packed buffers."""
    symp = np.asarray(symp, dtype=bool)
    M, two_n = symp.shape
    n = two_n // 2
    W = max(1, (n + 63) // 64)
    out = np.zeros((M, 2 * W), dtype=np.uint64)
    for blk in range(2):
        bits = symp[:, blk * n:(blk + 1) * n]
        pad = np.zeros((M, W * 64), dtype=np.uint8)
        pad[:, :n] = bits
        # little-endian bit order inside each byte, little-endian bytes inside each word
        by = np.packbits(pad, axis=1, bitorder="little")
        out[:, blk * W:(blk + 1) * W] = by.view("<u8")
    return out


def unpack_bits(xz: np.ndarray, n: int) -> np.ndarray:
    xz = np.ascontiguousarray(xz, dtype=np.uint64)
    M = xz.shape[0]
    W = xz.shape[1] // 2
    out = np.zeros((M, 2 * n), dtype=bool)
    for blk in range(2):
        by = xz[:, blk * W:(blk + 1) * W].copy().view(np.uint8)
        bits = np.unpackbits(by, axis=1, bitorder="little")[:, :n]
        out[:, blk * n:(blk + 1) * n] = bits.astype(bool)
    return out


def canonical_order(symp: np.ndarray) -> np.ndarray:
    """Permutation putting rows in the reference's 'lex' order (`np.lexsort(symp.T)`,
    base.py:469-470: the LAST column is the primary key)."""
    symp = np.asarray(symp, dtype=bool)
    if symp.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    # lexsort over 2n keys is slow for wide matrices; an equivalent order is a byte-wise sort of
    # the rows with the columns reversed (last column most significant).
    rev = np.ascontiguousarray(symp[:, ::-1]).view(np.uint8)
    view = rev.view(np.dtype((np.void, rev.shape[1]))).ravel()
    return np.argsort(view, kind="stable")


def canonical(symp: np.ndarray, coeff: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    order = canonical_order(symp)
    return np.asarray(symp, dtype=bool)[order], np.asarray(coeff, dtype=complex)[order]


def compare_term_sets(symp_a, coeff_a, symp_b, coeff_b, scale: float = 1.0,
                      rtol: float = 1e-12) -> Tuple[bool, str]:
    """Term-set parity of SURVEY.md §8 'parity caveats' 1-4: rows with |c| > tau must match
    bit-exactly (tau = rtol*scale), coefficients within rtol (+ atol tau); rows at or below tau on
    either side are ignored."""
    tau = rtol * scale
    symp_a = np.asarray(symp_a, dtype=bool)
    symp_b = np.asarray(symp_b, dtype=bool)
    coeff_a = np.asarray(coeff_a, dtype=complex)
    coeff_b = np.asarray(coeff_b, dtype=complex)
    ka = np.abs(coeff_a) > tau
    kb = np.abs(coeff_b) > tau
    sa, ca = canonical(symp_a[ka], coeff_a[ka])
    sb, cb = canonical(symp_b[kb], coeff_b[kb])
    if sa.shape != sb.shape:
        # a row may sit just either side of tau: retry on the union of rows
        return _compare_by_dict(symp_a, coeff_a, symp_b, coeff_b, tau, rtol)
    if not np.array_equal(sa, sb):
        return _compare_by_dict(symp_a, coeff_a, symp_b, coeff_b, tau, rtol)
    if not np.allclose(ca, cb, rtol=rtol, atol=tau):
        worst = np.max(np.abs(ca - cb))
        return False, f"coefficients differ (max abs err {worst:.3e})"
    return True, "ok"


def _compare_by_dict(symp_a, coeff_a, symp_b, coeff_b, tau, rtol):
    da = {r.tobytes(): c for r, c in zip(np.asarray(symp_a, dtype=bool), coeff_a)}
    db = {r.tobytes(): c for r, c in zip(np.asarray(symp_b, dtype=bool), coeff_b)}
    for k in set(da) | set(db):
        a = da.get(k, 0.0)
        b = db.get(k, 0.0)
        if abs(a - b) > tau + rtol * max(abs(a), abs(b)):
            return False, f"term mismatch: {a} vs {b}"
    return True, "ok"


# ----------------------------------------------------------------------------------------------
# a3  Y_count                                                         base.py:604-615
# ----------------------------------------------------------------------------------------------

def y_count(symp: np.ndarray) -> np.ndarray:
    symp = np.asarray(symp, dtype=bool)
    n = symp.shape[1] // 2
    return np.count_nonzero(symp[:, :n] & symp[:, n:], axis=1).astype(np.int64)


# ----------------------------------------------------------------------------------------------
# a5  dedup + coefficient reduction
# ----------------------------------------------------------------------------------------------

def unordered_unique(arr: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """qiskit 1.2.4 `unordered_unique` (call site utils.py:271): scan the rows in order; the first
    time a row value is seen it gets the next output slot. Returns (first-occurrence row indices
    int64[U], inverse map int64[T])."""
    arr = np.ascontiguousarray(arr)
    T = arr.shape[0]
    if T == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    lib = _clib()
    if lib is not None:
        first = np.empty(T, dtype=np.int64)
        inv = np.empty(T, dtype=np.int64)
        U = lib.orc_unordered_unique(arr.ctypes.data, T, arr.strides[0], first.ctypes.data,
                                     inv.ctypes.data)
        return first[:U].copy(), inv
    row_bytes = arr.shape[1] * arr.dtype.itemsize
    view = arr.view(np.dtype((np.void, row_bytes))).ravel()
    _, first_sorted, inv_sorted = np.unique(view, return_index=True, return_inverse=True)
    # np.unique orders by value; re-label groups by order of first occurrence
    rank = np.argsort(first_sorted, kind="stable")
    relabel = np.empty_like(rank)
    relabel[rank] = np.arange(rank.size)
    return first_sorted[rank].astype(np.int64), relabel[inv_sorted.ravel()].astype(np.int64)


def symplectic_cleanup(symp: np.ndarray, coeff: np.ndarray,
                       zero_threshold: Optional[float] = None) -> Tuple[np.ndarray, np.ndarray]:
    """utils.py:230-279: unique rows in first-occurrence order, duplicates' coefficients summed in
    input order (np.add.at), then keep only |c| > zero_threshold (strict)."""
    symp = np.asarray(symp, dtype=bool)
    coeff = np.asarray(coeff, dtype=complex)
    first, inv = unordered_unique(symp.astype("uint16"))   # the reference's widening copy (:271)
    rows = symp[first]
    acc = np.zeros(first.shape[0], dtype=complex)
    lib = _clib()
    if lib is not None and coeff.size:
        c = np.ascontiguousarray(coeff)
        lib.orc_add_at(acc.ctypes.data, inv.ctypes.data, c.ctypes.data, c.size)
    else:
        np.add.at(acc, inv, coeff)
    if zero_threshold is not None:
        keep = np.abs(acc) > zero_threshold
        rows, acc = rows[keep], acc[keep]
    return rows, acc


def cleanup(symp: np.ndarray, coeff: np.ndarray, zero_threshold: float = 1e-15):
    """PauliwordOp.cleanup, base.py:617-638 (n_qubits == 0 and n_terms == 0 special cases)."""
    symp = np.asarray(symp, dtype=bool)
    coeff = np.asarray(coeff, dtype=complex)
    if symp.shape[1] == 0:
        return np.zeros((1, 0), dtype=bool), np.array([coeff.sum()], dtype=complex)
    if symp.shape[0] == 0:
        return np.zeros((1, symp.shape[1]), dtype=bool), np.zeros(1, dtype=complex)
    return symplectic_cleanup(symp, coeff, zero_threshold)


# ----------------------------------------------------------------------------------------------
# a4  all-pairs product with phases                                   base.py:764-794, 821-859
# ----------------------------------------------------------------------------------------------

def cross_terms(a_symp, a_coeff, b_symp, b_coeff) -> Tuple[np.ndarray, np.ndarray]:
    """Pre-cleanup cross terms of A*B in the reference's flattened order t = q*M + p (q indexes B,
    p indexes A). base.py:783-792."""
    a_symp = np.asarray(a_symp, dtype=bool)
    b_symp = np.asarray(b_symp, dtype=bool)
    a_coeff = np.asarray(a_coeff, dtype=complex)
    b_coeff = np.asarray(b_coeff, dtype=complex)
    M, two_n = a_symp.shape
    N = b_symp.shape[0]
    n = two_n // 2
    assert b_symp.shape[1] == two_n, "PauliwordOps defined for different number of qubits"
    b3 = b_symp.reshape(N, 1, two_n)
    prod = a_symp ^ b3                                                   # :784
    y_in = y_count(a_symp) + y_count(b_symp).reshape(-1, 1)             # :785
    y_out = np.sum(prod[:, :, :n] & prod[:, :, n:], axis=2)              # :786
    sign = (-1) ** (np.sum(a_symp[:, :n] & b3[:, :, n:], axis=2) % 2)    # :787
    phase = sign * (1j) ** ((3 * y_in + y_out) % 4)                      # :788
    coeff = (phase * np.outer(a_coeff, b_coeff).T).reshape(-1)           # :792
    return prod.reshape(-1, two_n), coeff


def multiply_by_operator(a_symp, a_coeff, b_symp, b_coeff, zero_threshold: float = 1e-15):
    rows, coeff = cross_terms(a_symp, a_coeff, b_symp, b_coeff)
    return symplectic_cleanup(rows, coeff, zero_threshold)


def multiply(a_symp, a_coeff, b_symp, b_coeff, zero_threshold: float = 1e-15):
    """PauliwordOp.__mul__ for two operators, including the dagger swap that makes the smaller
    operand the outer loop (base.py:846-851)."""
    a_coeff = np.asarray(a_coeff, dtype=complex)
    b_coeff = np.asarray(b_coeff, dtype=complex)
    if np.asarray(a_symp).shape[0] < np.asarray(b_symp).shape[0]:
        rows, c = multiply_by_operator(b_symp, b_coeff.conjugate(), a_symp, a_coeff.conjugate(),
                                       zero_threshold)
        return rows, c.conjugate()
    return multiply_by_operator(a_symp, a_coeff, b_symp, b_coeff, zero_threshold)


# ----------------------------------------------------------------------------------------------
# a7  commutes_termwise / adjacency                                   base.py:938-971, utils.py:9-78
# ----------------------------------------------------------------------------------------------

def matmul_gf2(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """utils.py:63-78: float64 GEMM as a stand-in for GF(2), then mod 2."""
    return np.asarray(np.dot(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)) % 2,
                      dtype=bool)


def commutes_termwise(a_symp, b_symp) -> np.ndarray:
    """True where A[i] commutes with B[j]; shape (M, N). base.py:971."""
    a_symp = np.asarray(a_symp, dtype=bool)
    b_symp = np.asarray(b_symp, dtype=bool)
    n = a_symp.shape[1] // 2
    assert b_symp.shape[1] == 2 * n, "Pauliwords defined for different number of qubits"
    omega_b = np.hstack((b_symp[:, n:], b_symp[:, :n])).T
    return ~matmul_gf2(a_symp, omega_b)


def qubitwise_commutes_termwise(a_symp, b_symp) -> np.ndarray:
    """True where A[i] and B[j] commute qubit by qubit; shape (M, N). base.py:985-1009: on the qubits where
    both terms are non-trivial (`non_I`) the X bits and the Z bits must agree."""
    a_symp = np.asarray(a_symp, dtype=bool)
    b_symp = np.asarray(b_symp, dtype=bool)
    n = a_symp.shape[1] // 2
    assert b_symp.shape[1] == 2 * n, "Pauliwords defined for different number of qubits"
    ax, az = a_symp[:, :n], a_symp[:, n:]
    columns = []
    for x_term, z_term in zip(b_symp[:, :n], b_symp[:, n:]):
        non_i = (ax | az) & (x_term | z_term)                            # :1001
        x_match = np.all((ax & non_i) == (x_term & non_i), axis=1)       # :1003
        z_match = np.all((az & non_i) == (z_term & non_i), axis=1)       # :1004
        columns.append((x_match & z_match).reshape(-1, 1))               # :1006
    if not columns:
        return np.zeros((a_symp.shape[0], 0), dtype=bool)
    return np.hstack(columns)


def reindex(symp, qubit_map) -> np.ndarray:
    """PauliwordOp.reindex, base.py:493-521: `new[:, old_indices] = old[:, new_indices]` on both blocks."""
    symp = np.asarray(symp, dtype=bool)
    n = symp.shape[1] // 2
    if isinstance(qubit_map, list):
        old_indices, new_indices = sorted(qubit_map), qubit_map
    else:
        old_indices, new_indices = zip(*qubit_map.items())
    old_indices, new_indices = list(old_indices), list(new_indices)
    assert len(new_indices) == len(set(new_indices)), 'Duplicated index'
    assert not set(old_indices).difference(new_indices), 'Assignment conflict'
    x, z = symp[:, :n].copy(), symp[:, n:].copy()
    x[:, old_indices] = symp[:, :n][:, new_indices]
    z[:, old_indices] = symp[:, n:][:, new_indices]
    return np.hstack([x, z])


def tensor(a_symp, a_coeff, b_symp, b_coeff):
    """PauliwordOp.tensor, base.py:1188-1204: pad both factors with identities, then the ordinary product."""
    a_symp = np.asarray(a_symp, dtype=bool)
    b_symp = np.asarray(b_symp, dtype=bool)
    nl, nr = a_symp.shape[1] // 2, b_symp.shape[1] // 2
    pad_l = np.zeros((a_symp.shape[0], nr), dtype=bool)
    pad_r = np.zeros((b_symp.shape[0], nl), dtype=bool)
    left = np.hstack([a_symp[:, :nl], pad_l, a_symp[:, nl:], pad_l])
    right = np.hstack([pad_r, b_symp[:, :nr], pad_r, b_symp[:, nr:]])
    return multiply(left, a_coeff, right, b_coeff)


# ----------------------------------------------------------------------------------------------
# a8  rotations                                                       base.py:1090-1186
# ----------------------------------------------------------------------------------------------

def rotate_by_single_pword(symp, coeff, q_symp, angle=None, threshold: float = 1e-18):
    """R P R^dagger with R = exp(i*angle/2*Q), Q a single Pauli with coefficient 1.
    Returns (symp, coeff) WITHOUT the trailing cleanup that perform_rotations applies."""
    symp = np.asarray(symp, dtype=bool)
    coeff = np.asarray(coeff, dtype=complex)
    q_symp = np.asarray(q_symp, dtype=bool).reshape(1, -1)
    if angle is None:
        angle = np.pi / 2
    angle = complex(angle).real
    commute = commutes_termwise(symp, q_symp).reshape(-1)                # :1130
    if np.all(commute):
        return symp, coeff                                               # :1131-1133
    com_s, com_c = symp[commute], coeff[commute]
    ac_s, ac_c = symp[~commute], coeff[~commute]
    one = np.ones(1, dtype=complex)
    multiple = angle * 2 / np.pi
    int_part = round(multiple)
    if abs(int_part - multiple) <= threshold:                            # Clifford, :1141-1154
        if int_part % 2 == 0:
            part_s, part_c = ac_s, ac_c
        else:
            part_s, part_c = multiply(ac_s, ac_c, q_symp, one)
            part_c = part_c * (-1j)
        if int_part in [2, 3]:
            part_c = part_c * (-1)
        return np.vstack([part_s, com_s]), np.hstack([part_c, com_c])
    # general angle, :1159-1161 (a `*`, a `+` and a second `+`, each with its own cleanup)
    pq_s, pq_c = multiply(ac_s, ac_c, q_symp, one)
    part_s, part_c = cleanup(np.vstack([ac_s, pq_s]),
                             np.hstack([ac_c * np.cos(angle), pq_c * (-1j * np.sin(angle))]))
    return cleanup(np.vstack([com_s, part_s]), np.hstack([com_c, part_c]))


def perform_rotations(symp, coeff, rotations: Sequence[Tuple[np.ndarray, Optional[float]]]):
    """base.py:1163-1186: sequential rotations, cleanup after each."""
    symp = np.asarray(symp, dtype=bool).copy()
    coeff = np.asarray(coeff, dtype=complex).copy()
    if len(rotations) == 0:
        return cleanup(symp, coeff)
    for q_symp, angle in rotations:
        symp, coeff = rotate_by_single_pword(symp, coeff, q_symp, angle)
        symp, coeff = cleanup(symp, coeff)
    return symp, coeff


# ----------------------------------------------------------------------------------------------
# a9  sparse matrix                                                   base.py:1458-1510, utils.py:182-228
# ----------------------------------------------------------------------------------------------

def zx_to_matrix_sparse(x, z, phases, coeffs):
    """qiskit 1.2.4 `to_matrix_sparse` value semantics (call site base.py:1508): column index 0 of
    x/z is the LEAST significant bit of the basis index; entry (r, r^x) accumulates
    coeff * (-i)^((phase + wt(x&z)) mod 4) * (-1)^popcount(r & z). Returns canonical CSR arrays."""
    x = np.asarray(x, dtype=bool)
    z = np.asarray(z, dtype=bool)
    coeffs = np.asarray(coeffs, dtype=complex)
    phases = np.asarray(phases).astype(np.int64)
    M, n = x.shape
    assert n <= 63
    weights = (np.int64(1) << np.arange(n, dtype=np.int64))
    x_int = (x.astype(np.int64) * weights).sum(axis=1)
    z_int = (z.astype(np.int64) * weights).sum(axis=1)
    ny = np.count_nonzero(x & z, axis=1)
    c = coeffs * (-1j) ** ((phases + ny) % 4)
    side = 1 << n
    lib = _clib()
    if lib is not None and M > 0:
        # group equal x masks (they share a column per row), sorted so columns ascend per row? no:
        # r ^ x is not monotone in x, so the C helper sorts per row.
        ux = np.unique(x_int)
        nnz = side * ux.size
        data = np.empty(nnz, dtype=complex)
        indices = np.empty(nnz, dtype=np.int64)
        indptr = np.empty(side + 1, dtype=np.int64)
        xi = np.ascontiguousarray(x_int)
        zi = np.ascontiguousarray(z_int)
        cc = np.ascontiguousarray(c)
        got = lib.orc_to_matrix_sparse(xi.ctypes.data, zi.ctypes.data, cc.ctypes.data, M, n,
                                       data.ctypes.data, indices.ctypes.data, indptr.ctypes.data, nnz)
        assert got == nnz
        return data, indices, indptr
    rows = np.arange(side, dtype=np.int64)
    acc = sps.csr_matrix((side, side), dtype=complex)
    for t in range(M):
        par = _parity64(rows & z_int[t])
        vals = c[t] * (1 - 2 * par.astype(np.int64))
        acc = acc + sps.csr_matrix((vals, (rows, rows ^ x_int[t])), shape=(side, side))
    acc.sum_duplicates()
    acc.sort_indices()
    return acc.data, acc.indices, acc.indptr


def _parity64(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64)
    for s in (32, 16, 8, 4, 2, 1):
        v = v ^ (v >> np.uint64(s))
    return (v & np.uint64(1)).astype(np.uint8)


def to_sparse_matrix(symp, coeff) -> sps.csr_matrix:
    """PauliwordOp.to_sparse_matrix (base.py:1458-1510): qubit 0 is the MOST significant bit of the
    basis index (the reference reverses columns before calling qiskit, :1502-1503)."""
    symp = np.asarray(symp, dtype=bool)
    coeff = np.asarray(coeff, dtype=complex)
    n = symp.shape[1] // 2
    if n == 0:
        return sps.csr_matrix(coeff)
    phase = np.zeros(symp.shape[0], dtype=np.uint8)
    data, indices, indptr = zx_to_matrix_sparse(symp[:, :n][:, ::-1], symp[:, n:][:, ::-1], phase, coeff)
    side = 1 << n
    return sps.csr_matrix((data, indices, indptr), shape=(side, side))


def single_term_matrix(symp_vec, coeff) -> sps.csr_matrix:
    """In-repo single-term form, utils.py:182-228 (cross-check for the qiskit restatement)."""
    symp_vec = np.asarray(symp_vec, dtype=bool)
    n = symp_vec.size // 2
    xb, zb = symp_vec[:n], symp_vec[n:]
    phase = (-1j) ** int(np.count_nonzero(xb & zb))
    w = (np.int64(1) << np.arange(n - 1, -1, -1, dtype=np.int64))
    x_int = int((xb.astype(np.int64) * w).sum())
    z_int = int((zb.astype(np.int64) * w).sum())
    rows = np.arange(1 << n, dtype=np.int64)
    vals = phase * (1 - 2 * _parity64(rows & z_int).astype(np.int64))
    return coeff * sps.csr_matrix((vals, (rows, rows ^ x_int)), shape=(1 << n, 1 << n), dtype=complex)


def pauli_apply_dense(symp, coeff, psi: np.ndarray) -> np.ndarray:
    """Matrix-free y = (sum_t c_t P_t) psi with the index convention of `to_sparse_matrix`
    (qubit 0 = MSB). Restates the same entry formula row by row; used as the oracle of the
    matrix-free device kernels where the CSR cannot be materialised."""
    symp = np.asarray(symp, dtype=bool)
    coeff = np.asarray(coeff, dtype=complex)
    psi = np.asarray(psi, dtype=complex)
    n = symp.shape[1] // 2
    assert psi.size == 1 << n
    w = (np.int64(1) << np.arange(n - 1, -1, -1, dtype=np.int64))
    rows = np.arange(1 << n, dtype=np.int64)
    y = np.zeros_like(psi)
    for t in range(symp.shape[0]):
        xb, zb = symp[t, :n], symp[t, n:]
        x_int = int((xb.astype(np.int64) * w).sum())
        z_int = int((zb.astype(np.int64) * w).sum())
        ph = coeff[t] * (-1j) ** int(np.count_nonzero(xb & zb))
        sgn = 1 - 2 * _parity64(rows & z_int).astype(np.int64)
        y += ph * sgn * psi[rows ^ x_int]
    return y


def expval_dense(symp, coeff, psi: np.ndarray) -> complex:
    """<psi| H |psi> for a dense state vector (VQE path: variational_optimization.py:115-117 does
    state^dagger @ H.to_sparse_matrix @ state)."""
    return complex(np.vdot(psi, pauli_apply_dense(symp, coeff, psi)))


# ----------------------------------------------------------------------------------------------
# a11 GF(2) elimination                                               utils.py:292-359, 504-519
# ----------------------------------------------------------------------------------------------

def _rref_binary(matrix: np.ndarray) -> np.ndarray:
    """utils.py:292-315. Row-driven pivot rule, no row swaps: for each row i in order (reading the
    CURRENT contents of row i), pivot = its first set column; XOR row i into every other row that
    has a 1 in that column."""
    m = np.array(matrix, dtype=bool, copy=True)
    for i in range(m.shape[0]):
        row = m[i]
        nz = np.flatnonzero(row)
        if nz.size:
            p = nz[0]
            hit = np.flatnonzero(m[:, p])
            hit = hit[hit != i]
            m[hit] ^= row
    return m


def rref_binary(matrix: np.ndarray) -> np.ndarray:
    """utils.py:317-335: order rows by pivot column; all-zero rows last."""
    red = _rref_binary(matrix)
    piv = [(i, int(np.flatnonzero(r)[0])) for i, r in enumerate(red) if r.any()]
    piv.sort(key=lambda t: t[1])
    order = [i for i, _ in piv]
    seen = set(order)
    order += [i for i in range(red.shape[0]) if i not in seen]
    return red[order]


def _cref_binary(matrix: np.ndarray) -> np.ndarray:
    return _rref_binary(np.asarray(matrix).T).T                          # utils.py:337-347


def cref_binary(matrix: np.ndarray) -> np.ndarray:
    return rref_binary(np.asarray(matrix).T).T                           # utils.py:349-359


def check_independent(symp: np.ndarray) -> bool:
    """utils.py:504-519."""
    symp = np.asarray(symp, dtype=bool)
    if symp.shape[0] > symp.shape[1]:
        return False
    red = _rref_binary(symp)
    return bool(~np.any(np.all(~red, axis=1)))


def symmetry_generator_rows(symp: np.ndarray) -> np.ndarray:
    """The GF(2) part of IndependentOp.symmetry_generators, independent_op.py:123-126: column
    reduction of [[Z X],[I]] and read-out of the kernel basis S (rows = generators)."""
    symp = np.asarray(symp, dtype=bool)
    M, two_n = symp.shape
    n = two_n // 2
    stack = np.vstack([np.hstack([symp[:, n:], symp[:, :n]]), np.eye(two_n, dtype=bool)])
    red = _cref_binary(stack)
    null_cols = np.all(~red[:M], axis=0)
    return red[M:, null_cols].T


def generator_reconstruction(gen_symp: np.ndarray, op_symp: np.ndarray):
    """base.py:523-560 (without the independence assert)."""
    gen_symp = np.asarray(gen_symp, dtype=bool)
    op_symp = np.asarray(op_symp, dtype=bool)
    dim = gen_symp.shape[0]
    red = cref_binary(np.vstack([gen_symp, op_symp]))
    mask = np.all(~red[dim:, dim:], axis=1)
    return red[dim:, :dim].astype(int), mask


# ----------------------------------------------------------------------------------------------
# f2  stabilizer-subspace projection                                  projection/base.py:44-84
# ----------------------------------------------------------------------------------------------

def project_onto_stabilizers(symp, coeff, stab_symp, stab_coeff, free_qubits):
    """S3Projection._perform_projection (projection/base.py:44-84) on plain arrays: `symp` is the
    operator AFTER the stabilizer rotations, `stab_symp`/`stab_coeff` the rotated single-qubit
    stabilizers with their +/-1 eigenvalues, `free_qubits` the qubit positions that survive."""
    symp = np.asarray(symp, dtype=bool)
    coeff = np.asarray(coeff, dtype=complex)
    stab_symp = np.asarray(stab_symp, dtype=bool)
    n = symp.shape[1] // 2
    keep = np.all(commutes_termwise(symp, stab_symp), axis=1)            # :63
    kept_s, kept_c = symp[keep], coeff[keep]                             # :64-65
    idx = np.where(stab_symp)[1]                                         # :69
    eig = kept_s[:, idx] * np.asarray(stab_coeff)                        # :70
    eig[eig == 0] = 1                                                    # :71
    kept_c = kept_c * np.prod(eig, axis=1).T                             # :72
    free = np.asarray(free_qubits, dtype=int)
    proj = kept_s[:, np.hstack([free, free + n])]                        # :75-77
    if proj.shape[1]:
        return cleanup(proj, kept_c)                                     # :82
    return np.zeros((1, 0), dtype=bool), np.array([np.sum(kept_c)])      # :84


def taper(symp, coeff, rotations, stab_symp, stab_coeff, free_qubits):
    """S3Projection.perform_projection (projection/base.py:116-124) given the rotation list and the
    rotated stabilizers: Clifford rotations (angle None = pi/2), then the projection."""
    if len(rotations):
        symp, coeff = perform_rotations(symp, coeff, [(r, None) for r in rotations])
    return project_onto_stabilizers(symp, coeff, stab_symp, stab_coeff, free_qubits)


# ----------------------------------------------------------------------------------------------
# a2  benchmark input generator                                       base.py:82-107, utils.py:281-290
# ----------------------------------------------------------------------------------------------

def random_operator(n_qubits: int, n_terms: int, seed: Optional[int] = None, density: float = 0.3,
                    complex_coeffs: bool = True, diagonal: bool = False):
    """Same draws, in the same order, from the same global NumPy RNG as PauliwordOp.random."""
    if seed is not None:
        np.random.seed(seed)
    if diagonal:
        zb = np.random.choice([True, False], size=[n_terms, n_qubits], p=[density / 2, 1 - density / 2])
        symp = np.hstack([np.zeros_like(zb), zb])
    else:
        symp = np.random.choice([True, False], size=[n_terms, 2 * n_qubits], p=[density, 1 - density])
    coeff = np.random.randn(n_terms).astype(complex)
    if complex_coeffs:
        coeff += 1j * np.random.randn(n_terms)
    return symp, coeff


_PAULI_X = {"I": 0, "X": 1, "Y": 1, "Z": 0}
_PAULI_Z = {"I": 0, "X": 0, "Y": 1, "Z": 1}


def from_strings(paulis: List[str], coeffs=None):
    """utils.py:140-163 (string_to_symplectic) applied row by row."""
    n = len(paulis[0]) if paulis else 0
    symp = np.zeros((len(paulis), 2 * n), dtype=bool)
    for i, s in enumerate(paulis):
        assert len(s) == n
        symp[i, :n] = [_PAULI_X[ch] for ch in s]
        symp[i, n:] = [_PAULI_Z[ch] for ch in s]
    if coeffs is None:
        coeffs = np.ones(len(paulis))
    return symp, np.asarray(coeffs, dtype=complex)


def to_strings(symp: np.ndarray) -> List[str]:
    symp = np.asarray(symp, dtype=bool)
    n = symp.shape[1] // 2
    lut = np.array(list("IXZY"))
    code = symp[:, :n].astype(int) + 2 * symp[:, n:].astype(int)
    return ["".join(lut[r]) for r in code]
