"""Array-level seams of the reference (symmer/operators/utils.py) on the B200 engine.

Same names, argument meaning and return types (NumPy in, NumPy out) as the reference functions they
replace, so `symmer_b200.patch.install()` can bind them into an unmodified symmer. Each call copies
its inputs to the device, runs the CUDA kernels through the C ABI and copies the result back; code
that wants to stay on the device uses `symmer_b200.base.PauliwordOp` / `symmer_b200.ops` instead.
"""
from typing import Tuple

import os

import numpy as np
import torch

from . import ops


def _to_dev_bool(a: np.ndarray) -> torch.Tensor:
    a = np.array(a, dtype=bool, order='C', copy=True)      # private writable copy (torch wants writable)
    return torch.from_numpy(a).to(ops.device(), non_blocking=True)


def symplectic_cleanup(symp_matrix: np.ndarray, coeff_vec: np.ndarray,
                       zero_threshold: float = None) -> Tuple[np.ndarray, np.ndarray]:
    """utils.py:230-279. Unique rows (first-occurrence order), duplicate coefficients summed in input
    order, then |c| > zero_threshold."""
    symp_matrix = np.asarray(symp_matrix, dtype=bool)
    coeff_vec = np.asarray(coeff_vec, dtype=complex)
    T, two_n = symp_matrix.shape
    n = two_n // 2
    if T == 0:
        return symp_matrix.copy(), coeff_vec.copy()
    xz = ops.pack(_to_dev_bool(symp_matrix), n)
    c = torch.from_numpy(np.ascontiguousarray(coeff_vec)).to(xz.device)
    oxz, oc = ops.cleanup(xz, c, zero_threshold)
    return ops.unpack(oxz, n).cpu().numpy(), oc.cpu().numpy()


def matmul_GF2(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """utils.py:9-26: (A @ B) mod 2 for boolean A[M,K], B[K,N]. Runs the bit-packed symplectic
    inner-product kernel with A's rows as X blocks and B's columns as Z blocks."""
    A = np.asarray(A, dtype=bool)
    B = np.asarray(B, dtype=bool)
    M, K = A.shape
    K2, N = B.shape
    assert K == K2
    dev = ops.device()
    if M == 0 or N == 0:
        return np.zeros((M, N), dtype=bool)
    a_rows = torch.zeros((M, 2 * K), dtype=torch.bool, device=dev)
    a_rows[:, :K] = _to_dev_bool(A)
    b_rows = torch.zeros((N, 2 * K), dtype=torch.bool, device=dev)
    b_rows[:, K:] = _to_dev_bool(B.T)
    commute = ops.commute(ops.pack(a_rows, K), ops.pack(b_rows, K))
    return (~commute).cpu().numpy()


def commutes_termwise(a_symp: np.ndarray, b_symp: np.ndarray) -> np.ndarray:
    """base.py:938-971 on plain arrays: True where A[i] commutes with B[j]."""
    a_symp = np.asarray(a_symp, dtype=bool)
    b_symp = np.asarray(b_symp, dtype=bool)
    n = a_symp.shape[1] // 2
    assert b_symp.shape[1] == 2 * n, 'Pauliwords defined for different number of qubits'
    a = ops.pack(_to_dev_bool(a_symp), n)
    b = ops.pack(_to_dev_bool(b_symp), n)
    return ops.commute(a, b).cpu().numpy()


def _rref_device(matrix: np.ndarray):
    matrix = np.asarray(matrix, dtype=bool)
    if matrix.ndim != 2:
        raise ValueError("matrix must be 2-dimensional")
    R, C = matrix.shape
    if R == 0 or C == 0:
        return matrix.copy(), np.full(R, -1, dtype=np.int32)
    bits = ops.pack_matrix(_to_dev_bool(matrix))
    piv = ops.rref_packed(bits, C)
    return ops.unpack_matrix(bits, C).cpu().numpy(), piv.cpu().numpy()


def _rref_binary(matrix: np.ndarray) -> np.ndarray:
    """utils.py:292-315: row reduction over GF(2), rows not reordered."""
    return _rref_device(matrix)[0]


def rref_binary(matrix: np.ndarray) -> np.ndarray:
    """utils.py:317-335: rows ordered by pivot column, zero rows last."""
    red, piv = _rref_device(matrix)
    nz = np.flatnonzero(piv >= 0)
    order = list(nz[np.argsort(piv[nz], kind="stable")])
    seen = set(order)
    order += [i for i in range(red.shape[0]) if i not in seen]
    return red[order]


def _cref_binary(matrix: np.ndarray) -> np.ndarray:
    return _rref_binary(np.asarray(matrix).T).T            # utils.py:337-347


def cref_binary(matrix: np.ndarray) -> np.ndarray:
    return rref_binary(np.asarray(matrix).T).T             # utils.py:349-359


def check_independent(operators) -> bool:
    """utils.py:504-519 (accepts anything with n_terms, n_qubits, symp_matrix)."""
    if operators.n_terms > 2 * operators.n_qubits:
        return False
    if hasattr(operators, 'device_rows') and operators.n_terms and operators.n_qubits:
        # the packed rows ARE a packed bit matrix in the column order [X | padding | Z | padding]: all-zero padding
        # columns never pivot, so reducing a copy of them in place is the reduction of bool[M, 2n]
        rows = operators.device_rows.clone()
        piv = ops.rref_packed(rows, 64 * rows.shape[1])
        return bool((piv >= 0).all().item())
    _, piv = _rref_device(operators.symp_matrix)
    return bool(np.all(piv >= 0))


def check_jordan_independent(operators) -> bool:
    """utils.py:521-565: independence under the Jordan product (anticommuting elements may not be multiplied).
    The commutation matrix and both GF(2) reductions run on the device."""
    if operators.n_terms > 3 * operators.n_qubits:
        return False
    comm_mask = np.sum(operators.commutes_termwise(operators), axis=1) == operators.n_terms
    comm_part = operators[comm_mask]
    if comm_part.n_terms and not check_independent(comm_part):
        return False
    X_block, Z_block = operators.X_block, operators.Z_block
    Y_block = np.logical_and(Z_block, X_block)
    XZY_block = np.hstack((np.logical_xor(X_block, Y_block), np.logical_xor(Z_block, Y_block), Y_block))
    _, piv = _rref_device(XZY_block)
    return bool(np.all(piv >= 0))


def check_adjmat_noncontextual(adjmat) -> bool:
    """utils.py:567-589 — host logic on the (small) set of distinct commutation characters."""
    adjmat = np.asarray(adjmat, dtype=bool)
    mask = np.where(~np.all(adjmat, axis=1))[0]
    unique = np.unique(adjmat[mask, :][:, mask], axis=0)
    return bool(np.all(np.count_nonzero(unique, axis=0) == 1))


def random_symplectic_matrix(n_qubits, n_terms, diagonal=False, density=0.3):
    """utils.py:281-290: same draws from the global NumPy RNG as the reference."""
    if diagonal:
        Z_block = np.random.choice([True, False], size=[n_terms, n_qubits], p=[density / 2, 1 - density / 2])
        return np.hstack([np.zeros_like(Z_block), Z_block])
    return np.random.choice([True, False], size=[n_terms, 2 * n_qubits], p=[density, 1 - density])


_X_OF = {"I": False, "X": True, "Y": True, "Z": False}
_Z_OF = {"I": False, "X": False, "Y": True, "Z": True}


def string_to_symplectic(pauli_str, n_qubits):
    """utils.py:140-163."""
    assert (len(pauli_str) == n_qubits), 'Number of qubits is incompatible with pauli string'
    assert (set(pauli_str).issubset({'I', 'X', 'Y', 'Z'})), 'pauliword must only contain X,Y,Z,I terms'
    out = np.zeros(2 * n_qubits, dtype=int)
    out[:n_qubits] = [_X_OF[ch] for ch in pauli_str]
    out[n_qubits:] = [_Z_OF[ch] for ch in pauli_str]
    return out


def strings_to_symplectic(pauli_terms, n_qubits) -> np.ndarray:
    """Vectorised ingest of many Pauli strings: one pass over a byte view instead
    of a Python loop per character."""
    if len(pauli_terms) == 0:
        return np.zeros((0, 2 * n_qubits), dtype=bool)
    for s in pauli_terms:
        assert (len(s) == n_qubits), 'Number of qubits is incompatible with pauli string'
    chars = np.frombuffer("".join(pauli_terms).encode("ascii"), dtype=np.uint8).reshape(len(pauli_terms), n_qubits)
    is_x, is_y, is_z, is_i = chars == ord("X"), chars == ord("Y"), chars == ord("Z"), chars == ord("I")
    assert bool(np.all(is_x | is_y | is_z | is_i)), 'pauliword must only contain X,Y,Z,I terms'
    return np.hstack([is_x | is_y, is_z | is_y])


_LUT = np.array(list("IXZY"))


def symplectic_to_string(symp_vec) -> str:
    """utils.py:80-107."""
    symp_vec = np.asarray(symp_vec, dtype=bool)
    n = symp_vec.size // 2
    return "".join(_LUT[symp_vec[:n].astype(int) + 2 * symp_vec[n:].astype(int)])


# ------------------------------------------------------------------ packed on-disk operators
PACKED_FORMAT_VERSION = 1


def save_packed(path, symp_matrix, coeff_vec, **extra) -> None:
    """Write an operator as a compressed .npz holding the engine's own layout: uint64[M, 2W] X|Z rows
    (qubit q = bit q%64 of word q//64 of its block) + complex128[M]. 8x smaller than the bool matrix
    and ~30x smaller than the reference's JSON dictionaries (tests/hamiltonian_data/*.json); loading
    needs no per-string parsing and uploads straight into device memory."""
    symp_matrix = np.asarray(symp_matrix, dtype=bool)
    M, two_n = symp_matrix.shape
    n = two_n // 2
    path = _npz_path(path)
    np.savez_compressed(path, format_version=np.array([PACKED_FORMAT_VERSION]), n_qubits=np.array([n]),
                        xz=pack_rows_host(symp_matrix), coeff=np.asarray(coeff_vec, dtype=complex),
                        **{k: np.asarray(v) for k, v in extra.items()})


def _npz_path(path):
    """np.savez appends '.npz' to a name without it: save and load agree on the name actually written."""
    path = os.fspath(path) if not hasattr(path, "write") and not hasattr(path, "read") else path
    if isinstance(path, str) and not path.endswith(".npz"):
        path += ".npz"
    return path


def load_packed(path):
    """(xz uint64[M, 2W], coeff complex128[M], n_qubits, extra dict) from a file written by save_packed. The
    contents are checked before they reach a kernel: dtypes, shapes (2W words per row, one coefficient per row)
    and zero padding bits above qubit n-1 (dedup, equality and the GF(2) reductions rely on them)."""
    path = _npz_path(path)
    d = np.load(path)
    if int(d["format_version"][0]) != PACKED_FORMAT_VERSION:
        raise ValueError(f"unsupported packed operator format {int(d['format_version'][0])}")
    xz, coeff, n = d["xz"], d["coeff"], int(d["n_qubits"][0])
    W = max(1, (n + 63) // 64)
    if xz.dtype != np.uint64 or xz.ndim != 2 or xz.shape[1] != 2 * W:
        raise ValueError(f"packed rows must be uint64[M, {2 * W}] for {n} qubits, got {xz.dtype}{list(xz.shape)}")
    coeff = np.ascontiguousarray(coeff, dtype=complex)
    if coeff.ndim != 1 or coeff.shape[0] != xz.shape[0]:
        raise ValueError(f"{xz.shape[0]} packed rows but {coeff.shape} coefficients")
    used = n - 64 * (W - 1)                                     # qubits in the last word of each block
    if n == 0:
        pad = np.uint64(0xFFFFFFFFFFFFFFFF)
    else:
        pad = np.uint64(0) if used == 64 else np.uint64(0xFFFFFFFFFFFFFFFF) << np.uint64(used)
    if pad and xz.shape[0] and (np.any(xz[:, W - 1] & pad) or np.any(xz[:, 2 * W - 1] & pad)):
        raise ValueError("packed rows carry bits above the last qubit (padding must be zero)")
    extra = {k: d[k] for k in d.files if k not in ("format_version", "n_qubits", "xz", "coeff")}
    return np.ascontiguousarray(xz), coeff, n, extra


def pack_rows_host(symp_matrix: np.ndarray) -> np.ndarray:
    """bool[M, 2n] -> uint64[M, 2W] in the device layout (host-side twin of sym_pack, used for files)."""
    symp_matrix = np.asarray(symp_matrix, dtype=bool)
    M, two_n = symp_matrix.shape
    n = two_n // 2
    W = max(1, (n + 63) // 64)
    out = np.zeros((M, 2 * W), dtype=np.uint64)
    for blk in range(2):
        bits = np.zeros((M, W * 64), dtype=np.uint8)
        bits[:, :n] = symp_matrix[:, blk * n:(blk + 1) * n]
        by = np.packbits(bits.reshape(M, W, 8, 8), axis=-1, bitorder="little").reshape(M, W, 8)
        out[:, blk * W:(blk + 1) * W] = by.astype(np.uint64).dot(np.uint64(1) << (np.arange(8, dtype=np.uint64) * np.uint64(8)))
    return out


def unpack_rows_host(xz: np.ndarray, n_qubits: int) -> np.ndarray:
    """uint64[M, 2W] -> bool[M, 2n] (host-side twin of sym_unpack)."""
    xz = np.asarray(xz, dtype=np.uint64)
    M, two_w = xz.shape
    W = two_w // 2
    by = (xz[:, :, None] >> (np.arange(8, dtype=np.uint64) * np.uint64(8))).astype(np.uint8)      # little-endian bytes
    bits = np.unpackbits(by, axis=-1, bitorder="little").reshape(M, 2, W * 64)
    return np.hstack([bits[:, 0, :n_qubits], bits[:, 1, :n_qubits]]).astype(bool)


# ---------------------------------------------------------------------------------------------- small helpers
def symplectic_to_sparse_matrix(symp_vec, coeff):
    """utils.py:182-228: CSR matrix of one Pauli term (through the device CSR emitter of `to_sparse_matrix`)."""
    from .base import PauliwordOp
    return PauliwordOp(np.asarray(symp_vec, dtype=bool).reshape(1, -1), [coeff]).to_sparse_matrix


def mul_symplectic(symp_vec1, coeff1, symp_vec2, coeff2):
    """utils.py:429-470: product of two Pauli terms with its phase — one cross term of the device product kernel."""
    from .base import PauliwordOp
    left = PauliwordOp(np.asarray(symp_vec1, dtype=bool).reshape(1, -1), [coeff1])
    right = PauliwordOp(np.asarray(symp_vec2, dtype=bool).reshape(1, -1), [coeff2])
    out = left.cross_terms(right)
    return out.symp_matrix[0], out.coeff_vec[0]


def safe_PauliwordOp_to_dict(op):
    """utils.py:401-413: {pauli string: (real, imag)}."""
    return {term: (c.real, c.imag) for term, c in op.to_dictionary.items()}


def safe_QuantumState_to_dict(psi):
    """utils.py:415-427: {bit string: (real, imag)}."""
    return {bits: (c.real, c.imag) for bits, c in psi.to_dictionary.items()}


def count1_in_int_bitstring(i) -> int:
    """utils.py:165-180: number of set bits of a 32-bit integer."""
    return bin(int(i) & 0xFFFFFFFF).count('1')


def binary_array_to_int(bin_arr) -> np.ndarray:
    """utils.py:618-638: rows of bits (most significant first) as integers; floats from 64 columns on, like the reference."""
    bin_arr = np.asarray(bin_arr)
    width = bin_arr.shape[1]
    weights = 2 ** np.arange(width - 1, -1, -1) if width < 64 else 2 ** np.arange(width - 1, -1, -1, dtype=float)
    return bin_arr @ weights


def unit_n_sphere_cartesian_coords(angles) -> np.ndarray:
    """utils.py:472-485: the n+1 Cartesian coordinates of the point of the unit n-sphere with the given n angles."""
    angles = np.asarray(angles, dtype=float)
    sines = np.concatenate([[1.0], np.cumprod(np.sin(angles))])
    return np.concatenate([sines[:-1] * np.cos(angles), sines[-1:]])


def binomial_coefficient(n, k):
    """utils.py:487-502: n choose k for non-integer n."""
    out = 1
    for r in range(k):
        out *= (n - r) / (k - r)
    return out


def perform_noncontextual_sweep(operator):
    """utils.py:592-616: one pass over an ordered operator keeping every term that leaves the kept set noncontextual.
    The reference grows the commutation matrix term by term; here the whole matrix comes from ONE device call and
    the sweep indexes into it."""
    if operator.n_terms == 0:
        return operator
    adjacency = operator.adjacency_matrix
    kept = [0]
    for index in range(1, operator.n_terms):
        trial = kept + [index]
        if check_adjmat_noncontextual(adjacency[np.ix_(trial, trial)]):
            kept = trial
    return operator[kept]
