"""PauliwordOp / QuantumState with the reference's API (symmer/operators/base.py) on the B200 engine.

Operators are held on the GPU as bit-packed uint64 X|Z rows (`torch.int64[M, 2W]`) plus complex128
coefficients; `symp_matrix` / `coeff_vec` are lazy host views with the reference's dtypes and
shapes. Every algebraic method runs hand-written sm_100a kernels through the C ABI
(`symmer_b200.ops`); only host-side glue (validation, string conversion, sort-by-key) is NumPy.

Reference citations are relative to symmer/operators/base.py unless stated otherwise.
"""
import warnings
from copy import deepcopy
from functools import reduce
from numbers import Number
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch
from scipy.sparse import csr_matrix

from . import ops
from .utils import (check_adjmat_noncontextual, random_symplectic_matrix, rref_binary, strings_to_symplectic,
                    symplectic_to_string, _rref_device)

warnings.simplefilter('always', UserWarning)

# above this many qubits a dense 2^n state no longer fits comfortably next to the operator
DENSE_STATE_MAX_QUBITS = 30
# general rotations of operators with at least this many terms run rotation + dedup as one block-list product
# (ops.rotate_dedup: measured 4.7 vs 6.0 ms at 1e7 rows, equal at 1e6); smaller ones keep the two-step form,
# which has fewer launches (0.30 vs 0.39 ms at 1e5 rows)
FUSED_ROTATION_MIN_TERMS = 1 << 21


class PauliwordOp:
    """Weighted sum of Pauli strings in the symplectic representation (class at base.py:33)."""
    sigfig = 3

    # ------------------------------------------------------------------ construction (base.py:42-74)
    def __init__(self, symp_matrix, coeff_vec) -> None:
        symp_matrix = np.asarray(symp_matrix)
        if symp_matrix.dtype == int:
            assert (set(np.unique(symp_matrix)).issubset({0, 1})), 'symplectic matrix not defined with 0 and 1 only'
            symp_matrix = symp_matrix.astype(bool)
        assert (symp_matrix.dtype == bool), 'Symplectic matrix must be defined over bools'
        if len(symp_matrix.shape) == 1:
            symp_matrix = symp_matrix.reshape([1, len(symp_matrix)])
        assert symp_matrix.shape[-1] % 2 == 0, 'symplectic matrix must have even number of columns'
        assert len(symp_matrix.shape) == 2, 'symplectic matrix must be 2 dimensional only'
        coeff = np.asarray(coeff_vec, dtype=complex)
        n_terms = symp_matrix.shape[0]
        assert (n_terms == len(coeff)), 'coeff list and Pauliwords not same length'   # TypeError for a scalar
        self.n_qubits = symp_matrix.shape[1] // 2
        self.n_terms = n_terms
        dev = ops.device()
        symp_matrix = np.array(symp_matrix, dtype=bool, order='C', copy=True)   # private, read-only host view
        self._xz = ops.pack(torch.from_numpy(symp_matrix).to(dev), self.n_qubits)
        self._c = torch.from_numpy(np.ascontiguousarray(coeff)).to(dev)
        self._symp_host = symp_matrix
        self._symp_host.setflags(write=False)
        self._c_host = None          # becomes authoritative once handed out (callers mutate it in place)
        self._cache = {}

    @classmethod
    def _from_device(cls, xz: torch.Tensor, c: torch.Tensor, n_qubits: int) -> "PauliwordOp":
        """Wrap device tensors without a host round trip."""
        self = cls.__new__(cls)
        self.n_qubits = int(n_qubits)
        self.n_terms = int(xz.shape[0])
        self._xz = xz.contiguous()
        self._c = c.contiguous()
        self._symp_host = None
        self._c_host = None
        self._cache = {}
        return self

    # ------------------------------------------------------------------ lazy host views
    @property
    def symp_matrix(self) -> np.ndarray:
        if self._symp_host is None:
            if self.n_qubits == 0:
                self._symp_host = np.zeros((self.n_terms, 0), dtype=bool)
            else:
                self._symp_host = ops.unpack(self._xz, self.n_qubits).cpu().numpy()
            self._symp_host.setflags(write=False)
        return self._symp_host

    @property
    def coeff_vec(self) -> np.ndarray:
        if self._c_host is None:
            self._c_host = self._c.cpu().numpy()
        return self._c_host

    @coeff_vec.setter
    def coeff_vec(self, value) -> None:
        value = np.asarray(value, dtype=complex)
        assert len(value) == self.n_terms, 'coeff list and Pauliwords not same length'
        self._c_host = value
        self._cache.clear()

    @property
    def X_block(self) -> np.ndarray:
        return self.symp_matrix[:, :self.n_qubits]

    @property
    def Z_block(self) -> np.ndarray:
        return self.symp_matrix[:, self.n_qubits:]

    def _coeff_dev(self) -> torch.Tensor:
        """Device coefficients; refreshed from the host view if one was handed out (the reference's
        callers mutate coeff_vec in place: base.py:746, 1821; independent_op.py:35, 295)."""
        if self._c_host is not None:
            self._c = torch.from_numpy(np.ascontiguousarray(self._c_host, dtype=complex)).to(self._xz.device)
        return self._c

    @property
    def device_rows(self) -> torch.Tensor:
        return self._xz

    @property
    def device_coeffs(self) -> torch.Tensor:
        return self._coeff_dev()

    # ------------------------------------------------------------------ constructors
    @classmethod
    def random(cls, n_qubits: int, n_terms: int, diagonal: bool = False, complex_coeffs: bool = True,
               density: float = 0.3) -> "PauliwordOp":
        """base.py:82-107 — same draws from the global NumPy RNG as the reference."""
        symp_matrix = random_symplectic_matrix(n_qubits, n_terms, diagonal, density=density)
        coeff_vec = np.random.randn(n_terms).astype(complex)
        if complex_coeffs:
            coeff_vec += 1j * np.random.randn(n_terms)
        return cls(symp_matrix, coeff_vec)

    @classmethod
    def from_list(cls, pauli_terms: List[str], coeff_vec: List[complex] = None) -> "PauliwordOp":
        """base.py:128-160."""
        n_rows = len(pauli_terms)
        if coeff_vec is None:
            coeff_vec = np.ones(n_rows)
        else:
            coeff_vec = np.array(coeff_vec)
            if len(coeff_vec.shape) == 2:
                assert (coeff_vec.shape[1] == 2), 'Only tuples of size two allowed (real and imaginary components)'
                coeff_vec = coeff_vec[:, 0] + 1j * coeff_vec[:, 1]
        if pauli_terms:
            symp_matrix = strings_to_symplectic(list(pauli_terms), len(pauli_terms[0]))
        else:
            symp_matrix = np.array([[]], dtype=bool)
        return cls(symp_matrix, coeff_vec)

    @classmethod
    def from_dictionary(cls, operator_dict: Dict[str, complex]) -> "PauliwordOp":
        """base.py:162-177."""
        pauli_terms, coeff_vec = zip(*operator_dict.items())
        return cls.from_list(list(pauli_terms), coeff_vec)

    @classmethod
    def from_packed_file(cls, path) -> "PauliwordOp":
        """Load an operator written by `to_packed_file` (packed uint64 rows + complex128 coefficients)
        straight into device memory: no string parsing, no bool matrix."""
        from .utils import load_packed
        xz, coeff, n, _ = load_packed(path)
        dev = ops.device()
        return cls._from_device(torch.from_numpy(xz.view(np.int64).copy()).to(dev),
                                torch.from_numpy(np.ascontiguousarray(coeff)).to(dev), n)

    def to_packed_file(self, path, **extra) -> None:
        from .utils import save_packed
        save_packed(path, self.symp_matrix, self.coeff_vec, **extra)

    @classmethod
    def empty(cls, n_qubits: int) -> "PauliwordOp":
        """base.py:223-236: 0 * I...I."""
        return cls.from_dictionary({'I' * n_qubits: 0})

    def copy(self) -> "PauliwordOp":
        return deepcopy(self)

    def __deepcopy__(self, memo):
        new = PauliwordOp._from_device(self._xz.clone(), self._coeff_dev().clone(), self.n_qubits)
        new.__class__ = self.__class__
        return new

    # ------------------------------------------------------------------ printing / conversion
    def __str__(self) -> str:
        if self.n_qubits == 0:
            return f'{self.coeff_vec[0]: .{self.sigfig}f}'
        out = ''
        for row, c in zip(self.symp_matrix, self.coeff_vec):
            out += f'{c: .{self.sigfig}f} {symplectic_to_string(row)} +\n'
        return out[:-3]

    def __repr__(self) -> str:
        return str(self)

    @property
    def to_dictionary(self) -> Dict[str, complex]:
        """base.py:1403-1416 (cleans first: duplicates would overwrite each other)."""
        op = self.cleanup()
        return {symplectic_to_string(r): c for r, c in zip(op.symp_matrix, op.coeff_vec)}

    def __hash__(self) -> int:
        return hash(tuple(self.to_dictionary.items()))

    # ------------------------------------------------------------------ ordering / indexing
    def sort(self, by: str = 'magnitude', key: str = 'decreasing') -> "PauliwordOp":
        """base.py:453-489. Sort keys are computed on the host; rows are gathered on the device."""
        symp = self.symp_matrix
        if by == 'magnitude':
            order = np.argsort(-abs(self.coeff_vec))
        elif by == 'lex':
            order = np.lexsort(symp.T) if symp.shape[1] else np.arange(self.n_terms)
        elif by == 'weight':
            order = np.argsort(-np.sum(symp.astype(int), axis=1))
        elif by == 'support':
            pos = np.logical_or(self.X_block, self.Z_block)
            view = np.ascontiguousarray(pos).view(np.dtype((np.void, pos.dtype.itemsize * pos.shape[1])))
            order = np.argsort(view.ravel())[::-1]
        elif by == 'Z':
            order = np.argsort(np.sum((self.n_qubits + 1) * self.X_block.astype(int) + self.Z_block.astype(int), axis=1))
        elif by == 'X':
            order = np.argsort(np.sum(self.X_block.astype(int) + (self.n_qubits + 1) * self.Z_block.astype(int), axis=1))
        elif by == 'Y':
            order = np.argsort(np.sum(abs(self.X_block.astype(int) - self.Z_block.astype(int)), axis=1))
        else:
            raise ValueError('Only permitted sort by values are magnitude, weight, X, Y or Z')
        if key == 'increasing':
            order = order[::-1]
        elif key != 'decreasing':
            raise ValueError('Only permitted sort by values are increasing or decreasing')
        return self._take(np.ascontiguousarray(order))

    def _take(self, index) -> "PauliwordOp":
        idx = torch.as_tensor(np.asarray(index, dtype=np.int64), device=self._xz.device)
        return PauliwordOp._from_device(self._xz.index_select(0, idx), self._coeff_dev().index_select(0, idx),
                                        self.n_qubits)

    def __getitem__(self, key) -> "PauliwordOp":
        """base.py:894-928."""
        if isinstance(key, (int, np.integer)):
            key = int(key)
            if key < 0:
                key += self.n_terms
            assert (key < self.n_terms), 'Index out of range'
            mask = [key]
        elif isinstance(key, slice):
            start, stop = key.start, key.stop
            start = 0 if start is None else start
            stop = self.n_terms if stop is None else stop
            mask = np.arange(start, stop, key.step)
        elif isinstance(key, (list, np.ndarray)):
            mask = np.asarray(key)
            if mask.dtype == bool:
                mask = np.flatnonzero(mask)
        else:
            raise ValueError(f'Unrecognised input {type(key)}, must be an integer, slice, list or np.array')
        return self._take(mask)

    def __iter__(self):
        return iter([self[i] for i in range(self.n_terms)])

    # ------------------------------------------------------------------ a3 Y_count (base.py:604-615)
    @property
    def Y_count(self) -> np.ndarray:
        if 'Y_count' not in self._cache:
            self._cache['Y_count'] = ops.ycount(self._xz).cpu().numpy().astype(np.int64)
        return self._cache['Y_count']

    # ------------------------------------------------------------------ a5 cleanup (base.py:617-638)
    def cleanup(self, zero_threshold: float = 1e-15) -> "PauliwordOp":
        if self.n_qubits == 0:
            return PauliwordOp(np.zeros((1, 0), dtype=bool), [np.sum(self.coeff_vec)])
        if self.n_terms == 0:
            return PauliwordOp(np.zeros((1, 2 * self.n_qubits), dtype=bool), [0])
        xz, c = ops.cleanup(self._xz, self._coeff_dev(), zero_threshold)
        return PauliwordOp._from_device(xz, c, self.n_qubits)

    def __eq__(self, Pword: "PauliwordOp") -> bool:
        """base.py:640-662: equal term sets after cleanup and lexicographic sort."""
        check_1 = self.cleanup().sort('lex')
        check_2 = Pword.cleanup().sort('lex')
        if check_1.n_qubits != check_2.n_qubits:
            raise ValueError('Operators defined over differing numbers of qubits.')
        if check_1.n_terms != check_2.n_terms:
            return False
        return bool(torch.equal(check_1._xz, check_2._xz) and np.allclose(check_1.coeff_vec, check_2.coeff_vec))

    # ------------------------------------------------------------------ a6 append / add / sub (base.py:682-748)
    def append(self, PwordOp: "PauliwordOp") -> "PauliwordOp":
        assert (self.n_qubits == PwordOp.n_qubits), 'Pauliwords defined for different number of qubits'
        return PauliwordOp._from_device(torch.cat([self._xz, PwordOp._xz], dim=0),
                                        torch.cat([self._coeff_dev(), PwordOp._coeff_dev()], dim=0), self.n_qubits)

    def __add__(self, PwordOp: "PauliwordOp") -> "PauliwordOp":
        return self.append(PwordOp).cleanup()

    def __radd__(self, add_obj) -> "PauliwordOp":
        if isinstance(add_obj, Number) and add_obj == 0:
            return self
        return self + add_obj

    def __sub__(self, PwordOp: "PauliwordOp") -> "PauliwordOp":
        neg = PauliwordOp._from_device(PwordOp._xz, -PwordOp._coeff_dev(), PwordOp.n_qubits)
        return self + neg

    def multiply_by_constant(self, const: complex) -> "PauliwordOp":
        return PauliwordOp._from_device(self._xz, self._coeff_dev() * complex(const), self.n_qubits)

    @property
    def dagger(self) -> "PauliwordOp":
        return PauliwordOp._from_device(self._xz, self._coeff_dev().conj().resolve_conj(), self.n_qubits)

    # ------------------------------------------------------------------ a4 multiply (base.py:764-794, 821-859)
    def _multiply_by_operator(self, PwordOp: "PauliwordOp", zero_threshold: float = 1e-15) -> "PauliwordOp":
        assert (self.n_qubits == PwordOp.n_qubits), 'PauliwordOps defined for different number of qubits'
        xz, c = ops.mul_cleanup(self._xz, self._coeff_dev(), PwordOp._xz, PwordOp._coeff_dev(), zero_threshold)
        return PauliwordOp._from_device(xz, c, self.n_qubits)

    def cross_terms(self, PwordOp: "PauliwordOp") -> "PauliwordOp":
        """Pre-cleanup cross terms in the reference's order t = q*M + p (base.py:783-792)."""
        assert (self.n_qubits == PwordOp.n_qubits), 'PauliwordOps defined for different number of qubits'
        xz, c = ops.cross_mul(self._xz, self._coeff_dev(), PwordOp._xz, PwordOp._coeff_dev())
        return PauliwordOp._from_device(xz, c, self.n_qubits)

    def __mul__(self, mul_obj, zero_threshold: float = 1e-15):
        if isinstance(mul_obj, Number):
            return self.multiply_by_constant(mul_obj)
        if isinstance(mul_obj, QuantumState):
            assert (mul_obj.vec_type == 'ket'), 'cannot multiply a bra from the left'
            PwordOp = mul_obj.state_op
        else:
            PwordOp = mul_obj
        # more efficient to make the larger operator the inner loop (base.py:846-851)
        if self.n_terms < PwordOp.n_terms:
            out = PwordOp.dagger._multiply_by_operator(self.dagger, zero_threshold=zero_threshold).dagger
        else:
            out = self._multiply_by_operator(PwordOp, zero_threshold=zero_threshold)
        if isinstance(mul_obj, QuantumState):
            # identities are all mapped to Z in a state: rebuild from the X block (base.py:853-857)
            y = ops.ycount(out._xz).to(torch.int64)
            c = out._coeff_dev() * _i_pow(y)
            return QuantumState._from_x_rows(out._xz, c, out.n_qubits, 'ket').cleanup()
        return out

    def __rmul__(self, other):
        if isinstance(other, Number):
            return self.multiply_by_constant(other)
        return NotImplemented

    def __imul__(self, PwordOp):
        return self.__mul__(PwordOp)

    def __pow__(self, exponent: int) -> "PauliwordOp":
        assert (isinstance(exponent, int)), 'the exponent is not an integer'
        if exponent == 0:
            return PauliwordOp.from_list(['I' * self.n_qubits], [1])
        return reduce(lambda x, y: x * y, [self] * exponent)

    def commutator(self, PwordOp: "PauliwordOp") -> "PauliwordOp":
        return self * PwordOp - PwordOp * self

    def anticommutator(self, PwordOp: "PauliwordOp") -> "PauliwordOp":
        return self * PwordOp + PwordOp * self

    def commutes(self, PwordOp: "PauliwordOp") -> bool:
        commutator = self.commutator(PwordOp).cleanup()
        return (commutator.n_terms == 0 or np.all(commutator.coeff_vec[0] == 0))

    # ------------------------------------------------------------------ a7 commute (base.py:938-971, 1054-1062)
    def commutes_termwise_device(self, PwordOp: "PauliwordOp") -> torch.Tensor:
        assert (self.n_qubits == PwordOp.n_qubits), 'Pauliwords defined for different number of qubits'
        return ops.commute(self._xz, PwordOp._xz)

    def commutes_termwise(self, PwordOp: "PauliwordOp") -> np.ndarray:
        return self.commutes_termwise_device(PwordOp).cpu().numpy()

    def anticommutes_termwise(self, PwordOp: "PauliwordOp") -> np.ndarray:
        return ~self.commutes_termwise(PwordOp)

    @property
    def adjacency_matrix(self) -> np.ndarray:
        if 'adjacency_matrix' not in self._cache:
            self._cache['adjacency_matrix'] = self.commutes_termwise(self)
        return self._cache['adjacency_matrix']

    @property
    def is_noncontextual(self) -> bool:
        """base.py:1074-1088."""
        if self.n_terms < 4:
            return True
        return check_adjmat_noncontextual(self.adjacency_matrix)

    # ------------------------------------------------------------------ a8 rotations (base.py:1090-1186)
    def _rotation_step(self, Pword: "PauliwordOp", angle, threshold: float = 1e-18):
        """One rotation. Returns (operator, status): 'clifford' = rows relabelled, no dedup needed;
        'dirty' = general rotation without its dedup; 'clean' = general rotation already deduplicated."""
        if angle is None:
            angle = np.pi / 2
        if complex(angle).imag != 0:
            warnings.warn('Complex component in angle: this will be ignored.')
        angle = complex(angle).real
        assert (Pword.n_terms == 1), 'Only rotation by single Pauliword allowed here'
        if Pword.coeff_vec[0] != 1:
            warnings.warn(f'Pword coefficient {Pword.coeff_vec[0]: .8f} has been set to 1')
        assert (self.n_qubits == Pword.n_qubits), 'Pauliwords defined for different number of qubits'
        multiple = angle * 2 / np.pi
        int_part = round(multiple)
        c = self._coeff_dev()
        if abs(int_part - multiple) <= threshold:
            sign = -1.0 if int_part in [2, 3] else 1.0          # base.py:1148-1149
            xz, cc = ops.rotate(self._xz, c, Pword._xz, 0.0, 0.0, 1 if int_part % 2 else 2, sign)
            return PauliwordOp._from_device(xz, cc, self.n_qubits), 'clifford'
        if abs(angle) > 1e6:
            warnings.warn('Large angle can lead to precision errors: recommend using high-precision math library '
                          'such as mpmath or redefine angle in range [-pi, pi]')
        if self.n_terms >= FUSED_ROTATION_MIN_TERMS:
            # rotation + dedup as one block-list product: the rotated rows are never materialised in between
            xz, cc = ops.rotate_dedup(self._xz, c, Pword._xz, np.cos(angle), np.sin(angle))
            return PauliwordOp._from_device(xz, cc, self.n_qubits), 'clean'
        xz, cc = ops.rotate(self._xz, c, Pword._xz, np.cos(angle), np.sin(angle), 0)
        return PauliwordOp._from_device(xz, cc, self.n_qubits), 'dirty'

    def _rotate_by_single_Pword(self, Pword: "PauliwordOp", angle: float = None,
                                threshold: float = 1e-18) -> "PauliwordOp":
        """R P R^dagger with R = exp(i*angle/2*Q). Clifford angles return without dedup (like the
        reference); general angles are cleaned once."""
        op, status = self._rotation_step(Pword, angle, threshold)
        return op.cleanup() if status == 'dirty' else op

    def perform_rotations(self, rotations: List[Tuple["PauliwordOp", float]]) -> "PauliwordOp":
        """base.py:1163-1186. The reference cleans up after every rotation; a Clifford rotation is a
        bijection on Pauli rows, so its dedup is deferred to the next general rotation / the end."""
        op = self
        pending = True
        for pauli_rotation, angle in rotations:
            op, status = op._rotation_step(pauli_rotation, angle)
            if status == 'dirty':
                op = op.cleanup()
            if status != 'clifford':
                pending = False
        return op.cleanup() if pending else op

    # ------------------------------------------------------------------ a9/a10 matrix + expval
    def _terms_sorted(self):
        if 'terms_sorted' not in self._cache:
            assert 1 <= self.n_qubits <= 62, 'dense-index kernels need 1 <= n_qubits <= 62'
            self._cache['terms_sorted'] = ops.term_masks_sorted(self._xz, self._coeff_dev(), self.n_qubits)
        return self._cache['terms_sorted']

    @property
    def to_sparse_matrix(self) -> csr_matrix:
        """base.py:1458-1510: CSR of the operator, qubit 0 = most significant bit."""
        if 'to_sparse_matrix' not in self._cache:
            if self.n_qubits == 0:
                self._cache['to_sparse_matrix'] = csr_matrix(self.coeff_vec)
            elif self.n_terms == 0:
                side = 1 << self.n_qubits
                self._cache['to_sparse_matrix'] = csr_matrix((side, side), dtype=complex)
            else:
                xm, zm, cp = self._terms_sorted()
                data, indices, indptr = ops.to_csr(xm, zm, cp, self.n_qubits)
                side = 1 << self.n_qubits
                self._cache['to_sparse_matrix'] = csr_matrix(
                    (data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(side, side))
        return self._cache['to_sparse_matrix']

    def apply_dense(self, psi: torch.Tensor, row_begin: int = 0, row_end: Optional[int] = None) -> torch.Tensor:
        """Matrix-free (H psi)[row_begin:row_end] for a dense complex128 device vector of 2^n amplitudes."""
        xm, zm, cp = self._terms_sorted()
        return ops.apply_dense(xm, zm, cp, self.n_qubits, psi, row_begin, row_end)

    def expval_dense(self, psi: torch.Tensor, row_begin: int = 0, row_end: Optional[int] = None) -> torch.Tensor:
        """Matrix-free partial <psi|H|psi> over basis rows [row_begin, row_end) (device complex128 scalar)."""
        xm, zm, cp = self._terms_sorted()
        return ops.expval_dense(xm, zm, cp, self.n_qubits, psi, row_begin, row_end)

    def expval(self, psi: "QuantumState") -> complex:
        """base.py:796-819. Dense matrix-free kernel when 2^n fits, symbolic products otherwise."""
        assert self.n_qubits == psi.n_qubits
        if 1 <= self.n_qubits <= DENSE_STATE_MAX_QUBITS:
            dense = psi.to_dense_device()
            return complex(self.expval_dense(dense).cpu().numpy()).real
        return (psi.dagger * self * psi).real

    # ------------------------------------------------------------------ a11 GF(2)
    @property
    def generators(self) -> "PauliwordOp":
        """base.py:1436-1456: independent generating set via row reduction."""
        red, piv = _rref_device(self.symp_matrix)
        non_zero = red[piv >= 0]
        gens = PauliwordOp(non_zero, np.ones(non_zero.shape[0], dtype=complex))
        assert gens.n_terms <= 2 * self.n_qubits, 'cannot have an independent generating set of size greaterthan 2 time num qubits'
        return gens

    def generator_reconstruction(self, generators: "PauliwordOp", override_independence_check: bool = False):
        """base.py:523-560: column reduction of [[B],[M]] -> [[I 0],[R F]]; returns (R as int[M, dim],
        mask of the rows whose F part is zero).

        Device-resident: the packed rows of [B; M] are bit-transposed (the reference's cref_binary is
        rref_binary on the transpose, utils.py:349-359), reduced with the bit-exact GF(2) kernel, the
        pivot rows ordered by pivot column (rref_binary's order, zero rows last), and only R and the
        mask travel back to the host. Padding bit positions of the packed layout become all-zero
        rows of the transpose; they never pivot and sort last, so the result is unchanged."""
        from .utils import check_independent
        if not override_independence_check:
            assert check_independent(generators), 'Supplied generators are algebraically dependent'
        assert (self.n_qubits == generators.n_qubits), 'Pauliwords defined for different number of qubits'
        dim, M = generators.n_terms, self.n_terms
        if M == 0 or self.n_qubits == 0:
            return np.zeros((M, dim), dtype=int), np.ones(M, dtype=bool)
        stack = torch.cat([generators._xz, self._xz], dim=0).contiguous()
        T = ops.bit_transpose(stack)                                  # [2W*64 bit positions][ceil((dim+M)/64)]
        piv = ops.rref_packed(T, dim + M).cpu().numpy()
        nz = np.flatnonzero(piv >= 0)
        order = nz[np.argsort(piv[nz], kind='stable')]                # pivot rows by pivot column
        dev = T.device
        recon = np.zeros((M, dim), dtype=int)
        first = order[:dim]
        if len(first):
            sub = T.index_select(0, torch.as_tensor(first, dtype=torch.int64, device=dev)).contiguous()
            back = ops.bit_transpose(sub)                             # [columns of [B; M]][ceil(len(first)/64)]
            recon[:, :len(first)] = ops.unpack_matrix(back[dim:dim + M].contiguous(), len(first)).cpu().numpy()
        rest = order[dim:]
        if len(rest):
            used = ops.or_rows(T, torch.as_tensor(rest, dtype=torch.int32, device=dev))
            used = ops.unpack_matrix(used.reshape(1, -1), dim + M).cpu().numpy()[0, dim:]
            mask = ~used
        else:
            mask = np.ones(M, dtype=bool)
        return recon, mask


def _i_pow(k: torch.Tensor) -> torch.Tensor:
    """i^k as complex128 for an integer device tensor (exact)."""
    lut = torch.tensor([1, 1j, -1, -1j], dtype=torch.complex128, device=k.device)
    return lut[(k % 4).to(torch.int64)]


class QuantumState:
    """Sparse state vector identified with a state_op (|0> -> Z, |1> -> X); class at base.py:1564."""
    sigfig = 3

    def __init__(self, state_matrix, coeff_vector=None, vec_type: str = 'ket') -> None:
        """base.py:1586-1619."""
        if isinstance(state_matrix, list):
            state_matrix = np.array(state_matrix)
        if isinstance(coeff_vector, list):
            coeff_vector = np.array(coeff_vector)
        state_matrix = np.asarray(state_matrix)
        if len(state_matrix.shape) == 1:
            state_matrix = state_matrix.reshape([1, -1])
        state_matrix = state_matrix.astype(int)
        assert (set(state_matrix.flatten()).issubset({0, 1}))
        self.n_terms, self.n_qubits = state_matrix.shape
        if coeff_vector is None:
            coeff_vector = np.ones(self.n_terms) / np.sqrt(self.n_terms)
        self.vec_type = vec_type
        symp_matrix = np.hstack([state_matrix, 1 - state_matrix])
        self.state_op = PauliwordOp(symp_matrix, coeff_vector)
        self._state_host = state_matrix

    @classmethod
    def _from_x_rows(cls, xz: torch.Tensor, c: torch.Tensor, n_qubits: int, vec_type: str) -> "QuantumState":
        """Device constructor: take the X block of packed rows as the bit strings, Z block = complement."""
        W = xz.shape[1] // 2
        x = xz[:, :W]
        full = torch.full((W,), -1, dtype=torch.int64, device=xz.device)
        rem = n_qubits - 64 * (W - 1)
        if rem < 64:
            full[W - 1] = (1 << rem) - 1 if rem > 0 else 0
        z = (~x) & full
        self = cls.__new__(cls)
        self.n_terms, self.n_qubits = int(xz.shape[0]), int(n_qubits)
        self.vec_type = vec_type
        self.state_op = PauliwordOp._from_device(torch.cat([x, z], dim=1).contiguous(), c, n_qubits)
        self._state_host = None
        return self

    @property
    def state_matrix(self) -> np.ndarray:
        if self._state_host is None:
            self._state_host = self.state_op.X_block.astype(int)
        return self._state_host

    def copy(self) -> "QuantumState":
        return QuantumState._from_x_rows(self.state_op._xz.clone(), self.state_op._coeff_dev().clone(), self.n_qubits,
                                         self.vec_type)

    @property
    def dagger(self) -> "QuantumState":
        """base.py:1978-1992."""
        new_type = 'bra' if self.vec_type == 'ket' else 'ket'
        return QuantumState._from_x_rows(self.state_op._xz, self.state_op._coeff_dev().conj().resolve_conj(),
                                         self.n_qubits, new_type)

    def cleanup(self, zero_threshold=1e-15) -> "QuantumState":
        """base.py:1870-1886."""
        clean = self.state_op.cleanup(zero_threshold=zero_threshold)
        return QuantumState._from_x_rows(clean._xz, clean._coeff_dev(), self.n_qubits, self.vec_type)

    def __add__(self, Qstate: "QuantumState") -> "QuantumState":
        new = self.state_op + Qstate.state_op
        return QuantumState._from_x_rows(new._xz, new._coeff_dev(), self.n_qubits, self.vec_type)

    def __sub__(self, Qstate: "QuantumState") -> "QuantumState":
        new = self.state_op - Qstate.state_op
        return QuantumState._from_x_rows(new._xz, new._coeff_dev(), self.n_qubits, self.vec_type)

    @property
    def normalize(self) -> "QuantumState":
        c = self.state_op._coeff_dev()
        return QuantumState._from_x_rows(self.state_op._xz, c / torch.linalg.vector_norm(c), self.n_qubits,
                                         self.vec_type)

    def _is_normalized(self) -> bool:
        """base.py:1964-1976."""
        return bool(np.isclose(np.sum(abs(self.state_op.coeff_vec) ** 2), 1))

    def _bit_keys(self):
        """(sketch keys, packed X rows) used to join two states on equal bit strings."""
        W = self.state_op._xz.shape[1] // 2
        x = self.state_op._xz[:, :W].contiguous()
        padded = torch.cat([x, torch.zeros_like(x)], dim=1).contiguous()
        return ops.sketch(padded), x

    def __mul__(self, mul_obj):
        """base.py:1781-1830: bra * ket -> inner product, bra * PauliwordOp -> bra."""
        if isinstance(mul_obj, Number):
            return QuantumState._from_x_rows(self.state_op._xz, self.state_op._coeff_dev() * complex(mul_obj),
                                             self.n_qubits, self.vec_type)
        assert (self.n_qubits == mul_obj.n_qubits), 'Multiplication object defined for different number of qubits'
        assert (self.vec_type == 'bra'), 'Cannot multiply a ket from the right'
        if isinstance(mul_obj, QuantumState):
            assert (mul_obj.vec_type == 'ket'), 'Cannot multiply a bra with another bra'
            left, right = self.cleanup(zero_threshold=None), mul_obj.cleanup(zero_threshold=None)
            if left.n_terms == 0 or right.n_terms == 0:
                return 0
            # sorted join on the 64-bit row sketches, verified on the packed bits (exact)
            kl, xl = left._bit_keys()
            kr, xr = right._bit_keys()
            kr_sorted, perm = torch.sort(kr)
            pos = torch.searchsorted(kr_sorted, kl).clamp_(max=kr_sorted.numel() - 1)
            cand = perm[pos]
            hit = (kr_sorted[pos] == kl) & (xr[cand] == xl).all(dim=1)
            lc, rc = left.state_op._coeff_dev(), right.state_op._coeff_dev()
            return complex(torch.sum(lc[hit] * rc[cand[hit]]).cpu().numpy())
        if isinstance(mul_obj, PauliwordOp):
            new = self.state_op * mul_obj
            y = ops.ycount(new._xz).to(torch.int64)
            c = new._coeff_dev() * _i_pow(3 * y)                       # (-i)^Y
            return QuantumState._from_x_rows(new._xz, c, self.n_qubits, 'bra').cleanup()
        raise ValueError('Trying to multiply QuantumState by unrecognised object - must be another Quantum state or PauliwordOp')

    def __getitem__(self, key) -> "QuantumState":
        sub = self.state_op[key]
        return QuantumState._from_x_rows(sub._xz, sub._coeff_dev(), self.n_qubits, self.vec_type)

    def __iter__(self):
        return iter([self[i] for i in range(self.n_terms)])

    @property
    def to_dictionary(self) -> Dict[str, complex]:
        """base.py:2098-2111."""
        st = self.cleanup()
        return {"".join(str(int(b)) for b in row): c for row, c in zip(st.state_matrix, st.state_op.coeff_vec)}

    def basis_indices_device(self) -> torch.Tensor:
        """Basis index of every term (qubit 0 = most significant bit), n_qubits <= 62."""
        assert 1 <= self.n_qubits <= 62
        xm, _, _ = _unsorted_masks(self.state_op)
        return xm

    def to_dense_device(self) -> torch.Tensor:
        """complex128[2^n] device vector (duplicates summed)."""
        assert 1 <= self.n_qubits <= DENSE_STATE_MAX_QUBITS
        idx = self.basis_indices_device()
        dense = torch.zeros(1 << self.n_qubits, dtype=torch.complex128, device=idx.device)
        c = self.state_op._coeff_dev()
        dense_r = torch.view_as_real(dense)
        dense_r[:, 0].index_add_(0, idx, c.real.contiguous())
        dense_r[:, 1].index_add_(0, idx, c.imag.contiguous())
        return dense

    @property
    def to_sparse_matrix(self):
        """base.py:1994-2011: column (ket) or row (bra) sparse vector."""
        from scipy.sparse import csr_matrix as _csr
        idx = self.basis_indices_device().cpu().numpy()
        c = self.state_op.coeff_vec
        side = 1 << self.n_qubits
        if self.vec_type == 'ket':
            return _csr((c, (idx, np.zeros_like(idx))), shape=(side, 1), dtype=complex)
        return _csr((c, (np.zeros_like(idx), idx)), shape=(1, side), dtype=complex)

    def __str__(self) -> str:
        out = ''
        for row, c in zip(self.state_matrix, self.state_op.coeff_vec):
            bits = "".join(str(int(b)) for b in row)
            out += (f'{c: .{self.sigfig}f} |{bits}> +\n' if self.vec_type == 'ket'
                    else f'{c: .{self.sigfig}f} <{bits}| +\n')
        return out[:-3]

    def __repr__(self) -> str:
        return str(self)


def _unsorted_masks(op: PauliwordOp):
    """x/z basis-index masks in term order (qubit 0 = MSB)."""
    import ctypes
    from . import _cabi
    M = op.n_terms
    dev = op._xz.device
    xm = torch.empty(M, dtype=torch.int64, device=dev)
    zm = torch.empty(M, dtype=torch.int64, device=dev)
    cp = torch.empty(M, dtype=torch.complex128, device=dev)
    if M:
        _cabi.check(ops.lib().sym_term_masks(ctypes.c_void_p(op._xz.data_ptr()), ctypes.c_void_p(op._coeff_dev().data_ptr()),
                                             M, op.n_qubits, ctypes.c_void_p(xm.data_ptr()),
                                             ctypes.c_void_p(zm.data_ptr()), ctypes.c_void_p(cp.data_ptr()),
                                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return xm, zm, cp


def single_term_expval(P_op: PauliwordOp, psi: QuantumState) -> float:
    """base.py:2438-2471: <psi|P|psi> for a single Pauli term."""
    assert (P_op.n_terms == 1), 'Supplied multiple Pauli terms.'
    unit = PauliwordOp._from_device(P_op._xz, torch.ones(1, dtype=torch.complex128, device=P_op._xz.device),
                                    P_op.n_qubits)
    return unit.expval(psi)
