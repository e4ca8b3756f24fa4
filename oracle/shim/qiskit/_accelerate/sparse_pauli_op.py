"""Stand-in for the two qiskit 1.2.4 Rust entry points the reference calls (test infrastructure only).

The reference imports `unordered_unique` (symmer/operators/utils.py:6, called :271) and
`ZXPaulis` + `to_matrix_sparse` (symmer/operators/base.py:24-27, called :1500-1508). qiskit is not
installable here (no network, no wheel), so their published value semantics are restated in
`oracle/pauli_oracle.py`; this module only adapts the call signatures.
"""
import os
import sys

_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..", ".."))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import pauli_oracle as _po  # noqa: E402


def unordered_unique(arr):
    return _po.unordered_unique(arr)


class ZXPaulis:
    def __init__(self, x, z, phases, coeffs):
        self.x = x
        self.z = z
        self.phases = phases
        self.coeffs = coeffs


def to_matrix_sparse(zx, force_serial=False):
    return _po.zx_to_matrix_sparse(zx.x, zx.z, zx.phases, zx.coeffs)
