"""Packed on-disk operator format: host-side twins of sym_pack / sym_unpack and the
.npz container; the GPU round trip through PauliwordOp.{to,from}_packed_file."""
import os

import numpy as np
import pytest

from oracle import pauli_oracle as po
from symmer_b200 import utils as u


@pytest.mark.parametrize("n", [1, 5, 63, 64, 65, 130, 1000])
def test_host_pack_matches_device_layout(n, tmp_path):
    rng = np.random.default_rng(n)
    symp = rng.random((23, 2 * n)) < 0.3
    coeff = rng.standard_normal(23) + 1j * rng.standard_normal(23)
    xz = u.pack_rows_host(symp)
    assert xz.dtype == np.uint64 and np.array_equal(xz, po.pack_bits(symp))     # the layout of include/symmer_b200.h
    assert np.array_equal(u.unpack_rows_host(xz, n), symp)
    path = os.path.join(tmp_path, "op.npz")
    u.save_packed(path, symp, coeff, hf_array=np.arange(n) % 2)
    xz2, c2, n2, extra = u.load_packed(path)
    assert n2 == n and np.array_equal(xz2, xz) and np.array_equal(c2, coeff)
    assert np.array_equal(extra["hf_array"], np.arange(n) % 2)


def test_rejects_unknown_version(tmp_path):
    path = os.path.join(tmp_path, "bad.npz")
    np.savez(path, format_version=np.array([99]), n_qubits=np.array([1]), xz=np.zeros((1, 2), np.uint64),
             coeff=np.zeros(1, complex))
    with pytest.raises(ValueError):
        u.load_packed(path)


@pytest.mark.gpu
def test_operator_file_round_trip_on_device(tmp_path, hamiltonians):
    from symmer_b200 import PauliwordOp
    symp, coeff, _ = hamiltonians("H2O_STO3G")
    H = PauliwordOp(symp, coeff)
    path = os.path.join(tmp_path, "h2o.npz")
    H.to_packed_file(path)
    assert os.path.getsize(path) < symp.size // 4
    G = PauliwordOp.from_packed_file(path)
    assert G.n_qubits == H.n_qubits and np.array_equal(G.symp_matrix, symp) and np.array_equal(G.coeff_vec, coeff)
    assert (G * G) == (H * H)
