"""Development check (build container only: needs /root/reference): run the REFERENCE'S OWN test files, unmodified and
in place, against this engine — `symmer` is served by `symmer_b200.compat.install_as_symmer()`.

    python scripts/run_reference_tests.py            # CPU box: kernels swapped for the NumPy test double (host logic)
    python scripts/run_reference_tests.py --device   # on a B200: the CUDA kernels themselves

Tests that build openfermion / qiskit objects fail with the stand-ins of oracle/shim (those packages are not in
this image); everything else is expected to pass. Nothing of the reference is copied into this repository.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF_TESTS = "/root/reference/tests"
FILES = ["test_operators/test_base.py", "test_operators/test_independent_op.py", "test_operators/test_operator_utils.py",
         "test_projection/test_qubit_tapering.py", "test_evolution/test_evolution_gate_library.py",
         "test_evolution/test_circuit_symmerlator.py", "test_symmer_utils.py"]


def main():
    if not os.path.isdir(REF_TESTS):
        print("reference tests not available on this machine")
        return 0
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "shim")]
    if not hasattr(np, "product"):
        np.product = np.prod                                   # removed in NumPy 2; symmer's base.py:2038 uses it
    if "--device" not in sys.argv:
        from _host_double import host_double
        main.double = host_double()                            # kept alive: stays active for the whole run
        main.double.__enter__()
    from symmer_b200 import compat
    compat.install_as_symmer()
    args = [os.path.join(REF_TESTS, f) for f in FILES]
    return pytest.main(args + ["-q", "-p", "no:cacheprovider", "--rootdir", "/tmp", "-c", "/dev/null",
                               "-W", "ignore::UserWarning"])


if __name__ == "__main__":
    sys.exit(main())
