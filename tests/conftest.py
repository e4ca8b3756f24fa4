import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the oracle's C helpers are test infrastructure; build them if missing (gcc only, ~1 s)
    lib = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    # SYMMER_HOST_DOUBLE=1 (developer aid on the CPU box): run the API-level GPU tests against the NumPy test double
    # of the kernels (tests/_host_double.py) instead of skipping them — host logic only, never a parity claim.
    doubled = os.environ.get("SYMMER_HOST_DOUBLE") == "1"
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            if not (doubled and os.path.basename(str(item.fspath)) in _DOUBLED_FILES):
                item.add_marker(skip)


_DOUBLED_FILES = ("test_gpu_api.py", "test_gpu_api_ext.py", "test_gpu_ops_ext.py", "test_gpu_replay.py")


@pytest.fixture(autouse=True)
def _host_double_ops(request):
    import torch
    if (os.environ.get("SYMMER_HOST_DOUBLE") == "1" and not torch.cuda.is_available() and "gpu" in request.keywords
            and os.path.basename(str(request.node.fspath)) in _DOUBLED_FILES):
        from _host_double import host_double
        with host_double():
            yield
    else:
        yield


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "golden_vectors.npz")
    data = np.load(path)
    cases = {}
    for key in data.files:
        name, field = key.split("/", 1)
        cases.setdefault(name, {})[field] = data[key]
    return cases


@pytest.fixture(scope="session")
def taper_golden():
    """Tapering workflow vectors produced by the real reference (tests/golden/make_golden_taper.py)."""
    data = np.load(os.path.join(ROOT, "tests", "golden", "taper_vectors.npz"))
    cases = {}
    for key in data.files:
        name, field = key.split("/", 1)
        cases.setdefault(name, {})[field] = data[key]
    for g in cases.values():
        nq = int(g["n_out_qubits"][0])
        g["out_symp"] = np.unpackbits(g["out_symp"], axis=1)[:, :2 * nq].astype(bool)
    return cases


def load_hamiltonian(tag):
    d = np.load(os.path.join(ROOT, "tests", "golden", "hamiltonians", tag + ".npz"))
    n = int(d["n_qubits"][0])
    symp = np.unpackbits(d["symp"], axis=1)[:, :2 * n].astype(bool)
    return symp, d["coeff"], d


@pytest.fixture(scope="session")
def hamiltonians():
    return load_hamiltonian
