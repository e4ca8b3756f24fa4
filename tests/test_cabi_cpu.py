"""CPU-only checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/symmer_b200.h declares (and nothing the header does not), the ctypes table mirrors the
header, size queries work, and the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "symmer_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sym_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from symmer_b200 import _cabi
    lib = _cabi.load()
    names = header_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (sym_[a-z0-9_]+)", out)))
    assert exported == names, set(exported) ^ set(names)


def test_ctypes_table_matches_header():
    from symmer_b200 import _cabi
    assert sorted(_cabi.SIGNATURES) == header_symbols()
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, argtypes) in _cabi.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", text, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))


def test_size_queries_and_version():
    from symmer_b200 import _cabi
    lib = _cabi.load()
    assert lib.sym_abi_version() == 1
    small = lib.sym_mul_cleanup_ws_bytes(500, 500, 16)
    big = lib.sym_mul_cleanup_ws_bytes(12500, 10000, 16)
    assert 0 < small < big < 16 * 2**30
    assert lib.sym_cleanup_ws_bytes(10**6, 16) > 10**6 * 30
    assert lib.sym_launch_count() == 0


def test_argument_validation_without_gpu():
    from symmer_b200 import _cabi
    lib = _cabi.load()
    rc = lib.sym_mul_cleanup_count(None, None, 10**6, None, None, 10**6, 16, 1e-15, None, None, None, 0, None)
    assert rc == -1 and b"4e9" in lib.sym_last_error()
    rc = lib.sym_rotate(None, None, 5, 1, None, 0.0, 0.0, 7, 1.0, None, None, None, None, 0, None)
    assert rc == -1
    with pytest.raises(_cabi.SymmerB200Error):
        _cabi.check(rc)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import symmer_b200
    from symmer_b200 import utils
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        symmer_b200.PauliwordOp.from_list(["XX"], [1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        utils.symplectic_cleanup(np.zeros((2, 4), dtype=bool), np.ones(2))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under symmer_b200/ may import it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "symmer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("pauli_oracle", "oracle") or "import oracle" not in src, f
                assert "from oracle" not in src and "import oracle" not in src and "pauli_oracle" not in src, f
                # ... nor the NumPy test double of the kernels (tests/_host_double.py) or its developer switch
                assert "_host_double" not in src and "host_double" not in src and "SYMMER_HOST_DOUBLE" not in src, f


def test_host_string_ingest():
    from symmer_b200 import utils
    from oracle import pauli_oracle as po
    terms = ["XYZI", "IIII", "YYXZ"]
    got = utils.strings_to_symplectic(terms, 4)
    exp, _ = po.from_strings(terms)
    assert np.array_equal(got, exp)
    assert [utils.symplectic_to_string(r) for r in got] == terms
    assert np.array_equal(utils.string_to_symplectic("XYZI", 4).astype(bool), exp[0])
    with pytest.raises(AssertionError):
        utils.strings_to_symplectic(["XA"], 2)
    with pytest.raises(AssertionError):
        utils.strings_to_symplectic(["XX", "X"], 2)
