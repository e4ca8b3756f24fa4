"""Bind the B200 engine into an unmodified UCL-CCS/symmer install.

The reference has no plugin interface; its hot path goes through module-level array kernels that
`symmer.operators.base` and `symmer.operators.independent_op` import BY NAME (base.py:7-11,
independent_op.py:6), so the names are re-bound in every namespace that holds them. After
`install()` the reference's own PauliwordOp runs its cleanup, commutation matrix and GF(2)
reductions on the GPU (host arrays in, host arrays out); code that wants device-resident
operators uses `symmer_b200.PauliwordOp` directly. See INTEGRATION.md.
"""
import importlib

from . import utils as _u

_SEAMS = {
    "symplectic_cleanup": _u.symplectic_cleanup,      # utils.py:230
    "matmul_GF2": _u.matmul_GF2,                      # utils.py:9
    "_rref_binary": _u._rref_binary,                  # utils.py:292
    "rref_binary": _u.rref_binary,                    # utils.py:317
    "_cref_binary": _u._cref_binary,                  # utils.py:337
    "cref_binary": _u.cref_binary,                    # utils.py:349
    "check_independent": _u.check_independent,        # utils.py:504
    "check_jordan_independent": _u.check_jordan_independent,   # utils.py:521
}
_MODULES = ["symmer.operators.utils", "symmer.operators.base", "symmer.operators.independent_op",
            "symmer.operators.noncontextual_op", "symmer.operators.anticommuting_op"]
_saved = {}


def install():
    """Re-bind the seams; returns the list of (module, name) pairs that were patched."""
    import symmer  # noqa: F401  (must be importable; this package does not ship it)
    from symmer import process
    process.method = 'single_thread'   # the reference forks in expval (base.py:811): never after CUDA init
    done = []
    for modname in _MODULES:
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            continue
        for name, fn in _SEAMS.items():
            if hasattr(mod, name):
                _saved.setdefault((modname, name), getattr(mod, name))
                setattr(mod, name, fn)
                done.append((modname, name))
    return done


def uninstall():
    for (modname, name), fn in _saved.items():
        setattr(importlib.import_module(modname), name, fn)
    _saved.clear()
