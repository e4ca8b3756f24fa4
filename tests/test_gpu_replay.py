"""The API script of tests/replay_cases.py on the CUDA kernels: every call is compared with what the REAL reference
returned for the same call (tests/golden/replay_vectors.npz, recorded by tests/golden/make_golden_replay.py)."""
import os
import types

import numpy as np
import pytest

import replay_cases

pytestmark = pytest.mark.gpu


def test_replay_of_the_reference_on_device():
    import symmer_b200
    from symmer_b200 import ops
    stored = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "replay_vectors.npz")))
    api = types.SimpleNamespace(PauliwordOp=symmer_b200.PauliwordOp, QuantumState=symmer_b200.QuantumState,
                                IndependentOp=symmer_b200.IndependentOp, QubitTapering=symmer_b200.QubitTapering)
    before = ops.launch_count()
    check = replay_cases.Checker(stored)
    replay_cases.run(api, check)
    assert check.checked == sum(1 for k in stored if k.endswith("/kind") and not k.endswith("stab_input/kind"))
    assert ops.launch_count() > before
