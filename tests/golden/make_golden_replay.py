"""Record what the REAL reference returns for the API script of tests/replay_cases.py (test infrastructure; run in
the build container):

    python tests/golden/make_golden_replay.py

The script is written against the public API only, so it drives `symmer` (imported unmodified through oracle/shim)
here and `symmer_b200` in the tests; every result of the reference goes to tests/golden/replay_vectors.npz.
`np.product` (removed in NumPy 2) is aliased to `np.prod` for base.py:2038."""
import json
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "shim"), "/root/reference"]
warnings.simplefilter("ignore")
if not hasattr(np, "product"):
    np.product = np.prod

import symmer  # noqa: E402
from symmer import process  # noqa: E402
from symmer.operators import IndependentOp  # noqa: E402
import replay_cases  # noqa: E402

process.method = 'single_thread'
HAM_DIR = "/root/reference/tests/hamiltonian_data"


def loader(fname):
    def load():
        with open(os.path.join(HAM_DIR, fname)) as f:
            d = json.load(f)
        H = symmer.PauliwordOp.from_dictionary({k: complex(v[0], v[1]) for k, v in d["hamiltonian"].items()})
        return dict(symp=H.symp_matrix, coeff=H.coeff_vec, hf=np.asarray(d["data"]["hf_array"], dtype=int))
    return load


api = types.SimpleNamespace(PauliwordOp=symmer.PauliwordOp, QuantumState=symmer.QuantumState, IndependentOp=IndependentOp,
                            QubitTapering=symmer.QubitTapering)
hams = {"H3p_STO3G": loader("H3+_STO-3G_SINGLET_JW.json"), "H4_STO3G": loader("H4_STO-3G_SINGLET_JW.json"),
        "HeHp_321G": loader("HeH+_3-21G_SINGLET_JW.json"), "LiH_STO3G": loader("LiH_STO-3G_SINGLET_JW.json")}
rec = replay_cases.Recorder()
replay_cases.run(api, rec, hamiltonians=hams)
path = os.path.join(HERE, "replay_vectors.npz")
np.savez_compressed(path, **rec.out)
n_results = sum(1 for k in rec.out if k.endswith("/kind"))
print(f"recorded {n_results} results ({len(rec.out)} arrays) to {path} ({os.path.getsize(path) / 1024:.0f} KB)")
