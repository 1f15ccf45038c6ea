import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    if "seqm_parameters" in d:
        d["seqm_parameters"] = json.loads(str(d["seqm_parameters"]))
    if "n_scf_iter" in d:
        d["n_scf_iter"] = int(d["n_scf_iter"])
    return d


@pytest.fixture(scope="session")
def golden():
    return load_golden


# North-star tolerances (BASELINE.json): energies 1e-6 eV, density 1e-8, forces 1e-5 eV/A, equal iteration counts
TOL_E = 1.0e-6
TOL_DM = 1.0e-8
TOL_F = 1.0e-5
