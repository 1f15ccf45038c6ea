"""Measured margins of the SP2-route parity tests (what tests/test_large_molecule.py and test_sp2_route assert, printed):
    python tools/sp2_margins.py [cuda|hostemu]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from conftest import load_golden  # noqa: E402
from helpers import cuda_lib, golden_inputs, hostemu_lib, run_molecule  # noqa: E402
from test_large_molecule import stacked  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cuda"
lib, dev = (cuda_lib(), torch.device("cuda:0")) if which == "cuda" else (hostemu_lib(), torch.device("cpu"))


def diffs(mol, ref, keys):
    out = {}
    for k in keys:
        if k in ref:
            out[k] = float(np.abs(getattr(mol, k).detach().cpu().numpy() - ref[k]).max())
    return out


g = load_golden("cfg1_AM1_sp2")
charges, learned = golden_inputs(g, dev)
mol, es = run_molecule(lib, dev, g["species"], g["coordinates"], g["seqm_parameters"], charges=charges, learned=learned)
print("cfg1_AM1_sp2 n_scf_iter", mol.n_scf_iter, g["n_scf_iter"], diffs(mol, g, ("Etot", "Hf", "Eelec", "dm", "q", "e_mo", "e_gap", "force")))

import seqm_oracle as so  # noqa: E402

s2, c2 = stacked(3.5, 1.2)
sp = {"method": "AM1", "scf_eps": 1e-6, "scf_converger": [2], "sp2": [True, 1e-7]}
ref = so.single_point(s2, c2, sp)
mol, es = run_molecule(lib, dev, s2, c2, sp)
print("dimer n_scf_iter", mol.n_scf_iter, ref["n_scf_iter"], diffs(mol, ref, ("Etot", "dm", "force", "e_gap")))
if which == "cuda":
    g = load_golden("cfg4_C380_AM1_sp2")
    mol, es = run_molecule(lib, dev, g["species"], g["coordinates"], g["seqm_parameters"])
    print("C380 n_scf_iter", mol.n_scf_iter, g["n_scf_iter"], diffs(mol, g, ("Etot", "Enuc", "force", "q", "e_gap", "dm")))
