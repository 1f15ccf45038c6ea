"""Measured margins of the SP2-route C380 forward against the reference fixture (tests/golden/cfg4_C380_AM1_sp2.npz):
what tests/test_large_molecule.py::test_gpu_c380_against_reference asserts, printed.  python tools/c380_margins.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402
from helpers import cuda_lib, run_molecule  # noqa: E402

g = load_golden("cfg4_C380_AM1_sp2")
mol, es = run_molecule(cuda_lib(), torch.device("cuda:0"), g["species"], g["coordinates"], g["seqm_parameters"])
print("n_scf_iter", mol.n_scf_iter, "reference", int(g["n_scf_iter"]))
print("dEtot", abs(float(mol.Etot[0]) - float(g["Etot"][0])))
print("dEnuc", abs(float(mol.Enuc[0]) - float(g["Enuc"][0])))
print("dForce", np.abs(mol.force.cpu().numpy() - g["force"]).max())
print("dq", np.abs(mol.q.cpu().numpy() - g["q"]).max())
print("dgap", abs(float(mol.e_gap[0]) - float(g["e_gap"][0])))
if "dm" in g:
    print("dP", np.abs(mol.dm.cpu().numpy() - g["dm"]).max())
