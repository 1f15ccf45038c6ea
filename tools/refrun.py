"""Helpers that run the UNMODIFIED reference (/root/reference) in the build container.
Only tools/ scripts import this; nothing here exists on the GPU box."""
import contextlib
import io
import os
import re
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle", "h5py_stub"))
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore", category=SyntaxWarning)

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from seqm.ElectronicStructure import Electronic_Structure  # noqa: E402
from seqm.Molecule import Molecule  # noqa: E402
from seqm.seqm_functions.constants import Constants  # noqa: E402
from seqm.seqm_functions.read_xyz import read_xyz  # noqa: E402,F401


def run_reference(species, coordinates, seqm_parameters, P0=None, threads=None):
    """Run Electronic_Structure.forward on CPU; returns dict of numpy results + n_scf_iter
    (parsed from the reference's own verbose line, scf_loop.py:329-346/616-633/975-992)."""
    if threads:
        torch.set_num_threads(threads)
    sp = dict(seqm_parameters)
    species = torch.as_tensor(np.asarray(species), dtype=torch.int64)
    coordinates = torch.as_tensor(np.asarray(coordinates), dtype=torch.float64)
    const = Constants()
    mol = Molecule(const, sp, coordinates, species)
    es = Electronic_Structure(sp)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        es(mol, P0=None if P0 is None else torch.as_tensor(P0).clone())
    txt = buf.getvalue()
    m = re.findall(r"scf [a-z ]+:\s+(\d+) \|", txt)
    out = dict(
        Etot=mol.Etot, Hf=mol.Hf, Eelec=mol.Eelec, Enuc=mol.Enuc, Eiso=mol.Eiso, e_mo=mol.e_mo, e_gap=mol.e_gap,
        dm=mol.dm, q=mol.q, force=mol.force, notconverged=es.notconverged, w=mol.w,
        molecular_orbitals=mol.molecular_orbitals, charge=es.charge, dipole=mol.dipole,
    )  # fmt: skip
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}
    out["n_scf_iter"] = int(m[-1]) if m else -1
    out["stdout"] = txt
    return out
