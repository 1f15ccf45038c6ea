"""One PM6 d-orbital step (BASELINE configs[4] sample: 512 synthetic P/S/Cl organics) for ncu captures:
    ncu ... python tools/profile_pm6d.py [nmol] [nsteps]
"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pyseqm_b200 as seqm  # noqa: E402
from pyseqm_b200.synthetic import qm9_like_batch  # noqa: E402

nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
species, coords = qm9_like_batch(nmol, seed=0, hetero=(15, 16, 17))
sp = {"method": "PM6", "scf_eps": 1.0e-7, "scf_converger": [1], "sp2": [False], "b200_scf_max_iter": 40}
const = seqm.Constants().to(dev)
mol = seqm.Molecule(const, dict(sp), torch.as_tensor(coords, device=dev), torch.as_tensor(species, device=dev))
mol.verbose = False
es = seqm.Electronic_Structure(dict(sp))
import warnings  # noqa: E402

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for _ in range(nsteps):
        es(mol)
torch.cuda.synchronize()
print("done", mol.n_scf_iter, float(mol.Etot.sum()))
