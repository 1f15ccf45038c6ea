"""One C380 forward (configs[3]) for ncu captures:  ncu ... python tools/profile_c380.py"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pyseqm_b200 as seqm  # noqa: E402

torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
s, c = seqm.read_xyz([os.path.join(ROOT, "tests", "golden", "xyz", "C380.xyz")])
sp = {"method": "AM1", "scf_eps": 1.0e-6, "scf_converger": [2], "sp2": [True, 1.0e-5]}
const = seqm.Constants().to(dev)
mol = seqm.Molecule(const, dict(sp), torch.as_tensor(c, device=dev), torch.as_tensor(s, device=dev))
mol.verbose = False
es = seqm.Electronic_Structure(dict(sp))
es(mol)
torch.cuda.synchronize()
print("done", mol.n_scf_iter, float(mol.Etot[0]))
