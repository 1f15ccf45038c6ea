"""One bench step (configs[1]: 4096 QM9-size molecules, PM3, DIIS, forces) for ncu captures:
    ncu ... python tools/profile_step.py [nmol] [nsteps]
"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import pyseqm_b200 as seqm  # noqa: E402

nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
species, coords, sha = bench.workload(nmol, 0)
const = seqm.Constants().to(dev)
mol = seqm.Molecule(const, dict(bench.SP), torch.as_tensor(coords, device=dev), torch.as_tensor(species, device=dev))
mol.verbose = False
es = seqm.Electronic_Structure(dict(bench.SP))
for _ in range(nsteps):
    es(mol)
torch.cuda.synchronize()
print("done", mol.n_scf_iter, float(mol.Etot.sum()))
