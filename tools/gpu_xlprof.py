import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
import pyseqm_b200 as seqm
from pyseqm_b200._lib import get_lib
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0"); lib = get_lib()
nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
s, c = seqm.read_xyz([os.path.join(ROOT, "tests/golden/xyz/coronene.xyz")] * nrep)
sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [2], "sp2": [False]}
const = seqm.Constants().to(dev)
torch.manual_seed(0)
mol = seqm.Molecule(const, sp, torch.as_tensor(c, device=dev), torch.as_tensor(s, device=dev))
md = seqm.XL_BOMD(xl_bomd_params={"k": 6}, seqm_parameters=sp, timestep=0.4, Temp=300.0)
md.initialize(mol)
for i in range(3): md._do_integrator_step(i, mol, dict())
torch.cuda.synchronize()
lib.profile_enable(True); lib.jacobi_stats()
t = time.perf_counter()
for i in range(3, 13): md._do_integrator_step(i, mol, dict())
torch.cuda.synchronize(); dt = time.perf_counter() - t
prof = lib.profile_collect(); lib.profile_enable(False)
print("10 XL-BOMD steps: %.2f ms/step" % (dt * 100)); st = lib.jacobi_stats(); print("sweeps/mol/step", st["sweeps"] / max(st["molecules"], 1))
for k, v in prof.items():
    if v[1]: print("  %-18s %9.3f ms/step n=%d" % (k, v[0] / 10, v[1]))
