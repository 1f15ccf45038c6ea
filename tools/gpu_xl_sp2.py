"""XL-BOMD step rate with the two density routes (Jacobi eigensolver / in-SM SP2) on coronene replicas."""
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
const = seqm.Constants().to(dev)
for sp2 in ([False], [True, 1.0e-5]):
    sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [2], "sp2": sp2}
    torch.manual_seed(0)
    mol = seqm.Molecule(const, sp, torch.as_tensor(c, device=dev), torch.as_tensor(s, device=dev))
    md = seqm.XL_BOMD(xl_bomd_params={"k": 6}, seqm_parameters=sp, timestep=0.4, Temp=300.0)
    md.initialize(mol)
    E0 = (mol.Etot + md._kinetic_energy(mol)).clone() if hasattr(md, "_kinetic_energy") else None
    for i in range(3): md._do_integrator_step(i, mol, dict())
    torch.cuda.synchronize()
    lib.profile_enable(True)
    t = time.perf_counter()
    for i in range(3, 53): md._do_integrator_step(i, mol, dict())
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    prof = lib.profile_collect(); lib.profile_enable(False)
    E1 = mol.Etot + md._kinetic_energy(mol)
    print("sp2=%s: %.2f ms/step  %.0f replica-steps/s  max |dE_total| over 53 steps %.2e eV" % (sp2, dt * 20, nrep * 50 / dt, float((E1 - E0).abs().max())))
    for k, v in prof.items():
        if v[1]: print("    %-18s %9.3f ms/step" % (k, v[0] / 50))
