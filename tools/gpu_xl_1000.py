"""BASELINE configs[2] at full size: XL-BOMD NVE, 1000 steps, 1024 coronene replicas, SP2 density (eps 1e-5), k = 6,
dt = 0.4 fs, 300 K.  Prints throughput and the conservation of E(total) = E(potential, shadow) + E(kinetic)."""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
import pyseqm_b200 as seqm
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
nrep, nsteps = 1024, 1000
s, c = seqm.read_xyz([os.path.join(ROOT, "tests/golden/xyz/coronene.xyz")] * nrep)
sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [2], "sp2": [True, 1.0e-5]}
const = seqm.Constants().to(dev)
torch.manual_seed(0)
mol = seqm.Molecule(const, sp, torch.as_tensor(c, device=dev), torch.as_tensor(s, device=dev))
md = seqm.XL_BOMD(xl_bomd_params={"k": 6}, seqm_parameters=sp, timestep=0.4, Temp=300.0)
md.initialize(mol)
E = []
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(nsteps):
    md._do_integrator_step(i, mol, dict())
    if i % 10 == 0:
        E.append((mol.Etot + md._kinetic_energy(mol)).clone())
torch.cuda.synchronize(); dt = time.perf_counter() - t0
E = torch.stack(E).cpu()  # (samples, replicas)
dev_ = (E - E[0]).abs().max(dim=0).values
slope = (E[-10:].mean(dim=0) - E[:10].mean(dim=0)) / (0.4e-3 * (nsteps - 100))  # eV per ps
out = {"replicas": nrep, "steps": nsteps, "seconds": dt, "replica_steps_per_s": nrep * nsteps / dt, "ms_per_step": dt / nsteps * 1e3,
       "max_abs_dE_total_eV": float(dev_.max()), "mean_abs_dE_total_eV": float(dev_.mean()),
       "drift_eV_per_ps_mean": float(slope.mean()), "drift_eV_per_ps_max_abs": float(slope.abs().max()),
       "temperature_K_final_mean": float(md._temperature(mol).mean()) if hasattr(md, "_temperature") else None,
       "finite": bool(torch.isfinite(E).all())}
print(json.dumps(out))
