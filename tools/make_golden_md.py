#!/usr/bin/env python
"""Golden MD fixtures from the UNMODIFIED reference: XL-BOMD (eigensolver branch, xlbomd.py:361) and plain
BOMD velocity-Verlet trajectories driven step by step (no HDF5/XYZ output).

    python tools/make_golden_md.py        # build container only
Writes tests/golden/md_xl_bomd_*.npz, tests/golden/md_basic_*.npz with: inputs, velocities at t=0 (after the
reference's own initialisation), and per-step Etot, kinetic energy, plus final coordinates / velocities / force / dm.
"""
import contextlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from refrun import np, read_xyz, torch  # noqa: E402

from seqm.Molecule import Molecule  # noqa: E402
from seqm.MolecularDynamics import XL_BOMD, Molecular_Dynamics_Basic  # noqa: E402
from seqm.seqm_functions.constants import Constants  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")
XYZ = os.path.join(GOLD, "xyz")
OUT = {"molid": [0], "prefix": "/tmp/seqm_md_golden", "print every": 0, "checkpoint every": 0, "xyz": 0, "h5": {}}


def run(name, files, sp, steps, timestep, temp, k=None, jitter_copies=0, seed=0):
    species, coords = read_xyz([os.path.join(XYZ, f) for f in files])
    species = torch.as_tensor(species, dtype=torch.int64)
    coords = torch.as_tensor(coords, dtype=torch.float64)
    torch.manual_seed(seed)
    sp = dict(sp)  # the reference mutates and shares this dict (Molecule adds 'elements')
    mol = Molecule(Constants(), sp, coords.clone(), species)
    if k is not None:
        md = XL_BOMD(xl_bomd_params={"k": k}, damp=None, seqm_parameters=sp, timestep=timestep, Temp=temp, output=dict(OUT))
    else:
        md = Molecular_Dynamics_Basic(seqm_parameters=sp, timestep=timestep, Temp=temp, output=dict(OUT))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        md.initialize(mol, remove_com=None)
        v0 = mol.velocities.detach().clone().numpy()
        x0 = mol.coordinates.detach().clone().numpy()
        E0 = mol.Etot.detach().clone().numpy()
        Etot, Ek = [], []
        for i in range(steps):
            md._do_integrator_step(i, mol, dict())
            if torch.is_tensor(mol.coordinates.grad):
                mol.coordinates.grad.zero_()
            Ek.append(md._kinetic_energy(mol).detach().numpy().copy())
            Etot.append(mol.Etot.detach().numpy().copy())
    out = dict(
        species=species.numpy(), coordinates0=x0, velocities0=v0, Etot0=E0, Etot=np.stack(Etot), Ek=np.stack(Ek),
        coordinates=mol.coordinates.detach().numpy(), velocities=mol.velocities.detach().numpy(),
        force=mol.force.detach().numpy(), dm=mol.dm.detach().numpy(), timestep=timestep, temp=temp,
        k=-1 if k is None else k, steps=steps, seqm_parameters=json.dumps({k_: v_ for k_, v_ in sp.items() if k_ != 'elements'}),
    )  # fmt: skip
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    drift = (np.stack(Etot) + np.stack(Ek))
    print(name, "steps", steps, "E(total) first/last", drift[0], drift[-1])


if __name__ == "__main__":
    sp = {"method": "AM1", "scf_eps": 1.0e-7, "scf_converger": [1]}
    run("md_xl_bomd_methane_k6", ["methane.xyz"], sp, steps=30, timestep=0.5, temp=300.0, k=6)
    run("md_xl_bomd_mixed_k4", ["methane.xyz", "benzene.xyz"], sp, steps=20, timestep=0.5, temp=300.0, k=4)
    sp2 = {"method": "AM1", "scf_eps": 1.0e-7, "scf_converger": [2]}
    run("md_xl_bomd_coronene_k6", ["coronene.xyz"], sp2, steps=10, timestep=0.4, temp=300.0, k=6)
    run("md_basic_methanal", ["methanal.1.xyz", "methanal.2.xyz"], sp, steps=10, timestep=0.5, temp=300.0)
