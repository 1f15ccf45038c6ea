#!/usr/bin/env python
"""Golden KSA-XL-BOMD fixtures from the UNMODIFIED reference (KSA_XL_BOMD, MolecularDynamics.py:1608-1619; Krylov branch of
EnergyXL.forward, xlbomd.py:201-341): trajectories driven step by step, plus operator-level fixtures of Fermi_Q
(fermi_q.py:8-72) and Canon_DM_PRT (canon_dm_prt.py:6-39) on a mixed batch.

    python tools/make_golden_ksa.py        # build container only
Writes tests/golden/md_ksa_*.npz and tests/golden/ksa_operators.npz.
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
from seqm.MolecularDynamics import KSA_XL_BOMD  # noqa: E402
from seqm.seqm_functions.constants import Constants  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")
XYZ = os.path.join(GOLD, "xyz")
OUT = {"molid": [0], "prefix": "/tmp/seqm_ksa_golden", "print every": 0, "checkpoint every": 0, "xyz": 0, "h5": {}}


def run(name, files, sp, steps, timestep, temp, xl, seed=0):
    species, coords = read_xyz([os.path.join(XYZ, f) for f in files])
    species = torch.as_tensor(species, dtype=torch.int64)
    coords = torch.as_tensor(coords, dtype=torch.float64)
    torch.manual_seed(seed)
    sp = dict(sp)
    mol = Molecule(Constants(), sp, coords.clone(), species)
    md = KSA_XL_BOMD(xl_bomd_params=dict(xl), damp=None, seqm_parameters=sp, timestep=timestep, Temp=temp, output=dict(OUT))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        md.initialize(mol, remove_com=None)
        v0 = mol.velocities.detach().clone().numpy()
        x0 = mol.coordinates.detach().clone().numpy()
        E0 = mol.Etot.detach().clone().numpy()
        Etot, Ek, Ent, Err = [], [], [], []
        for i in range(steps):
            md._do_integrator_step(i, mol, dict())
            if torch.is_tensor(mol.coordinates.grad):
                mol.coordinates.grad.zero_()
            Ek.append(md._kinetic_energy(mol).detach().numpy().copy())
            Etot.append(mol.Etot.detach().numpy().copy())
            Ent.append(mol.Electronic_entropy.detach().numpy().copy())
            Err.append(mol.Krylov_Error.detach().numpy().copy())
    out = dict(
        species=species.numpy(), coordinates0=x0, velocities0=v0, Etot0=E0, Etot=np.stack(Etot), Ek=np.stack(Ek),
        Electronic_entropy=np.stack(Ent), Krylov_Error=np.stack(Err),
        coordinates=mol.coordinates.detach().numpy(), velocities=mol.velocities.detach().numpy(),
        force=mol.force.detach().numpy(), dm=mol.dm.detach().numpy(), dP2dt2=mol.dP2dt2.detach().numpy(),
        Fermi_occ=mol.Fermi_occ.detach().numpy(), e_gap=mol.e_gap.detach().numpy(),
        timestep=timestep, temp=temp, steps=steps, xl_bomd_params=json.dumps(xl),
        seqm_parameters=json.dumps({k_: v_ for k_, v_ in sp.items() if k_ != 'elements'}),
    )  # fmt: skip
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    tot = np.stack(Etot) + np.stack(Ek) + np.stack(Ent)
    print(name, "steps", steps, "E(total) first/last", tot[0], tot[-1], "Krylov error last", Err[-1])


def operators():
    """Fermi_Q and Canon_DM_PRT of the reference on the converged Fock matrices of {methane, benzene, toluene} (AM1)."""
    from seqm.seqm_functions.canon_dm_prt import Canon_DM_PRT
    from seqm.seqm_functions.fermi_q import Fermi_Q
    from seqm.seqm_functions.fock import fock
    from seqm.seqm_functions.G_XL_LR import G
    from seqm.seqm_functions.hcore import hcore
    from seqm.ElectronicStructure import Electronic_Structure

    species, coords = read_xyz([os.path.join(XYZ, f) for f in ("methane.xyz", "benzene.xyz", "toluene.xyz")])
    sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [2]}
    mol = Molecule(Constants(), sp, torch.as_tensor(coords), torch.as_tensor(species, dtype=torch.int64))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        Electronic_Structure(sp)(mol)
    M, w, rho0i, rho0j, riXH, ri = hcore(mol)
    par = mol.parameters
    W0 = torch.tensor([0])
    args = (w, W0, par["g_ss"], par["g_pp"], par["g_sp"], par["g_p2"], par["h_sp"], mol.method, par["s_orb_exp_tail"],
            par["p_orb_exp_tail"], par["d_orb_exp_tail"], mol.Z, par["F0SD"], par["G2SD"])  # fmt: skip
    F = fock(mol.nmol, mol.molsize, mol.dm, M, mol.maskd, mol.mask, mol.idxi, mol.idxj, *args)
    kB, T = 8.61739e-5, 1500.0
    D0, S, QQ, e, Fe, mu, occ_mask = Fermi_Q(F, T, mol.nocc, mol.nHeavy, mol.nHydro, kB, scf_backward=0)
    g = torch.Generator().manual_seed(3)
    X = torch.randn(mol.dm.shape, generator=g, dtype=torch.float64) * (mol.dm != 0)
    X = 0.5 * (X + X.transpose(1, 2))
    X = X / torch.linalg.norm(X, ord="fro", dim=(1, 2), keepdim=True)
    FO1 = G(mol.nmol, mol.molsize, X, M, mol.maskd, mol.mask, mol.idxi, mol.idxj, *args)
    PO1 = Canon_DM_PRT(FO1, T, mol.nHeavy, mol.nHydro, QQ, e, mu, 10, kB, occ_mask)
    t = lambda x: x.detach().numpy()  # noqa: E731
    Th = 20000.0  # fractional occupations and a non-zero entropy
    D0h, Sh, QQh, eh, Feh, muh, occ_mask_h = Fermi_Q(F, Th, mol.nocc, mol.nHeavy, mol.nHydro, kB, scf_backward=0)
    PO1h = Canon_DM_PRT(FO1, Th, mol.nHeavy, mol.nHydro, QQh, eh, muh, 10, kB, occ_mask_h)
    np.savez_compressed(os.path.join(GOLD, "ksa_operators.npz"), species=np.asarray(species), coordinates=np.asarray(coords),
                        F=t(F), T_el=T, kB=kB, D0=t(D0), S=t(S), e=t(e), Fe=t(Fe), mu=t(mu), X=t(X), FO1=t(FO1), PO1=t(PO1),
                        T_hot=Th, D0_hot=t(D0h), S_hot=t(Sh), Fe_hot=t(Feh), mu_hot=t(muh), PO1_hot=t(PO1h),
                        dm=t(mol.dm), nocc=t(mol.nocc), seqm_parameters=json.dumps(sp))  # fmt: skip
    print("ksa_operators: S", t(S), "S_hot", t(Sh), "mu", t(mu).ravel(), "|PO1|", float(PO1.abs().max()))


def cis_sigma():
    """makeA_pi_batched (rcis_batch.py:296-403) of the reference on two methanal geometries, three random non-symmetric AO
    matrices per molecule, both symmetry modes."""
    from seqm.seqm_functions.hcore import hcore
    from seqm.seqm_functions.rcis_batch import makeA_pi_batched

    species, coords = read_xyz([os.path.join(XYZ, f) for f in ("methanal.1.xyz", "methanal.2.xyz")])
    sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [2]}
    mol = Molecule(Constants(), sp, torch.as_tensor(coords), torch.as_tensor(species, dtype=torch.int64))
    M, w, *_ = hcore(mol)
    norb = int(mol.norb[0])
    g = torch.Generator().manual_seed(5)
    X = torch.randn(mol.nmol, 3, norb, norb, generator=g, dtype=torch.float64)
    F = makeA_pi_batched(mol, X.clone(), w, allSymmetric=False)
    Xs = 0.5 * (X + X.transpose(2, 3))
    Fs = makeA_pi_batched(mol, Xs.clone(), w, allSymmetric=True)
    t = lambda x: x.detach().numpy()  # noqa: E731
    np.savez_compressed(os.path.join(GOLD, "cis_sigma_methanal.npz"), species=np.asarray(species), coordinates=np.asarray(coords),
                        seqm_parameters=json.dumps({k_: v_ for k_, v_ in sp.items() if k_ != 'elements'}), X=t(X), F=t(F),
                        Xs=t(Xs), Fs=t(Fs))  # fmt: skip
    print("cis_sigma_methanal |F|", float(F.abs().max()), "antisymmetric part", float((F - F.transpose(2, 3)).abs().max()))


def scf_ksa_cases():
    """scf_converger = [3, {...}] (scf_forward3, scf_loop.py:1135-1381) single points; the iteration count is the number in the
    reference's own verbose line "scf KSA step : N"."""
    import re

    from seqm.ElectronicStructure import Electronic_Structure

    xl = {"max_rank": 3, "err_threshold": 0.0, "T_el": 1500}
    for name, files in (("ksa_scf_mixed", ["methane.xyz", "benzene.xyz"]), ("ksa_scf_methanal", ["methanal.1.xyz", "methanal.2.xyz"])):
        species, coords = read_xyz([os.path.join(XYZ, f) for f in files])
        sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [3, dict(xl)]}
        sp_in = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [3, dict(xl)]}  # the reference adds 'elements' to it
        mol = Molecule(Constants(), sp_in, torch.as_tensor(coords), torch.as_tensor(species, dtype=torch.int64))
        mol.verbose = True
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            Electronic_Structure(sp_in)(mol)
        n_iter = int(re.findall(r"scf KSA step\s*:\s+(\d+) \|", buf.getvalue())[-1])
        t = lambda x: x.detach().numpy()  # noqa: E731
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), species=np.asarray(species), coordinates=np.asarray(coords),
                            seqm_parameters=json.dumps(sp), n_scf_iter=n_iter, Etot=t(mol.Etot), Eelec=t(mol.Eelec), Hf=t(mol.Hf),
                            dm=t(mol.dm), force=t(mol.force), e_gap=t(mol.e_gap), q=t(mol.q))  # fmt: skip
        print(name, "iterations", n_iter, "Etot", t(mol.Etot))


if __name__ == "__main__":
    if os.environ.get("GOLDEN_ONLY") == "scf":
        scf_ksa_cases()
        sys.exit(0)
    if os.environ.get("GOLDEN_ONLY") == "cis":
        cis_sigma()
        sys.exit(0)
    scf_ksa_cases()
    cis_sigma()
    operators()
    if os.environ.get("GOLDEN_ONLY") == "operators":
        sys.exit(0)
    sp = {"method": "AM1", "scf_eps": 1.0e-7, "scf_converger": [1]}
    run("md_ksa_methane_k6", ["methane.xyz"], sp, steps=30, timestep=0.5, temp=300.0,
        xl={"k": 6, "max_rank": 3, "err_threshold": 0.0, "T_el": 1500})
    run("md_ksa_mixed_k4", ["methane.xyz", "benzene.xyz"], sp, steps=20, timestep=0.5, temp=300.0,
        xl={"k": 4, "max_rank": 3, "err_threshold": 0.0, "T_el": 1500})
    run("md_ksa_benzene_thr", ["benzene.xyz"], sp, steps=12, timestep=0.5, temp=300.0,
        xl={"k": 6, "max_rank": 4, "err_threshold": 0.05, "T_el": 3000})
