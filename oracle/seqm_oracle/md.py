"""XL-BOMD and plain BOMD steps (oracle).

Restates seqm/dynamics/xlbomd.py:73-570 (EnergyXL / ForceXL, eigensolver branch 361-365: one Fock build from the
field P, density D from F, shadow energy elec_energy_xl, force at fixed D and P) and the integrators
seqm/MolecularDynamics.py: Molecular_Dynamics_Basic.one_step 813-842, XL_BOMD.__init__ 1330-1371 (Niklasson
coefficients), _propagate_P 1418-1427 (c = 0.95), one_step 1437-1515, initialize 1530-1605.
"""
import numpy as np

from .density import density_from_fock, sp2_density
from .energy import elec_energy_xl, isolated_atom_energy, molecule_sums, pair_nuclear_energy
from .gradient import hf_gradient
from .hamiltonian import build_fock, build_hcore
from .integrals import atom_multipoles, rho0_eff
from .parser import parse
from .tables import Tables, method_parameters

ACC_SCALE = 0.009648532800137615  # eV/A/(g/mol) -> A/fs^2   (MolecularDynamics.py:27)
KINETIC_ENERGY_SCALE = 1.0364270099032438e2  # amu (A/fs)^2 -> eV  (MolecularDynamics.py:29)

XL_COEFFS = {  # kappa, alpha, c0..ck   (MolecularDynamics.py:1337-1345; Niklasson et al. JCP 130, 214109)
    3: [1.69, 150e-3, -2.0, 3.0, 0.0, -1.0],
    4: [1.75, 57e-3, -3.0, 6.0, -2.0, -2.0, 1.0],
    5: [1.82, 18e-3, -6.0, 14.0, -8.0, -3.0, 4.0, -1.0],
    6: [1.84, 5.5e-3, -14.0, 36.0, -27.0, -2.0, 12.0, -6.0, 1.0],
    7: [1.86, 1.6e-3, -36.0, 99.0, -88.0, 11.0, 32.0, -25.0, 8.0, -1.0],
    8: [1.88, 0.44e-3, -99.0, 286.0, -286.0, 78.0, 78.0, -90.0, 42.0, -10.0, 1.0],
    9: [1.89, 0.12e-3, -286.0, 858.0, -936.0, 364.0, 168.0, -300.0, 184.0, -63.0, 12.0, -1.0],
}


def xl_coefficients(k):
    kappa, alpha = XL_COEFFS[k][0], XL_COEFFS[k][1]
    tmp = np.asarray(XL_COEFFS[k][2:], dtype=np.float64) * alpha
    tmp[0] += 2.0 - kappa
    tmp[1] -= 1.0
    return kappa, np.concatenate([tmp, tmp])


def xl_forward(species, coordinates, seqm_parameters, field):
    """Electronic_Structure.forward(dm_prop="XL-BOMD", P0=field) -> dict(force, dm (=D), Etot, Hf, Eelec, Enuc, e_gap)."""
    T = Tables.get()
    method = seqm_parameters["method"]
    sp2 = seqm_parameters.get("sp2", [False])
    P = parse(species, coordinates)
    par = method_parameters(method, P.Z)
    mp = atom_multipoles(P.Z, par)
    hc = build_hcore(P, par, mp)
    H, w = hc["H"], hc["w"]
    F = build_fock(P, par, H, w, field)
    if sp2[0]:
        D = sp2_density(F, P.nHeavy, P.nHydro, P.nocc, sp2[1])
        e_gap = np.zeros(P.nmol)
    else:
        D, e, _ = density_from_fock(F, P.nHeavy, P.nHydro, P.nocc)
        ar = np.arange(P.nmol)
        e_gap = e[ar, P.nocc] - e[ar, P.nocc - 1]
    Eelec = elec_energy_xl(D, field, F, H)
    EnucAB = pair_nuclear_energy(method, P.ni, P.nj, P.idxi, P.idxj, P.rij, w[:, 0, 0], par, rho0=rho0_eff(par, mp))
    Enuc = molecule_sums(EnucAB, P.pair_molid, P.nmol)
    Etot = Eelec + Enuc
    Eiso = molecule_sums(isolated_atom_energy(P.Z, par), P.atom_molid, P.nmol)
    Hf = Etot - Eiso + molecule_sums(T.eheat[P.Z], P.atom_molid, P.nmol)
    force = -hf_gradient(P, par, method, D, mp, field=field)
    return dict(force=force, dm=D, Etot=Etot, Hf=Hf, Eelec=Eelec, Enuc=Enuc, Eiso=Eiso, e_gap=e_gap)


def kinetic_energy(mass, vel):
    return np.sum(0.5 * mass * vel**2, axis=(1, 2)) * KINETIC_ENERGY_SCALE


def run_md(species, coordinates, velocities, seqm_parameters, timestep, steps, k=None):
    """k=None: velocity-Verlet BOMD with an SCF per step restarted from the previous density;
    k=3..9: XL-BOMD.  Returns per-step Etot, Ek and the final state."""
    from .api import single_point

    T = Tables.get()
    species = np.asarray(species)
    x = np.array(coordinates, dtype=np.float64)
    v = np.array(velocities, dtype=np.float64)
    mass = T.mass[species][:, :, None]
    minv = np.where(species[:, :, None] > 0, 1.0 / np.where(mass > 0, mass, 1.0), 0.0)
    r = single_point(species, x, seqm_parameters)
    dm = r["dm"]
    force = r["force"]
    acc = force * minv * ACC_SCALE
    if k is not None:
        m = k + 1
        kappa, coeff = xl_coefficients(k)
        Pf = dm.copy()
        Pt = np.stack([dm.copy() for _ in range(m)])
    Etot, Ek = [], []
    for i in range(steps):
        v += 0.5 * acc * timestep
        x += v * timestep
        if k is None:
            r = single_point(species, x, seqm_parameters, P0=dm)
            dm = r["dm"]
        else:
            c = 0.95
            cindx = i % m
            Pf = kappa * (c * dm + (1.0 - c) * Pf) + np.tensordot(coeff[cindx : cindx + m], Pt, axes=(0, 0))
            Pt[m - 1 - cindx] = Pf
            r = xl_forward(species, x, seqm_parameters, Pf)
            dm = r["dm"]
        force = r["force"]
        acc = force * minv * ACC_SCALE
        v += 0.5 * acc * timestep
        Etot.append(r["Etot"].copy())
        Ek.append(kinetic_energy(mass, v))
    return dict(Etot=np.stack(Etot), Ek=np.stack(Ek), coordinates=x, velocities=v, force=force, dm=dm)
