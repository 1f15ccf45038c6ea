"""End-to-end single point: the oracle's counterpart of Electronic_Structure.forward(dm_prop="SCF").

Restates the orchestration of seqm/basics.py:813-1244 (Energy.forward, ground-state branch),
1260-1365 (Force.forward) and seqm/ElectronicStructure.py:57-127.
"""
import numpy as np

from .density import density_from_fock
from .energy import elec_energy, isolated_atom_energy, molecule_sums, pair_nuclear_energy
from .gradient import hf_gradient
from .hamiltonian import build_fock, build_hcore, initial_density
from .integrals import atom_multipoles, rho0_eff
from .parser import parse
from .scf import run_scf
from .tables import Tables, method_parameters

_SYM = {"H": 1, "He": 2, "Li": 3, "Be": 4, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9, "Ne": 10, "Na": 11, "Mg": 12,
        "Al": 13, "Si": 14, "P": 15, "S": 16, "Cl": 17, "Ar": 18}  # fmt: skip


def read_xyz(files):
    """seqm/seqm_functions/read_xyz.py:17-54 (sort=True: stable sort by descending Z, zero padding)."""
    mols = []
    for fn in files:
        with open(fn) as f:
            lines = f.readlines()
        n = int(lines[0])
        rows = []
        for L in lines[2 : 2 + n]:
            a, *xyz = L.split()
            z = int(a) if a.isdigit() else _SYM[a]
            rows.append([z] + [float(t) for t in xyz[:3]])
        d = np.asarray(rows, dtype=np.float64)
        d = d[np.argsort(-d[:, 0], kind="stable")]
        mols.append(d)
    K = max(m.shape[0] for m in mols)
    species = np.zeros((len(mols), K), dtype=np.int64)
    coords = np.zeros((len(mols), K, 3))
    for i, m in enumerate(mols):
        species[i, : m.shape[0]] = m[:, 0].astype(np.int64)
        coords[i, : m.shape[0]] = m[:, 1:]
    return species, coords


def single_point(species, coordinates, seqm_parameters, P0=None, do_force=True, charges=0, learned_parameters=None):
    """Returns a dict with the result contract of SURVEY 8(a15)."""
    T = Tables.get()
    method = seqm_parameters["method"]
    table = method
    if method == "PM6":
        # elements without a d shell (basics.py:240-269: not in the nSuperHeavy set) go through the same arithmetic
        # as PM6_SP with the PM6 parameter file; the reference only pads every atom to 9 orbital slots
        s = np.asarray(species)
        d_shell = (((s > 12) & (s < 18)) | ((s > 20) & (s < 30)) | ((s > 32) & (s < 36)) | ((s > 38) & (s < 48))
                   | ((s > 50) & (s < 54)) | ((s > 70) & (s < 80)) | (s == 57))  # fmt: skip
        if d_shell.any():
            from .pm6d import single_point_pm6d

            if learned_parameters:
                raise NotImplementedError("oracle: learned parameters with PM6 d-shell elements are not covered")
            return single_point_pm6d(species, coordinates, seqm_parameters, P0=P0, do_force=do_force, charges=charges)
        method = "PM6_SP"
    if method not in ("MNDO", "AM1", "PM3", "PM6_SP"):
        raise NotImplementedError(f"oracle covers MNDO/AM1/PM3/PM6_SP, not {method}")
    eps = float(seqm_parameters["scf_eps"])
    conv = seqm_parameters.get("scf_converger", [2])
    sp2 = seqm_parameters.get("sp2", [False])
    P = parse(species, coordinates, charges=charges, outer_cutoff=seqm_parameters.get("pair_outer_cutoff", 1.0e10))
    par = method_parameters(table, P.Z)
    for name in seqm_parameters.get("learned", []):  # basics.py:442-448: only the names listed in `learned` are taken
        par[name] = np.asarray(learned_parameters[name], dtype=np.float64)
    mp = atom_multipoles(P.Z, par)
    hc = build_hcore(P, par, mp)
    H, w = hc["H"], hc["w"]
    D0 = initial_density(P) if P0 is None else np.asarray(P0, dtype=np.float64)
    if conv[0] == 3:  # scf_forward3 (scf_loop.py:1135-1381): SCF by Krylov-subspace-approximated Newton steps
        from .ksa import scf_ksa

        D, notconv, n_iter = scf_ksa(P, par, H, w, D0, eps, conv[1])
    else:
        D, notconv, n_iter = run_scf(P, par, H, w, D0, eps, conv, sp2)
    F = build_fock(P, par, H, w, D)
    _, e_mo, V = density_from_fock(F, P.nHeavy, P.nHydro, P.nocc, want_eig=True)
    Eelec = elec_energy(D, F, H)
    EnucAB = pair_nuclear_energy(method, P.ni, P.nj, P.idxi, P.idxj, P.rij, w[:, 0, 0], par, rho0=rho0_eff(par, mp))
    Enuc = molecule_sums(EnucAB, P.pair_molid, P.nmol)
    Etot = Eelec + Enuc
    Eiso = molecule_sums(isolated_atom_energy(P.Z, par), P.atom_molid, P.nmol)
    Hf = Etot - Eiso
    if seqm_parameters.get("Hf_flag", True):
        Hf = Hf + molecule_sums(T.eheat[P.Z], P.atom_molid, P.nmol)
    ar = np.arange(P.nmol)
    e_gap = e_mo[ar, P.nocc] - e_mo[ar, P.nocc - 1]
    q = T.tore[P.species] - np.diagonal(D, axis1=1, axis2=2).reshape(P.nmol, P.molsize, 4).sum(axis=2)
    out = dict(
        Etot=Etot, Hf=Hf, Eelec=Eelec, Enuc=Enuc, Eiso=Eiso, e_mo=e_mo, e_gap=e_gap, dm=D, F=F, H=H, w=w,
        q=q, notconverged=notconv, n_scf_iter=n_iter, molecular_orbitals=V, parsed=P,
    )  # fmt: skip
    if do_force:
        out["force"] = -hf_gradient(P, par, method, D, mp)
    if table == "PM6":  # widen to the 9-slot layout (packd.py:195-218)
        m = P.molsize
        wide = np.zeros((P.nmol, m, 9, m, 9))
        wide[:, :, :4, :, :4] = D.reshape(P.nmol, m, 4, m, 4)
        out["dm"] = wide.reshape(P.nmol, 9 * m, 9 * m)
        e9 = np.zeros((P.nmol, 9 * m))
        e9[:, : e_mo.shape[1]] = e_mo
        out["e_mo"] = e9
    return out
