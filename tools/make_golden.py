#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

    python tools/make_golden.py            # in the build container only

Each .npz holds the inputs (species, coordinates, seqm_parameters as JSON) and the reference outputs
(Etot, Hf, Eelec, Enuc, Eiso, e_mo, e_gap, dm, q, force, notconverged, n_scf_iter) plus, for the
operator-level files, hcore()/fock()/sym_eig_trunc()/SP2() outputs (SURVEY 8(b) level B).
tests/golden/ref_json/*.json and tests/golden/xyz/*.xyz are verbatim copies of the reference's own test
fixtures (tests/reference/*.json, tests/data/*.xyz, examples/*.xyz).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
from refrun import np, read_xyz, run_reference, torch  # noqa: E402

import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("synthetic", os.path.join(HERE, "..", "pyseqm_b200", "synthetic.py"))
synthetic = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synthetic)

GOLD = os.path.join(HERE, "..", "tests", "golden")
XYZ = os.path.join(GOLD, "xyz")
KEEP = ["Etot", "Hf", "Eelec", "Enuc", "Eiso", "e_mo", "e_gap", "dm", "q", "force", "notconverged", "n_scf_iter", "dipole"]


def save(name, species, coords, sp, extra=None, drop=()):
    ref = run_reference(species, coords, sp)
    out = {k: ref[k] for k in KEEP if k not in drop}
    out["species"] = np.asarray(species, dtype=np.int64)
    out["coordinates"] = np.asarray(coords, dtype=np.float64)
    out["seqm_parameters"] = json.dumps(sp)
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: nmol={out['species'].shape[0]} iters={ref['n_scf_iter']} Etot[0]={ref['Etot'][0]:.10f}")
    return ref


def operator_level(species, coords, method):
    """hcore / fock / sym_eig_trunc / SP2 outputs of the reference on one batch."""
    from seqm.Molecule import Molecule
    from seqm.seqm_functions.constants import Constants
    from seqm.seqm_functions.diag import sym_eig_trunc
    from seqm.seqm_functions.fock import fock
    from seqm.seqm_functions.hcore import hcore
    from seqm.seqm_functions.SP2 import SP2
    from seqm.seqm_functions.pack import pack

    sp = {"method": method, "scf_eps": 1e-7, "scf_converger": [2]}
    mol = Molecule(Constants(), sp, torch.as_tensor(coords), torch.as_tensor(species, dtype=torch.int64))
    M, w, rho0i, rho0j, riXH, ri = hcore(mol)
    p = mol.parameters
    g = torch.Generator().manual_seed(7)
    nb = 4 * mol.molsize
    # a random symmetric "density" confined to the real orbitals, to exercise every Fock term
    X = torch.rand(mol.nmol, nb, nb, generator=g) - 0.5
    real = torch.zeros(mol.nmol, mol.molsize, 4, dtype=torch.bool)
    real[mol.species > 1] = True
    real[..., 0] |= mol.species == 1
    real = real.reshape(mol.nmol, nb)
    X = (X + X.transpose(1, 2)) * (real.unsqueeze(1) & real.unsqueeze(2))
    Wd = torch.tensor([0])
    F = fock(mol.nmol, mol.molsize, X, M, mol.maskd, mol.mask, mol.idxi, mol.idxj, w, Wd, p["g_ss"], p["g_pp"],
             p["g_sp"], p["g_p2"], p["h_sp"], method, p["zeta_s"], p["zeta_p"], p["zeta_d"], mol.Z, p["F0SD"], p["G2SD"])  # fmt: skip
    e, P, v = sym_eig_trunc(F, mol.nHeavy, mol.nHydro, mol.nocc)
    Psp2 = SP2(pack(F, mol.nHeavy, mol.nHydro), mol.nocc, 1.0e-5)
    t = lambda x: x.detach().numpy()  # noqa: E731
    return dict(op_M=t(M), op_w=t(w), op_rho0i=t(rho0i), op_rho0j=t(rho0j), op_riXH=t(riXH), op_ri=t(ri),
                op_X=t(X), op_F=t(F), op_e=t(e), op_P=t(P), op_sp2_packed=t(Psp2))  # fmt: skip


def main():
    cfg1 = [os.path.join(XYZ, f) for f in ("methane.xyz", "benzene.xyz", "toluene.xyz")]
    species, coords = read_xyz(cfg1)
    only = os.environ.get("GOLDEN_ONLY")  # e.g. GOLDEN_ONLY=PM6_SP regenerates just that method
    for method in ("AM1", "PM3", "MNDO", "PM6_SP"):
        if only and method != only:
            continue
        for tag, conv, eps in (("c2", [2], 1e-7), ("c1", [1], 1e-6), ("c0", [0, 0.3], 1e-7)):
            sp = {"method": method, "scf_eps": eps, "scf_converger": conv, "sp2": [False], "analytical_gradient": [True]}
            extra = operator_level(species, coords, method) if tag == "c2" else None
            save(f"cfg1_{method}_{tag}", species, coords, sp, extra)
    if only:
        if only == "PM6_SP":
            s4, c4 = synthetic.qm9_like_batch(24, seed=2)
            sp = {"method": "PM6_SP", "scf_eps": 1e-7, "scf_converger": [2], "sp2": [False], "analytical_gradient": [True]}
            save("cfg2_PM6_SP_24", s4, c4, sp, drop=("e_mo",))
        return
    # default (autograd) forces and the SP2 density route
    sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [2], "sp2": [False]}
    save("cfg1_AM1_autograd", species, coords, sp, drop=("dm", "e_mo"))
    sp = {"method": "AM1", "scf_eps": 1e-6, "scf_converger": [2], "sp2": [True, 1e-5], "analytical_gradient": [True]}
    save("cfg1_AM1_sp2", species, coords, sp)
    # the reference's own batch tests re-run with densities kept
    sp = {"method": "AM1", "scf_eps": 1e-6, "scf_converger": [1]}
    s2, c2 = read_xyz([os.path.join(XYZ, f) for f in ("methane.xyz", "benzene.xyz")])
    save("ref_batch_single_point_am1", s2, c2, sp)
    s3, c3 = read_xyz([os.path.join(XYZ, f"methanal.{i}.xyz") for i in (1, 2, 3)])
    sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [1], "analytical_gradient": [True]}
    save("ref_ground_force_methanal", s3, c3, sp)
    # cfg2 sample: 48 synthetic QM9-size molecules, PM3, DIIS, 1e-7
    s4, c4 = synthetic.qm9_like_batch(48, seed=0)
    sp = {"method": "PM3", "scf_eps": 1e-7, "scf_converger": [2], "sp2": [False], "analytical_gradient": [True]}
    save("cfg2_PM3_48", s4, c4, sp, extra={"sha256": synthetic.batch_sha256(s4, c4)})
    sp = {"method": "AM1", "scf_eps": 1e-6, "scf_converger": [1], "sp2": [False], "analytical_gradient": [True]}
    save("cfg2_AM1_48_c1", s4, c4, sp, drop=("dm", "e_mo"))
    # coronene pair (cfg3 geometry): SCF + SP2
    s5, c5 = read_xyz([os.path.join(XYZ, "coronene.xyz")] * 2)
    c5 = c5.copy()
    c5[1] += np.random.default_rng(3).normal(scale=0.02, size=c5[1].shape) * (s5[1] > 0)[:, None]
    sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [2], "sp2": [False], "analytical_gradient": [True]}
    save("cfg3_coronene_AM1", s5, c5, sp)


if __name__ == "__main__":
    main()
