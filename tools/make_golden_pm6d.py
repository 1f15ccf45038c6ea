#!/usr/bin/env python
"""Golden fixtures for PM6 WITH d orbitals (BASELINE configs[4]) from the UNMODIFIED reference.

    python tools/make_golden_pm6d.py          # build container only (needs /root/reference)

Writes tests/golden/pm6d_*.npz: inputs + Etot, Hf, Eelec, Enuc, Eiso, e_mo, e_gap, dm, q, force (autograd: the reference has
no analytic PM6 gradient, anal_grad.py:50-51), notconverged, n_scf_iter, and the operator-level outputs that pin each
kernel: hcore() -> op_M (nmol*molsize^2, 9, 9), op_w (npairs, 45, 45; the reference's [j-pair, i-pair] orientation),
diatom_overlap_matrixD -> op_di, calc_integral -> op_W243, fock() on a random symmetric density -> op_X, op_F.
tests/golden/ref_json/pm6_batch_notebook.json is the reference's own golden (tests/reference/), copied verbatim.
"""
import json
import os
import shutil
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
KEEP = ["Etot", "Hf", "Eelec", "Enuc", "Eiso", "e_mo", "e_gap", "dm", "q", "force", "notconverged", "n_scf_iter"]


def operator_level(species, coords):
    from seqm.Molecule import Molecule
    from seqm.seqm_functions.build_two_elec_one_center_int_D import calc_integral
    from seqm.seqm_functions.constants import Constants
    from seqm.seqm_functions.diat_overlapD import diatom_overlap_matrixD
    from seqm.seqm_functions.fock import fock
    from seqm.seqm_functions.hcore import hcore

    sp = {"method": "PM6", "scf_eps": 1e-7, "scf_converger": [1]}
    mol = Molecule(Constants(), sp, torch.as_tensor(coords), torch.as_tensor(species, dtype=torch.int64))
    M, w, rho0i, rho0j, _, _ = hcore(mol)
    p = mol.parameters
    zeta = torch.stack([p["zeta_s"], p["zeta_p"], p["zeta_d"]], dim=1)
    di = diatom_overlap_matrixD(mol.ni, mol.nj, mol.xij, mol.rij, zeta[mol.idxi], zeta[mol.idxj], mol.const.qn_int,
                                mol.const.qnD_int)  # fmt: skip
    nb = 9 * mol.molsize
    g = torch.Generator().manual_seed(11)
    X = torch.rand(mol.nmol, nb, nb, generator=g) - 0.5
    sh = (((mol.species > 12) & (mol.species < 18)) | ((mol.species > 20) & (mol.species < 30))
          | ((mol.species > 32) & (mol.species < 36)))  # fmt: skip
    real = torch.zeros(mol.nmol, mol.molsize, 9, dtype=torch.bool)
    real[..., :4] |= (mol.species > 1).unsqueeze(-1)
    real[..., 0] |= mol.species == 1
    real[..., 4:] |= sh.unsqueeze(-1)
    real = real.reshape(mol.nmol, nb)
    X = (X + X.transpose(1, 2)) * (real.unsqueeze(1) & real.unsqueeze(2))
    W = calc_integral(p["s_orb_exp_tail"], p["p_orb_exp_tail"], p["d_orb_exp_tail"], mol.Z,
                      mol.nmol * mol.molsize * mol.molsize, mol.maskd, X, p["F0SD"], p["G2SD"])  # fmt: skip
    F = fock(mol.nmol, mol.molsize, X, M, mol.maskd, mol.mask, mol.idxi, mol.idxj, w, W, p["g_ss"], p["g_pp"], p["g_sp"],
             p["g_p2"], p["h_sp"], "PM6", p["zeta_s"], p["zeta_p"], p["zeta_d"], mol.Z, p["F0SD"], p["G2SD"])  # fmt: skip
    t = lambda x: x.detach().numpy()  # noqa: E731
    return dict(op_M=t(M), op_w=t(w), op_di=t(di), op_W243=t(W[mol.maskd]), op_X=t(X), op_F=t(F), op_rho0i=t(rho0i),
                op_rho0j=t(rho0j))  # fmt: skip


def save(name, species, coords, sp, ops=True, drop=()):
    ref = run_reference(species, coords, sp)
    out = {k: ref[k] for k in KEEP if k not in drop}
    out["species"] = np.asarray(species, dtype=np.int64)
    out["coordinates"] = np.asarray(coords, dtype=np.float64)
    out["seqm_parameters"] = json.dumps(sp)
    if ops:
        out.update(operator_level(species, coords))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: nmol={out['species'].shape[0]} iters={ref['n_scf_iter']} notconv={int(ref['notconverged'].sum())} "
          f"Etot[0]={ref['Etot'][0]:.10f}")
    return ref


def rotated(coords, seed):
    q = np.random.default_rng(seed).normal(size=(3, 3))
    Q, _ = np.linalg.qr(q)
    return coords @ Q.T


def main():
    only = os.environ.get("GOLDEN_ONLY")
    os.makedirs(os.path.join(GOLD, "ref_json"), exist_ok=True)
    shutil.copy("/root/reference/tests/reference/pm6_batch_notebook.json", os.path.join(GOLD, "ref_json", "pm6_batch_notebook.json"))
    # 1. the reference's own PM6 test batch (tests/unit/test_pm6_batch.py): S2, Ti2, TiS, BrCl, CrTi along y, their settings
    species = np.array([[16, 16], [22, 22], [22, 16], [35, 17], [24, 22]])
    coords = np.zeros((5, 2, 3))
    coords[:, 1, 1] = 1.2
    sp = {"method": "PM6", "scf_eps": 1.0e-5, "scf_converger": [0, 0.2], "sp2": [False, 1.0e-5], "pair_outer_cutoff": 1.0e10,
          "eig": True, "Hf_flag": True, "scf_backward": 0, "UHF": False}  # fmt: skip
    save("pm6d_notebook_diatomics", species, coords, sp)
    # the same diatomics in general orientation at more relaxed distances, tight SCF
    # (Cr-Ti has several SCF solutions and does not converge reproducibly; Ti-O and H-Br stand in: transition metal with
    # an sp-only partner, 4th-row main group with hydrogen)
    species = np.array([[16, 16], [22, 22], [22, 16], [35, 17], [22, 8], [35, 1]])
    c2 = np.zeros((6, 2, 3))
    c2[:, 1, 1] = [1.9, 2.1, 2.0, 2.14, 1.62, 1.41]
    c2 = np.stack([rotated(c2[i], 40 + i) for i in range(6)])
    sp = {"method": "PM6", "scf_eps": 1.0e-8, "scf_converger": [1], "sp2": [False]}
    save("pm6d_diatomics_rotated", species, c2, sp)
    # 2. small S / P / Cl organics, randomly oriented
    files = ["h2s.xyz", "ch3cl.xyz", "ch3sh.xyz", "ph3.xyz", "pcl3.xyz"]
    s, c = read_xyz([os.path.join(XYZ, f) for f in files])
    c = np.stack([rotated(c[i], 7 + i) * (s[i] > 0)[:, None] for i in range(len(files))])
    for tag, conv, eps in (("c1", [1], 1e-7), ("c2", [2], 1e-7), ("c0", [0, 0.3], 1e-7)):
        sp = {"method": "PM6", "scf_eps": eps, "scf_converger": conv, "sp2": [False]}
        save(f"pm6d_organics_{tag}", s, c, sp, ops=(tag == "c1"))
    # 3. configs[4] sample: synthetic QM9-size organics with P / S / Cl at one or two heavy sites
    s4, c4 = synthetic.qm9_like_batch(16, seed=0, hetero=(15, 16, 17))
    sp = {"method": "PM6", "scf_eps": 1e-7, "scf_converger": [1], "sp2": [False]}
    save("pm6d_cfg5_16", s4, c4, sp, ops=False, drop=("e_mo",))


if __name__ == "__main__":
    main()
