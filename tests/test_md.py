"""XL-BOMD / BOMD: oracle against reference-generated trajectories (CPU), the kernels through host emulation
(CPU) and the CUDA path (GPU).  Fixtures: tools/make_golden_md.py drove the unmodified reference's XL_BOMD
(eigensolver branch, xlbomd.py:361) and Molecular_Dynamics_Basic step by step."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

CASES = ["md_xl_bomd_methane_k6", "md_xl_bomd_mixed_k4", "md_basic_methanal", "md_xl_bomd_coronene_k6"]


def load_md(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    g["seqm_parameters"] = json.loads(str(g["seqm_parameters"]))
    g["k"] = int(g["k"])
    return g


@pytest.mark.parametrize("name", CASES)
def test_oracle_md_matches_reference(name):
    import seqm_oracle as so

    g = load_md(name)
    out = so.run_md(g["species"], g["coordinates0"], g["velocities0"], g["seqm_parameters"], float(g["timestep"]),
                    int(g["steps"]), None if g["k"] < 0 else g["k"])  # fmt: skip
    assert np.abs(out["Etot"] - g["Etot"]).max() < 1e-6
    assert np.abs(out["Ek"] - g["Ek"]).max() < 1e-6
    assert np.abs(out["coordinates"] - g["coordinates"]).max() < 1e-7
    assert np.abs(out["force"] - g["force"]).max() < 1e-5
    assert np.abs(out["dm"] - g["dm"]).max() < 1e-8


def run_product(lib, device, g):
    import pyseqm_b200 as seqm

    torch.set_default_dtype(torch.float64)
    sp = dict(g["seqm_parameters"])
    mol = seqm.Molecule(seqm.Constants().to(device), sp, torch.as_tensor(g["coordinates0"], device=device).clone(),
                        torch.as_tensor(g["species"], device=device), _lib=lib)  # fmt: skip
    mol.velocities = torch.as_tensor(g["velocities0"], device=device).clone()
    if g["k"] > 0:
        md = seqm.XL_BOMD(xl_bomd_params={"k": g["k"]}, seqm_parameters=sp, timestep=float(g["timestep"]), Temp=float(g["temp"]))
    else:
        md = seqm.Molecular_Dynamics_Basic(seqm_parameters=sp, timestep=float(g["timestep"]), Temp=float(g["temp"]))
    md.run(mol, int(g["steps"]))
    Etot = torch.stack(md.history["Etot"]).cpu().numpy()
    Ek = torch.stack(md.history["Ek"]).cpu().numpy()
    assert np.abs(Etot - g["Etot"]).max() < 1e-6
    assert np.abs(Ek - g["Ek"]).max() < 1e-6
    assert np.abs(mol.coordinates.detach().cpu().numpy() - g["coordinates"]).max() < 1e-8
    assert np.abs(mol.force.cpu().numpy() - g["force"]).max() < 1e-5
    assert np.abs(mol.dm.cpu().numpy() - g["dm"]).max() < 1e-8
    return mol, md


@pytest.mark.parametrize("name", CASES)
def test_hostemu_md(name):
    from helpers import hostemu_lib

    run_product(hostemu_lib(), torch.device("cpu"), load_md(name))


def test_hostemu_xl_forward_public_api():
    """Electronic_Structure.forward(dm_prop='XL-BOMD', P0=P) with the dense padded field, as the reference's
    XL_BOMD.one_step calls it (MolecularDynamics.py:1487-1496), against the oracle."""
    import seqm_oracle as so
    from helpers import hostemu_lib, run_molecule

    g = load_md("md_xl_bomd_mixed_k4")
    sp = g["seqm_parameters"]
    ref0 = so.single_point(g["species"], g["coordinates0"], sp)
    x1 = g["coordinates0"] + 0.01 * np.random.default_rng(0).normal(size=g["coordinates0"].shape) * (g["species"] > 0)[:, :, None]
    ref = so.xl_forward(g["species"], x1, sp, ref0["dm"])
    import pyseqm_b200 as seqm

    lib = hostemu_lib()
    mol = seqm.Molecule(seqm.Constants(), dict(sp), torch.as_tensor(x1), torch.as_tensor(g["species"]), _lib=lib)
    es = seqm.Electronic_Structure(dict(sp))
    es(mol, P0=torch.as_tensor(ref0["dm"]).clone(), dm_prop="XL-BOMD", xl_bomd_params={"k": 4})
    assert np.abs(mol.Etot.numpy() - ref["Etot"]).max() < 1e-6
    assert np.abs(mol.dm.numpy() - ref["dm"]).max() < 1e-8
    assert np.abs(mol.force.numpy() - ref["force"]).max() < 1e-5
    assert np.abs(mol.e_gap.numpy() - ref["e_gap"]).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_md(name):
    from helpers import cuda_lib

    run_product(cuda_lib(), torch.device("cuda:0"), load_md(name))


@pytest.mark.gpu
def test_gpu_xl_bomd_energy_conservation_replicas():
    """64 coronene replicas, 40 XL-BOMD steps: every replica conserves E(total) within the reference's drift tolerance and the
    replicas stay independent (identical replicas give identical trajectories)."""
    import pyseqm_b200 as seqm
    from helpers import cuda_lib

    torch.set_default_dtype(torch.float64)
    dev = torch.device("cuda:0")
    s, c = seqm.read_xyz([os.path.join(GOLDEN, "xyz", "coronene.xyz")] * 64)
    sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [2]}
    mol = seqm.Molecule(seqm.Constants().to(dev), sp, torch.as_tensor(c, device=dev), torch.as_tensor(s, device=dev), _lib=cuda_lib())
    torch.manual_seed(0)
    md = seqm.XL_BOMD(xl_bomd_params={"k": 6}, seqm_parameters=sp, timestep=0.4, Temp=300.0)
    md.set_dof(mol)
    md.initialize_velocity(mol)
    mol.velocities[1] = mol.velocities[0]  # two identical replicas
    md.run(mol, 40)
    E = (torch.stack(md.history["Etot"]) + torch.stack(md.history["Ek"])).cpu().numpy()
    assert np.abs(E - E[0]).max() < 5e-2  # the reference's own drift tolerance (tests/unit/test_md_suite.py:232)
    assert np.abs(E[:, 0] - E[:, 1]).max() < 1e-9


@pytest.mark.gpu
def test_gpu_xl_bomd_sp2_route_tracks_the_eigensolver_route():
    """BASELINE configs[2] uses the SP2 density (eps 1e-5) in XL-BOMD; the reference cannot run that combination
    (xlbomd.py:359), so the route is pinned against this package's eigensolver route, which is itself pinned against
    the reference trajectories: same start, 30 steps, 16 coronene replicas."""
    import pyseqm_b200 as seqm
    from helpers import cuda_lib

    torch.set_default_dtype(torch.float64)
    dev = torch.device("cuda:0")
    s, c = seqm.read_xyz([os.path.join(GOLDEN, "xyz", "coronene.xyz")] * 16)
    out = {}
    for tag, sp2 in (("eig", [False]), ("sp2", [True, 1.0e-5])):
        sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [2], "sp2": sp2}
        mol = seqm.Molecule(seqm.Constants().to(dev), sp, torch.as_tensor(c, device=dev), torch.as_tensor(s, device=dev), _lib=cuda_lib())
        torch.manual_seed(0)
        md = seqm.XL_BOMD(xl_bomd_params={"k": 6}, seqm_parameters=sp, timestep=0.4, Temp=300.0)
        md.run(mol, 30)
        E = (torch.stack(md.history["Etot"]) + torch.stack(md.history["Ek"])).cpu().numpy()
        out[tag] = (E, mol.coordinates.detach().cpu().numpy().copy())
        assert np.abs(E - E[0]).max() < 5e-2
    assert np.abs(out["eig"][0] - out["sp2"][0]).max() < 2e-3   # total energy along the trajectory, eV
    assert np.abs(out["eig"][1] - out["sp2"][1]).max() < 1e-4   # final coordinates, Angstrom


# ---- KSA-XL-BOMD: fixtures from tools/make_golden_ksa.py (unmodified reference) ---------------------
KSA_CASES = ["md_ksa_methane_k6", "md_ksa_mixed_k4", "md_ksa_benzene_thr"]


def load_ksa(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    g["seqm_parameters"] = json.loads(str(g["seqm_parameters"]))
    g["xl_bomd_params"] = json.loads(str(g["xl_bomd_params"]))
    return g


def test_oracle_ksa_operators_match_reference():
    """Fermi_Q (fermi_q.py:8-72) and Canon_DM_PRT (canon_dm_prt.py:6-39) of the oracle against the reference's outputs on
    {methane, benzene, toluene} at 1500 K (integer occupations) and 20000 K (fractional occupations, non-zero entropy)."""
    from seqm_oracle import ksa
    from seqm_oracle.parser import parse

    g = dict(np.load(os.path.join(GOLDEN, "ksa_operators.npz")))
    P = parse(g["species"], g["coordinates"])
    for T, suf in ((float(g["T_el"]), ""), (float(g["T_hot"]), "_hot")):
        D0, S, eig, f, mu = ksa.fermi_q(g["F"], T, P.nocc, P.nHeavy, P.nHydro)
        assert np.abs(D0 - g["D0" + suf]).max() < 1e-12
        assert np.abs(S - g["S" + suf]).max() < 1e-15
        assert np.abs(mu - g["mu" + suf].ravel()).max() < 1e-11
        assert np.abs(f - g["Fe" + suf][:, : f.shape[1]]).max() < 1e-12
        assert np.abs(ksa.canon_dm_prt(g["FO1"], T, eig, mu) - g["PO1" + suf]).max() < 1e-13
    assert g["S_hot"].min() > 1e-6  # the hot fixture does exercise the entropy


@pytest.mark.parametrize("name", KSA_CASES)
def test_oracle_ksa_md_matches_reference(name):
    from seqm_oracle import ksa

    g = load_ksa(name)
    out = ksa.run_ksa_md(g["species"], g["coordinates0"], g["velocities0"], g["seqm_parameters"], float(g["timestep"]),
                         int(g["steps"]), g["xl_bomd_params"])  # fmt: skip
    for k, tol in (("Etot", 1e-6), ("Ek", 1e-6), ("Electronic_entropy", 1e-9), ("Krylov_Error", 1e-8), ("coordinates", 1e-7),
                   ("force", 1e-5), ("dm", 1e-8), ("dP2dt2", 1e-8)):  # fmt: skip
        assert np.abs(out[k] - g[k]).max() < tol, k


def check_ksa_operators(lib, device):
    """The product's Fermi_Q / Canon_DM_PRT (ForceXL.fermi_density / density_response: eigensolver + seqm_ksa.cu kernels on the
    packed layout) against the reference's outputs."""
    import pyseqm_b200 as seqm
    from pyseqm_b200 import engine
    from pyseqm_b200.basics import ForceXL

    torch.set_default_dtype(torch.float64)
    g = dict(np.load(os.path.join(GOLDEN, "ksa_operators.npz")))
    sp = json.loads(str(g["seqm_parameters"]))
    mol = seqm.Molecule(seqm.Constants().to(device), dict(sp), torch.as_tensor(g["coordinates"], device=device),
                        torch.as_tensor(g["species"], device=device), _lib=lib)  # fmt: skip
    plan = mol._plan
    fx = ForceXL(dict(sp))
    F = engine.op_pack(plan, torch.as_tensor(g["F"], device=device))
    FO1 = engine.op_pack(plan, torch.as_tensor(g["FO1"], device=device))
    for T, suf in ((float(g["T_el"]), ""), (float(g["T_hot"]), "_hot")):
        e, Q, f, mu, D, S = fx.fermi_density(plan, F, T)
        assert np.abs(engine.op_unpack(plan, D).cpu().numpy() - g["D0" + suf]).max() < 1e-10
        assert np.abs(S.cpu().numpy() - g["S" + suf]).max() < 1e-12
        assert np.abs(mu.cpu().numpy() - g["mu" + suf].ravel()).max() < 1e-9
        assert np.abs(f.cpu().numpy() - g["Fe" + suf][:, : plan.nmax]).max() < 1e-10
        P1 = fx.density_response(plan, FO1, Q, e, mu, 1.0 / (fx.KB * T))
        assert np.abs(engine.op_unpack(plan, P1).cpu().numpy() - g["PO1" + suf]).max() < 1e-10
    # the generic per-molecule product against numpy, all four transpose combinations
    A = engine.op_pack(plan, torch.as_tensor(g["FO1"], device=device))
    B = engine.op_pack(plan, torch.as_tensor(g["D0"], device=device) + 0.3 * torch.as_tensor(g["X"], device=device).triu())
    Ad, Bd = engine.op_unpack(plan, A).cpu().numpy(), engine.op_unpack(plan, B).cpu().numpy()
    for ta in (False, True):
        for tb in (False, True):
            Cd = engine.op_unpack(plan, engine.op_packed_gemm(plan, A, B, ta, tb)).cpu().numpy()
            ref = np.matmul(Ad.transpose(0, 2, 1) if ta else Ad, Bd.transpose(0, 2, 1) if tb else Bd)
            assert np.abs(Cd - ref).max() < 1e-12, (ta, tb)


def run_ksa_product(lib, device, g):
    import pyseqm_b200 as seqm

    torch.set_default_dtype(torch.float64)
    sp = dict(g["seqm_parameters"])
    mol = seqm.Molecule(seqm.Constants().to(device), sp, torch.as_tensor(g["coordinates0"], device=device).clone(),
                        torch.as_tensor(g["species"], device=device), _lib=lib)  # fmt: skip
    mol.velocities = torch.as_tensor(g["velocities0"], device=device).clone()
    md = seqm.KSA_XL_BOMD(xl_bomd_params=dict(g["xl_bomd_params"]), seqm_parameters=sp, timestep=float(g["timestep"]),
                          Temp=float(g["temp"]))  # fmt: skip
    md.run(mol, int(g["steps"]))
    Epot = torch.stack(md.history["Etot"]).cpu().numpy()  # XL-BOMD's thermodynamic potential: Etot + entropy term
    Ek = torch.stack(md.history["Ek"]).cpu().numpy()
    assert np.abs(Epot - (g["Etot"] + g["Electronic_entropy"])).max() < 1e-6
    assert np.abs(Ek - g["Ek"]).max() < 1e-6
    assert np.abs(mol.Electronic_entropy.cpu().numpy() - g["Electronic_entropy"][-1]).max() < 1e-9
    assert np.abs(mol.Krylov_Error.cpu().numpy() - g["Krylov_Error"][-1]).max() < 1e-7
    assert np.abs(mol.coordinates.detach().cpu().numpy() - g["coordinates"]).max() < 1e-8
    assert np.abs(mol.force.cpu().numpy() - g["force"]).max() < 1e-5
    assert np.abs(mol.dm.cpu().numpy() - g["dm"]).max() < 1e-8
    assert np.abs(mol.dP2dt2.cpu().numpy() - g["dP2dt2"]).max() < 1e-8
    assert np.abs(mol.Fermi_occ.cpu().numpy() - g["Fermi_occ"][:, : mol.Fermi_occ.shape[1]]).max() < 1e-9


def test_hostemu_ksa_operators():
    from helpers import hostemu_lib

    check_ksa_operators(hostemu_lib(), torch.device("cpu"))


@pytest.mark.parametrize("name", KSA_CASES)
def test_hostemu_ksa_md(name):
    from helpers import hostemu_lib

    run_ksa_product(hostemu_lib(), torch.device("cpu"), load_ksa(name))


def test_hostemu_ksa_public_api():
    """Electronic_Structure.forward(dm_prop='XL-BOMD', xl_bomd_params with max_rank) as KSA_XL_BOMD.one_step of the reference
    calls it (dense padded field in, dense dm / dP2dt2 out) against the oracle."""
    import pyseqm_b200 as seqm
    import seqm_oracle as so
    from helpers import hostemu_lib
    from seqm_oracle import ksa

    g = load_ksa("md_ksa_mixed_k4")
    sp, xl = g["seqm_parameters"], g["xl_bomd_params"]
    ref0 = so.single_point(g["species"], g["coordinates0"], sp)
    x1 = g["coordinates0"] + 0.01 * np.random.default_rng(1).normal(size=g["coordinates0"].shape) * (g["species"] > 0)[:, :, None]
    ref = ksa.ksa_forward(g["species"], x1, sp, ref0["dm"], xl)
    mol = seqm.Molecule(seqm.Constants(), dict(sp), torch.as_tensor(x1), torch.as_tensor(g["species"]), _lib=hostemu_lib())
    es = seqm.Electronic_Structure(dict(sp))
    es(mol, P0=torch.as_tensor(ref0["dm"]).clone(), dm_prop="XL-BOMD", xl_bomd_params=dict(xl))
    assert np.abs(mol.Etot.numpy() - ref["Etot"]).max() < 1e-6
    assert np.abs(mol.dm.numpy() - ref["dm"]).max() < 1e-8
    assert np.abs(mol.dP2dt2.numpy() - ref["dP2dt2"]).max() < 1e-8
    assert np.abs(mol.force.numpy() - ref["force"]).max() < 1e-5
    assert np.abs(mol.Krylov_Error.numpy() - ref["Krylov_Error"]).max() < 1e-8


@pytest.mark.gpu
def test_gpu_ksa_operators():
    from helpers import cuda_lib

    check_ksa_operators(cuda_lib(), torch.device("cuda:0"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", KSA_CASES)
def test_gpu_ksa_md(name):
    from helpers import cuda_lib

    run_ksa_product(cuda_lib(), torch.device("cuda:0"), load_ksa(name))


KSA_SCF_CASES = ["ksa_scf_mixed", "ksa_scf_methanal"]


def load_ksa_scf(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    g["seqm_parameters"] = json.loads(str(g["seqm_parameters"]))
    g["n_scf_iter"] = int(g["n_scf_iter"])
    return g


@pytest.mark.parametrize("name", KSA_SCF_CASES)
def test_oracle_ksa_scf_matches_reference(name):
    """scf_converger = [3, {...}] (scf_forward3, scf_loop.py:1135-1381): oracle against the reference single points."""
    import seqm_oracle as so

    g = load_ksa_scf(name)
    r = so.single_point(g["species"], g["coordinates"], g["seqm_parameters"])
    assert r["n_scf_iter"] == g["n_scf_iter"]
    for k, tol in (("Etot", 1e-6), ("Eelec", 1e-6), ("Hf", 1e-6), ("dm", 1e-8), ("force", 1e-5), ("e_gap", 1e-6), ("q", 1e-8)):
        assert np.abs(r[k] - g[k]).max() < tol, k


def check_ksa_scf(lib, device, name):
    from helpers import run_molecule

    g = load_ksa_scf(name)
    mol, es = run_molecule(lib, device, g["species"], g["coordinates"], g["seqm_parameters"])
    assert mol.n_scf_iter == g["n_scf_iter"]
    assert not bool(es.notconverged.any())
    for k, tol in (("Etot", 1e-6), ("Eelec", 1e-6), ("Hf", 1e-6), ("dm", 1e-8), ("force", 1e-5), ("e_gap", 1e-6), ("q", 1e-8)):
        assert np.abs(getattr(mol, k).cpu().numpy() - g[k]).max() < tol, k


@pytest.mark.parametrize("name", KSA_SCF_CASES)
def test_hostemu_ksa_scf(name):
    from helpers import hostemu_lib

    check_ksa_scf(hostemu_lib(), torch.device("cpu"), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", KSA_SCF_CASES)
def test_gpu_ksa_scf(name):
    from helpers import cuda_lib

    check_ksa_scf(cuda_lib(), torch.device("cuda:0"), name)


# ---- CIS sigma-vector building block (makeA_pi_batched, rcis_batch.py:296-403) ------------------------
def _dense_from_packed(Xp, nheavy, nhyd, molsize):
    """(nmol, nroots, norb, norb) in packed orbital order -> dense padded (nmol * nroots, 4 molsize, 4 molsize)"""
    idx = np.concatenate([np.arange(4 * nheavy), 4 * nheavy + 4 * np.arange(nhyd)])
    nmol, nr = Xp.shape[:2]
    D = np.zeros((nmol, nr, 4 * molsize, 4 * molsize))
    D[:, :, idx[:, None], idx[None, :]] = Xp
    return D, idx


def test_oracle_cis_sigma_matches_reference():
    from seqm_oracle import cis
    from seqm_oracle.hamiltonian import build_hcore
    from seqm_oracle.integrals import atom_multipoles
    from seqm_oracle.parser import parse
    from seqm_oracle.tables import method_parameters

    g = dict(np.load(os.path.join(GOLDEN, "cis_sigma_methanal.npz")))
    P = parse(g["species"], g["coordinates"])
    par = method_parameters("AM1", P.Z)
    w = build_hcore(P, par, atom_multipoles(P.Z, par))["w"]
    nh, ny = int(P.nHeavy[0]), int(P.nHydro[0])
    for key_x, key_f, sym in (("X", "F", False), ("Xs", "Fs", True)):
        Xd, idx = _dense_from_packed(g[key_x], nh, ny, P.molsize)
        for r in range(Xd.shape[1]):
            F = cis.sigma_ao(P, par, w, Xd[:, r], all_symmetric=sym)
            assert np.abs(F[:, idx[:, None], idx[None, :]] - g[key_f][:, r]).max() < 1e-12


def check_cis_sigma(lib, device):
    import pyseqm_b200 as seqm
    from pyseqm_b200.seqm_functions.hcore import hcore
    from pyseqm_b200.seqm_functions.rcis_batch import makeA_pi_batched

    torch.set_default_dtype(torch.float64)
    g = dict(np.load(os.path.join(GOLDEN, "cis_sigma_methanal.npz")))
    sp = json.loads(str(g["seqm_parameters"]))
    mol = seqm.Molecule(seqm.Constants().to(device), dict(sp), torch.as_tensor(g["coordinates"], device=device),
                        torch.as_tensor(g["species"], device=device), _lib=lib)  # fmt: skip
    w = hcore(mol)[1]
    F = makeA_pi_batched(mol, torch.as_tensor(g["X"], device=device), w, allSymmetric=False)
    assert np.abs(F.cpu().numpy() - g["F"]).max() < 1e-11
    Fs = makeA_pi_batched(mol, torch.as_tensor(g["Xs"], device=device), w, allSymmetric=True)
    assert np.abs(Fs.cpu().numpy() - g["Fs"]).max() < 1e-11
    with pytest.raises(ValueError):
        makeA_pi_batched(mol, torch.zeros((2, 1, 7, 7), dtype=torch.float64, device=device), w)


def test_hostemu_cis_sigma():
    from helpers import hostemu_lib

    check_cis_sigma(hostemu_lib(), torch.device("cpu"))


@pytest.mark.gpu
def test_gpu_cis_sigma():
    from helpers import cuda_lib

    check_cis_sigma(cuda_lib(), torch.device("cuda:0"))
