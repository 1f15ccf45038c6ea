"""CPU-only check of the kernel LOGIC: the kernel sources compiled with -DSEQM_HOSTEMU (every CTA run
sequentially with one thread) driven through the same C ABI and Python host code as the CUDA build, against
the reference-generated fixtures.  This is test infrastructure, not a product path."""
import numpy as np
import pytest
import torch

from helpers import (check_golden_case, check_mo_match, check_mo_match_vs_oracle, check_operator_level, check_pm6_sp_elements,
                     check_two_forwards_match_orbitals, hostemu_lib, run_molecule)

CPU = torch.device("cpu")


@pytest.fixture(scope="module")
def lib():
    return hostemu_lib()


@pytest.mark.parametrize("method", ["AM1", "PM3", "MNDO", "PM6_SP"])
def test_operator_level(lib, method):
    check_operator_level(lib, CPU, method)


@pytest.mark.parametrize("method", ["AM1", "PM6_SP"])
def test_reference_operator_signatures(lib, method):
    from helpers import check_level_b_signatures

    check_level_b_signatures(lib, CPU, method)


@pytest.mark.parametrize(
    "name",
    ["cfg1_AM1_c2", "cfg1_AM1_c1", "cfg1_AM1_c0", "cfg1_PM3_c2", "cfg1_PM3_c1", "cfg1_MNDO_c2", "cfg1_MNDO_c0",
     "ref_batch_single_point_am1", "ref_ground_force_methanal", "cfg2_PM3_48", "cfg1_PM6_SP_c2", "cfg1_PM6_SP_c1",
     "cfg2_PM6_SP_24", "opt_charged_AM1", "opt_learned_PM3", "opt_flags_MNDO", "opt_cutoff_AM1", "thirdrow_PM3_c2", "thirdrow_AM1_c2", "thirdrow_MNDO_c2", "thirdrow_PM6_SP_c2", "cfg3_coronene_AM1"],
)  # fmt: skip
def test_single_point_golden(lib, name):
    check_golden_case(lib, CPU, name)


@pytest.mark.parametrize("name", ["op_momatch_mixed", "op_momatch_uniform"])
def test_mo_crossing_matcher(lib, name):
    check_mo_match(lib, CPU, name)


def test_mo_crossing_matcher_against_oracle(lib):
    check_mo_match_vs_oracle(lib, CPU)


def test_second_forward_continues_the_orbitals(lib):
    check_two_forwards_match_orbitals(lib, CPU)


def test_pm6_on_elements_without_d_shell(lib):
    check_pm6_sp_elements(lib, CPU)


@pytest.mark.parametrize("name", ["pm6d_organics_c1", "pm6d_diatomics_rotated", "pm6d_notebook_diatomics"])
def test_pm6_d_orbital_operators(lib, name):
    from helpers import check_pm6d_operators

    check_pm6d_operators(lib, CPU, name)


def test_pm6_d_orbital_single_points(lib):
    from helpers import PM6D_CASES, check_pm6d_case

    for name in PM6D_CASES:
        check_pm6d_case(lib, CPU, name)


def test_sp2_route(lib):
    # mixed batch: the reference zero-pads packed matrices inside SP2 (pack.py:76-77), we purify at native
    # size, so agreement is at the SP2 tolerance; the unpadded (largest) molecule agrees tightly.
    mol = check_golden_case(lib, CPU, "cfg1_AM1_sp2", sp2_tolerant=True)
    from conftest import load_golden

    g = load_golden("cfg1_AM1_sp2")
    assert abs(float(mol.Etot[2]) - g["Etot"][2]) < 1e-6


def test_user_P0_is_updated_in_place(lib):
    from conftest import load_golden

    g = load_golden("cfg1_AM1_c2")
    P0 = torch.as_tensor(g["dm"]).clone()
    mol, es = run_molecule(lib, CPU, g["species"], g["coordinates"], g["seqm_parameters"], P0=P0)
    assert mol.dm.data_ptr() == P0.data_ptr()
    assert mol.n_scf_iter <= 3  # restart from the converged density
    assert np.abs(mol.Etot.numpy() - g["Etot"]).max() < 1e-6


def test_lazy_orbital_charge_table(lib):
    """esdriver.charge (scf_loop.py:2346-2387) is built lazily from the eigenvectors; every row of the
    orthogonal eigenvector matrix carries unit weight in total."""
    from conftest import load_golden

    g = load_golden("cfg1_AM1_c2")
    mol, es = run_molecule(lib, CPU, g["species"], g["coordinates"], g["seqm_parameters"])
    c = es.charge
    assert tuple(c.shape) == (3, 4 * mol.molsize, mol.molsize)
    norb = (4 * mol.nHeavy + mol.nHydro).tolist()
    for m in range(3):
        assert np.abs(c[m, : norb[m]].sum(dim=1).numpy() - 1.0).max() < 1e-12
        assert float(c[m, norb[m] :].abs().max()) == 0.0


@pytest.mark.parametrize("nmol", [1, 2, 3])
def test_forced_half_batches_on_tiny_batches(lib, nmol, monkeypatch):
    """seqm_scf_opts_t.pipeline = 2 on 1-3 molecules (empty / unequal halves) gives the single-stream result."""
    from conftest import load_golden

    g = load_golden("cfg1_AM1_c2")
    sp_, xyz = g["species"][:nmol], g["coordinates"][:nmol]
    res = {}
    for mode in ("1", "2"):
        monkeypatch.setenv("SEQM_B200_PIPELINE", mode)
        mol, _ = run_molecule(lib, CPU, sp_, xyz, g["seqm_parameters"])
        res[mode] = (mol.n_scf_iter, mol.Etot.numpy().copy(), mol.dm.numpy().copy())
    assert res["1"][0] == res["2"][0]
    assert np.array_equal(res["1"][1], res["2"][1]) and np.array_equal(res["1"][2], res["2"][2])
    if nmol == 3:
        assert res["1"][0] == g["n_scf_iter"]


def test_device_batch_plan_against_numpy(lib):
    from helpers import check_device_batch_plan

    check_device_batch_plan(lib, CPU)


def test_largest_shared_memory_size_class(lib):
    from helpers import check_largest_in_sm_class

    check_largest_in_sm_class(lib, CPU)


@pytest.mark.parametrize("pipe", ["1", "2"])
def test_scf_iteration_cap_reports_not_converged(lib, pipe, monkeypatch):
    """seqm_scf with max_iter = 3 (the reference's MAX_ITER is 1000, scf_loop.py:29): the loop runs iterations
    0..max_iter, leaves the not-converged flags set and reports max_iter + 1, on both host-loop variants."""
    from conftest import load_golden
    from pyseqm_b200 import engine

    monkeypatch.setenv("SEQM_B200_PIPELINE", pipe)
    g = load_golden("cfg1_AM1_c2")
    plan = engine.BatchPlan(lib, torch.as_tensor(g["species"]), "AM1")
    xyz = plan.real_xyz(torch.as_tensor(g["coordinates"]))
    w, hab = engine.op_pair_integrals(plan, xyz)
    H = engine.op_hcore(plan, w, hab)
    P = engine.op_initial_density(plan)
    F, E, nc, n_iter = engine.op_scf(plan, H, w, P, 1e-7, [2], max_iter=3)
    assert n_iter == 4 and bool(nc.all()) and torch.isfinite(E).all()
    P2 = engine.op_initial_density(plan)
    F2, E2, nc2, n_iter2 = engine.op_scf(plan, H, w, P2, 1e-7, [2])
    assert n_iter2 == g["n_scf_iter"] and not bool(nc2.any())


def test_isolated_atoms(lib):
    from helpers import check_isolated_atoms

    check_isolated_atoms(lib, CPU)
