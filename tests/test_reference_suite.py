"""The reference's own unit tests for the in-scope path, run against pyseqm_b200 with `import pyseqm_b200 as seqm`
in place of `import seqm`: tests/unit/test_smoke_single_point.py, test_batch_single_point.py, test_force_methods.py
(ground state, all three force modes), test_invariants.py (ground state), test_uhf.py (the RHF error), and the
input validation of Molecule.py:188-206 -- against the reference's own JSON goldens (tests/golden/ref_json) at the
reference's tolerances.  Each test runs twice: on the CUDA library (-m gpu) and, as kernel-logic coverage for GPU-less
CI, on the host-emulation build of the same kernel sources (test infrastructure, never loaded by the package)."""
import json
import os

import numpy as np
import pytest
import torch

import pyseqm_b200 as seqm
import pyseqm_b200._lib as seqm_lib
from pyseqm_b200.ElectronicStructure import Electronic_Structure
from pyseqm_b200.Molecule import Molecule
from pyseqm_b200.seqm_functions.constants import Constants
from pyseqm_b200.seqm_functions.read_xyz import read_xyz

from conftest import GOLDEN
from helpers import hostemu_lib

XYZ = os.path.join(GOLDEN, "xyz")


@pytest.fixture(params=["hostemu", pytest.param("cuda", marks=pytest.mark.gpu)])
def device(request, monkeypatch):
    torch.set_default_dtype(torch.float64)
    if request.param == "hostemu":
        monkeypatch.setattr(seqm_lib, "_LIB", hostemu_lib())  # same kernel sources, run sequentially on the host
        return torch.device("cpu")
    monkeypatch.setattr(seqm_lib, "_LIB", None)
    return torch.device("cuda")


def _load(files, device):
    torch.manual_seed(0)
    species, coordinates = read_xyz([os.path.join(XYZ, f) for f in files])
    return (torch.as_tensor(species, dtype=torch.int64, device=device),
            torch.as_tensor(coordinates, dtype=torch.float64, device=device))  # fmt: skip


@pytest.fixture
def methane_molecule_data(device):
    return _load(["methane.xyz"], device)


@pytest.fixture
def batch_molecule_data(device):
    return _load(["methane.xyz", "benzene.xyz"], device)


@pytest.fixture
def methanal_batch_data(device):
    return _load(["methanal.1.xyz", "methanal.2.xyz", "methanal.3.xyz"], device)


def reference(name):
    with open(os.path.join(GOLDEN, "ref_json", name + ".json")) as f:
        return json.load(f)


def assert_allclose(actual, expected, rtol=1e-6, atol=1e-6):
    np.testing.assert_allclose(np.asarray(actual, dtype=float), np.asarray(expected, dtype=float), rtol=rtol, atol=atol)


@pytest.mark.parametrize("method", ["MNDO", "AM1", "PM3", "PM6", "PM6_SP"])
def test_single_point_runs_for_all_methods(method, device, methane_molecule_data):
    species, coordinates = methane_molecule_data
    const = Constants().to(device)
    seqm_parameters = {"method": method, "scf_eps": 1.0e-6, "scf_converger": [1]}
    molecule = Molecule(const, seqm_parameters, coordinates, species).to(device)
    esdriver = Electronic_Structure(seqm_parameters).to(device)
    esdriver(molecule)
    assert torch.isfinite(molecule.Etot).all()
    assert torch.isfinite(molecule.Eelec).all()
    assert torch.isfinite(molecule.Enuc).all()
    assert molecule.force is not None
    assert molecule.force.shape == coordinates.shape
    ref = reference(f"smoke_single_point_{method}")
    assert_allclose(float(molecule.Etot.item()), ref["Etot"], rtol=1e-5, atol=1e-5)
    assert_allclose(float(molecule.Eelec.item()), ref["Eelec"], rtol=1e-5, atol=1e-5)
    assert_allclose(float(molecule.Enuc.item()), ref["Enuc"], rtol=1e-5, atol=1e-5)
    assert_allclose(molecule.force.detach().cpu().tolist(), ref["force"], rtol=1e-5, atol=1e-5)


def test_batch_single_point_am1(device, batch_molecule_data):
    species, coordinates = batch_molecule_data
    const = Constants().to(device)
    seqm_parameters = {"method": "AM1", "scf_eps": 1.0e-6, "scf_converger": [1]}
    molecule = Molecule(const, seqm_parameters, coordinates, species).to(device)
    esdriver = Electronic_Structure(seqm_parameters).to(device)
    esdriver(molecule)
    assert torch.isfinite(molecule.Etot).all()
    assert molecule.force.shape == coordinates.shape
    ref = reference("batch_single_point_am1")
    assert_allclose(molecule.Etot.detach().cpu().tolist(), ref["Etot"], rtol=1e-5, atol=1e-5)
    assert_allclose(molecule.force.detach().cpu().tolist(), ref["force"], rtol=1e-4, atol=1e-4)


_FORCE_MODES = [
    ("autodiff", {}),
    ("analytical", {"analytical_gradient": [True]}),
    ("semi_numerical", {"analytical_gradient": [True, "numerical"]}),
]


def _run_ground_force(device, species, coordinates, mode_overrides):
    const = Constants().to(device)
    seqm_parameters = {"method": "AM1", "scf_eps": 1.0e-7, "scf_converger": [1]}
    seqm_parameters.update(mode_overrides)
    molecule = Molecule(const, seqm_parameters, coordinates, species).to(device)
    esdriver = Electronic_Structure(seqm_parameters).to(device)
    esdriver(molecule)
    return molecule.force.detach().cpu().tolist()


@pytest.mark.parametrize("mode_name, mode_overrides", _FORCE_MODES)
def test_ground_force_methods_single_molecule(device, methane_molecule_data, mode_name, mode_overrides):
    species, coordinates = methane_molecule_data
    force = _run_ground_force(device, species, coordinates, mode_overrides)
    assert_allclose(force, reference("ground_force_methane")["force"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("mode_name, mode_overrides", _FORCE_MODES)
def test_ground_force_methods_batch_same_species(device, methanal_batch_data, mode_name, mode_overrides):
    species, coordinates = methanal_batch_data
    force = _run_ground_force(device, species, coordinates, mode_overrides)
    assert_allclose(force, reference("ground_force_batch_methanal")["force"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("mode_name, mode_overrides", _FORCE_MODES)
def test_ground_force_methods_batch_mixed(device, batch_molecule_data, mode_name, mode_overrides):
    species, coordinates = batch_molecule_data
    force = _run_ground_force(device, species, coordinates, mode_overrides)
    assert_allclose(force, reference("ground_force_batch_mixed")["force"], rtol=1e-5, atol=1e-5)


def _rotation_matrix_z(theta):
    c, s = torch.cos(theta), torch.sin(theta)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([torch.stack([c, -s, z]), torch.stack([s, c, z]), torch.stack([z, z, o])])


def _rotate(coords, R):
    return torch.einsum("...i,ij->...j", coords, R)


def test_rotation_invariance_ground_state(device, methane_molecule_data):
    species, coordinates = methane_molecule_data
    const = Constants().to(device)
    seqm_parameters = {"method": "AM1", "scf_eps": 1.0e-7, "scf_converger": [1]}
    molecule = Molecule(const, seqm_parameters, coordinates.clone(), species).to(device)
    esdriver = Electronic_Structure(seqm_parameters).to(device)
    esdriver(molecule)
    E0 = molecule.Etot.detach().cpu()
    F0 = molecule.force.detach().cpu()
    theta = torch.tensor(0.7, dtype=coordinates.dtype)
    R = _rotation_matrix_z(theta).to(coordinates.device)
    coords_rot = _rotate(coordinates, R)
    molecule_rot = Molecule(const, seqm_parameters, coords_rot, species).to(device)
    esdriver(molecule_rot)
    assert_allclose(E0, molecule_rot.Etot.detach().cpu(), rtol=1e-5, atol=1e-5)
    assert_allclose(_rotate(F0, R.cpu()), molecule_rot.force.detach().cpu(), rtol=1e-4, atol=1e-4)


def test_rhf_rejects_odd_electron_counts(device):
    species = torch.as_tensor([[6, 1, 1, 1]], dtype=torch.int64, device=device)
    coordinates = torch.tensor([[[0.0, 0.0, 0.0], [1.08, 0.0, 0.0], [-0.54, 0.935, 0.0], [-0.54, -0.935, 0.0]]],
                               dtype=torch.float64, device=device)  # fmt: skip
    const = Constants().to(device)
    seqm_parameters = {"method": "AM1", "scf_eps": 1.0e-6, "scf_converger": [1]}
    with pytest.raises(ValueError) as excinfo:
        Molecule(const, seqm_parameters, coordinates, species).to(device)
    assert str(excinfo.value) == reference("rhf_odd_electron_error")["message"]


def test_unsorted_species_rows_are_rejected(device):
    species = torch.as_tensor([[6, 1, 1, 1, 1], [1, 6, 1, 1, 1]], dtype=torch.int64, device=device)
    coordinates = torch.randn(2, 5, 3, dtype=torch.float64, device=device)
    with pytest.raises(ValueError, match="species must be non-increasing along each row, but row 1 is not sorted."):
        Molecule(Constants().to(device), {"method": "AM1", "scf_eps": 1.0e-6, "scf_converger": [1]}, coordinates, species)
