"""The C-ABI library loads and exports every symbol include/seqm_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "seqm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(seqm_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_hot_path():
    names = declared_symbols()
    for must in ("seqm_pair_integrals", "seqm_hcore", "seqm_fock", "seqm_eig_density", "seqm_sp2_density", "seqm_scf",
                 "seqm_gradient", "seqm_nuclear_energy", "seqm_elec_energy"):  # fmt: skip
        assert must in names


def test_cuda_library_exports_all_declared_symbols():
    import __graft_entry__ as ge
    from pyseqm_b200._lib import LIB_PATH, build_library

    build_library()
    dll = ctypes.CDLL(LIB_PATH)
    for name in declared_symbols():
        assert hasattr(dll, name), name
    dll.seqm_abi_version.restype = ctypes.c_int
    assert dll.seqm_abi_version() == 3
    # the host-emulation build exposes the same ABI
    emu = ctypes.CDLL(ge.build_hostemu())
    for name in declared_symbols():
        assert hasattr(emu, name), name


def test_product_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import pyseqm_b200 as seqm
    from pyseqm_b200._lib import SeqmError

    species = torch.tensor([[6, 1, 1, 1, 1]])
    coords = torch.randn(1, 5, 3, dtype=torch.float64)
    sp = {"method": "AM1", "scf_eps": 1e-6, "scf_converger": [2]}
    with pytest.raises(SeqmError, match="no CPU fallback"):
        seqm.Molecule(seqm.Constants(), sp, coords, species)


def test_option_matrix_rejections():
    import torch

    import pyseqm_b200 as seqm

    species = torch.tensor([[6, 1, 1, 1, 1]])
    coords = torch.randn(1, 5, 3, dtype=torch.float64)
    with pytest.raises(NotImplementedError, match="learned parameters with PM6 d-shell"):
        seqm.Molecule(seqm.Constants(), {"method": "PM6", "scf_eps": 1e-6, "scf_converger": [2], "learned": ["U_ss"]},
                      torch.randn(1, 3, 3, dtype=torch.float64), torch.tensor([[16, 1, 1]]),
                      learned_parameters={"U_ss": torch.zeros(3, dtype=torch.float64)})  # fmt: skip
    for bad in ({"method": "PM7"}, {"UHF": True}, {"excited_states": {"n_states": 2}}, {"scf_backward": 1},
                {"scf_converger": [3, 0.1]}, {"dispersion": True}):  # fmt: skip
        sp = {"method": "AM1", "scf_eps": 1e-6, "scf_converger": [2]}
        sp.update(bad)
        with pytest.raises(NotImplementedError):
            seqm.Molecule(seqm.Constants(), sp, coords, species)
    with pytest.raises(ValueError, match="non-increasing"):
        seqm.Molecule(seqm.Constants(), {"method": "AM1", "scf_eps": 1e-6, "scf_converger": [2]}, coords,
                      torch.tensor([[1, 6, 1, 1, 1]]))  # fmt: skip
