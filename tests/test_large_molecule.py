"""Large-molecule path (n > 118 orbitals): global-memory Fock build, SP2 density by the FP64 GEMM, GEMM-based DIIS.
SP2 (eps = 1e-5) leaves O(eps) noise in the density, so SCF iteration paths of two implementations differ in
summation order and may not hit the 1e-6 energy criterion at the same iteration; energies are variational and
still agree far below 1e-6 eV.  The tolerances below are the ones SP2 itself supports."""
import numpy as np
import pytest
import torch

from conftest import load_golden


def coronene_dimer():
    import os

    import pyseqm_b200 as seqm
    from conftest import GOLDEN

    s, c = seqm.read_xyz([os.path.join(GOLDEN, "xyz", "coronene.xyz")])
    heavy, hyd = s[0] > 1, s[0] == 1
    sh = c[0] + np.array([0.0, 0.0, 6.0])
    s2 = np.concatenate([s[0][heavy], s[0][heavy], s[0][hyd], s[0][hyd]])[None]
    c2 = np.concatenate([c[0][heavy], sh[heavy], c[0][hyd], sh[hyd]])[None]
    return s2, c2  # 72 atoms, 216 orbitals


def check_dimer(lib, device):
    import seqm_oracle as so
    from helpers import run_molecule

    s2, c2 = coronene_dimer()
    sp = {"method": "AM1", "scf_eps": 1e-6, "scf_converger": [2], "sp2": [True, 1e-7]}
    ref = so.single_point(s2, c2, sp)
    mol, es = run_molecule(lib, device, s2, c2, sp)
    assert not bool(es.notconverged.any())
    # no iteration-count assertion: with SP2 noise the DIIS tail wanders (see the module docstring)
    assert np.abs(mol.Etot.cpu().numpy() - ref["Etot"]).max() < 2e-6
    assert np.abs(mol.dm.cpu().numpy() - ref["dm"]).max() < 2e-3  # SP2-limited (weakly coupled stacked dimer)
    assert np.abs(mol.force.cpu().numpy() - ref["force"]).max() < 1e-2
    # the eigensolver route is refused for this size instead of silently doing something else
    with pytest.raises(NotImplementedError, match="SP2"):
        run_molecule(lib, device, s2, c2, {"method": "AM1", "scf_eps": 1e-6, "scf_converger": [2]})


def test_hostemu_large_path_dimer():
    from helpers import hostemu_lib

    check_dimer(hostemu_lib(), torch.device("cpu"))


@pytest.mark.gpu
def test_gpu_large_path_dimer():
    from helpers import cuda_lib

    check_dimer(cuda_lib(), torch.device("cuda:0"))


@pytest.mark.gpu
def test_gpu_dgemm_against_torch():
    """The hand-written FP64 GEMM (through SP2's first product X0^2) against torch.matmul."""
    import ctypes as C

    from helpers import cuda_lib
    from pyseqm_b200 import engine

    lib = cuda_lib()
    dev = torch.device("cuda:0")
    s2, c2 = coronene_dimer()
    plan = engine.BatchPlan(lib, torch.as_tensor(s2, device=dev), "AM1")
    n = plan.nmax
    g = torch.Generator(device="cpu").manual_seed(0)
    A = torch.randn(n, n, generator=g, dtype=torch.float64)
    F = ((A + A.T) * 0.5).to(dev).reshape(-1).contiguous()
    P, nit = engine.op_sp2_density(plan, F, 1e-7)
    Pm = P[: n * n].reshape(n, n)
    assert float((Pm - Pm.T).abs().max()) < 1e-10
    assert abs(float(Pm.diagonal().sum()) - 2.0 * float(plan.nocc[0])) < 1e-5
    assert float((Pm @ Pm - 2.0 * Pm).abs().max()) < 1e-4  # idempotent to the SP2 tolerance
    Fm = F.reshape(n, n)
    assert float((Fm @ Pm - Pm @ Fm).abs().max()) < 1e-3  # commutes with F


@pytest.mark.gpu
def test_gpu_c380_against_reference():
    """BASELINE configs[3]: C380 fullerene, 1520 orbitals, AM1, SCF 1e-6 (DIIS), SP2 1e-5.  Reference: 41 iterations."""
    from helpers import cuda_lib, run_molecule

    g = load_golden("cfg4_C380_AM1_sp2")
    mol, es = run_molecule(cuda_lib(), torch.device("cuda:0"), g["species"], g["coordinates"], g["seqm_parameters"])
    assert not bool(es.notconverged.any())
    assert abs(mol.n_scf_iter - g["n_scf_iter"]) <= 10  # 41 in the reference; equal in practice, not guaranteed under SP2 noise
    assert abs(float(mol.Etot[0]) - float(g["Etot"][0])) < 1e-5
    assert abs(float(mol.Enuc[0]) - float(g["Enuc"][0])) < 1e-6
    assert np.abs(mol.force.cpu().numpy() - g["force"]).max() < 2e-3
    assert np.abs(mol.q.cpu().numpy() - g["q"]).max() < 1e-4
    assert abs(float(mol.e_gap[0]) - float(g["e_gap"][0])) < 1e-4
