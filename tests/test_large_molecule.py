"""Large-molecule path (n > 118 orbitals): global-memory Fock build, SP2 density by the FP64 GEMM, GEMM-based DIIS;
and the mid-size (119..256 orbital) eigensolver route.  All cases are pinned at the north-star tolerances (1e-6 eV,
1e-8, 1e-5 eV/A) with equal SCF iteration counts; the measured margins are quoted at each assertion."""
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
    """Large path (matrices in global memory, SP2 by the symmetric DMMA product, DIIS through GEMM commutators) on a
    216-orbital coronene dimer, stacked 3.5 A apart and shifted sideways, against the oracle at the north-star tolerances
    with equal iteration counts (measured margins, tools/sp2_margins.py: 1.5e-11 eV, 7e-13, 3e-8 eV/A; 40 = 40 iterations).
    The unshifted dimer of round 1 (coronene_dimer, two exactly superposed copies 6 A apart) has exactly degenerate
    frontier orbitals: there the DIIS tail wanders for 100-190 iterations in BOTH implementations and only the energy
    agrees tightly, so it is no parity case."""
    import seqm_oracle as so
    from helpers import run_molecule

    s2, c2 = stacked(3.5, 1.2)
    for eps in (1e-7, 1e-5):
        sp = {"method": "AM1", "scf_eps": 1e-6, "scf_converger": [2], "sp2": [True, eps]}
        ref = so.single_point(s2, c2, sp)
        mol, es = run_molecule(lib, device, s2, c2, sp)
        assert int(mol._plan.nmax) == 216 and mol._plan.large
        assert not bool(es.notconverged.any())
        assert mol.n_scf_iter == ref["n_scf_iter"]
        assert np.abs(mol.Etot.cpu().numpy() - ref["Etot"]).max() < 1e-6
        assert np.abs(mol.dm.cpu().numpy() - ref["dm"]).max() < 1e-8
        assert np.abs(mol.force.cpu().numpy() - ref["force"]).max() < 1e-5
        assert np.abs(mol.e_gap.cpu().numpy() - ref["e_gap"]).max() < 1e-6


def stacked(dz, shift, second="coronene.xyz"):
    """coronene with a second molecule stacked dz above it and shifted sideways (no exact degeneracies)."""
    import os

    import pyseqm_b200 as seqm
    from conftest import GOLDEN

    s, c = seqm.read_xyz([os.path.join(GOLDEN, "xyz", "coronene.xyz")])
    t, d = seqm.read_xyz([os.path.join(GOLDEN, "xyz", second)])
    Z = np.concatenate([s[0], t[0]])
    X = np.concatenate([c[0], d[0] + np.array([shift, 0.3 * shift, dz])])
    keep = Z > 0
    Z, X = Z[keep], X[keep]
    o = np.argsort(-Z, kind="stable")
    return Z[o][None], X[o][None]


def check_mid_eigensolver(lib, device):
    """Eigensolver route between 119 and 256 orbitals (north_star: one-sided Jacobi up to about 256 orbitals;
    sym_eig_trunc, diag.py:110-241) against the oracle at north-star tolerances with equal iteration counts:
    coronene + benzene (138 orbitals, matrix in shared memory) and a shifted coronene dimer (216 orbitals, matrix in L2)."""
    import seqm_oracle as so
    from helpers import run_molecule

    sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [2]}
    for S, C, norb in ((*stacked(3.4, 0.7, "benzene.xyz"), 138), (*stacked(3.5, 1.2), 216)):
        ref = so.single_point(S, C, sp)
        mol, es = run_molecule(lib, device, S, C, sp)
        assert int(mol._plan.nmax) == norb and mol._plan.large
        assert not bool(es.notconverged.any())
        assert mol.n_scf_iter == ref["n_scf_iter"]
        assert np.abs(mol.Etot.cpu().numpy() - ref["Etot"]).max() < 1e-6
        assert np.abs(mol.dm.cpu().numpy() - ref["dm"]).max() < 1e-8
        assert np.abs(mol.force.cpu().numpy() - ref["force"]).max() < 1e-5
        assert np.abs(mol.e_gap.cpu().numpy() - ref["e_gap"]).max() < 1e-6
        n = norb
        assert np.abs(mol.e_mo.cpu().numpy()[:, :n] - ref["e_mo"][:, :n]).max() < 1e-6
    # beyond 256 orbitals the eigensolver route is refused instead of silently doing something else
    s3, c3 = stacked(3.5, 1.2)
    s3 = np.concatenate([s3[0][s3[0] > 1], s3[0][s3[0] > 1], s3[0][s3[0] == 1]])[None]
    c3 = np.concatenate([c3[0][: (s3[0] > 1).sum() // 2], c3[0][: (s3[0] > 1).sum() // 2] + np.array([0.0, 0.0, 7.0]),
                         c3[0][(s3[0] > 1).sum() // 2:]])[None]
    with pytest.raises(NotImplementedError, match="SP2"):
        run_molecule(lib, device, s3, c3, sp)


def test_hostemu_mid_size_eigensolver():
    from helpers import hostemu_lib

    check_mid_eigensolver(hostemu_lib(), torch.device("cpu"))


@pytest.mark.gpu
def test_gpu_mid_size_eigensolver():
    from helpers import cuda_lib

    check_mid_eigensolver(cuda_lib(), torch.device("cuda:0"))


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
@pytest.mark.parametrize("n", [96, 216, 777, 1520])
def test_gpu_square_product_against_torch(n):
    """seqm_square_product: the general DMMA GEMM and the symmetric upper-triangle kernel against torch.matmul (cuBLAS)."""
    from helpers import cuda_lib
    from pyseqm_b200._lib import ptr

    lib = cuda_lib()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(n)
    A = torch.randn(n, n, generator=g, dtype=torch.float64).to(dev)
    B = torch.randn(n, n, generator=g, dtype=torch.float64).to(dev)
    X = ((A + A.T) * 0.5).contiguous()
    C1, C2 = torch.full_like(A, float("nan")), torch.full_like(A, float("nan"))
    lib.check(lib.dll.seqm_square_product(n, ptr(A), ptr(B), ptr(C1), None), "seqm_square_product")
    lib.check(lib.dll.seqm_square_product(n, ptr(X), None, ptr(C2), None), "seqm_square_product")
    torch.cuda.synchronize()
    tol = 1e-13 * n * float(A.abs().max()) ** 2
    assert float((C1 - A @ B).abs().max()) < tol
    assert float((C2 - X @ X).abs().max()) < tol
    assert float((C2 - C2.T).abs().max()) == 0.0  # mirrored tiles and symmetric diagonal tiles: exactly symmetric


@pytest.mark.gpu
def test_gpu_c380_against_reference():
    """BASELINE configs[3]: C380 fullerene, 1520 orbitals, AM1, SCF 1e-6 (DIIS), SP2 1e-5.  Reference: 41 iterations."""
    from helpers import cuda_lib, run_molecule

    g = load_golden("cfg4_C380_AM1_sp2")
    mol, es = run_molecule(cuda_lib(), torch.device("cuda:0"), g["species"], g["coordinates"], g["seqm_parameters"])
    assert not bool(es.notconverged.any())
    # north-star tolerances; measured margins on a B200 (tools/sp2_margins.py): 41 = 41 iterations, dEtot 9e-10 eV,
    # dEnuc 3e-9 eV, dForce 2.6e-9 eV/A, dq 7e-11, dgap 3e-11 eV
    assert mol.n_scf_iter == g["n_scf_iter"] == 41
    assert abs(float(mol.Etot[0]) - float(g["Etot"][0])) < 1e-6
    assert abs(float(mol.Enuc[0]) - float(g["Enuc"][0])) < 1e-6
    assert np.abs(mol.force.cpu().numpy() - g["force"]).max() < 1e-5
    assert np.abs(mol.q.cpu().numpy() - g["q"]).max() < 1e-8
    assert abs(float(mol.e_gap[0]) - float(g["e_gap"][0])) < 1e-6
