"""sym_eig_trunc with the reference's signature (seqm/seqm_functions/diag.py:110-241), served by the batched
shared-memory Jacobi eigensolver (`seqm_eig_density`)."""
import torch

from .. import engine
from ._plans import matrix_plan


def sym_eig_trunc(x, nheavyatom, nH, nocc, eig_only=False):
    """x: dense padded Fock matrices (nmol, 4 molsize, 4 molsize) or one (N, N) matrix.  Returns (e, P, v) --
    eigenvalues zero-padded to 4 molsize, P = 2 C_occ C_occ^T in the dense layout, v (nmol, nmax, nmax) -- or
    (e, v) when `eig_only`."""
    if x.dim() == 4:
        raise NotImplementedError("unrestricted (nmol, 2, N, N) input is not on the B200 path")
    single = x.dim() == 2
    xb = x.unsqueeze(0) if single else x
    dev = x.device
    nh = torch.as_tensor(nheavyatom, device=dev).reshape(-1)
    ny = torch.as_tensor(nH, device=dev).reshape(-1)
    no = torch.as_tensor(nocc, device=dev).reshape(-1)
    plan = matrix_plan(nh, ny, no)
    if plan.large:
        raise NotImplementedError(f"{plan.nmax} orbitals exceed the shared-memory eigensolver; use SP2")
    if xb.shape[1] != 4 * plan.molsize:  # the stand-in species are as wide as the largest molecule only
        from .pack import pack, unpack

        F = engine.op_pack(plan, unpack(pack(xb, nh, ny), nh, ny, 4 * plan.molsize))
    else:
        F = engine.op_pack(plan, xb)
    e_n, P, Cm = engine.op_eig_density(plan, F, want_P=not eig_only, want_C=True)
    e = torch.zeros((xb.shape[0], xb.shape[1]), dtype=x.dtype, device=dev)
    e[:, : plan.nmax] = e_n
    v = engine.op_orbitals_dense(plan, Cm)
    if single:
        e, v = e[0], v[0]
    if eig_only:
        return e, v
    Pd = engine.op_unpack(plan, P)
    if Pd.shape[1] != xb.shape[1]:
        from .pack import pack, unpack

        Pd = unpack(pack(Pd, nh, ny), nh, ny, xb.shape[1])
    return e, (Pd[0] if single else Pd), v
