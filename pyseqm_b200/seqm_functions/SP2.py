"""SP2 with the reference's signature (seqm/seqm_functions/SP2.py:9-85): second-order spectral projection
of zero-padded packed Fock matrices.  Each molecule is purified at its native size (the reference's padding
to the largest active molecule is not reproduced; DESIGN.md, deviations)."""
import torch

from .. import engine
from ._plans import matrix_plan


def SP2(a, nocc, eps=1.0e-4, factor=2.0, nHeavy=None, nHydro=None):
    """a: (nmol, n, n) packed Fock matrices, zero beyond each molecule's size; nocc: occupied orbitals.  The true
    sizes come from `nHeavy`/`nHydro` when given, else from the zero padding of the diagonal."""
    if a.dtype != torch.float64:
        raise NotImplementedError("the B200 path is fp64 only")
    dev = a.device
    nmol, n = a.shape[0], a.shape[1]
    no = torch.as_tensor(nocc, device=dev).reshape(-1)
    if nHeavy is None:
        nz = (a.abs().sum(dim=2) > 0).to(torch.int64)
        size = (nz * torch.arange(1, n + 1, device=dev)).amax(dim=1)
        nh, ny = torch.zeros_like(size), size  # only the matrix size matters here
    else:
        nh = torch.as_tensor(nHeavy, device=dev).reshape(-1)
        ny = torch.as_tensor(nHydro, device=dev).reshape(-1)
    plan = matrix_plan(nh, ny, no)
    flat = plan.new_mat()
    sizes = plan.norb.tolist()
    mat0 = plan.t["mol_mat0"].tolist()
    for m in range(nmol):  # boundary conversion, not the hot path
        k = sizes[m]
        flat[mat0[m] : mat0[m] + k * k] = a[m, :k, :k].reshape(-1)
    eps = min(max(float(eps), 1.0e-7), 1.0e-3)  # SP2.py:28-31
    P, _ = engine.op_sp2_density(plan, flat, eps)
    out = torch.zeros_like(a)
    for m in range(nmol):
        k = sizes[m]
        out[m, :k, :k] = P[mat0[m] : mat0[m] + k * k].reshape(k, k)
    return out * (factor / 2.0)
