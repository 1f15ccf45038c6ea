"""pack / unpack with the reference's signatures (seqm/seqm_functions/pack.py:64-96): dense padded
(…, 4 molsize, 4 molsize) <-> zero-padded packed (…, nmax, nmax).  Boundary conversions only: the kernels
work on the flat packed layout (`seqm_pack` / `seqm_unpack`) and never see either of these."""
import torch

from ._plans import dense_index, matrix_plan


def _as_batch(nHeavy, nHydro, device):
    nh = torch.as_tensor(nHeavy, device=device).reshape(-1)
    ny = torch.as_tensor(nHydro, device=device).reshape(-1)
    return nh, ny


def pack(x, nHeavy, nHydro):
    single = x.dim() == 2
    xb = x.unsqueeze(0) if single else x
    nh, ny = _as_batch(nHeavy, nHydro, x.device)
    plan = matrix_plan(nh, ny)
    idx, valid = dense_index(plan)
    rows = xb.gather(1, idx.unsqueeze(2).expand(-1, -1, xb.shape[2]))
    out = rows.gather(2, idx.unsqueeze(1).expand(-1, idx.shape[1], -1))
    out = out * (valid.unsqueeze(2) & valid.unsqueeze(1))
    return out[0] if single else out


def unpack(x0, nHeavy, nHydro, size):
    single = x0.dim() == 2
    xb = x0.unsqueeze(0) if single else x0
    nh, ny = _as_batch(nHeavy, nHydro, x0.device)
    plan = matrix_plan(nh, ny)
    idx, valid = dense_index(plan)
    n = idx.shape[1]
    src = xb[:, :n, :n] * (valid.unsqueeze(2) & valid.unsqueeze(1))
    tmp = torch.zeros((xb.shape[0], n, size), dtype=x0.dtype, device=x0.device)
    tmp.scatter_add_(2, idx.unsqueeze(1).expand(-1, n, -1), src)
    out = torch.zeros((xb.shape[0], size, size), dtype=x0.dtype, device=x0.device)
    out.scatter_add_(1, idx.unsqueeze(2).expand(-1, -1, size), tmp)
    return out[0] if single else out
