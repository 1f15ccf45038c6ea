"""Multi-GPU plumbing: molecules (and MD replicas) are independent units, so a batch is dealt to the ranks
of one node and nothing is exchanged inside the SCF; the only collective gathers per-molecule results
(SURVEY 8(e)).  One process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
import math

import torch
import torch.distributed as dist


def shard_indices(cost, world_size, rank):
    """Deal molecules to ranks round-robin in order of decreasing cost (~ n_orbitals^3) so that every GPU
    gets the same size mix.  Returns the sorted global indices owned by `rank`."""
    order = torch.argsort(torch.as_tensor(cost), descending=True, stable=True)
    mine = order[rank::world_size]
    return torch.sort(mine).values


_GATHER_PLANS = {}


def _gather_plan(index, nmol_total, group):
    """Index exchange of one sharding (done once, cached): how many molecules every rank owns, where each row of
    the padded all-gathered buffer goes in global molecule order."""
    ws = dist.get_world_size(group)
    key = (index.data_ptr(), int(index.shape[0]), int(nmol_total), ws, str(index.device))
    plan = _GATHER_PLANS.get(key)
    if plan is None:
        dev = index.device
        n_local = torch.tensor([index.shape[0]], device=dev, dtype=torch.int64)
        counts = [torch.zeros_like(n_local) for _ in range(ws)]
        dist.all_gather(counts, n_local, group=group)
        counts = [int(c) for c in counts]
        nmax = max(counts)
        pad_idx = torch.full((nmax,), -1, device=dev, dtype=torch.int64)
        pad_idx[: index.shape[0]] = index
        all_idx = [torch.empty_like(pad_idx) for _ in range(ws)]
        dist.all_gather(all_idx, pad_idx, group=group)
        src = torch.cat([r * nmax + torch.arange(counts[r], device=dev) for r in range(ws)])
        dest = torch.cat([all_idx[r][: counts[r]] for r in range(ws)])
        plan = (nmax, src, dest)
        if len(_GATHER_PLANS) > 16:
            _GATHER_PLANS.clear()
        _GATHER_PLANS[key] = plan
    return plan


def gather_results(local, index, nmol_total, group=None):
    """all_gather per-molecule tensors (first dim = local molecules) back into global molecule order: every tensor
    is packed into one (molecules, width) fp64 buffer, so a step costs ONE collective and one scatter.
    `local`: dict name -> tensor; `index`: global indices of the local molecules."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return {k: v for k, v in local.items()}
    ws = dist.get_world_size(group)
    nmax, src, dest = _gather_plan(index, nmol_total, group)
    dev = index.device
    nloc = index.shape[0]
    widths = {k: int(math.prod(v.shape[1:])) for k, v in local.items()}
    width = sum(widths.values())
    buf = torch.zeros((nmax, width), device=dev, dtype=torch.float64)
    col = 0
    for k, v in local.items():
        buf[:nloc, col : col + widths[k]] = v.reshape(nloc, widths[k]).to(torch.float64)
        col += widths[k]
    gathered = torch.empty((ws * nmax, width), device=dev, dtype=torch.float64)
    if dev.type == "cuda":
        dist.all_gather_into_tensor(gathered, buf, group=group)
    else:  # gloo (CPU tests)
        parts = [torch.empty_like(buf) for _ in range(ws)]
        dist.all_gather(parts, buf, group=group)
        gathered = torch.cat(parts, dim=0)
    full = torch.zeros((nmol_total, width), device=dev, dtype=torch.float64)
    full[dest] = gathered[src]
    out, col = {}, 0
    for k, v in local.items():
        out[k] = full[:, col : col + widths[k]].reshape((nmol_total,) + tuple(v.shape[1:])).to(v.dtype)
        col += widths[k]
    return out


def run_sharded(species, coordinates, seqm_parameters, const, make_molecule, make_driver, group=None):
    """Shard a global batch over the ranks, run the drop-in forward on the local shard, gather
    Etot / Hf / force / notconverged in global order.  Iteration counts are per shard (the reference's DIIS
    reset is a batch-global decision)."""
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    nheavy = (species > 1).sum(dim=1)
    nhyd = (species == 1).sum(dim=1)
    cost = (4 * nheavy + nhyd).to(torch.float64) ** 3
    idx = shard_indices(cost.cpu(), ws, rank).to(species.device)
    mol = make_molecule(const, dict(seqm_parameters), coordinates[idx].contiguous(), species[idx].contiguous())
    drv = make_driver(dict(seqm_parameters))
    drv(mol)
    local = dict(Etot=mol.Etot, Hf=mol.Hf, force=mol.force, notconverged=drv.notconverged.to(torch.int32))
    out = gather_results(local, idx, species.shape[0], group)
    out["n_scf_iter_local"] = mol.n_scf_iter
    return out
