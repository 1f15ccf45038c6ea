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


def gather_results(local, index, nmol_total, group=None, nmax=None):
    """all_gather per-molecule tensors (first dim = local molecules) back into global molecule order.

    Every tensor is packed into one (rows, 1 + width) fp64 buffer whose column 0 is the molecule's GLOBAL index
    (-1 for padding rows), so a step costs ONE collective and one scatter and the destination of every row travels
    with the row: nothing is cached between calls and every rank issues exactly the same collectives whatever its
    local state (a cache keyed on local pointers could be hit on one rank and missed on another).
    `local`: dict name -> tensor; `index`: global indices of the local molecules; `nmax`: rows per rank (>= the
    largest local count; `ceil(nmol_total / world)` for the round-robin deal of `shard_indices`).  When None it is
    agreed with one extra MAX all-reduce per call."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return {k: v for k, v in local.items()}
    ws = dist.get_world_size(group)
    dev = index.device
    nloc = int(index.shape[0])
    if nmax is None:
        t = torch.tensor([nloc], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        nmax = int(t)
    if nloc > nmax:
        raise ValueError(f"gather_results: {nloc} local molecules exceed nmax={nmax}")
    widths = {k: int(math.prod(v.shape[1:])) for k, v in local.items()}
    width = 1 + sum(widths.values())
    buf = torch.zeros((nmax, width), device=dev, dtype=torch.float64)
    buf[:, 0] = -1.0
    buf[:nloc, 0] = index.to(torch.float64)  # exact below 2^53
    col = 1
    for k, v in local.items():
        buf[:nloc, col : col + widths[k]] = v.reshape(nloc, widths[k]).to(torch.float64)
        col += widths[k]
    if dev.type == "cuda":
        gathered = torch.empty((ws * nmax, width), device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(gathered, buf, group=group)
    else:  # gloo (CPU tests)
        parts = [torch.empty_like(buf) for _ in range(ws)]
        dist.all_gather(parts, buf, group=group)
        gathered = torch.cat(parts, dim=0)
    dest = gathered[:, 0].to(torch.int64)
    # padding rows (-1) are routed to one extra row that is dropped
    full = torch.zeros((nmol_total + 1, width - 1), device=dev, dtype=torch.float64)
    full[torch.where(dest < 0, nmol_total, dest)] = gathered[:, 1:]
    out, col = {}, 0
    for k, v in local.items():
        out[k] = full[:nmol_total, col : col + widths[k]].reshape((nmol_total,) + tuple(v.shape[1:])).to(v.dtype)
        col += widths[k]
    return out


class ShardedBatch:
    """One global batch dealt to the ranks (size-sorted round-robin): the local shard is built once and
    stays resident; `forward()` runs the drop-in driver on it and gathers Etot / Hf / force / notconverged in global
    molecule order with ONE collective.  Iteration counts are per shard: the reference's DIIS reset is a batch-global
    decision (scf_loop.py:1027), so parity for a shard is defined against the oracle run on that shard's molecules."""

    def __init__(self, species, coordinates, seqm_parameters, const, make_molecule, make_driver, group=None):
        self.group = group
        self.ws = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nmol_total = int(species.shape[0])
        nheavy = (species > 1).sum(dim=1)
        nhyd = (species == 1).sum(dim=1)
        cost = (4 * nheavy + nhyd).to(torch.float64) ** 3
        self.index = shard_indices(cost.cpu(), self.ws, self.rank).to(species.device)
        self.molecule = make_molecule(const, dict(seqm_parameters), coordinates[self.index].contiguous(),
                                      species[self.index].contiguous())  # fmt: skip
        self.driver = make_driver(dict(seqm_parameters))
        self.nmax = -(-self.nmol_total // self.ws)

    def forward(self):
        mol, drv = self.molecule, self.driver
        drv(mol)
        local = dict(Etot=mol.Etot, Hf=mol.Hf, force=mol.force, notconverged=drv.notconverged.to(torch.int32))
        out = gather_results(local, self.index, self.nmol_total, self.group, nmax=self.nmax)
        out["n_scf_iter_local"] = mol.n_scf_iter
        return out


def run_sharded(species, coordinates, seqm_parameters, const, make_molecule, make_driver, group=None):
    """Shard a global batch over the ranks, run the drop-in forward on the local shard, gather
    Etot / Hf / force / notconverged in global order (one-shot form of `ShardedBatch`)."""
    return ShardedBatch(species, coordinates, seqm_parameters, const, make_molecule, make_driver, group).forward()
