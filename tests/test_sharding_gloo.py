"""N>1 path on CPU: two gloo ranks shard one batch by molecule, run the drop-in forward on their shard
(host-emulation kernels stand in for the GPU here) and all_gather energies / forces back into global order.
Checks the result against the reference-generated fixture, i.e. sharding changes nothing per molecule."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_default_dtype(torch.float64)
    import pyseqm_b200 as seqm
    from helpers import hostemu_lib
    from pyseqm_b200.sharding import run_sharded

    lib = hostemu_lib()
    g = load_golden("cfg2_PM3_48")
    species = torch.as_tensor(g["species"])
    coords = torch.as_tensor(g["coordinates"])
    sp = dict(g["seqm_parameters"])
    sp["scf_converger"] = [0, 0.3]  # batch-independent iteration path (DIIS resets are batch-global)
    sp["scf_eps"] = 1e-8

    def make_mol(const, p, c, s):
        m = seqm.Molecule(const, p, c, s, _lib=lib)
        m.verbose = False
        return m

    out = run_sharded(species, coords, sp, seqm.Constants(), make_mol, seqm.Electronic_Structure)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), Etot=out["Etot"].numpy(), force=out["force"].numpy(),
                 nc=out["notconverged"].numpy())  # fmt: skip
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_forward(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    g = load_golden("cfg2_PM3_48")
    assert got["nc"].sum() == 0
    assert np.abs(got["Etot"] - g["Etot"]).max() < 1e-6
    assert np.abs(got["force"] - g["force"]).max() < 1e-5


def _gather_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyseqm_b200.sharding import gather_results, shard_indices

    n = 11
    ok = True
    # two shardings of the same size with different cost orders, gathered back to back from freshly allocated index
    # tensors (the round-1 cache keyed on data_ptr could serve the first plan to the second call), plus an uneven one
    for trial, cost in enumerate([torch.arange(n, dtype=torch.float64), torch.arange(n, dtype=torch.float64).flip(0) ** 2,
                                  torch.tensor([3.0, 1, 4, 1, 5, 9, 2, 6, 5, 3, 5])]):
        idx = shard_indices(cost, world, rank)
        vals = {"a": idx.to(torch.float64) * 10.0 + trial, "b": torch.stack((idx, 2 * idx), dim=1).to(torch.float64),
                "flag": (idx % 2).to(torch.int32)}
        out = gather_results(vals, idx, n, nmax=None if trial == 2 else -(-n // world))
        g = torch.arange(n)
        ok &= bool(torch.equal(out["a"], g.to(torch.float64) * 10.0 + trial))
        ok &= bool(torch.equal(out["b"], torch.stack((g, 2 * g), dim=1).to(torch.float64)))
        ok &= bool(torch.equal(out["flag"], (g % 2).to(torch.int32)))
        del idx
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([ok]))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_two_orderings_back_to_back(tmp_path):
    port = _free_port()
    mp.spawn(_gather_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert bool(np.load(os.path.join(str(tmp_path), f"ok{r}.npy"))[0])


def test_shard_indices_balance_and_cover():
    from pyseqm_b200.sharding import shard_indices

    cost = torch.tensor([5.0, 1.0, 9.0, 3.0, 7.0, 2.0, 8.0])
    parts = [shard_indices(cost, 3, r) for r in range(3)]
    allidx = torch.sort(torch.cat(parts)).values
    assert allidx.tolist() == list(range(7))
    sums = [float(cost[p].sum()) for p in parts]
    assert max(sums) - min(sums) <= float(cost.max())
