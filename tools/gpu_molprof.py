import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch, cProfile, pstats
import bench
import pyseqm_b200 as seqm
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
species, coords, sha = bench.workload(4096, 0)
const = seqm.Constants().to(dev)
s_d = torch.as_tensor(species, device=dev); c_d = torch.as_tensor(coords, device=dev)
for _ in range(2): mol = seqm.Molecule(const, dict(bench.SP), c_d, s_d)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5): mol = seqm.Molecule(const, dict(bench.SP), c_d, s_d)
torch.cuda.synchronize(); print("Molecule() %.3f ms" % ((time.perf_counter() - t) / 5 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(5): mol = seqm.Molecule(const, dict(bench.SP), c_d, s_d)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
