"""Where the eigensolver's cycles go over one full bench step (thread-0 clock64 spans summed over CTAs)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
import bench
import pyseqm_b200 as seqm
from pyseqm_b200._lib import get_lib
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
lib = get_lib()
species, coords, sha = bench.workload(4096, 0)
const = seqm.Constants().to(dev)
mol = seqm.Molecule(const, dict(bench.SP), torch.as_tensor(coords, device=dev), torch.as_tensor(species, device=dev))
mol.verbose = False
es = seqm.Electronic_Structure(dict(bench.SP))
for _ in range(2):
    es(mol)
torch.cuda.synchronize()
lib.jacobi_stats(reset=True)
es(mol)
torch.cuda.synchronize()
st = lib.jacobi_stats(reset=True)
tot = st["cycles_total"]
print(st)
print("solves %d  sweeps/solve %.2f  first-order finishes %.1f%%  no-sweep %.1f%%" % (
    st["molecules"], st["sweeps"] / st["molecules"], 100 * st["first_order_finishes"] / st["molecules"], 100 * st["no_sweep"] / st["molecules"]))
for k in ("transform", "sweeps", "epilogue"):
    print("  %-10s %5.1f %% of CTA cycles" % (k, 100.0 * st["cycles_" + k] / tot))
print("  cycles per solve %.0f ; per sweep %.0f" % (tot / st["molecules"], st["cycles_sweeps"] / max(st["sweeps"], 1)))
