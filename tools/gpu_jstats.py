import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
import bench
import pyseqm_b200 as seqm
from pyseqm_b200 import engine
from pyseqm_b200._lib import get_lib
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
lib = get_lib()
species, coords, sha = bench.workload(4096, 0)
plan = engine.BatchPlan(lib, torch.as_tensor(species, device=dev), "PM3")
xyz = plan.real_xyz(torch.as_tensor(coords, device=dev))
w, hab = engine.op_pair_integrals(plan, xyz); H = engine.op_hcore(plan, w, hab)
P = engine.op_initial_density(plan)
F = engine.op_fock(plan, P, H, w)
lib.jacobi_stats()
def timed(fn):
    torch.cuda.synchronize(); t=time.perf_counter(); r=fn(); torch.cuda.synchronize(); return r, (time.perf_counter()-t)*1e3
for it in range(6):
    (e, P1, C1), ms = timed(lambda: engine.op_eig_density(plan, F, want_C=True, Cguess=(C1 if it else None)))
    st = lib.jacobi_stats()
    print(f"iter {it}: {ms:.3f} ms  sweeps/mol {st['sweeps']/st['molecules']:.2f} first-order finishes {st['first_order_finishes']/st['molecules']:.2f}  avg n {float(plan.norb.double().mean()):.1f}")
    F = engine.op_fock(plan, 0.5*P + 0.5*P1 if it == 0 else P1, H, w); P = P1
