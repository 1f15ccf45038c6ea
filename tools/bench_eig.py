"""Eigensolver alone on the configs[1] batch: cold solves of F(P0) for every molecule (sweep dominated), ms per call.
    python tools/bench_eig.py [nmol]      (SEQM_B200_LIB selects a library variant)"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pyseqm_b200 as seqm  # noqa: E402
from pyseqm_b200 import engine  # noqa: E402

nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
species, coords, _ = bench.workload(nmol, 0)
mol = seqm.Molecule(seqm.Constants().to(dev), dict(bench.SP), torch.as_tensor(coords, device=dev), torch.as_tensor(species, device=dev))
plan = mol._plan
xyz = mol._refresh_geometry()
w, hab = engine.op_pair_integrals(plan, xyz)
H = engine.op_hcore(plan, w, hab)
F = engine.op_fock(plan, engine.op_initial_density(plan), H, w)
lib = plan.lib
for warm in (False, True):
    Cg = None
    if warm:
        _, _, Cg = engine.op_eig_density(plan, F, want_P=True, want_C=True)
        F2 = F + 0.02 * engine.op_fock(plan, engine.op_initial_density(plan), plan.new_mat(), w)  # a nearby Fock matrix
    Fx = F2 if warm else F
    for _ in range(3):
        engine.op_eig_density(plan, Fx, want_P=True, want_C=True, Cguess=Cg, want_e=False)
    lib.jacobi_stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        engine.op_eig_density(plan, Fx, want_P=True, want_C=True, Cguess=Cg, want_e=False)
    e1.record()
    torch.cuda.synchronize()
    st = lib.jacobi_stats(reset=True)
    print("warm" if warm else "cold", "ms per call %.3f" % (e0.elapsed_time(e1) / 10), "sweeps per solve %.2f" % (st["sweeps"] / max(st["molecules"], 1)))
