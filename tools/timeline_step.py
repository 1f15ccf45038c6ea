"""GPU timeline of ONE configs[1] forward (pipelined DIIS): busy time (union of kernel intervals), summed kernel time, idle gaps.
    python tools/timeline_step.py"""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pyseqm_b200 as seqm  # noqa: E402

torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
species, coords, _ = bench.workload(4096, 0)
mol = seqm.Molecule(seqm.Constants().to(dev), dict(bench.SP), torch.as_tensor(coords, device=dev), torch.as_tensor(species, device=dev))
mol.verbose = False
es = seqm.Electronic_Structure(dict(bench.SP))
for _ in range(3):
    es(mol)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    es(mol)
    torch.cuda.synchronize()
ev = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort()
t0, t1 = ev[0][0], max(e[1] for e in ev)
tot = sum(e[1] - e[0] for e in ev)
busy, cur_s, cur_e = 0.0, None, None
gaps = []
for s, e, _ in ev:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, cur_e - t0))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print("wall %.3f ms  busy (union) %.3f ms  summed kernel time %.3f ms  idle %.3f ms in %d gaps" % ((t1 - t0) / 1e3, busy / 1e3, tot / 1e3, (t1 - t0 - busy) / 1e3, len(gaps)))
gaps.sort(reverse=True)
print("largest gaps (us, at ms):", [(round(g, 1), round(a / 1e3, 2)) for g, a in gaps[:12]])
jac = [(s, e) for s, e, n in ev if "jacobi" in n]
jb, cs, ce = 0.0, None, None
for s, e in sorted(jac):
    if ce is None or s > ce:
        if ce is not None:
            jb += ce - cs
        cs, ce = s, e
    else:
        ce = max(ce, e)
jb += ce - cs
print("eigensolver: union %.3f ms, summed %.3f ms" % (jb / 1e3, sum(e - s for s, e in jac) / 1e3))
