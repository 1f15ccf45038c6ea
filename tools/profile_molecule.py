"""GPU kernels and wall time of Molecule() construction (parser + parameter gather) on the configs[1] batch."""
import collections
import os
import sys
import time

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pyseqm_b200 as seqm  # noqa: E402

torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
species, coords, _ = bench.workload(4096, 0)
const = seqm.Constants().to(dev)
s_d, c_d = torch.as_tensor(species, device=dev), torch.as_tensor(coords, device=dev)
for _ in range(3):
    mol = seqm.Molecule(const, dict(bench.SP), c_d, s_d)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    mol = seqm.Molecule(const, dict(bench.SP), c_d, s_d)
torch.cuda.synchronize()
print("Molecule() wall ms", (time.perf_counter() - t0) / 20 * 1e3)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    mol = seqm.Molecule(const, dict(bench.SP), c_d, s_d)
    torch.cuda.synchronize()
cnt, dur = collections.Counter(), collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        cnt[ev.name] += 1
        dur[ev.name] += ev.device_time
print("launches", sum(cnt.values()), "GPU ms", sum(dur.values()) / 1e3)
for k, n in sorted(cnt.items(), key=lambda kv: -dur[kv[0]])[:14]:
    print(f"{n:4d} x {dur[k] / 1e3:8.3f} ms  {k[:110]}")
