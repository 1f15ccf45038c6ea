"""Per-kernel GPU time of ONE XL-BOMD step (SP2 route, configs[2]) (coronene replicas): python tools/profile_ksa.py [nrep]"""
import collections
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import pyseqm_b200 as seqm  # noqa: E402

nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
s, c = seqm.read_xyz([os.path.join(ROOT, "tests", "golden", "xyz", "coronene.xyz")] * nrep)
sp = {"method": "AM1", "scf_eps": 1.0e-7, "scf_converger": [2], "sp2": [True, 1.0e-5]}
mol = seqm.Molecule(seqm.Constants().to(dev), sp, torch.as_tensor(c, device=dev), torch.as_tensor(s, device=dev))
torch.manual_seed(0)
md = seqm.XL_BOMD(xl_bomd_params={"k": 6}, seqm_parameters=sp,
                      timestep=0.4, Temp=300.0)
md.initialize(mol)
for i in range(3):
    md._do_integrator_step(i, mol, dict())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3, 13):
    md._do_integrator_step(i, mol, dict())
e1.record()
torch.cuda.synchronize()
print("ms per step", e0.elapsed_time(e1) / 10)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    md._do_integrator_step(13, mol, dict())
    torch.cuda.synchronize()
cnt, dur = collections.Counter(), collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        cnt[ev.name] += 1
        dur[ev.name] += ev.device_time
print("launches", sum(cnt.values()), "GPU ms", sum(dur.values()) / 1e3)
for k, n in sorted(cnt.items(), key=lambda kv: -dur[kv[0]])[:18]:
    print(f"{n:5d} x {dur[k] / 1e3:9.3f} ms  {k[:100]}")
