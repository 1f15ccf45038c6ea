"""Count the GPU kernel launches of ONE Electronic_Structure forward (SCF + force) on the configs[1] batch, split into
this library's kernels and eager ATen kernels.  Output: profiles/forward_launches_<tag>.txt.

    python tools/count_launches.py r02 [nmol]
"""
import collections
import sys

import torch

sys.path.insert(0, ".")
import pyseqm_b200 as seqm  # noqa: E402
import bench  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
    nmol = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    dev = torch.device("cuda:0")
    torch.set_default_dtype(torch.float64)
    species, coords, _ = bench.workload(nmol, 0)
    const = seqm.Constants().to(dev)
    es = seqm.Electronic_Structure(dict(bench.SP))
    mol = seqm.Molecule(const, dict(bench.SP), torch.as_tensor(coords, device=dev), torch.as_tensor(species, device=dev))
    mol.verbose = False
    for _ in range(2):
        es(mol)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        es(mol)
        torch.cuda.synchronize()
    cnt = collections.Counter()
    dur = collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            cnt[ev.name] += 1
            dur[ev.name] += ev.device_time
    ours = {k for k in cnt if "seqm" in k or "_kernel" in k and "at::" not in k and "elementwise" not in k}
    lines = [f"one Electronic_Structure forward (bench.SP: configs[1]), {nmol} QM9-like molecules, n_scf_iter={mol.n_scf_iter}",
             f"library kernels: {sum(cnt[k] for k in ours)} launches, eager ATen / memcpy / memset: {sum(cnt[k] for k in cnt if k not in ours)} launches",
             ""]  # fmt: skip
    for k, n in sorted(cnt.items(), key=lambda kv: -dur[kv[0]]):
        lines.append(f"{'LIB ' if k in ours else 'ATEN'} {n:5d} x {dur[k] / 1e3:10.3f} ms  {k[:110]}")
    out = "\n".join(lines)
    print(out)
    with open(f"profiles/forward_launches_{tag}.txt", "w") as f:
        f.write(out + "\n")


if __name__ == "__main__":
    main()
