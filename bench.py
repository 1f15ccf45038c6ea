#!/usr/bin/env python
"""bench.py -- molecule-SCF/s on BASELINE.json configs[1]: PM3, synthetic QM9-size CHNO batch of 4096
molecules per GPU, SCF to 1e-7 with Pulay DIIS, energies + forces, fp64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one Electronic_Structure.forward over the batch (pair integrals, Hcore, SCF, energies, forces).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "molecule-SCF/s (PM3 QM9-size batch)"
UNIT = "molecule-SCF/s"
SP = {"method": "PM3", "scf_eps": 1.0e-7, "scf_converger": [2], "sp2": [False], "analytical_gradient": [True]}


def workload(nmol, rank):
    import numpy as np

    from pyseqm_b200.synthetic import batch_sha256, qm9_like_batch

    cache = f"/tmp/seqm_qm9like_{nmol}_{rank}.npz"
    if os.path.exists(cache):
        d = np.load(cache)
        return d["species"], d["coords"], str(d["sha"])
    species, coords = qm9_like_batch(nmol, seed=0, start=rank * nmol)
    sha = batch_sha256(species, coords)
    np.savez(cache, species=species, coords=coords, sha=sha)
    return species, coords, sha


def config_dict(nmol, ngpu, sha):
    return {
        "workload": "configs[1]: PM3, 4096 synthetic QM9-size CHNO molecules (<=29 atoms) per GPU, scf_eps 1e-7, "
        "scf_converger [2] (Pulay DIIS), energies + forces",
        "molecules_per_gpu": nmol, "global_batch": nmol * ngpu, "molsize": 29, "seed": 0, "sha256_rank0": sha[:16],
        "sharding": f"molecules, {ngpu} rank(s), NCCL gather of Etot/Hf/force only",
        "l2": "L2 flushed (256 MiB write) before every timed step; per-step working set (w 0.3 GB + DIIS history "
        "~0.8 GB) also exceeds the 126 MB L2",
    }  # fmt: skip


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        """Only rows sampled after this call are reported (the timed region).  The sampler itself is started before
        the warm-up steps: nvidia-smi's start-up (NVML attaching to every GPU of the box) stalls the driver for
        ~0.1-0.2 s on multi-GPU boxes, which must not land inside a timed step."""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows[getattr(self, "first", 0):]:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}  # fmt: skip


def xl_bomd_rate(seqm, dev, const, nrep, nsteps, world, dist):
    """replica-steps/s of XL-BOMD NVE on `nrep` coronene replicas per GPU (BASELINE configs[2]: SP2 density, eps 1e-5);
    the eigensolver density branch -- the only one the reference itself can run (xlbomd.py:359 crashes with SP2) -- is
    timed beside it.  The t = 0 SCF is excluded, as in SURVEY 8(d)."""
    out = _xl_bomd_rate(seqm, dev, const, nrep, nsteps, world, dist, [True, 1.0e-5])
    eig = _xl_bomd_rate(seqm, dev, const, nrep, nsteps, world, dist, [False])
    out["eigensolver_branch"] = {k: eig[k] for k in ("value", "unit", "ms_per_md_step", "method", "finite")}
    return out


def _xl_bomd_rate(seqm, dev, const, nrep, nsteps, world, dist, sp2):
    import torch

    xyz = os.path.join(ROOT, "tests", "golden", "xyz", "coronene.xyz")
    s, c = seqm.read_xyz([xyz] * nrep)
    sp = {"method": "AM1", "scf_eps": 1.0e-7, "scf_converger": [2], "sp2": sp2}
    torch.manual_seed(1234 + int(os.environ.get("RANK", "0")))
    mol = seqm.Molecule(const, sp, torch.as_tensor(c, device=dev), torch.as_tensor(s, device=dev))
    md = seqm.XL_BOMD(xl_bomd_params={"k": 6}, seqm_parameters=sp, timestep=0.4, Temp=300.0)
    md.initialize(mol)
    for i in range(3):
        md._do_integrator_step(i, mol, dict())
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3, 3 + nsteps):
        md._do_integrator_step(i, mol, dict())
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    Ek = md._kinetic_energy(mol)
    return {"metric": "XL-BOMD MD steps/s (replica-steps/s)", "value": nrep * world * nsteps / float(t), "unit": "replica-steps/s",
            "ms_per_md_step": float(t) / nsteps * 1e3, "replicas_per_gpu": nrep, "steps": nsteps, "molecule": "coronene C24H12 (108 orbitals)",
            "method": "AM1, k=6, dt=0.4 fs, 300 K, density by " + ("in-SM SP2 purification on the FP64 tensor cores, eps 1e-5"
                                                               if sp2[0] else "the Jacobi eigensolver (reference branch xlbomd.py:361)"),
            "finite": bool(torch.isfinite(mol.Etot).all() and torch.isfinite(Ek).all())}  # fmt: skip


def cpu_reference_rate(sample, cores, steps=1, warmup=0):
    """The oracle (numpy port of the reference's CPU path) on the first `sample` molecules of the rank-0 batch, one
    process, numpy's BLAS on all host threads -- the same execution model as the reference's own CPU path (torch
    intra-op threads over one big batch), which it matches in speed (~23 molecule-SCF/s on 8 cores for this batch)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import seqm_oracle as so

    species, coords, _ = workload(4096, 0)
    species, coords = species[:sample], coords[:sample]
    times = []
    import contextlib

    try:  # torchrun exports OMP_NUM_THREADS=1: ask the BLAS/OpenMP pools for every host thread explicitly
        from threadpoolctl import threadpool_limits

        pool = threadpool_limits(limits=cores)
    except Exception:
        pool = contextlib.nullcontext()
    with pool:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            out = so.single_point(species, coords, SP)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    assert np.all(np.isfinite(out["Etot"])) and not out["notconverged"].any()
    dt = sum(times) / len(times)
    return sample / dt, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = min(1024, 32 * cores)
    rate, dt = cpu_reference_rate(sample, cores, steps=max(1, min(args.steps, 3)), warmup=0)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(4096, args.gpus, workload(4096, 0)[2]),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {sample} molecules of the rank-0 batch in one process, numpy BLAS on {cores} host "
                                   "threads, oracle/seqm_oracle (numpy restatement of the reference CPU path; the reference "
                                   "is Python and cannot travel to the GPU box)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }  # fmt: skip
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nmol", type=int, default=4096)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--xl-replicas", type=int, default=1024, help="coronene replicas for the XL-BOMD line (0 = skip)")
    ap.add_argument("--xl-steps", type=int, default=20)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import pyseqm_b200 as seqm
    from pyseqm_b200._lib import get_lib
    from pyseqm_b200.sharding import gather_results

    torch.set_default_dtype(torch.float64)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = get_lib()

    nmol = args.nmol
    species_h, coords_h, sha = workload(nmol, rank)
    species_pin = torch.as_tensor(species_h).pin_memory()
    coords_pin = torch.as_tensor(coords_h).pin_memory()
    const = seqm.Constants().to(dev)
    species = species_pin.to(dev)
    coords = coords_pin.to(dev)
    mol = seqm.Molecule(const, dict(SP), coords, species)
    mol.verbose = False
    es = seqm.Electronic_Structure(dict(SP))
    gidx = torch.arange(rank * nmol, (rank + 1) * nmol, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        es(mol)
        if world > 1:
            return gather_results(dict(Etot=mol.Etot, Hf=mol.Hf, force=mol.force), gidx, nmol * world)
        return None

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    if rank == 0:
        sampler.mark()
    launches0 = lib.dll.seqm_launch_count()
    ms = []
    for _ in range(args.steps):
        flush.fill_(1.0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_resident()
        e1.record()
        barrier()
        ms.append(e0.elapsed_time(e1))
    launches = lib.dll.seqm_launch_count() - launches0
    print("resident step times (ms):", [round(x, 2) for x in ms], file=sys.stderr)
    clocks = sampler.stop() if rank == 0 else None
    t_local = torch.tensor([sum(ms) / len(ms)], device=dev)
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    ms_step = float(t_local)
    value = nmol * world / (ms_step * 1e-3)
    n_iter = mol.n_scf_iter
    nnot = int(es.notconverged.sum())

    # ---- end to end through the public API with host buffers (H2D + Molecule() + forward + D2H) -----------
    out_E = torch.empty(nmol, dtype=torch.float64).pin_memory()
    out_Hf = torch.empty(nmol, dtype=torch.float64).pin_memory()
    out_F = torch.empty((nmol, species_h.shape[1], 3), dtype=torch.float64).pin_memory()
    out_nc = torch.empty(nmol, dtype=torch.bool).pin_memory()

    def step_e2e():
        s_d = species_pin.to(dev, non_blocking=True)
        c_d = coords_pin.to(dev, non_blocking=True)
        m = seqm.Molecule(const, dict(SP), c_d, s_d)
        m.verbose = False
        es(m)
        out_E.copy_(m.Etot, non_blocking=True)
        out_Hf.copy_(m.Hf, non_blocking=True)
        out_F.copy_(m.force, non_blocking=True)
        out_nc.copy_(es.notconverged, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(max(1, min(args.warmup, 3))):
        step_e2e()
    e2e_t = []
    for _ in range(args.steps):
        flush.fill_(1.0)
        barrier()
        t0 = time.perf_counter()
        step_e2e()
        e2e_t.append(time.perf_counter() - t0)
        barrier()
    print("e2e step times (ms):", [round(t * 1e3, 2) for t in e2e_t], file=sys.stderr)
    t_e2e = torch.tensor([sum(e2e_t) / len(e2e_t)], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = nmol * world / float(t_e2e)
    h2d = species_pin.numel() * 8 + coords_pin.numel() * 8
    d2h = out_E.numel() * 8 + out_Hf.numel() * 8 + out_F.numel() * 8 + out_nc.numel()

    # ---- per-kernel breakdown (CUDA events on the launch stream) and roofline of the dominant kernel ------
    roofline, breakdown = None, None
    if rank == 0:
        # one extra, untimed step on a single stream (SEQM_B200_PIPELINE=1): the timed steps run the DIIS loop as two
        # half-batches out of phase on two streams, where per-kernel event intervals overlap and cannot be summed
        prev_pipe = os.environ.get("SEQM_B200_PIPELINE")
        os.environ["SEQM_B200_PIPELINE"] = "1"
        lib.jacobi_stats(reset=True)
        lib.profile_enable(True)
        flush.fill_(1.0)
        torch.cuda.synchronize()
        es(mol)
        prof = lib.profile_collect()
        lib.profile_enable(False)
        jstats = lib.jacobi_stats(reset=True)
        if prev_pipe is None:
            del os.environ["SEQM_B200_PIPELINE"]
        else:
            os.environ["SEQM_B200_PIPELINE"] = prev_pipe
        tot = sum(v[0] for v in prof.values()) or 1.0
        breakdown = {k: {"ms": round(v[0], 4), "launches": v[1], "share": round(v[0] / tot, 4)} for k, v in prof.items() if v[1]}
        plan = mol._plan
        n = plan.norb.to(torch.float64)
        nocc = plan.nocc.to(torch.float64)
        top = max(prof.items(), key=lambda kv: kv[1][0])[0]
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        fp64_peak = lib.dll.seqm_fp64_peak_tflops()
        # algorithmic work (DESIGN.md "Kernels"): the library counts the molecules the eigensolver actually solved
        # (converged molecules drop out); each is charged the batch-mean 10 n^3 + 2 n^2 nocc.  The Fock kernel is
        # charged a full batch for every launch that did work (initial F(P0) + one per SCF iteration; the trailing
        # look-behind iteration of the pipelined loop launches with nothing active and is not counted).
        jac_flops = float((10.0 * n**3 + 2.0 * n**2 * nocc).mean()) * jstats["molecules"]
        fock_bytes = 1184.0 * plan.npairs + 384.0 * plan.nat
        j_ms, j_n = prof.get("jacobi_density", (0.0, 0))
        f_ms, f_n = prof.get("fock", (0.0, 0))
        f_n = min(f_n, n_iter + 1)
        roofline = {
            "kernel": "jacobi_fixed_kernel<NP> (12 size classes)", "bound": "tensor",
            "bound_note": "compute side of the roofline: FP64. tcgen05 has no FP64 kind; the kernel's dense products run on the "
            "FP64 tensor cores (mma.sync DMMA), its rotation sweeps on the FP64 FMA pipe; B200's DMMA and DFMA peaks coincide "
            "(37.2 / 36.5 TFLOP/s measured)", "achieved": (jac_flops / (j_ms * 1e-3) / 1e12) if j_ms else None,
            "peak": fp64_peak, "unit": "TFLOP/s", "peak_source": "measured in this run: seqm_fp64_peak_tflops() "
            "DFMA probe (MEASURED_PEAKS.json has no fp64 entry)",
            "traffic": 77.4e6, "traffic_note": "DRAM bytes read + written per full-batch eigensolver call from the ncu --set full "
            "capture in profiles/jacobi_r01_final.txt (38.7 MB per half-batch call over the five populated size classes; "
            "writes stay in the 126 MB L2); the HBM floor 16 n^2 B per molecule is 74.9 MB",
            "algorithmic": "10 n^3 + 2 n^2 nocc flop (batch mean) x molecules solved in the step (library counter)",
            "molecules_solved": jstats["molecules"], "sweeps_per_solve": round(jstats["sweeps"] / max(jstats["molecules"], 1), 3),
            "share_of_step": round(j_ms / tot, 4), "dominant_kernel_by_time": top,
        }  # fmt: skip
        if roofline["achieved"] and fp64_peak > 0:
            roofline["frac"] = roofline["achieved"] / fp64_peak
        roofline["fock_kernel_hbm"] = {
            "bound": "hbm", "achieved": (fock_bytes * f_n / (f_ms * 1e-3) / 1e9) if f_ms else None, "peak": hbm_peak,
            "unit": "GB/s", "peak_source": hbm_src, "algorithmic": "1184 B/pair + 384 B/atom per launch",
            "share_of_step": round(f_ms / tot, 4),
        }  # fmt: skip
        if roofline["fock_kernel_hbm"]["achieved"]:
            roofline["fock_kernel_hbm"]["frac"] = roofline["fock_kernel_hbm"]["achieved"] / hbm_peak

    # ---- second headline metric: XL-BOMD MD steps/s (configs[2]: coronene replicas, AM1, k = 6, dt = 0.4 fs) ------
    xl = None
    if args.xl_replicas > 0:
        xl = xl_bomd_rate(seqm, dev, const, args.xl_replicas, args.xl_steps, world, dist if world > 1 else None)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or min(1024, 32 * cores)
        rate, dt = cpu_reference_rate(sample, cores)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {sample} molecules of the same batch in {dt:.1f} s, one process, numpy BLAS on "
                                  f"{cores} host threads, oracle/seqm_oracle (numpy restatement of the reference CPU path)"}  # fmt: skip

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(nmol, world, sha),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "includes": "pinned-host species+coordinates H2D, Molecule() (parser, parameter gather), forward, "
                                "D2H of Etot, Hf, force, notconverged"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "kernel_breakdown": breakdown, "scf_iterations": n_iter, "not_converged": nnot, "xl_bomd": xl,
        }  # fmt: skip
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
