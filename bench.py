#!/usr/bin/env python
"""bench.py -- molecule-SCF/s on BASELINE.json configs[1]: PM3, synthetic QM9-size CHNO batch of 4096 molecules per
GPU, SCF to 1e-7 with Pulay DIIS, energies + forces, fp64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--extras all|none|a,b,...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one Electronic_Structure.forward over the batch (pair integrals, Hcore, SCF, energies, forces).
Prints ONE JSON line (rank 0).  DESIGN.md section 5 documents every field.

`--impl reference` times the UNMODIFIED reference (oracle/_ref, installed by oracle/ref_runner.py at build()) on the
host cores: exactly K timed + W warm-up steps, each step one reference forward over a bounded sample (the first
REF_SAMPLE molecules of the same batch), and reports the steps / sample it actually ran.
"""
import argparse
import contextlib
import io
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
REF_SAMPLE = 128  # molecules per step of the reference arm (one reference forward ~2-3 s on 8-16 host threads)
PARITY_SAMPLE = 512  # molecules of the parity-at-size check / cpu_baseline sample (~10 s of reference CPU work)
ALL_EXTRAS = ("parity", "ref_gpu", "xl_bomd", "c380", "pm6", "strong")


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


def host_threads():
    import contextlib

    cores = os.cpu_count() or 1
    try:  # torchrun exports OMP_NUM_THREADS=1: ask the BLAS/OpenMP pools for every host thread explicitly
        from threadpoolctl import threadpool_limits

        return cores, threadpool_limits(limits=cores)
    except Exception:
        return cores, contextlib.nullcontext()


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


# ---------------------------------------------------------------------------------------------------------
# reference arm / CPU baselines
def reference_rate(species, coords, sp, steps, warmup, device="cpu"):
    """`steps` timed + `warmup` untimed forwards of the unmodified reference (oracle/_ref).  Returns (molecules/s,
    seconds per step, last result dict)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_runner

    cores, pool = host_threads()
    times, out = [], None
    with pool:
        for it in range(warmup + steps):
            out, dt = ref_runner.run_reference(species, coords, sp, device=device, threads=cores)
            if it >= warmup:
                times.append(dt)
    dt = sum(times) / len(times)
    return len(species) / dt, dt, out


def port_rate(species, coords, sp, steps, warmup):
    """Fallback when oracle/_ref is absent: the numpy restatement (oracle/seqm_oracle), kind "port"."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import seqm_oracle as so

    cores, pool = host_threads()
    times, out = [], None
    with pool:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            out = so.single_point(species, coords, sp)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return len(species) / dt, dt, out


def have_reference():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_runner

    return ref_runner.reference_available()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    species, coords, sha = workload(4096, 0)
    sample = REF_SAMPLE
    s, c = species[:sample], coords[:sample]
    if have_reference():
        rate, dt, out = reference_rate(s, c, SP, steps=args.steps, warmup=args.warmup)
        kind, what = "reference", "the unmodified reference (oracle/_ref: lanl/PYSEQM v2.0.0 installed by oracle/ref_runner.py), device='cpu'"
    else:
        rate, dt, out = port_rate(s, c, SP, steps=args.steps, warmup=args.warmup)
        kind, what = "port", "oracle/seqm_oracle (numpy restatement; oracle/_ref was not installed on this box)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(4096, args.gpus, sha),
        "molecules_per_step": sample, "scf_iterations": int(out["n_scf_iter"]),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"each of the {args.steps} timed (+{args.warmup} warm-up) steps is one Electronic_Structure.forward "
                                   f"of {what} over the first {sample} molecules of the rank-0 batch, torch intra-op threads = "
                                   f"{cores}; ms_per_step is the time of that {sample}-molecule step, value = {sample} / it"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }  # fmt: skip
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# extras: the other BASELINE configs and the honesty checks, each returning a dict for the JSON line
def parity_at_size(seqm, dev, const, species_h, coords_h, want_ref_gpu):
    """configs[1] parity at BASELINE scale: the first PARITY_SAMPLE molecules of the bench batch run as their OWN batch
    on the GPU (so the batch-global DIIS reset sees the same set) against the unmodified reference on the host cores;
    the reference timing is the cpu_baseline.  north_star tolerances: 1e-6 eV, 1e-8, 1e-5 eV/A, equal iterations."""
    import numpy as np
    import torch

    n = PARITY_SAMPLE
    s, c = species_h[:n], coords_h[:n]
    mol = seqm.Molecule(const, dict(SP), torch.as_tensor(c, device=dev), torch.as_tensor(s, device=dev))
    mol.verbose = False
    es = seqm.Electronic_Structure(dict(SP))
    es(mol)
    torch.cuda.synchronize()
    cores = os.cpu_count() or 1
    if have_reference():
        rate, dt, ref = reference_rate(s, c, SP, steps=1, warmup=0)
        kind, what = "reference", "unmodified reference (oracle/_ref), device='cpu'"
    else:
        rate, dt, ref = port_rate(s, c, SP, steps=1, warmup=0)
        kind, what = "port", "oracle/seqm_oracle numpy restatement (oracle/_ref absent)"
    dE = float(np.abs(mol.Etot.cpu().numpy() - ref["Etot"]).max())
    dH = float(np.abs(mol.Hf.cpu().numpy() - ref["Hf"]).max())
    dP = float(np.abs(mol.dm.cpu().numpy() - ref["dm"]).max())
    dF = float(np.abs(mol.force.cpu().numpy() - ref["force"]).max())
    ok_num = dE < 1e-6 and dH < 1e-6 and dP < 1e-8 and dF < 1e-5
    par = {"molecules": n, "against": what, "max_abs_dEtot_eV": dE, "max_abs_dHf_eV": dH, "max_abs_dP": dP,
           "max_abs_dForce_eV_per_A": dF, "n_scf_iter": int(mol.n_scf_iter), "n_scf_iter_reference": int(ref["n_scf_iter"]),
           "not_converged": int(es.notconverged.sum()), "not_converged_reference": int(np.asarray(ref["notconverged"]).sum()),
           "tolerances": {"Etot_eV": 1e-6, "dm": 1e-8, "force_eV_per_A": 1e-5, "n_scf_iter": "equal"},
           "pass": bool(ok_num and int(mol.n_scf_iter) == int(ref["n_scf_iter"]))}  # fmt: skip
    cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
           "sample": f"one forward of the {what} over the first {n} molecules of the same batch: {dt:.1f} s on {cores} host threads"}  # fmt: skip
    ref_gpu = None
    if want_ref_gpu and have_reference():
        try:  # reference-on-the-same-B200: PyTorch eager + cuSOLVER/cuBLAS path of the reference itself
            r1, d1, o1 = reference_rate(s, c, SP, steps=2, warmup=1, device=str(dev))
            ref_gpu = {"value": r1, "unit": UNIT, "molecules": n, "seconds_per_forward": d1, "n_scf_iter": int(o1["n_scf_iter"]),
                       "what": "unmodified reference (oracle/_ref) with device='cuda' on this B200, same 512-molecule batch, mean of 2 after 1 warm-up"}  # fmt: skip
        except Exception as e:  # the reference's own CUDA path is not ours to fix: report, do not fail the bench
            ref_gpu = {"error": f"{type(e).__name__}: {e}"[:300]}
    if not ok_num:
        print("PARITY FAILURE at BASELINE size:", json.dumps(par), file=sys.stderr)
    return par, cpu, ref_gpu, ok_num


def xl_bomd_rate(seqm, dev, const, nrep, nsteps, world, dist, sp2, ksa=None):
    """replica-steps/s of XL-BOMD NVE on `nrep` coronene replicas per GPU (configs[2]; t = 0 SCF excluded).
    ksa: xl_bomd_params of a KSA_XL_BOMD run (rank-m Krylov kernel at electronic temperature T_el) instead."""
    import torch

    xyz = os.path.join(ROOT, "tests", "golden", "xyz", "coronene.xyz")
    s, c = seqm.read_xyz([xyz] * nrep)
    sp = {"method": "AM1", "scf_eps": 1.0e-7, "scf_converger": [2], "sp2": sp2}
    torch.manual_seed(1234 + int(os.environ.get("RANK", "0")))
    mol = seqm.Molecule(const, sp, torch.as_tensor(c, device=dev), torch.as_tensor(s, device=dev))
    if ksa:
        md = seqm.KSA_XL_BOMD(xl_bomd_params=dict(ksa), seqm_parameters=sp, timestep=0.4, Temp=300.0)
    else:
        md = seqm.XL_BOMD(xl_bomd_params={"k": 6}, seqm_parameters=sp, timestep=0.4, Temp=300.0)
    md.initialize(mol)
    for i in range(3):
        md._do_integrator_step(i, mol, dict())
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Es = []
    e0.record()
    for i in range(3, 3 + nsteps):
        md._do_integrator_step(i, mol, dict())
        if i % 50 == 0:
            Es.append(md._thermo_potential(mol) + md._kinetic_energy(mol))
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    Ek = md._kinetic_energy(mol)
    drift = float((torch.stack(Es) - Es[0]).abs().max()) if len(Es) > 1 else None
    return {"metric": "XL-BOMD MD steps/s (replica-steps/s)", "value": nrep * world * nsteps / float(t), "unit": "replica-steps/s",
            "ms_per_md_step": float(t) / nsteps * 1e3, "replicas_per_gpu": nrep, "steps": nsteps, "molecule": "coronene C24H12 (108 orbitals)",
            "method": "AM1, k=6, dt=0.4 fs, 300 K, density by " + (
                f"Fermi occupations at T_el = {ksa['T_el']} K, rank-{ksa['max_rank']} Krylov kernel (KSA_XL_BOMD, xlbomd.py:201-341)" if ksa
                else "in-SM SP2 purification on the FP64 tensor cores, eps 1e-5" if sp2[0]
                else "the Jacobi eigensolver (reference branch xlbomd.py:361)"),
            "max_abs_total_energy_change_eV": drift,
            "finite": bool(torch.isfinite(mol.Etot).all() and torch.isfinite(Ek).all())}  # fmt: skip


def c380_run(seqm, lib, dev, const):
    """configs[3]: C380 fullerene (1520 orbitals), AM1, SCF 1e-6 (DIIS), SP2 1e-5, energies + forces on one GPU; the
    FP64 GEMM rate of the SP2 products next to cuBLAS DGEMM (torch.matmul) of the same shape measured in the same run."""
    import torch

    xyz = os.path.join(ROOT, "tests", "golden", "xyz", "C380.xyz")
    s, c = seqm.read_xyz([xyz])
    sp = {"method": "AM1", "scf_eps": 1.0e-6, "scf_converger": [2], "sp2": [True, 1.0e-5]}
    mol = seqm.Molecule(const, dict(sp), torch.as_tensor(c, device=dev), torch.as_tensor(s, device=dev))
    mol.verbose = False
    es = seqm.Electronic_Structure(dict(sp))
    es(mol)  # warm-up
    torch.cuda.synchronize()
    walls = []
    for _ in range(3):
        t0 = time.perf_counter()
        es(mol)
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
    lib.profile_enable(True)
    es(mol)
    prof = lib.profile_collect()
    lib.profile_enable(False)
    n = int(mol._plan.nmax)
    g_ms, g_n = prof.get("dgemm", (0.0, 0))
    from pyseqm_b200._lib import ptr

    def ours(sym, reps=50):
        A = torch.randn(n, n, dtype=torch.float64, device=dev)
        A = ((A + A.T) * 0.5).contiguous()
        B = torch.empty_like(A)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            lib.check(lib.dll.seqm_square_product(n, ptr(A), None if sym else ptr(A), ptr(B), st), "seqm_square_product")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            lib.dll.seqm_square_product(n, ptr(A), None if sym else ptr(A), ptr(B), st)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms_gen, ms_sym = ours(False), ours(True)
    nb = (n + 95) // 96
    flop_sym = 2.0 * 96 * 96 * n * (nb * (nb + 1) // 2)

    def cublas(nn, reps):
        A = torch.randn(nn, nn, dtype=torch.float64, device=dev)
        B = torch.empty_like(A)
        for _ in range(3):
            torch.matmul(A, A, out=B)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(A, A, out=B)
        e1.record()
        torch.cuda.synchronize()
        return 2.0 * nn**3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12

    return {"workload": "configs[3]: C380 fullerene, 1520 orbitals, AM1, scf_eps 1e-6, scf_converger [2], sp2 [True, 1e-5], energies + forces",
            "wall_s_best_of_3": min(walls), "wall_s_all": [round(w, 4) for w in walls], "n_scf_iter": int(mol.n_scf_iter),
            "not_converged": int(es.notconverged.sum()), "Etot_eV": float(mol.Etot[0]),
            "dgemm": {"sp2_loop_launches_incl_voided": int(g_n), "sp2_loop_ms": g_ms,
                      "sym_x2_ms": ms_sym, "sym_x2_tflops_full_product_equivalent": 2.0 * n**3 / (ms_sym * 1e-3) / 1e12,
                      "sym_x2_tflops_executed": flop_sym / (ms_sym * 1e-3) / 1e12,
                      "general_ms": ms_gen, "general_tflops": 2.0 * n**3 / (ms_gen * 1e-3) / 1e12,
                      "cublas_dgemm_tflops_same_shape": cublas(n, 50), "cublas_dgemm_tflops_8192": cublas(8192, 3),
                      "note": "X^2 of the SP2 loop: dgemm_sym_kernel (mma.sync DMMA, upper-triangle 96x96 tiles + mirror, "
                              "0.53 of the full product's FLOPs); general products: dgemm_dmma_kernel; both timed alone "
                              "through seqm_square_product next to torch.matmul fp64 (cuBLAS) in this run"},
            "kernel_ms": {k: round(v[0], 3) for k, v in prof.items() if v[1]},
            "reference_cpu": "84-131 s wall on 8 host threads (BASELINE.md section 2; not re-run here: one forward exceeds the bench budget)"}  # fmt: skip


def pm6_run(seqm, lib, dev, const, nmol, steps):
    """configs[4]: PM6 sp+d batch of 2048 organics with S/P/Cl, scf_converger [1], energies + forces."""
    import numpy as np
    import torch

    from pyseqm_b200.synthetic import qm9_like_batch

    # "accept only molecules the reference converges".  Adaptive mixing (scf_converger [1], the configs[4]
    # setting) stalls on a few percent of the synthetic P/S/Cl organics, so candidates are screened once, untimed, with
    # a 150-iteration cap (iteration counts equal the reference's, see parity_64) and the first `nmol` survivors kept.
    import warnings

    ncand = int(nmol * 1.25) + 32
    cs, cc = qm9_like_batch(ncand, seed=0, start=0, hetero=(15, 16, 17))
    sp = {"method": "PM6", "scf_eps": 1.0e-7, "scf_converger": [1], "sp2": [False]}
    sp_screen = dict(sp, b200_scf_max_iter=150)
    ms_ = seqm.Molecule(const, dict(sp_screen), torch.as_tensor(cc, device=dev), torch.as_tensor(cs, device=dev))
    ms_.verbose = False
    es_ = seqm.Electronic_Structure(dict(sp_screen))
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        es_(ms_)
    keep = (~es_.notconverged).nonzero(as_tuple=False).squeeze(1).cpu().numpy()
    dropped = ncand - keep.size
    if keep.size < nmol:
        raise RuntimeError(f"only {keep.size} of {ncand} PM6 candidates converge")
    species, coords = cs[keep[:nmol]], cc[keep[:nmol]]
    del ms_, es_
    mol = seqm.Molecule(const, dict(sp), torch.as_tensor(coords, device=dev), torch.as_tensor(species, device=dev))
    mol.verbose = False
    es = seqm.Electronic_Structure(dict(sp))
    for _ in range(2):
        es(mol)
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        es(mol)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = sorted(ms)[len(ms) // 2]  # median of the timed forwards (all of them are in ms_all)
    lib.profile_enable(True)
    es(mol)
    prof = lib.profile_collect()
    lib.profile_enable(False)
    out = {"workload": f"configs[4]: PM6 (d orbitals on P/S/Cl), {nmol} synthetic organics (<=29 atoms), scf_eps 1e-7, scf_converger [1], "
                       f"energies + forces; the first {nmol} of {ncand} generated molecules that adaptive mixing converges within 150 iterations "
                       f"({dropped} candidates dropped)",
           "value": nmol / (t * 1e-3), "unit": UNIT, "ms_per_step": t, "ms_all": [round(x, 2) for x in ms], "steps": steps,
           "n_scf_iter": int(mol.n_scf_iter),
           "not_converged": int(es.notconverged.sum()), "molecules_with_d_shell": int((np.isin(species, (15, 16, 17))).any(axis=1).sum()),
           "orbitals_max": int(mol._plan.nmax), "pairs_with_d_atom": int(mol._plan.n_ypairs), "pairs": int(mol._plan.npairs),
           "d_integral_doubles": int(mol._plan.wd_total),
           "kernel_ms": {k: round(v[0], 3) for k, v in prof.items() if v[1]}}  # fmt: skip
    if have_reference():
        n = 64
        rate, dt, ref = reference_rate(species[:n], coords[:n], sp, steps=1, warmup=0)
        m2 = seqm.Molecule(const, dict(sp), torch.as_tensor(coords[:n], device=dev), torch.as_tensor(species[:n], device=dev))
        m2.verbose = False
        es(m2)
        out["reference_cpu"] = {"value": rate, "unit": UNIT, "molecules": n, "seconds": dt, "cores": os.cpu_count(), "kind": "reference"}
        out["parity_64"] = {"max_abs_dEtot_eV": float(np.abs(m2.Etot.cpu().numpy() - ref["Etot"]).max()),
                            "max_abs_dForce_eV_per_A": float(np.abs(m2.force.cpu().numpy() - ref["force"]).max()),
                            "n_scf_iter": int(m2.n_scf_iter), "n_scf_iter_reference": int(ref["n_scf_iter"])}  # fmt: skip
    return out


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nmol", type=int, default=4096)
    ap.add_argument("--extras", default="all", help="all | none | comma list of " + ",".join(ALL_EXTRAS))
    ap.add_argument("--xl-replicas", type=int, default=1024, help="coronene replicas for the XL-BOMD line")
    ap.add_argument("--xl-steps", type=int, default=0, help="0 = 1000 steps (configs[2]) on one GPU, 100 under torchrun")
    ap.add_argument("--pm6-nmol", type=int, default=2048)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    extras = set(ALL_EXTRAS) if args.extras == "all" else set(x for x in args.extras.split(",") if x and x != "none")

    import torch
    import torch.distributed as dist

    import pyseqm_b200 as seqm
    from pyseqm_b200._lib import get_lib
    from pyseqm_b200.sharding import ShardedBatch, gather_results

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
            return gather_results(dict(Etot=mol.Etot, Hf=mol.Hf, force=mol.force), gidx, nmol * world, nmax=nmol)
        return None

    def timed(fn, steps):
        ms = []
        for _ in range(steps):
            flush.fill_(1.0)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            barrier()
            ms.append(e0.elapsed_time(e1))
        t_local = torch.tensor([sum(ms) / len(ms)], device=dev)
        if world > 1:
            dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
        return float(t_local), ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    warm = max(3, args.warmup)
    for _ in range(warm):
        step_resident()
    barrier()
    if rank == 0:
        sampler.mark()
    launches0 = lib.dll.seqm_launch_count()
    ms_step, ms = timed(step_resident, args.steps)
    launches = lib.dll.seqm_launch_count() - launches0
    print("resident step times (ms):", [round(x, 2) for x in ms], file=sys.stderr)
    clocks = sampler.stop() if rank == 0 else None
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
        profs = []
        for _ in range(3):  # three profile steps, per-kernel median: one step's event timing wanders by +-10 % between boxes
            lib.jacobi_stats(reset=True)
            lib.profile_enable(True)
            flush.fill_(1.0)
            torch.cuda.synchronize()
            es(mol)
            profs.append(lib.profile_collect())
            lib.profile_enable(False)
            jstats = lib.jacobi_stats(reset=True)
        prof = {k: (sorted(p_.get(k, (0.0, 0))[0] for p_ in profs)[1], profs[0][k][1]) for k in profs[0]}
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
        # algorithmic work: the library counts the molecules the eigensolver actually solved
        # (converged molecules drop out); each is charged the batch-mean 10 n^3 + 2 n^2 nocc.
        solved = jstats["molecules"]
        jac_flops = float((10.0 * n**3 + 2.0 * n**2 * nocc).mean()) * solved
        # Fock: launched on the same shrinking active set as the eigensolver (+ the initial full-batch F(P0)); charged
        # the batch-mean 1184 B/pair + 384 B/atom per ACTIVE molecule (round 1 charged every launch a full batch).
        fock_bytes_full = 1184.0 * plan.npairs + 384.0 * plan.nat
        fock_active_molecules = plan.nmol + solved
        fock_bytes = fock_bytes_full * fock_active_molecules / plan.nmol
        j_ms, j_n = prof.get("jacobi_density", (0.0, 0))
        f_ms, f_n = prof.get("fock", (0.0, 0))
        roofline = {
            "kernel": "jacobi_fixed_kernel<NP> (12 size classes)", "bound": "fp64 issue / shared-memory latency",
            "bound_enum": "tensor",
            "bound_note": "compute side of the roofline, FP64: tcgen05 has no FP64 kind, so the peak is the FP64 pipe (DFMA = DMMA "
            "rate on B200, 36.5 / 37.2 TFLOP/s measured).  ncu (profiles/) shows the kernel limited by dependent-chain latency and "
            "shared-memory wavefronts, not by the tensor pipe (DMMA sub-pipe < 3 %); instruction budget of a step: 4.9 G warp "
            "instructions, 0.9 G of them FP64 = 45 % of the issue slots and 17 % of the FP64 pipe in the time the solver takes "
            "(profiles/jacobi_inst_r02.txt)",
            "achieved": (jac_flops / (j_ms * 1e-3) / 1e12) if j_ms else None,
            "peak": fp64_peak, "unit": "TFLOP/s", "peak_source": "measured in this run: seqm_fp64_peak_tflops() "
            "DFMA probe (MEASURED_PEAKS.json has no fp64 entry)",
            "traffic": 77.4e6, "traffic_note": "DRAM bytes read + written per full-batch eigensolver call from the ncu --set full "
            "captures in profiles/jacobi_r01_final.txt and profiles/hot_r02_final.txt (writes stay in the 126 MB L2); the HBM floor "
            "16 n^2 B per molecule is 74.9 MB",
            "algorithmic": "10 n^3 + 2 n^2 nocc flop (batch mean) x molecules solved in the step (library counter)",
            "molecules_solved": solved, "sweeps_per_solve": round(jstats["sweeps"] / max(solved, 1), 3),
            "share_of_step": round(j_ms / tot, 4), "dominant_kernel_by_time": top,
        }  # fmt: skip
        if roofline["achieved"] and fp64_peak > 0:
            roofline["frac"] = roofline["achieved"] / fp64_peak
        roofline["fock_kernel_hbm"] = {
            "bound": "hbm", "achieved": (fock_bytes / (f_ms * 1e-3) / 1e9) if f_ms else None, "peak": hbm_peak,
            "unit": "GB/s", "peak_source": hbm_src,
            "algorithmic": "(1184 B/pair + 384 B/atom, batch mean per molecule) x active molecules summed over the launches",
            "active_molecule_launches": int(fock_active_molecules), "launches": int(f_n),
            "traffic": 212.9e6, "traffic_note": "ncu dram__bytes_read+write per half-batch launch (profiles/hot_r02_final.txt: 198.3 MB "
            "read + 14.5 MB written in 84 us, cold): below the algorithmic bytes because H-H / X-H pairs touch 8 / 80 B of "
            "their 800 B w block",
            "share_of_step": round(f_ms / tot, 4),
        }  # fmt: skip
        if roofline["fock_kernel_hbm"]["achieved"]:
            roofline["fock_kernel_hbm"]["frac"] = roofline["fock_kernel_hbm"]["achieved"] / hbm_peak

    # ---- strong scaling: ONE global 4096-molecule batch through sharding.ShardedBatch (the shipped sharder) ------
    strong = None
    if world > 1 and "strong" in extras:
        gs, gc, gsha = workload(nmol, 0)  # every rank generates the same global batch
        sb = ShardedBatch(torch.as_tensor(gs, device=dev), torch.as_tensor(gc, device=dev), dict(SP), const,
                          lambda cst, p, c, s: _quiet(seqm.Molecule(cst, p, c, s)), seqm.Electronic_Structure)  # fmt: skip
        for _ in range(3):
            sb.forward()
        t_s, ms_s = timed(sb.forward, min(args.steps, 10))
        it = torch.tensor([sb.molecule.n_scf_iter], device=dev)
        its = [torch.zeros_like(it) for _ in range(world)]
        dist.all_gather(its, it)
        strong = {"scaling": "strong", "global_batch": nmol, "molecules_per_gpu": nmol // world, "value": nmol / (t_s * 1e-3),
                  "unit": UNIT, "ms_per_step": t_s, "steps": len(ms_s), "n_scf_iter_per_shard": [int(x) for x in its],
                  "path": "sharding.ShardedBatch: size-sorted round-robin deal, resident shards, one packed all_gather per step"}  # fmt: skip

    # ---- configs[2]: XL-BOMD MD steps/s (coronene replicas, AM1, k = 6, dt = 0.4 fs) --------------------------------
    xl = None
    if "xl_bomd" in extras and args.xl_replicas > 0:
        nsteps = args.xl_steps or (1000 if world == 1 else 100)
        d = dist if world > 1 else None
        xl = xl_bomd_rate(seqm, dev, const, args.xl_replicas, nsteps, world, d, [True, 1.0e-5])
        eig = xl_bomd_rate(seqm, dev, const, args.xl_replicas, min(nsteps, 100), world, d, [False])
        xl["eigensolver_branch"] = {k: eig[k] for k in ("value", "unit", "ms_per_md_step", "steps", "method", "finite")}
        xl["reference_cpu"] = "21.5 replica-steps/s (16 replicas x 10 steps, 8 host threads; BASELINE.md section 2)"
        try:  # KSA-XL-BOMD on the same replicas (a quarter of them: every step is 3 response solves)
            kp = {"k": 6, "max_rank": 3, "err_threshold": 0.0, "T_el": 1500}
            ks = xl_bomd_rate(seqm, dev, const, max(args.xl_replicas // 4, 1), min(nsteps, 50), world, d, [False], ksa=kp)
            xl["ksa_branch"] = {k: ks[k] for k in ("value", "unit", "ms_per_md_step", "replicas_per_gpu", "steps", "method",
                                                   "max_abs_total_energy_change_eV", "finite")}  # fmt: skip
        except Exception as e:
            xl["ksa_branch"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    parity = cpu_baseline = ref_gpu = c380 = pm6 = None
    parity_ok = True
    if rank == 0 and world == 1:
        if "parity" in extras:
            parity, cpu_baseline, ref_gpu, parity_ok = parity_at_size(seqm, dev, const, species_h, coords_h, "ref_gpu" in extras)
        if "c380" in extras:
            c380 = c380_run(seqm, lib, dev, const)
        if "pm6" in extras:
            try:
                pm6 = pm6_run(seqm, lib, dev, const, args.pm6_nmol, 5)
            except Exception as e:  # configs[4] must never take the configs[1] line down with it
                pm6 = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(nmol, world, sha),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "includes": "pinned-host species+coordinates H2D, Molecule() (parser, parameter gather), forward, "
                                "D2H of Etot, Hf, force, notconverged"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "kernel_breakdown": breakdown, "scf_iterations": n_iter, "not_converged": nnot,
            "parity_at_size": parity, "reference_on_gpu": ref_gpu, "strong_scaling": strong, "xl_bomd": xl,
            "c380": c380, "pm6_d": pm6,
        }  # fmt: skip
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not parity_ok:
        sys.exit(3)


def _quiet(m):
    m.verbose = False
    return m


if __name__ == "__main__":
    main()
