"""Wall-clock phases of one forward (synchronised), to find host-side overheads."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
import bench
import pyseqm_b200 as seqm
from pyseqm_b200 import engine, basics
torch.set_default_dtype(torch.float64)
dev = torch.device("cuda:0")
species, coords, sha = bench.workload(4096, 0)
const = seqm.Constants().to(dev)
t0 = time.perf_counter()
mol = seqm.Molecule(const, dict(bench.SP), torch.as_tensor(coords, device=dev), torch.as_tensor(species, device=dev))
torch.cuda.synchronize(); print("Molecule() %.2f ms" % ((time.perf_counter() - t0) * 1e3))
mol.verbose = False
es = seqm.Electronic_Structure(dict(bench.SP))
for _ in range(2): es(mol)
torch.cuda.synchronize()
plan = mol._plan
def T(label, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print("%-28s %8.3f ms" % (label, (time.perf_counter() - t) * 1e3)); return r
for rep in range(2):
    print("--- rep", rep)
    T("full forward", lambda: es(mol))
    xyz = T("refresh_geometry", mol._refresh_geometry)
    w, hab = T("pair_integrals", lambda: engine.op_pair_integrals(plan, xyz))
    H = T("hcore", lambda: engine.op_hcore(plan, w, hab))
    P = T("initial_density", lambda: engine.op_initial_density(plan))
    F, E, nc, nit = T("scf", lambda: engine.op_scf(plan, H, w, P, 1e-7, [2], [False]))
    e, _, Cm = T("final eig", lambda: engine.op_eig_density(plan, F, want_P=False, want_C=True))
    V = T("orbitals_dense", lambda: engine.op_orbitals_dense(plan, Cm))
    T("nuclear", lambda: engine.op_nuclear_energy(plan, xyz, w))
    g = T("gradient", lambda: engine.op_gradient(plan, xyz, P))
    Pd = T("unpack", lambda: engine.op_unpack(plan, P))
    T("dipole", lambda: basics._ground_dipole(mol, Pd))
    T("q", lambda: const.tore[mol.species] - es.atomic_charges(Pd))
