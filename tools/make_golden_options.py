"""Fixtures for the option matrix of SURVEY 8(b): charged molecules, learned per-atom parameters, Hf_flag / eig off.
Run in the build container only (imports the unmodified reference from /root/reference)."""
import contextlib
import io
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from refrun import Constants, Electronic_Structure, Molecule, np, read_xyz, torch  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
XYZ = os.path.join(OUT, "xyz")


def run(species, coords, sp, charges=0, learned=None):
    species = torch.as_tensor(species, dtype=torch.int64)
    coords = torch.as_tensor(coords, dtype=torch.float64)
    sp = dict(sp)
    kw = {}
    if learned is not None:
        kw["learned_parameters"] = {k: torch.as_tensor(v) for k, v in learned.items()}
    mol = Molecule(Constants(), sp, coords, species, charges=charges, **kw)
    es = Electronic_Structure(sp)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        es(mol, **({"learned_parameters": kw["learned_parameters"]} if learned is not None else {}))
    m = re.findall(r"scf [a-z ]+:\s+(\d+) \|", buf.getvalue())
    out = dict(Etot=mol.Etot, Hf=mol.Hf, Eelec=mol.Eelec, Enuc=mol.Enuc, Eiso=mol.Eiso, dm=mol.dm, q=mol.q, force=mol.force,
               notconverged=es.notconverged)  # fmt: skip
    if mol.e_mo is not None:
        out.update(e_mo=mol.e_mo, e_gap=mol.e_gap)
    out = {k: v.detach().numpy() for k, v in out.items()}
    out["n_scf_iter"] = int(m[-1]) if m else -1
    return out


def save(name, out, species, coords, sp, **extra):
    sp = {k: v for k, v in sp.items() if k != "elements"}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), species=species, coordinates=coords,
                        seqm_parameters=json.dumps(sp), **out, **extra)  # fmt: skip
    print(name, "iters", out["n_scf_iter"], "notconv", int(out["notconverged"].sum()), "Etot", out["Etot"])


# 1. ions: NH4+, OH-, H3O+, CH4
species = np.array([[7, 1, 1, 1, 1], [8, 1, 0, 0, 0], [8, 1, 1, 1, 0], [6, 1, 1, 1, 1]])
t = 1.03 / np.sqrt(3.0)
c = 1.09 / np.sqrt(3.0)
coords = np.zeros((4, 5, 3))
coords[0, 1:] = [[t, t, t], [t, -t, -t], [-t, t, -t], [-t, -t, t]]
coords[1, 1] = [0.0, 0.0, 0.97]
coords[2, 1:4] = [[0.93, 0.0, 0.28], [-0.465, 0.805, 0.28], [-0.465, -0.805, 0.28]]
coords[3, 1:] = [[c, c, c], [c, -c, -c], [-c, c, -c], [-c, -c, c]]
charges = np.array([1, -1, 1, 0])
sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [2], "sp2": [False]}
save("opt_charged_AM1", run(species, coords, sp, charges=torch.as_tensor(charges)), species, coords, sp, charges=charges)

# 2. learned per-atom parameters (no gradients): +-2 % around the PM3 tables
species, coords = read_xyz([os.path.join(XYZ, f) for f in ("methane.xyz", "benzene.xyz", "toluene.xyz")])
sp = {"method": "PM3", "scf_eps": 1e-6, "scf_converger": [1], "sp2": [False], "learned": ["U_ss", "zeta_s", "beta_p", "g_ss"]}
base = Molecule(Constants(), {k: v for k, v in sp.items() if k != "learned"}, torch.as_tensor(coords), torch.as_tensor(species))
rng = np.random.default_rng(3)
learned = {k: base.parameters[k].detach().numpy() * (1.0 + 0.02 * rng.uniform(-1, 1, base.parameters[k].shape[0])) for k in sp["learned"]}
save("opt_learned_PM3", run(species, coords, sp, learned=learned), species, coords, sp, **{"learned_" + k: v for k, v in learned.items()})

# 3. Hf_flag / eig off, constant mixing
sp = {"method": "MNDO", "scf_eps": 1e-6, "scf_converger": [0, 0.3], "sp2": [False], "Hf_flag": False, "eig": False}
save("opt_flags_MNDO", run(species, coords, sp), species, coords, sp)

# 4. pair_outer_cutoff that removes pairs (Parser, basics.py:209, 326): 3.5 Angstrom drops every para C...C, most
#    C...H and H...H pairs of benzene / toluene; methane keeps all of its pairs
sp = {"method": "AM1", "scf_eps": 1e-7, "scf_converger": [2], "sp2": [False], "pair_outer_cutoff": 3.5}
save("opt_cutoff_AM1", run(species, coords, sp), species, coords, sp)
