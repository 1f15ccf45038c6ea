"""Fixtures for the MO crossing matcher (Energy._crossing_match_molecular_orbitals[_grouped], basics.py:596-719).
Run in the build container only (imports the unmodified reference from /root/reference).

op_momatch_{mixed,uniform}.npz: crafted old/new orbital sets (random orthogonal bases, block-internal permutations,
random signs, small and large mixing so that both the row-argmax route and the greedy repair are exercised) and the
reference's matched orbitals / eigenvalues.  md_momatch_two_forwards.npz: two consecutive reference forwards on the
same Molecule with displaced coordinates (the matcher runs in the second one)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from refrun import Constants, Electronic_Structure, Molecule, np, read_xyz, torch  # noqa: E402
from seqm.basics import Energy  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
XYZ = os.path.join(OUT, "xyz")
rng = np.random.default_rng(11)


def crafted(norb, nocc, nmax, sigma):
    q, _ = np.linalg.qr(rng.standard_normal((norb, norb)))
    perm = np.concatenate([rng.permutation(nocc), nocc + rng.permutation(norb - nocc)])
    sign = rng.choice([-1.0, 1.0], norb)
    new = q[:, perm] * sign
    # mixing inside each block keeps the occupied / virtual spaces (what an MD step does to first order)
    for lo, hi in ((0, nocc), (nocc, norb)):
        r = hi - lo
        if r > 1:
            g, _ = np.linalg.qr(np.eye(r) + sigma * rng.standard_normal((r, r)))
            new[:, lo:hi] = new[:, lo:hi] @ g
    old = np.eye(nmax)
    old[:norb, :norb] = q
    nw = np.eye(nmax)
    nw[:norb, :norb] = new
    e = np.zeros(nmax)
    e[:norb] = np.sort(rng.uniform(-40.0, 5.0, norb))
    return old, nw, e


def not_bijective(old, new, lo, hi):
    s = np.abs(old[:, lo:hi].T @ new[:, lo:hi])
    return len(set(np.argmax(s, axis=1).tolist())) != hi - lo


def make(name, species, sigmas):
    species = np.asarray(species)
    nheavy = (species > 1).sum(1)
    nhyd = (species == 1).sum(1)
    norb = 4 * nheavy + nhyd
    tore = {1: 1, 6: 4, 7: 5, 8: 6}
    nocc = np.array([sum(tore[z] for z in row if z) // 2 for row in species.tolist()])
    nmax = int(norb.max())
    olds, news, es, greedy = [], [], [], 0
    for m in range(species.shape[0]):
        o, n, e = crafted(int(norb[m]), int(nocc[m]), nmax, sigmas[m % len(sigmas)])
        greedy += not_bijective(o[: norb[m], : norb[m]], n[: norb[m], : norb[m]], 0, nocc[m])
        greedy += not_bijective(o[: norb[m], : norb[m]], n[: norb[m], : norb[m]], nocc[m], norb[m])
        olds.append(o), news.append(n), es.append(e)
    V_old, V_new, e = (torch.as_tensor(np.stack(x)) for x in (olds, news, es))
    uniform = bool((species == species[0]).all())
    if uniform:
        V_out, e_out = Energy._crossing_match_molecular_orbitals(V_new, V_old, int(nocc[0]), e.clone())
    else:
        V_out, e_out = Energy._crossing_match_molecular_orbitals_grouped(V_new, V_old, torch.as_tensor(nocc), torch.as_tensor(norb), e)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), species=species, V_old=V_old.numpy(), V_new=V_new.numpy(), e=e.numpy(),
                        V_out=V_out.numpy(), e_out=e_out.numpy(), nocc=nocc, norb=norb)  # fmt: skip
    print(name, "molecules", species.shape[0], "blocks needing the greedy repair", greedy,
          "orbitals moved", int((e_out != e).sum()))


species, _ = read_xyz([os.path.join(XYZ, f) for f in ("methane.xyz", "benzene.xyz", "toluene.xyz")])
H2 = np.zeros((1, species.shape[1]), dtype=species.dtype)
H2[0, :2] = 1
mixed = np.concatenate([species, species, H2, species], axis=0)
make("op_momatch_mixed", mixed, [0.02, 0.6, 0.3, 1.0, 0.1])
make("op_momatch_uniform", np.repeat(species[1:2], 6, axis=0), [0.05, 0.8, 0.4])

# two consecutive forwards on one Molecule (unsymmetric molecules: no degenerate orbitals)
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("synthetic", os.path.join(OUT, "..", "..", "pyseqm_b200", "synthetic.py"))
synthetic = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synthetic)
sp_, co_ = synthetic.qm9_like_batch(6, seed=5)
sp = {"method": "PM3", "scf_eps": 1e-8, "scf_converger": [2], "sp2": [False]}
co_ = np.array(co_)
mol = Molecule(Constants(), sp, torch.as_tensor(co_.copy()), torch.as_tensor(sp_, dtype=torch.int64))
es = Electronic_Structure(sp)
es(mol)
V1, e1 = mol.molecular_orbitals.detach().clone(), mol.e_mo.detach().clone()
disp = 0.05 * rng.standard_normal(co_.shape) * (np.asarray(sp_) > 0)[..., None]
with torch.no_grad():
    mol.coordinates += torch.as_tensor(disp)
es(mol)
np.savez_compressed(os.path.join(OUT, "md_momatch_two_forwards.npz"), species=np.asarray(sp_), coordinates=np.asarray(co_),
                    displacement=disp, V1=V1.numpy(), e1=e1.numpy(), V2=mol.molecular_orbitals.detach().numpy(),
                    e2=mol.e_mo.detach().numpy(), e_gap2=mol.e_gap.detach().numpy(), Etot2=mol.Etot.detach().numpy())  # fmt: skip
print("two forwards: e_mo reordered in", int((np.diff(mol.e_mo.detach().numpy()[:, :8], axis=1) < 0).any(1).sum()), "molecules")
