"""Synthetic QM9-size CHNO molecule batches (BASELINE.json configs[1]; SURVEY 8(d) generator).

The reference ships no dataset, so the benchmark workload is generated: acyclic saturated CHNO
skeletons with 4..9 heavy atoms (C weight 0.7, N/O 0.15 each), ideal bond lengths, tetrahedral angles,
random dihedrals, valences filled with H, Gaussian jitter sigma = 0.02 A, at most 29 atoms, rows sorted by
descending Z and zero-padded -- i.e. exactly the (species, coordinates) layout `Molecule` takes.
Molecule i depends only on (seed, i), so any sub-batch or shard can be generated independently.
"""
import hashlib

import numpy as np

_VALENCE = {6: 4, 7: 3, 8: 2, 15: 3, 16: 2, 17: 1}
_NVAL_E = {1: 1, 6: 4, 7: 5, 8: 6, 15: 5, 16: 6, 17: 7}
_BOND = {(6, 6): 1.53, (6, 7): 1.47, (6, 8): 1.43, (7, 7): 1.45, (7, 8): 1.40, (8, 8): 1.48,
         (1, 6): 1.09, (1, 7): 1.01, (1, 8): 0.96,
         # third-row substituents of BASELINE configs[4] (PM6 with d orbitals on P / S / Cl)
         (6, 15): 1.84, (6, 16): 1.82, (6, 17): 1.77, (7, 15): 1.70, (7, 16): 1.68, (7, 17): 1.75, (8, 15): 1.63,
         (8, 16): 1.60, (8, 17): 1.70, (15, 15): 2.21, (15, 16): 2.10, (15, 17): 2.04, (16, 16): 2.05, (16, 17): 2.01,
         (17, 17): 1.99, (1, 15): 1.42, (1, 16): 1.34, (1, 17): 1.27}  # fmt: skip
_TET = np.array([[1.0, 1.0, 1.0], [1.0, -1.0, -1.0], [-1.0, 1.0, -1.0], [-1.0, -1.0, 1.0]]) / np.sqrt(3.0)


def _bond(a, b):
    return _BOND[(min(a, b), max(a, b))]


def _frame(rng, back=None):
    """Four tetrahedral unit vectors; if `back` is given the first one equals it (random dihedral)."""
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])  # fmt: skip
    d = _TET @ R.T
    if back is None:
        return d
    # rotate so that d[0] -> back (Rodrigues), keeping the random twist about it
    a, b = d[0], back / np.linalg.norm(back)
    v = np.cross(a, b)
    c = float(a @ b)
    if c < -1 + 1e-12:
        return -d
    vx = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    Rm = np.eye(3) + vx + vx @ vx / (1.0 + c)
    return d @ Rm.T


def _one_molecule(rng, max_atoms, hetero=None):
    for _attempt in range(200):
        nheavy = int(rng.integers(4, 10))
        Z = [int(rng.choice([6, 7, 8], p=[0.7, 0.15, 0.15])) for _ in range(nheavy)]
        Z[0] = 6
        if hetero:  # configs[4]: one or two heavy sites (never the root) carry an element of `hetero`
            nsub = 1 + int(rng.integers(2))
            for site in rng.choice(np.arange(1, nheavy), size=min(nsub, nheavy - 1), replace=False):
                Z[int(site)] = int(rng.choice(list(hetero)))
        pos = [np.zeros(3)]
        dirs = [list(_frame(rng))]  # unused bond directions per heavy atom
        nb = [0]
        bonds = []
        ok = True
        for k in range(1, nheavy):
            cand = [i for i in range(k) if nb[i] < _VALENCE[Z[i]] and dirs[i]]
            cand = [i for i in cand if nb[i] < _VALENCE[Z[i]] - 0]  # may use every valence for skeleton
            if not cand:
                ok = False
                break
            p = int(rng.choice(cand))
            d = dirs[p].pop(int(rng.integers(len(dirs[p]))))
            pos.append(pos[p] + d * _bond(Z[p], Z[k]))
            fr = _frame(rng, back=-d)
            dirs.append(list(fr[1:]))
            nb[p] += 1
            nb.append(1)
            bonds.append((p, k))
        if not ok:
            continue
        allZ, allpos, bonded = list(Z), list(pos), set(bonds)
        for i in range(nheavy):
            for _ in range(_VALENCE[Z[i]] - nb[i]):
                d = dirs[i].pop(0)
                allpos.append(pos[i] + d * _bond(1, Z[i]))
                allZ.append(1)
                bonded.add((i, len(allZ) - 1))
        n = len(allZ)
        if n > max_atoms or (sum(_NVAL_E[z] for z in allZ) % 2):
            continue
        X = np.asarray(allpos)
        dist = np.linalg.norm(X[:, None] - X[None], axis=-1)
        good = True
        for a in range(n):
            for b in range(a + 1, n):
                if (a, b) in bonded:
                    continue
                lim = 2.0 if (allZ[a] > 1 and allZ[b] > 1) else 1.6
                if allZ[a] > 10 or allZ[b] > 10:
                    lim += 0.3
                if dist[a, b] < lim:
                    good = False
                    break
            if not good:
                break
        if not good:
            continue
        X = X + rng.normal(scale=0.02, size=X.shape)
        order = np.argsort(-np.asarray(allZ), kind="stable")
        return np.asarray(allZ)[order], X[order]
    raise RuntimeError("synthetic generator failed to place a molecule")


def qm9_like_batch(nmol, seed=0, molsize=29, start=0, hetero=None):
    """Returns (species int64 (nmol, molsize), coordinates float64 (nmol, molsize, 3)).
    hetero: e.g. (15, 16, 17) substitutes P / S / Cl at one or two heavy sites of every molecule (BASELINE configs[4]);
    None (default) leaves the CHNO stream of configs[1] bit-for-bit unchanged."""
    species = np.zeros((nmol, molsize), dtype=np.int64)
    coords = np.zeros((nmol, molsize, 3), dtype=np.float64)
    for i in range(nmol):
        rng = np.random.default_rng([seed, start + i])
        z, x = _one_molecule(rng, molsize, hetero)
        species[i, : z.shape[0]] = z
        coords[i, : z.shape[0]] = x
    return species, coords


def batch_sha256(species, coords):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(species, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(coords, dtype=np.float64).tobytes())
    return h.hexdigest()
