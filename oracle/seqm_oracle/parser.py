"""Pair list and index maps.  Restates seqm/basics.py:219-403 (Parser.forward)."""
import numpy as np

from .tables import Tables


class Parsed:
    pass


def parse(species, coordinates, charges=0, outer_cutoff=1.0e10):
    """species (nmol, molsize) int, rows sorted by descending Z, 0 padded; coordinates in Angstrom.

    Returns an object with the tensors of basics.py:345-403: Z, maskd, mask, atom_molid, pair_molid,
    ni, nj, idxi, idxj, xij (unit vector i->j), rij (bohr, a0 = 0.529167), nocc, nHeavy, nHydro.
    """
    T = Tables.get()
    species = np.asarray(species, dtype=np.int64)
    coordinates = np.asarray(coordinates, dtype=np.float64)
    nmol, molsize = species.shape
    # Molecule.py:188-206 check_input
    if not np.all(species[:, :-1] >= species[:, 1:]):
        bad = np.nonzero(~np.all(species[:, :-1] >= species[:, 1:], axis=1))[0].tolist()
        rows = ", ".join(map(str, bad))
        raise ValueError(
            f"species must be non-increasing along each row, but {'row' if len(bad) == 1 else 'rows'} {rows} "
            f"{'is' if len(bad) == 1 else 'are'} not sorted."
        )
    p = Parsed()
    p.nmol, p.molsize = nmol, molsize
    p.species, p.coordinates = species, coordinates
    real = species.reshape(-1) > 0
    p.real_atoms = np.nonzero(real)[0]
    p.Z = species.reshape(-1)[p.real_atoms]
    p.nHeavy = np.sum(species > 1, axis=1)
    p.nHydro = np.sum(species == 1, axis=1)
    n_el = np.sum(T.tore[species], axis=1).astype(np.int64) - (np.zeros(nmol, dtype=np.int64) + charges)
    if np.any(n_el % 2 == 1):
        raise ValueError("RHF setting requires closed shell systems (even number of electrons)")
    p.nocc = n_el // 2
    p.norb = 4 * p.nHeavy + p.nHydro
    mol_of = np.repeat(np.arange(nmol), molsize)
    pos_of = np.tile(np.arange(molsize), nmol)
    p.atom_molid = mol_of[p.real_atoms]
    p.atom_pos = pos_of[p.real_atoms]
    p.maskd = p.atom_molid * molsize * molsize + p.atom_pos * (molsize + 1)
    # all i<j real pairs inside each molecule, ordered by (molecule, i, j)  (basics.py:306-343)
    natoms = np.sum(species > 0, axis=1)
    first = np.concatenate([[0], np.cumsum(natoms)[:-1]])
    ii, jj = [], []
    for m in range(nmol):
        n = natoms[m]
        a, b = np.triu_indices(n, 1)
        ii.append(a + first[m])
        jj.append(b + first[m])
    idxi = np.concatenate(ii) if ii else np.zeros(0, dtype=np.int64)
    idxj = np.concatenate(jj) if jj else np.zeros(0, dtype=np.int64)
    xyz = coordinates.reshape(-1, 3)[p.real_atoms]
    d = xyz[idxj] - xyz[idxi]
    dist = np.sqrt(np.sum(d * d, axis=1))
    keep = dist * dist < outer_cutoff**2
    idxi, idxj, d, dist = idxi[keep], idxj[keep], d[keep], dist[keep]
    p.idxi, p.idxj = idxi, idxj
    p.ni, p.nj = p.Z[idxi], p.Z[idxj]
    p.xij = d / dist[:, None]
    p.rij = dist * (1.0 / T.a0)
    p.pair_molid = p.atom_molid[idxi]
    p.mask = p.atom_molid[idxi] * molsize * molsize + p.atom_pos[idxi] * molsize + p.atom_pos[idxj]
    return p
