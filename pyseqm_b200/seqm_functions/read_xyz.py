"""read_xyz with the reference's contract (seqm/seqm_functions/read_xyz.py:17-54): atoms stably sorted by
descending atomic number, molecules zero-padded to the largest one."""
import numpy as np

_SYMBOLS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
            "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe").split()  # fmt: skip
_element_dict = {s: i for i, s in enumerate(_SYMBOLS) if i > 0}


def read_xyz(files, sort=True):
    mols = []
    for fn in files:
        with open(fn) as f:
            lines = f.readlines()
        n = int(lines[0])
        data = np.zeros((n, 4), float)
        for i, line in enumerate(lines[2 : 2 + n]):
            a, *xyz = line.split()
            data[i, 0] = int(a) if a.isdigit() else _element_dict[a]
            data[i, 1:4] = [float(t) for t in xyz[:3]]
        if sort:
            data = data[np.argsort(-data[:, 0], kind="stable")]
        mols.append(data)
    K = max(m.shape[0] for m in mols)
    species = np.zeros((len(mols), K), int)
    coords = np.zeros((len(mols), K, 3), float)
    for i, m in enumerate(mols):
        species[i, : m.shape[0]] = m[:, 0].astype(int)
        coords[i, : m.shape[0]] = m[:, 1:]
    return species, coords
