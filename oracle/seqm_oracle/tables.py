"""Element and method parameter tables (data extracted from the reference by tools/make_tables.py).

Restates: seqm/seqm_functions/constants.py:26-213 (Constants tables),
          seqm/seqm_functions/parameters.py:4-88 (params / PWCCT CSV loaders),
          seqm/basics.py:35-196 (parameterlist per method).
"""
import json
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "pyseqm_b200", "data")


class Tables:
    """Per-element constant tables; index = atomic number (constants.py:57-213)."""

    _inst = None

    def __init__(self):
        with open(os.path.join(_DATA, "element_tables.json")) as f:
            d = json.load(f)
        self.ev = d["ev"]  # 27.21 (constants.py:4)
        self.a0 = d["a0"]  # 0.529167 (constants.py:9)
        self.overlap_cutoff = d["overlap_cutoff"]  # 40 bohr (constants.py:23)
        self.to_debye = d["to_debye"]
        self.debye_to_AU = d["debye_to_AU"]
        for k in ["atomic_num", "tore", "qn", "ussc", "uppc", "gssc", "gspc", "hspc", "gp2c", "gppc", "eheat", "mass"]:
            setattr(self, k, np.asarray(d[k], dtype=np.float64))
        self.qn_int = np.asarray(d["qn_int"], dtype=np.int64)

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst


_NGAUSS = {"MNDO": 0, "AM1": 4, "PM3": 2, "PM6_SP": 4, "PM6": 4}


def method_parameters(method, Z):
    """Per-atom parameter vectors p[name][atom] (basics.py:442-448 `self.p[Z, i]`)."""
    with open(os.path.join(_DATA, f"params_{method}.json")) as f:
        d = json.load(f)
    cols = d["columns"]
    zmax = int(max(int(k) for k in d["rows"]))
    tab = np.zeros((zmax + 1, len(cols)))
    for k, v in d["rows"].items():
        tab[int(k)] = v
    out = {}
    for j, c in enumerate(cols):
        out[c] = tab[Z, j].copy()
    out["_ngauss"] = _NGAUSS.get(method, 0)
    if "pairwise_alpha_chi" in d:  # PWCCT (parameters.py:49-88)
        zm = max(max(t[0], t[1]) for t in d["pairwise_alpha_chi"])
        alp = np.zeros((zm + 1, zm + 1))
        chi = np.zeros((zm + 1, zm + 1))
        for zi, zj, a, c in d["pairwise_alpha_chi"]:
            alp[zi, zj] = a
            chi[zi, zj] = c
        out["_alp"], out["_chi"] = alp, chi
    return out
