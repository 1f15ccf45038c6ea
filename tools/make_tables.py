#!/usr/bin/env python
"""Extract the element / parameter DATA tables of the reference into JSON.

Run in the build container (needs /root/reference):
    python tools/make_tables.py

Reads   /root/reference/seqm/seqm_functions/constants.py  (Constants: tore, qn, eheat, mass, ...)
        /root/reference/seqm/params/parameters_<METHOD>_MOPAC.csv, PWCCT_<METHOD>_MOPAC.csv
Writes  pyseqm_b200/data/element_tables.json, pyseqm_b200/data/params_<METHOD>.json

Only numbers travel: the product never reads /root/reference at run time.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle", "h5py_stub"))
sys.path.insert(0, "/root/reference")
OUT = os.path.join(HERE, "..", "pyseqm_b200", "data")

import torch  # noqa: E402

torch.set_default_dtype(torch.float64)
from seqm.seqm_functions import constants as C  # noqa: E402

const = C.Constants()
el = {
    "ev": C.ev,
    "a0": C.a0,
    "ev_kcalpmol": C.ev_kcalpmol,
    "overlap_cutoff": C.overlap_cutoff,
    "to_debye": C.to_debye,
    "debye_to_AU": C.debye_to_AU,
    "label": [s.strip() for s in const.label],
}
for name in ["atomic_num", "tore", "iso", "qn", "ussc", "uppc", "gssc", "gspc", "hspc", "gp2c", "gppc", "eheat", "mass"]:
    el[name] = [float(x) for x in getattr(const, name).tolist()]
for name in ["qn_int", "qnD_int"]:
    el[name] = [int(x) for x in getattr(const, name).tolist()]
with open(os.path.join(OUT, "element_tables.json"), "w") as f:
    json.dump(el, f, indent=0)

pdir = "/root/reference/seqm/params/"
for method in ["MNDO", "AM1", "PM3", "PM6", "PM6_SP"]:
    with open(pdir + f"parameters_{method}_MOPAC.csv") as f:
        header = f.readline().strip().replace(" ", "").split(",")
        cols = header[2:]
        rows = {}
        for line in f:
            t = line.strip().replace(" ", "").split(",")
            if len(t) < 3:
                continue
            try:
                vals = [float(x) for x in t[2:]]
            except ValueError:
                print("skipping malformed row for Z =", t[0])
                continue
            vals = (vals + [0.0] * len(cols))[: len(cols)]
            if any(v != 0.0 for v in vals):
                rows[int(t[0])] = vals
    out = {"method": method, "columns": cols, "rows": rows}
    pw = pdir + f"PWCCT_{method}_MOPAC.csv"
    if os.path.exists(pw):
        trip = []
        with open(pw) as f:
            for line in f:
                t = line.strip().replace(" ", "").split(",")
                if len(t) >= 4:
                    trip.append([int(t[0]), int(t[1]), float(t[2]), float(t[3])])
        out["pairwise_alpha_chi"] = trip
    with open(os.path.join(OUT, f"params_{method}.json"), "w") as f:
        json.dump(out, f)
    print(method, len(rows), "elements", len(cols), "columns")
