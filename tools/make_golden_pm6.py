"""method="PM6" fixtures restricted to elements that carry no d shell in PM6 (Z <= 12): the reference runs them through
its 9-orbital-per-atom code path (dm (nmol, 9 molsize, 9 molsize), w (npairs, 45, 45)); pyseqm_b200 serves them with
the sp kernels and widens the results.  Run in the build container only (imports /root/reference)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from refrun import np, read_xyz, run_reference  # noqa: E402

from pyseqm_b200.synthetic import qm9_like_batch  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
XYZ = os.path.join(OUT, "xyz")
KEEP = ["Etot", "Hf", "Eelec", "Enuc", "Eiso", "e_mo", "e_gap", "dm", "q", "force", "notconverged", "n_scf_iter", "dipole"]

sp = {"method": "PM6", "scf_eps": 1e-7, "scf_converger": [2], "sp2": [False]}
species, coords = read_xyz([os.path.join(XYZ, f) for f in ("ch3f.xyz", "methane.xyz", "benzene.xyz", "toluene.xyz")])
ref = run_reference(species, coords, sp)
out = {k: ref[k] for k in KEEP if ref[k] is not None}
w = ref["w"]
assert np.abs(w[:, 10:, :]).max() == 0.0 and np.abs(w[:, :, 10:]).max() == 0.0  # sp pairs are the first 10 of 45
out["w_sp"] = w[:, :10, :10]
out.update(species=species, coordinates=coords, seqm_parameters=json.dumps(sp))
np.savez_compressed(os.path.join(OUT, "pm6_sp_elements_c2.npz"), **out)
print("cfg1+CH3F PM6 iters", ref["n_scf_iter"], "Etot", ref["Etot"])

species, coords = qm9_like_batch(12, seed=5, molsize=20)
sp = {"method": "PM6", "scf_eps": 1e-7, "scf_converger": [1], "sp2": [False]}
ref = run_reference(species, coords, sp)
out = {k: ref[k] for k in KEEP if k != "dm" and ref[k] is not None}
out["dm_diag"] = np.diagonal(ref["dm"], axis1=1, axis2=2).copy()
out.update(species=species, coordinates=coords, seqm_parameters=json.dumps(sp))
np.savez_compressed(os.path.join(OUT, "pm6_sp_elements_qm9_12_c1.npz"), **out)
print("qm9-like PM6 iters", ref["n_scf_iter"], "notconv", ref["notconverged"].sum())
