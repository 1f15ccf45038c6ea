"""Third-row (Si, P, S, Cl: qn=3 overlap branches) fixtures from the UNMODIFIED reference, forces in the reference's
default (autograd) mode: its analytical_gradient=[True] route disagrees with its own autograd gradient by up to
0.23 eV/A on Cl (PM3) and Si-F (PM6_SP) pairs, so it is not used as the truth here (deviations)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from refrun import np, read_xyz, run_reference
XYZ="/root/repo/tests/golden/xyz/"
files=["h2s.xyz","ch3cl.xyz","ph3.xyz","sih3f.xyz","ch3sh.xyz","pcl3.xyz"]
species, coords = read_xyz([XYZ+f for f in files])
print(species)
KEEP = ["Etot","Hf","Eelec","Enuc","Eiso","e_mo","e_gap","dm","q","force","notconverged","n_scf_iter","dipole"]
for method in ["PM3","AM1","MNDO","PM6_SP"]:
    sp = {"method": method, "scf_eps": 1e-7, "scf_converger": [2], "sp2": [False]}
    try:
        ref = run_reference(species, coords, sp)
    except Exception as e:
        print(method, "reference failed:", repr(e)[:200]); continue
    out = {k: ref[k] for k in KEEP}
    out.update(species=species, coordinates=coords, seqm_parameters=json.dumps(sp))
    np.savez_compressed(f"/root/repo/tests/golden/thirdrow_{method}_c2.npz", **out)
    print(method, "iters", ref["n_scf_iter"], "notconv", ref["notconverged"].sum(), "Etot", ref["Etot"])
