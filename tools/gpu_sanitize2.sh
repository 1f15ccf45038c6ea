cat > /tmp/san2.py <<'PY'
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
torch.set_default_dtype(torch.float64)
import pyseqm_b200 as seqm
from pyseqm_b200.synthetic import qm9_like_batch
dev = torch.device("cuda:0")
s, c = seqm.read_xyz(["tests/golden/xyz/coronene.xyz", "tests/golden/xyz/benzene.xyz"])
for sp2 in ([False], [True, 1e-5]):
    sp = {"method": "AM1", "scf_eps": 1e-6, "scf_converger": [2], "sp2": sp2}
    mol = seqm.Molecule(seqm.Constants().to(dev), dict(sp), torch.as_tensor(c, device=dev), torch.as_tensor(s, device=dev)); mol.verbose = False
    md = seqm.XL_BOMD(xl_bomd_params={"k": 6}, seqm_parameters=sp, timestep=0.4, Temp=300.0)
    torch.manual_seed(0)
    md.run(mol, 3)
    print(sp2, float(mol.Etot.sum()))
species, coords = qm9_like_batch(40, seed=9)
sp = {"method": "PM3", "scf_eps": 1e-7, "scf_converger": [2], "sp2": [False]}
mol = seqm.Molecule(seqm.Constants().to(dev), dict(sp), torch.as_tensor(coords, device=dev), torch.as_tensor(species, device=dev)); mol.verbose = False
seqm.Electronic_Structure(dict(sp))(mol); torch.cuda.synchronize(); print(mol.n_scf_iter, float(mol.Etot.sum()))
PY
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python /tmp/san2.py > gpurun_out/racecheck2.log 2>&1
grep -E "Race reported|RACECHECK SUMMARY" gpurun_out/racecheck2.log | sed "s/void //; s/(seqm_batch.*)+0x[0-9a-f]*//" | sort | uniq -c | sort -rn | head -20
tail -4 gpurun_out/racecheck2.log
