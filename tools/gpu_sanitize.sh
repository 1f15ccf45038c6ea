cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
torch.set_default_dtype(torch.float64)
import pyseqm_b200 as seqm
from pyseqm_b200.synthetic import qm9_like_batch
dev = torch.device("cuda:0")
n = int(sys.argv[1])
species, coords = qm9_like_batch(n, seed=3)
for sp in ({"method": "PM3", "scf_eps": 1e-7, "scf_converger": [2], "sp2": [False]},
           {"method": "AM1", "scf_eps": 1e-6, "scf_converger": [2], "sp2": [True, 1e-5]}):
    mol = seqm.Molecule(seqm.Constants().to(dev), dict(sp), torch.as_tensor(coords, device=dev), torch.as_tensor(species, device=dev))
    mol.verbose = False
    es = seqm.Electronic_Structure(dict(sp)); es(mol); torch.cuda.synchronize()
    print(sp["sp2"], mol.n_scf_iter, float(mol.Etot.sum()), int(es.notconverged.sum()))
PY
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py 300 2>&1 | tail -6
echo "memcheck rc=$?"
timeout 500 compute-sanitizer --tool racecheck --racecheck-report all python /tmp/san.py 8 > gpurun_out/racecheck.log 2>&1; grep -E "Race reported|RACECHECK SUMMARY" gpurun_out/racecheck.log | sed "s/void //; s/(seqm_batch.*)+0x[0-9a-f]*//" | sort | uniq -c | sort -rn | head -30
echo "racecheck rc=$?"
