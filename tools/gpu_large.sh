set -x
python -m pytest tests/test_large_molecule.py tests/test_md.py -m gpu -x -q 2>&1 | tail -6
python - <<'PY'
import sys, time, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
torch.set_default_dtype(torch.float64)
from conftest import load_golden
import pyseqm_b200 as seqm
from pyseqm_b200._lib import get_lib
lib = get_lib()
dev = torch.device("cuda:0")
g = load_golden("cfg4_C380_AM1_sp2")
sp = dict(g["seqm_parameters"])
mol = seqm.Molecule(seqm.Constants().to(dev), dict(sp), torch.as_tensor(g["coordinates"], device=dev), torch.as_tensor(g["species"], device=dev)); mol.verbose=False
es = seqm.Electronic_Structure(dict(sp))
es(mol); torch.cuda.synchronize()
lib.profile_enable(True)
t = time.perf_counter(); es(mol); torch.cuda.synchronize(); dt = time.perf_counter() - t
prof = lib.profile_collect(); lib.profile_enable(False)
print("C380 forward %.3f s, iters %d (ref %d in %.0f s on 8 CPU threads), Etot %.8f ref %.8f" % (dt, mol.n_scf_iter, g["n_scf_iter"], float(g["reference_seconds_8threads"]), float(mol.Etot[0]), float(g["Etot"][0])))
for k, v in prof.items():
    if v[1]: print("  %-18s %9.2f ms n=%d" % (k, v[0], v[1]))
n = 1520
gm = prof["dgemm"]
print("dgemm TFLOP/s: %.2f" % (2.0 * n**3 * gm[1] / (gm[0] * 1e-3) / 1e12))
PY
