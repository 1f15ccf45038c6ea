cat > /tmp/c380.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
torch.set_default_dtype(torch.float64)
from conftest import load_golden
import pyseqm_b200 as seqm
dev = torch.device("cuda:0")
g = load_golden("cfg4_C380_AM1_sp2"); sp = dict(g["seqm_parameters"])
mol = seqm.Molecule(seqm.Constants().to(dev), dict(sp), torch.as_tensor(g["coordinates"], device=dev), torch.as_tensor(g["species"], device=dev)); mol.verbose = False
seqm.Electronic_Structure(dict(sp))(mol); torch.cuda.synchronize(); print(mol.n_scf_iter, float(mol.Etot[0]))
PY
ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma_kernel -s 50 -c 2 -o gpurun_out/dgemm_r01_final python /tmp/c380.py > gpurun_out/prof_dgemm.log 2>&1
tail -2 gpurun_out/prof_dgemm.log
