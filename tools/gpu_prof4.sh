set -x
python tools/profile_step.py 4096 1 > /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01_final.csv python tools/profile_step.py 4096 1 > gpurun_out/prof_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jacobi_fixed_kernel -s 20 -c 6 -o gpurun_out/jacobi_r01_final python tools/profile_step.py 4096 1 > gpurun_out/prof_jacobi4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fock_pair_kernel|diis_store_kernel|pair_gradient_kernel|pair_integrals_kernel" -s 6 -c 8 -o gpurun_out/others_r01_final python tools/profile_step.py 4096 1 > gpurun_out/prof_others4.log 2>&1
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
ls -la gpurun_out | tail -8
