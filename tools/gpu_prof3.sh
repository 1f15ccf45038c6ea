set -x
ncu --set full --clock-control none --import-source on -k regex:jacobi_fixed_kernel -s 24 -c 8 -o gpurun_out/jacobi_r01_v3 python tools/profile_step.py 4096 1 > gpurun_out/prof_jacobi3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pair_gradient_kernel|fock_kernel|diis_store_kernel|pair_integrals_kernel" -s 5 -c 4 -o gpurun_out/others_r01_v3 python tools/profile_step.py 4096 1 > gpurun_out/prof_others3.log 2>&1
