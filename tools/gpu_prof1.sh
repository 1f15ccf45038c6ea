set -x
python tools/profile_step.py 4096 1 > /dev/null   # populate /tmp workload cache
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01.csv python tools/profile_step.py 4096 1 > gpurun_out/prof_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jacobi_density_kernel -s 2 -c 2 -o gpurun_out/jacobi_r01 python tools/profile_step.py 4096 1 > gpurun_out/prof_jacobi.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fock_kernel|diis_solve_kernel|pair_gradient_kernel" -s 3 -c 3 -o gpurun_out/others_r01 python tools/profile_step.py 4096 1 > gpurun_out/prof_others.log 2>&1
ls -la gpurun_out
