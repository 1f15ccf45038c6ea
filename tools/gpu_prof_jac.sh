ncu --set full --clock-control none --import-source on -k regex:jacobi_fixed_kernel -s 20 -c 6 -o gpurun_out/jacobi_r01_final python tools/profile_step.py 4096 1 > gpurun_out/prof_jacobi4.log 2>&1
tail -1 gpurun_out/prof_jacobi4.log
