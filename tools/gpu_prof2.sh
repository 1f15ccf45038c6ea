set -x
python tools/gpu_jstats.py
ncu --set full --clock-control none --import-source on -k regex:jacobi_density_kernel -s 3 -c 1 -o gpurun_out/jacobi_r01_v2 python tools/profile_step.py 4096 1 > gpurun_out/prof_jacobi2.log 2>&1
