ncu --set full --clock-control none --import-source on -k regex:"pair_gradient_kernel|pair_integrals_kernel" -s 12 -c 6 -o gpurun_out/pairs_xl_r01 python tools/gpu_xlprof.py 1024 > gpurun_out/prof_xl.log 2>&1
tail -2 gpurun_out/prof_xl.log
