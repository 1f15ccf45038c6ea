mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "pm6" > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_q.log
timeout 600 python bench.py --steps 5 --warmup 3 --extras pm6 > gpurun_out/bench_pm6d_c.json 2> gpurun_out/bench_pm6d_c.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spd_pair_gradient -c 1 -o gpurun_out/spdgrad_r02 python tools/profile_pm6d.py 512 1 > gpurun_out/prof_spdgrad.log 2>&1; tail -1 gpurun_out/prof_spdgrad.log
