# parity suite after the translation-unit split, C380 with the symmetric X^2 kernel, instruction counts of one configs[1] step
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 --extras c380 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum --clock-control none -c 3000 --csv --log-file gpurun_out/inst_$TAG.csv python tools/profile_step.py 4096 1 > gpurun_out/prof_inst_$TAG.log 2>&1; tail -1 gpurun_out/prof_inst_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dgemm_sym_kernel" -s 20 -c 1 -o gpurun_out/dgemm_sym_$TAG python tools/profile_c380.py > gpurun_out/prof_c380_$TAG.log 2>&1; tail -1 gpurun_out/prof_c380_$TAG.log
timeout 300 python tools/c380_margins.py > gpurun_out/c380_margins_$TAG.log 2>&1; cat gpurun_out/c380_margins_$TAG.log | tail -8
