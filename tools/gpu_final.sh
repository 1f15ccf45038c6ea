set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01_final.csv python tools/profile_step.py 4096 1 > gpurun_out/prof_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jacobi_fixed_kernel -s 20 -c 6 -o gpurun_out/jacobi_r01_final python tools/profile_step.py 4096 1 > gpurun_out/prof_jacobi4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fock_pair_kernel|diis_store_kernel|pair_gradient_kernel|pair_integrals_kernel" -s 6 -c 8 -o gpurun_out/others_r01_final python tools/profile_step.py 4096 1 > gpurun_out/prof_others4.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
