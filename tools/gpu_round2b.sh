# parity suite, PM6 bench extra, launch lists (configs[1] step and PM6-d step), ncu --set full of the spd kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_b.log
timeout 600 python bench.py --steps 5 --warmup 3 --extras pm6 > gpurun_out/bench_pm6d_b.json 2> gpurun_out/bench_pm6d_b.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_pm6d_b.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r02_pm6d.csv python tools/profile_pm6d.py 512 1 > gpurun_out/prof_pm6d_launch.log 2>&1; tail -1 gpurun_out/prof_pm6d_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spd_ -c 8 -o gpurun_out/spd_r02 python tools/profile_pm6d.py 512 1 > gpurun_out/prof_spd_full.log 2>&1; tail -1 gpurun_out/prof_spd_full.log
