set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01_s2.csv python tools/profile_step.py 4096 2 > gpurun_out/prof_launch_s2.log 2>&1
tail -2 gpurun_out/prof_launch_s2.log
python bench.py > gpurun_out/bench_s6.json 2> gpurun_out/bench_s6.err
