# round-2 final lease: parity suite, smoke, default bench (both arms), PM6 extra, forward launch count, ncu launch list and
# --set full captures of the eigensolver and the Fock kernel.  Usage: gpurun --timeout 3300 -- 'bash tools/gpu_lease.sh TAG'
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_$TAG.err
SEQM_B200_FOCK_BULK=0 timeout 600 python bench.py --steps 20 --warmup 5 --extras none > gpurun_out/bench_${TAG}_nobulk.json 2> gpurun_out/bench_${TAG}_nobulk.err; echo "bench nobulk rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --extras pm6 > gpurun_out/bench_pm6d_$TAG.json 2> gpurun_out/bench_pm6d_$TAG.err; echo "pm6 rc=$?"
timeout 300 python tools/count_launches.py $TAG > gpurun_out/count_$TAG.log 2>&1; head -3 gpurun_out/count_$TAG.log; cp profiles/forward_launches_$TAG.txt gpurun_out/ 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py 4096 1 > gpurun_out/prof_launch_$TAG.log 2>&1; tail -1 gpurun_out/prof_launch_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fock_pair_kernel|jacobi" -s 6 -c 4 -o gpurun_out/hot_$TAG python tools/profile_step.py 4096 1 > gpurun_out/prof_full_$TAG.log 2>&1; tail -1 gpurun_out/prof_full_$TAG.log
