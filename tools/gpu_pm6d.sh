# PM6 d-orbital path on the GPU: the whole parity suite, then the bench line with configs[4]
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_pm6d.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_pm6d.log
timeout 600 python bench.py --steps 5 --warmup 3 --extras pm6 > gpurun_out/bench_pm6d.json 2> gpurun_out/bench_pm6d.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_pm6d.err
