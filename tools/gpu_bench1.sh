set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -3 gpurun_out/bench1.err
cat gpurun_out/bench1.json
