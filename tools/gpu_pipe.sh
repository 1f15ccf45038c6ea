set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for p in 1 2; do
SEQM_B200_PIPELINE=$p timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --xl-replicas 0 > gpurun_out/bench_p$p.json 2> gpurun_out/bench_p$p.err
tail -2 gpurun_out/bench_p$p.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_p$p.json'))
print("pipeline $p value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "iters", d["scf_iterations"], "launches", d["gpu_launches"])
for k,v in d["kernel_breakdown"].items(): print(f"  {k:18s} {v['ms']:9.3f} ms  n={v['launches']:3d}  {v['share']:.3f}")
PY
done
