set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "iters", d["scf_iterations"], "launches", d["gpu_launches"])
for k,v in d["kernel_breakdown"].items(): print(f"  {k:18s} {v['ms']:9.3f} ms  n={v['launches']:3d}  {v['share']:.3f}")
print(d["roofline"]["achieved"], d["roofline"]["peak"], d["roofline"].get("frac"))
print(d["xl_bomd"])
PY
