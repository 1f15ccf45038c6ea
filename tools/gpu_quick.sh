# kernel experiment: full GPU suite, eigensolver micro-bench, default bench line without extras
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python tools/bench_eig.py 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --extras none > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - $TAG <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "iters", d.get("scf_iterations"))
print({k:v["ms"] for k,v in d["kernel_breakdown"].items()})
PY
