# eigensolver / SCF parity subset + default bench line (no extras) for a kernel experiment
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 --extras none > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_$TAG.err
python - $TAG <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_%s.json" % sys.argv[1] if len(sys.argv)>1 else "gpurun_out/bench_q.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "iters", d.get("scf_iterations"))
print({k:v["ms"] for k,v in d["kernel_breakdown"].items()})
PY
