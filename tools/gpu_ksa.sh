# KSA-XL-BOMD on the GPU: parity tests + the XL-BOMD extras of the bench (incl. the ksa_branch key)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_md.py tests/test_abi.py -m gpu -x -q > gpurun_out/pytest_ksa.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_ksa.log
timeout 900 python bench.py --steps 5 --warmup 3 --extras xl_bomd --xl-steps 100 > gpurun_out/bench_ksa.json 2> gpurun_out/bench_ksa.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_ksa.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_ksa.json").read().strip().splitlines()[-1])
x=d["xl_bomd"]
print("xl", x["value"], "eig", x["eigensolver_branch"]["value"], "ksa", x.get("ksa_branch"))
PY
