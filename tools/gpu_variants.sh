# bench line (no extras) for every library variant under variants/ (kernel experiments; SEQM_B200_LIB selects the library)
mkdir -p gpurun_out
for so in "" $(ls variants/lib_*.so 2>/dev/null); do
  tag=$(basename "${so:-base}" .so)
  SEQM_B200_LIB=${so:+$PWD/$so} timeout 300 python bench.py --steps 10 --warmup 3 --extras none > gpurun_out/var_$tag.json 2> gpurun_out/var_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/var_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    kb = d["kernel_breakdown"]
    print(sys.argv[1], "value %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), "jacobi %.3f fock %.3f iters %s" % (kb["jacobi_density"]["ms"], kb["fock"]["ms"], d.get("scf_iterations")))
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
