# configs[4] (PM6-d) bench key for every library variant under variants/
mkdir -p gpurun_out
for so in "" $(ls variants/lib_*.so 2>/dev/null); do
  tag=$(basename "${so:-base}" .so)
  SEQM_B200_LIB=${so:+$PWD/$so} timeout 400 python bench.py --steps 5 --warmup 3 --extras pm6 > gpurun_out/varp_$tag.json 2> gpurun_out/varp_$tag.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/varp_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])["pm6_d"]
    print(sys.argv[1], "value %.0f ms/step %.2f" % (d["value"], d["ms_per_step"]), d.get("ms_all"), d["kernel_ms"])
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
