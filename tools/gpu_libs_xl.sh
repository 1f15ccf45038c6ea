for f in variants/*.so; do
SEQM_B200_LIB=$PWD/$f python tools/gpu_xlprof.py 1024 2>/dev/null | grep -E "XL-BOMD steps|pair_integrals|gradient" | tr '\n' ' '; echo " <- $f"
done
