for f in variants/*.so; do
SEQM_B200_LIB=$PWD/$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline --xl-replicas 0 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$f', round(d['ms_per_step'],3), round(d['e2e']['value']), 'jacobi', d['kernel_breakdown']['jacobi_density']['ms'])"
done
