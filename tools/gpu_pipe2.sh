run() {
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --xl-replicas 0 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), round(d['e2e']['value']))"
}
SEQM_B200_PIPELINE=2 run base
SEQM_B200_PIPELINE=2 SEQM_PIPE_NOPRIO=1 run noprio
SEQM_B200_PIPELINE=2 SEQM_PIPE_NOSTAGGER=1 run stagger_first_only
SEQM_B200_PIPELINE=2 SEQM_PIPE_NOSTAGGER=2 run no_stagger
SEQM_B200_PIPELINE=2 SEQM_PIPE_NOSTAGGER=2 SEQM_PIPE_NOPRIO=1 run no_stagger_noprio
SEQM_B200_PIPELINE=1 run single
