python -m pytest tests -m gpu -q > gpurun_out/gt.log 2>&1; tail -5 gpurun_out/gt.log
run() { l=$1; b=$2; shift; shift; env "$@" python bench.py --no-cpu --steps 10 --preroll 150 --batch $b 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$l B=$b', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['health']['reward_mean'], d['health']['blown_up_envs'])
except Exception as e: print('$l failed', e)"; }
run auto 3072 X=1
run fused 3072 AVSIM_SPLIT=0
run auto 2560 X=1
run split 2560 AVSIM_SPLIT=1
