run() { l=$1; b=$2; shift; shift; env "$@" python bench.py --no-cpu --steps 10 --preroll 150 --batch $b 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$l B=$b', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['health']['reward_mean'], d['health']['blown_up_envs'])
except Exception as e: print('$l failed', e)"; }
run auto 128 X=1
run auto 512 X=1
run auto 1024 X=1
run old 1024 AVSIM_ENVW=11 AVSIM_SOLVE_WARPS=8
run auto 2048 X=1
run old 2048 AVSIM_ENVW=11 AVSIM_SOLVE_WARPS=8
run auto 4096 X=1
run old 4096 AVSIM_ENVW=11 AVSIM_SOLVE_WARPS=8
run e9 4096 AVSIM_ENVW=9 AVSIM_SOLVE_WARPS=8
