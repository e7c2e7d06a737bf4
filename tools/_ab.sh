run() { l=$1; b=$2; shift; shift; env "$@" python bench.py --no-cpu --steps 10 --batch $b 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$l B=$b', round(d['ms_per_step'],3), round(d['value']), d['config']['launch_shape'])
except Exception as e: print('$l failed', e)"; }
run e5 4096 AVSIM_ENVW=5
run e7 4096 AVSIM_ENVW=7
run e5_g2 4096 AVSIM_ENVW=5 AVSIM_GROUPS=2
run e7_g2 4096 AVSIM_ENVW=7 AVSIM_GROUPS=2
run e7_g4 4096 AVSIM_ENVW=7 AVSIM_GROUPS=4
