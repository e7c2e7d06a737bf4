run() { b=$1; shift; env "$@" python tools/eval_rollout.py --batch $b --rollouts 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$b $*', round(d['device_s']/300*1e3,2), 'ms/step', round(d['env_steps_per_s']))"; }
run 512 AVSIM_ENVW=2
run 512 AVSIM_ENVW=3
run 512 AVSIM_ENVW=4
run 1024 AVSIM_ENVW=4
run 1024 AVSIM_ENVW=5
run 1024 AVSIM_ENVW=7
run 2048 AVSIM_ENVW=7
run 2048 AVSIM_ENVW=11
