#!/bin/bash
# strong scaling + the config-5 leg after the batch-size-dependent launch shapes
N=${1:-2}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --steps 20 --warmup 3 --scaling strong > $O/r2_bench_n${N}_strong.json 2> $O/r2_bench_n${N}_strong.err
$TR tools/eval_rollout.py --batch 1024 --rollouts 1 --steps 300 --cameras 4 > $O/r2_eval_rollout_n${N}.json 2> $O/r2_eval_rollout_n${N}.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/r2_bench_n{n}_strong.json").read().strip().splitlines()[-1])
print("strong n_gpus", d["n_gpus"], round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "env-steps/s e2e", round(d["e2e"]["value"]), d["config"]["launch_shape"])
print(open(f"gpurun_out/r2_eval_rollout_n{n}.json").read().strip().splitlines()[-1][:420])
PY
