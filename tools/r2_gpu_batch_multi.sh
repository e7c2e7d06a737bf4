#!/bin/bash
# Round-2 multi-GPU batch: weak and strong scaling of the bench line, the reference arm under torchrun, and the config-5 leg.
N=${1:-2}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --steps 20 --warmup 3 > $O/r2_bench_n${N}_weak.json 2> $O/r2_bench_n${N}_weak.err
$TR bench.py --gpus $N --steps 20 --warmup 3 --scaling strong > $O/r2_bench_n${N}_strong.json 2> $O/r2_bench_n${N}_strong.err
$TR tools/eval_rollout.py --batch 1024 --rollouts 1 --steps 300 --cameras 4 > $O/r2_eval_rollout_n${N}.json 2> $O/r2_eval_rollout_n${N}.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for k in ("weak", "strong"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_n{n}_{k}.json").read().strip().splitlines()[-1])
        print(k, "n_gpus", d["n_gpus"], round(d["ms_per_step"], 2), "ms/step", round(d["value"]), "env-steps/s e2e", round(d["e2e"]["value"]), d["config"]["batch_per_gpu"], "envs/GPU")
    except Exception as e:
        print(k, "failed", e)
try:
    print(open(f"gpurun_out/r2_eval_rollout_n{n}.json").read().strip().splitlines()[-1][:700])
except Exception as e:
    print("eval failed", e)
PY
