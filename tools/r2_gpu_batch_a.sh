#!/bin/bash
# Round-2 measurement batch A (GPU box): GPU test suite, smoke, the bench line, the reference arm, the TMA A/B.
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2_gputests.log 2>&1; tail -3 $O/r2_gputests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_smoke.log 2>&1; tail -2 $O/r2_smoke.log
python bench.py > $O/r2_bench.json 2> $O/r2_bench.err; cut -c1-400 $O/r2_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err; cut -c1-300 $O/r2_bench_ref.json
for i in 1 2; do
  python bench.py --no-cpu --steps 20 > $O/r2_ab_tma_$i.json 2>/dev/null
  AVSIM_LIB=$PWD/av_aloha_b200/csrc/libavsim_notma.so python bench.py --no-cpu --steps 20 > $O/r2_ab_notma_$i.json 2>/dev/null
done
python - <<'PY'
import json
for n in ("tma_1", "notma_1", "tma_2", "notma_2"):
    try:
        d = json.load(open(f"gpurun_out/r2_ab_{n}.json"))
        print(n, round(d["ms_per_step"], 3), "ms/step", round(d["value"]), "env-steps/s  e2e", round(d["e2e"]["value"]))
    except Exception as e:
        print(n, "failed", e)
PY
