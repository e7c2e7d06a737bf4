#!/bin/sh
# Sweep of the heavy-task schedule (BatchState::heavy_tasks / heavy_warps) on the steady-state bench workload:
#   sh tools/heavy_sweep.sh "0:0 16:7 32:7"      (tasks:warps; 0:0 = every task takes a full block)
cd "$(dirname "$0")/.."
for h in ${1:-0:0 16:7 32:7 64:7}; do
    echo "== AVSIM_HEAVY=$h"
    AVSIM_HEAVY=$h timeout 200 python tools/variant_bench.py 4096 8 ${2:-13:1} 2>&1 | tail -2
done
