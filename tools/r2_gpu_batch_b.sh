#!/bin/bash
# Round-2 measurement batch B (GPU box): environment-group / solve-warp sweeps, the ncu launch list of the bench command,
# one full ncu capture of each of the two step kernels.
O=gpurun_out
run() { # label, env...
  l=$1; shift
  env "$@" python bench.py --no-cpu --steps 10 > $O/r2_sweep_$l.json 2>/dev/null
  python - "$l" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r2_sweep_{sys.argv[1]}.json"))
    print(f"{sys.argv[1]:28s} {d['ms_per_step']:8.3f} ms/step {d['value']:10.0f} env-steps/s   e2e {d['e2e']['value']:10.0f}")
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
}
run groups1 AVSIM_GROUPS=1
run groups2 AVSIM_GROUPS=2
run groups3_warps8_default AVSIM_GROUPS=3
run groups4 AVSIM_GROUPS=4
run groups6 AVSIM_GROUPS=6
run warps4 AVSIM_SOLVE_WARPS=4
run warps12 AVSIM_SOLVE_WARPS=12
run warps16 AVSIM_SOLVE_WARPS=16
run fused_newton AVSIM_SPLIT=0
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:avsim_ -c 1500 --csv \
    --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 --preroll 0 --no-cpu > $O/r2_launches_bench.log 2>&1
tail -c 300 $O/r2_launches_bench.log
# steady-state launches: skip the first 6 env.steps' worth (120 launches each)
ncu --set full --clock-control none --import-source on -k regex:avsim_substep_kernel -s 400 -c 1 -o $O/r2_substep_kernel_final \
    python bench.py --steps 2 --warmup 8 --preroll 0 --no-cpu > $O/r2_ncu_substep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:avsim_solve_kernel -s 400 -c 1 -o $O/r2_solve_kernel_final \
    python bench.py --steps 2 --warmup 8 --preroll 0 --no-cpu > $O/r2_ncu_solve.log 2>&1
ls -la $O/*.ncu-rep
