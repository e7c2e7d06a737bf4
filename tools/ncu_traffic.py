"""Summarise an ncu launch list of bench.py into profiles/: per-kernel launch counts, average durations, the share of the
step each kernel takes, and the DRAM bytes one env.step moves (-> profiles/r2_traffic.json, read by bench.py's
roofline.traffic).

capture (GPU box; a number printed by bench.py under ncu is not a bench value):
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:avsim_ \\
      -c 1440 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --preroll 0 --no-cpu
then here:
  python tools/ncu_traffic.py gpurun_out/r2_launches.csv 4096 [groups=3]
"""
import csv, json, os, sys
from collections import defaultdict

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
src, batch = sys.argv[1], int(sys.argv[2])
groups = int(sys.argv[3]) if len(sys.argv) > 3 else 3
NSUB = 20
rows = [r for r in csv.reader(open(src, newline="")) if len(r) > 10 and r[0].isdigit()]
per = defaultdict(lambda: defaultdict(list))
for r in rows:
    name, metric, unit, val = r[4].split("(")[0], r[-3], r[-2], float(r[-1].replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    per[name][metric].append(val * scale)
n_sub = len(per["avsim_substep_kernel"]["gpu__time_duration.sum"])
steps = n_sub / (NSUB * groups)
lines = [f"ncu launch list of `bench.py --steps 2 --warmup 3 --preroll 0 --no-cpu` (batch {batch}, {groups} environment groups): "
         f"{n_sub} substep launches = {steps:.2f} env.steps captured",
         "per-launch times under ncu are serialised and cold-cache: compare the SHARES with the live CUDA-event numbers, not the absolutes",
         "", f"{'kernel':34s} {'launches':>8s} {'avg ms':>9s} {'ms/step':>9s} {'share':>7s} {'dram MB/step':>13s}"]
tot_ms = sum(sum(v["gpu__time_duration.sum"]) for v in per.values())
dram_step = 0.0
out = {}
for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
    t = v["gpu__time_duration.sum"]
    d = sum(v.get("dram__bytes_read.sum", [0])) + sum(v.get("dram__bytes_write.sum", [0]))
    lines.append(f"{name:34s} {len(t):8d} {sum(t) / len(t):9.4f} {sum(t) / steps:9.3f} {sum(t) / tot_ms * 100:6.1f}% {d / steps / 1e6:13.2f}")
    out[name] = {"launches_per_step": len(t) / steps, "avg_ms": sum(t) / len(t), "share": sum(t) / tot_ms, "dram_bytes_per_step": d / steps}
    if name in ("avsim_substep_kernel", "avsim_solve_kernel", "avsim_step_kernel"):
        dram_step += d / steps
lines += ["", f"DRAM read+write of one env.step's substep+solve kernels: {dram_step / 1e6:.1f} MB = {dram_step / batch:.0f} B per environment-step "
          f"(algorithmic state traffic: 1032 B; the rest is the head records and contact scratch passed between the two kernels of "
          f"each substep -- ~9 KB x in/out x 2 kernels x 20 substeps.  ncu flushes the caches before every launch, so this is the cold "
          f"upper bound: live, a group's records (1365 x 9 KB = 12 MB) are still in the 126 MB L2 when the next kernel reads them)"]
open(os.path.join(ROOT, "profiles", "r2_launches_summary.txt"), "w").write("\n".join(lines) + "\n")
json.dump({"batch": batch, "groups": groups, "dram_bytes_per_step": dram_step, "steps_captured": steps, "per_kernel": out,
           "source": "tools/ncu_traffic.py over " + os.path.basename(src)}, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print("\n".join(lines))
