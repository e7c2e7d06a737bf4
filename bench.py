#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched env.step() hot path (BASELINE.json metric).

Workload (BASELINE.json configs[1]): SlotInsertion-3Arms-v0, 4096 environments per GPU in
lockstep, no render.  Actions are a synthetic scripted policy that performs the task (av_aloha_b200/workload.py: reach,
pinch a slot rail with the left hand, grasp the stick with the right, lift, carry over the slot, lower into the gap;
waypoints IK-solved per environment for object placements drawn from the reference's reset ranges, + N(0, 0.01 rad)
joint noise, seed 1234).  Solver setting: 8 PGS sweeps + 3 noslip per substep with the per-constraint force-cache warm
start -- 3-20x closer to the converged trajectory on this workload than 20 sweeps with MuJoCo-style warm start, the setting the round
started with (profiles/r1_warmstart_accuracy.txt).  One "step" = one env.step over the whole batch = 20 physics substeps of 2 ms + reward +
agent_pos (reference gym_guided_vision/env.py:203-226), including the auto-reset of the environments whose 300-step
episode (sim_slot_insertion_3arms.yaml:17) ended on that step.  Episodes are STAGGERED: environment e starts at script
phase (e * 300) // B, so every timed step sees the whole episode's mix of free motion, grasping and insertion and
costs the episode average (an untimed 300-step pre-roll brings the batch to that steady state).

Arms:
  default            : the CUDA path through the C-ABI (libavsim.so).  `value` = device-resident actions;
                       `e2e` = avsim_step_host with HOST numpy buffers (H2D of actions, D2H of agent_pos + reward
                       inside the timed region, every step).
  --impl reference   : the CPU path on this box's host cores (the fp64 oracle, "CPU restatement -- MuJoCo is not
                       installable offline"), all host threads, on a bounded sample of the same workload.

Multi-GPU (torchrun, one rank per GPU): envs are independent, every rank steps its own 4096 (weak scaling, no
data-path collective); one NCCL all_gather of per-env success at the end.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TASK, ARMS = "slot_insertion", 3
EPISODE_LEN = 300
ALGO_BYTES_PER_ENV_STEP = 1032   # fp32 state read + written per env.step
HOME = np.array([0, -0.082, 1.06, 0, -0.953, 0, 0.02239] * 2 + [0, -0.8, 0.8, 0, 0.5, 0, 0], np.float64)
METRIC = "env-steps/sec SlotInsertion-3Arms batch=4096"
NCU_DRAM_BYTES_PER_LAUNCH = 590.7e6   # dram read 39.8 MB + write 550.9 MB per launch at B=4096 (profiles/r1_step_kernel_ncu.txt)


def make_workload(B, seed):
    """(object positions [B,2,3] f64, staggered actions [T,B,21] f32, reset masks [T,B] u8, phases [B])."""
    from av_aloha_b200 import workload

    obj = workload.sample_object_positions(B, seed)
    acts = workload.slot_insertion_script(EPISODE_LEN, obj, seed)
    phase = (np.arange(B) * EPISODE_LEN) // B
    idx = (np.arange(EPISODE_LEN)[:, None] + phase[None, :]) % EPISODE_LEN          # script index of env e at loop step t
    stag = np.take_along_axis(acts, idx[:, :, None], axis=0)
    masks = (idx == 0).astype(np.uint8)                                              # episode of env e (re)starts at step t
    return obj, np.ascontiguousarray(stag), masks, phase


def script_actions(T, B, seed):
    """un-staggered stream [T, B, 21] (tools/ and tests use it)"""
    from av_aloha_b200 import workload

    return workload.slot_insertion_script(T, workload.sample_object_positions(B, seed), seed)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx = float(p[1])
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------ CPU arm (oracle)
def cpu_steps_per_sec(n_envs, n_steps, threads, solver_iters, seed=1234, warmstart=2):
    """The CPU path (fp64 oracle) on `threads` host threads: a sample of n_envs environments of the bench workload, their
    episode phases spread evenly over the 300-step script.  Each environment is first rolled (untimed) from reset to
    its phase, then n_steps consecutive env.steps are timed.  ctypes releases the GIL, so threads run in parallel."""
    from concurrent.futures import ThreadPoolExecutor

    from av_aloha_b200 import model_io, workload
    from oracle.oracle import OracleEnv, OracleModel

    path = model_io.model_path(TASK, ARMS)
    om = OracleModel(path)
    obj = workload.sample_object_positions(n_envs, seed)
    acts = workload.slot_insertion_script(EPISODE_LEN, obj, seed).astype(np.float64)
    phase = (np.arange(n_envs) * EPISODE_LEN) // n_envs
    envs = []
    for e in range(n_envs):
        o = OracleEnv(om)
        o.set_options(max_iter=solver_iters, tol=0.0, warmstart=warmstart)
        o.reset(free_pos=obj[e])
        envs.append(o)

    def preroll(e):
        for t in range(phase[e]):
            envs[e].step(acts[t, e])

    def run(e):
        for k in range(n_steps):
            t = (phase[e] + k) % EPISODE_LEN
            if t == 0:
                envs[e].reset(free_pos=obj[e])
            envs[e].step(acts[t, e])

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(preroll, range(n_envs)))
        t0 = time.perf_counter()
        list(ex.map(run, range(n_envs)))
        dt = time.perf_counter() - t0
    return n_envs * n_steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_envs = 2 * max(cores, 4)
    per_step_envs = n_envs
    vals = []
    total = args.warmup + args.steps
    # each "step" of this arm = one env.step of a bounded sample of `n_envs` environments on all host threads
    v, dt = cpu_steps_per_sec(n_envs, total, cores, args.solver_iters, warmstart=args.warmstart)
    vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_step_envs / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"SlotInsertion-3Arms-v0 no render, scripted grasp/insert policy, sample of {n_envs} envs "
                               f"x {total} consecutive env.steps, episode phases spread over the 300-step script "
                               f"(of the B=4096 staggered workload)",
                   "solver_iters": args.solver_iters, "nsubsteps": 20,
                   "warmstart": "force cache" if args.warmstart == 2 else "qacc map"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{n_envs} envs x {total} env.steps at staggered episode phases, fp64 CPU restatement of the "
                                   f"pipeline (MuJoCo not installable offline), {args.solver_iters} PGS sweeps + 3 noslip"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from av_aloha_b200 import capi, model_io

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    model = capi.Model(model_io.model_path(TASK, ARMS), local)
    batch = capi.Batch(model, B, seed=1234 + rank)
    batch.set_options(solver_iters=args.solver_iters)
    batch.set_warmstart(args.warmstart)
    batch.set_solver(args.solver, args.newton_iters, args.newton_ls, args.newton_tol)
    obj, acts_np, masks_np, phase = make_workload(B, 1234 + rank)
    acts = torch.as_tensor(acts_np, device=dev)                       # [T, B, 21] resident in HBM
    masks = torch.as_tensor(masks_np, device=dev)                     # [T, B] u8: envs whose episode restarts at step t
    mask_any = masks_np.any(axis=1)
    fp_dev = torch.as_tensor(obj.astype(np.float32), device=dev)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = {"t": 0}

    def advance():
        t = state["t"] % EPISODE_LEN
        state["t"] += 1
        return t

    def dev_step(t):
        if mask_any[t]:
            batch.reset(mask=masks[t], free_pos=fp_dev)
        batch.step(acts[t])

    agent_out = np.empty((B, model.njoints), np.float32)
    rew_out = np.empty((B,), np.int32)

    def host_step(t):
        if mask_any[t]:
            batch.reset(mask=masks[t], free_pos=fp_dev)
        batch.step_host(acts_np[t], agent_pos_out=agent_out, reward_out=rew_out)

    # ---- untimed pre-roll: one full script length brings every environment through its first wrap (reset at script
    # index 0), after which the batch holds the steady-state mix of episode phases
    batch.reset(free_pos=fp_dev)
    p0 = time.perf_counter()
    for _ in range(args.preroll):
        dev_step(advance())
    torch.cuda.synchronize()
    preroll_s = time.perf_counter() - p0

    # ---- device-resident arm: per-step CUDA events on the launching (current) stream, L2 flushed between steps
    for _ in range(args.warmup):
        dev_step(advance())
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        sampler.start()
    l0 = batch.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    w0 = time.perf_counter()
    for k in range(args.steps):
        if not args.no_flush:
            flush.zero_()
        t = advance()
        ev[k][0].record()
        dev_step(t)
        ev[k][1].record()
    barrier()
    wall = time.perf_counter() - w0
    launches = batch.launch_count - l0
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(kern_ms))
    status = batch.get(capi.STATUS)
    n_bad = int((status & 1).sum().item())
    n_ovf = int((status & 2 != 0).sum().item())
    ncon_mean = float(batch.get(capi.NCON).float().mean().item())
    cyc = batch.get(capi.ENV_CYCLES).double()
    cyc_stats = {"mean": float(cyc.mean().item()), "p99": float(cyc.quantile(0.99).item()), "max": float(cyc.max().item())}
    rew = batch.get(capi.REWARD)
    rew_max, rew_mean = int(rew.max().item()), float(rew.float().mean().item())
    nws = batch.get(capi.SOLVER_STAT).double()
    solver_stat = {"newton_iters_per_substep_mean": float(nws[:, 0].mean().item()) / 20.0, "scaled_gradient_max": float(nws[:, 1].max().item()),
                   "scaled_gradient_p99": float(nws[:, 1].quantile(0.99).item()), "iters_one_solve_max": float(nws[:, 2].max().item()),
                   "capped_solves": float(nws[:, 3].sum().item())}

    # ---- end-to-end arm: host numpy buffers through avsim_step_host, copies inside the timed region
    for _ in range(min(args.warmup, 3)):
        host_step(advance())
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        host_step(advance())
    e1.record()
    barrier()
    e2e_ms = float(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None

    # success reduction: the only collective of the path (K10), one all_gather per rollout
    succ = batch.get(capi.SUCCESS).to(torch.uint8)
    if world > 1:
        t_all = torch.tensor([total_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t_all[0].item()), float(t_all[1].item())
        gathered = [torch.empty_like(succ) for _ in range(world)]
        dist.all_gather(gathered, succ)
        succ = torch.cat(gathered)
    n_succ = int(succ.sum().item())

    if rank == 0:
        value = world * B * args.steps / (total_ms * 1e-3)
        e2e = world * B * args.steps / (e2e_ms * 1e-3)
        peak, how = peaks()
        kern_avg_ms = total_ms / args.steps
        achieved = ALGO_BYTES_PER_ENV_STEP * B / (kern_avg_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": kern_avg_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"SlotInsertion-3Arms-v0 batch={B}/GPU lockstep, no render, scripted grasp/lift/insert policy "
                                   f"(IK-solved joint targets + N(0,0.01) noise), 300-step episodes staggered over the batch, "
                                   f"auto-reset inside the step", "batch_per_gpu": B, "nsubsteps": 20,
                       "solver_iters": args.solver_iters, "noslip_iters": 3,
                       "warmstart": "force cache" if args.warmstart == 2 else "qacc map", "parallelism": f"env-sharded x{world}",
                       "kernel_shape": "1 block/SM of 16 warps in lockstep: 11 own an environment slice (11.8 KB), 5 helper warps "
                                       "pull pooled narrowphase items (library defaults; AVSIM_WARPS / AVSIM_ENVW override)",
                       "l2": "flushed between timed steps (256 MiB memset, outside the event pairs)",
                       "timing": "sum of per-step CUDA event pairs on the launching stream, max over ranks"},
            "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": int(B * model.njoints * 4),
                    "d2h_bytes_per_step": int(B * model.njoints * 4 + B * 4), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH if B == 4096 else None, "peak_source": how, "kernel": "avsim_step_kernel",
                         "note": "instruction-issue / latency bound by construction (SURVEY.md 8d): 1032 algorithmic bytes per "
                                 "env-step x 4096 envs per launch; traffic = dram read+write of one launch from "
                                 "profiles/r1_step_kernel_ncu.txt (solver scratch spilling out of L2)"},
            "clocks": clocks,
            "health": {"blown_up_envs": n_bad, "contact_overflow_envs": n_ovf, "ncon_mean": ncon_mean,
                       "reward_max": rew_max, "reward_mean": rew_mean, "successes": n_succ, "wall_s": wall,
                       "preroll_steps": args.preroll, "preroll_s": preroll_s,
                       "step_ms_min": float(min(kern_ms)), "step_ms_max": float(max(kern_ms)), "env_sm_cycles": cyc_stats,
                       "solver": args.solver, "solver_stat": solver_stat},
        }
        if not args.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            n_envs, n_steps = 2 * max(cores, 4), 150
            v, dt = cpu_steps_per_sec(n_envs, n_steps, cores, args.solver_iters, warmstart=args.warmstart)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{n_envs} envs x {n_steps} env.steps at staggered episode phases, {dt:.1f} s timed, "
                                              f"fp64 CPU restatement (MuJoCo not installable offline)"}
        print(json.dumps(line))
    batch.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--solver-iters", type=int, default=8, dest="solver_iters")
    ap.add_argument("--warmstart", type=int, default=2, choices=[1, 2],
                    help="1: MuJoCo-style qacc map, 2: per-constraint force cache (profiles/r1_warmstart_accuracy.txt)")
    ap.add_argument("--solver", default="newton", choices=["newton", "pgs"])
    ap.add_argument("--newton-iters", type=int, default=0, dest="newton_iters", help="<= 0: the library default (30)")
    ap.add_argument("--newton-ls", type=int, default=0, dest="newton_ls", help="<= 0: the library default (20)")
    ap.add_argument("--newton-tol", type=float, default=0.0, dest="newton_tol", help="<= 0: the library default (3e-7)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-clocks", action="store_true", dest="no_clocks", help="diagnostic: do not poll nvidia-smi during the run")
    ap.add_argument("--no-flush", action="store_true", dest="no_flush", help="diagnostic: do not flush L2 between timed steps")
    ap.add_argument("--preroll", type=int, default=EPISODE_LEN, help="untimed steps that bring the staggered batch to steady state")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
