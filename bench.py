#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched env.step() hot path (BASELINE.json metric).

Workload (BASELINE.json configs[1]): SlotInsertion-3Arms-v0, 4096 environments per GPU, no render.
Actions are a synthetic scripted policy that performs the task (av_aloha_b200/workload.py: reach, pinch a slot rail with the left
hand, grasp the stick with the right, lift, carry over the slot, lower into the gap; waypoints IK-solved per environment for
object placements drawn from the reference's reset ranges, + N(0, 0.01 rad) joint noise, seed 1234).  One "step" = one env.step
over the whole batch = 20 physics substeps of 2 ms + reward + agent_pos (reference gym_guided_vision/env.py:203-226), including
the auto-reset of the environments whose 300-step episode (sim_slot_insertion_3arms.yaml:17) ended on that step.  Episodes are
STAGGERED: environment e starts at script phase (e * 300) // B, so every timed step sees the whole episode's mix of free motion,
grasping and insertion and costs the episode average (an untimed 300-step pre-roll brings the batch to that steady state).

Solver: Newton on the primal run to its tolerance (3e-7 scaled gradient, what fp32 reaches) + 3 noslip sweeps -- the reference's
solver (assets/aloha_sim.xml:4-6 leaves MuJoCo's default) and the setting the `-m gpu` parity tests pass at
(tests/test_solver_newton.py).  `solver_residual` in the line reports what the timed run actually reached.  `--solver pgs`
times the round-1 mode (8 block Gauss-Seidel sweeps, force-cache warm start): a fixed-cost approximation, not the reference's answer.

Arms:
  default            : the CUDA path.  `value` = device-resident actions through the C-ABI (avsim_step);
                       `e2e` = GuidedVisionVectorEnv.step -- the gym-facing call lerobot's rollout makes -- with HOST numpy
                       actions in and the observation dict / reward / truncation flags out (H2D, D2H and the dict building
                       inside the timed region, every step).
  --impl reference   : the reference's own CPU path on this box's host cores.  If MuJoCo + dm_control + gymnasium are importable
                       (or installed under baseline/_ref) the real gym_guided_vision env runs (cpu_baseline.kind "reference");
                       they are not installable offline, so today it is the fp64 oracle ("port": a CPU restatement, never to
                       be read as "MuJoCo"), all host threads, on a bounded sample of the same workload.

Multi-GPU (torchrun, one rank per GPU): envs are independent, every rank steps its own shard with no data-path collective; one
NCCL all_gather of per-env success at the end.  --scaling weak (default): 4096 environments per GPU; --scaling strong: 4096
environments in total, split over the ranks.
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


def ncu_traffic(B):
    """DRAM bytes (read + write) of one env.step's kernels at this batch size, from the committed ncu capture of this very
    command (profiles/r2_traffic.json, written by tools/ncu_traffic.py); None when there is no capture for this B."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as fh:
            t = json.load(fh)
        return float(t["dram_bytes_per_step"]) if int(t["batch"]) == int(B) else None
    except Exception:
        return None


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


# ------------------------------------------------------------------------------------------ CPU arm
def _oracle_solver_args(args):
    """the oracle runs the solver the GPU arm runs: Newton to 1e-10 (the reference's: MuJoCo Newton, tolerance 1e-8) or the
    fixed-sweep PGS of --solver pgs"""
    if args.solver == "newton":
        return dict(solver="newton", max_iter=100, tol=1e-10, warmstart=1)
    return dict(solver="pgs", max_iter=args.solver_iters, tol=0.0, warmstart=args.warmstart)


def cpu_steps_per_sec(n_envs, n_steps, threads, args, seed=1234):
    """The CPU restatement (fp64 oracle) on `threads` host threads: a sample of n_envs environments of the bench workload, their
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
    sa = _oracle_solver_args(args)
    envs = []
    for e in range(n_envs):
        o = OracleEnv(om)
        o.set_solver(sa["solver"])
        o.set_options(max_iter=sa["max_iter"], tol=sa["tol"], warmstart=sa["warmstart"])
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


def reference_env_available():
    """BASELINE.md section 3 probe: can the REAL reference env run here?  It needs mujoco + dm_control + gymnasium (third-party,
    no wheel in /opt/wheelhouse) and the gym_guided_vision package (pip-installed under baseline/_ref when the offline install
    works).  Returns the imported module or None; never raises."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import dm_control  # noqa: F401
        import gymnasium  # noqa: F401
        import mujoco  # noqa: F401
        import gym_guided_vision
        return gym_guided_vision
    except Exception:
        return None


def _ref_env_worker(job):
    """one process = one serial SyncVectorEnv-style loop over its share of reference environments (that is how the reference
    steps a batch: lerobot/common/envs/factory.py:50-56, use_async_envs false)"""
    n_envs, n_steps, phases, seed = job
    os.environ.setdefault("MUJOCO_GL", "egl")
    import gymnasium as gym
    import gym_guided_vision  # noqa: F401  (registers the ids)

    from av_aloha_b200 import workload
    obj = workload.sample_object_positions(len(phases), seed)
    acts = workload.slot_insertion_script(EPISODE_LEN, obj, seed)
    envs = [gym.make("gym_guided_vision/SlotInsertion-3Arms-v0", cameras=[], disable_env_checker=True) for _ in phases]
    for e, env in enumerate(envs):
        np.random.seed(seed + e)
        env.reset()
        for t in range(int(phases[e])):
            env.step(acts[t, e])
    t0 = time.perf_counter()
    for k in range(n_steps):
        for e, env in enumerate(envs):
            t = (int(phases[e]) + k) % EPISODE_LEN
            if t == 0:
                env.reset()
            env.step(acts[t, e])
    return time.perf_counter() - t0


def reference_env_steps_per_sec(n_envs, n_steps, procs, seed=1234):
    """the real gym_guided_vision env (MuJoCo Newton) on `procs` processes, the same bounded sample as the port"""
    import multiprocessing as mp
    phase = (np.arange(n_envs) * EPISODE_LEN) // n_envs
    shares = [phase[i::procs] for i in range(procs)]
    with mp.get_context("spawn").Pool(procs) as pool:
        times = pool.map(_ref_env_worker, [(n_envs, n_steps, sh, seed) for sh in shares if len(sh)])
    dt = max(times)
    return n_envs * n_steps / dt, dt


def cpu_baseline(args, n_envs, n_steps, cores):
    """(value, description dict) of the CPU arm: the real reference env when it is importable, else the fp64 restatement"""
    if reference_env_available() is not None:
        v, dt = reference_env_steps_per_sec(n_envs, n_steps, cores)
        return v, {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "reference",
                   "sample": f"{n_envs} gym_guided_vision SlotInsertion-3Arms envs (MuJoCo, cameras=[]) x {n_steps} env.steps at staggered "
                             f"episode phases on {cores} processes, {dt:.1f} s timed"}
    v, dt = cpu_steps_per_sec(n_envs, n_steps, cores, args)
    how = "Newton to 1e-10 + 3 noslip" if args.solver == "newton" else f"{args.solver_iters} PGS sweeps + 3 noslip"
    return v, {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
               "sample": f"{n_envs} envs x {n_steps} env.steps at staggered episode phases, {dt:.1f} s timed, fp64 CPU restatement of "
                         f"the pipeline ({how}; MuJoCo / dm_control are not installable offline: not a MuJoCo number)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_envs = 2 * max(cores, 4)
    total = args.warmup + args.steps
    # each "step" of this arm = one env.step of a bounded sample of `n_envs` environments on all host cores
    value, desc = cpu_baseline(args, n_envs, total, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_envs / value,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"SlotInsertion-3Arms-v0 no render, scripted grasp/insert policy, sample of {n_envs} envs "
                               f"x {total} consecutive env.steps, episode phases spread over the 300-step script "
                               f"(of the B=4096 staggered workload)",
                   "solver": args.solver, "nsubsteps": 20},
        "cpu_baseline": desc,
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from av_aloha_b200 import capi, env as avenv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # weak scaling: every rank steps `batch` environments; strong: `batch` environments in total, rank r owns a contiguous shard
    if args.scaling == "strong":
        from av_aloha_b200.sharding import shard_range
        lo, hi = shard_range(args.batch, rank, world)
        B, seed = hi - lo, 1234
    else:
        lo, B, seed = 0, args.batch, 1234 + rank
    total_envs = args.batch if args.scaling == "strong" else args.batch * world
    obj, acts_np, masks_np, phase = make_workload(args.batch if args.scaling == "strong" else B, seed)
    if args.scaling == "strong":
        obj, acts_np, masks_np, phase = obj[lo:lo + B], np.ascontiguousarray(acts_np[:, lo:lo + B]), np.ascontiguousarray(masks_np[:, lo:lo + B]), phase[lo:lo + B]
    # the gym-facing environment (what lerobot's rollout holds); the device arm drives its batch through the C-ABI directly
    venv = avenv.GuidedVisionVectorEnv(TASK, B, num_arms=ARMS, cameras=[], max_episode_steps=EPISODE_LEN, device=local,
                                       solver=args.solver, solver_iterations=args.solver_iters, warmstart=args.warmstart,
                                       seed=seed, free_pos=obj, episode_phase=phase)
    batch, model = venv._batch, venv._model
    if args.solver == "newton":
        batch.set_solver("newton", args.newton_iters, args.newton_ls, args.newton_tol)
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
    nw_acc = torch.zeros(4, dtype=torch.float64, device=dev)
    nw_max = torch.zeros(2, dtype=torch.float64, device=dev)
    barrier()
    w0 = time.perf_counter()
    for k in range(args.steps):
        if not args.no_flush:
            flush.zero_()
        t = advance()
        ev[k][0].record()
        dev_step(t)
        ev[k][1].record()
        st = batch.get(capi.SOLVER_STAT).double()                     # outside the event pair: what the timed solves reached
        nw_acc += torch.stack([st[:, 0].sum(), st[:, 3].sum(), st[:, 1].median(), (st[:, 1] > 1e-5).double().sum()])
        nw_max = torch.maximum(nw_max, torch.stack([st[:, 1].max(), st[:, 2].max()]))
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
    nsolves = B * 20.0 * args.steps
    solver_residual = {
        "what": "Newton solves inside the timed region (every environment, every substep): iterations and the scaled gradient "
                "|M acc - J'f| / trace(M) each solve ended with (stop: 3e-7, or no further decrease in fp32)",
        "iterations_mean": float(nw_acc[0].item()) / nsolves, "iterations_one_solve_max": float(nw_max[1].item()),
        "scaled_gradient_median_of_step_max": float(nw_acc[2].item()) / args.steps, "scaled_gradient_max": float(nw_max[0].item()),
        "env_steps_with_a_solve_above_1e-5": float(nw_acc[3].item()) / (B * args.steps),
        "solves_at_iteration_cap": float(nw_acc[1].item()),
        "parity_test": "tests/test_solver_newton.py::test_gpu_solver_parity_on_contact_states (|dqacc|inf <= 1e-2 max(1,|qacc|inf) vs the fp64 oracle's Newton at 1e-13, same contacts)",
    } if args.solver == "newton" else {"what": f"{args.solver_iters} PGS sweeps, not converged (profiles/r1_pgs_sweep.txt: median rel |dqacc| ~ 1 on grasp states)"}

    # ---- end-to-end arm: GuidedVisionVectorEnv.step with host numpy actions; observation dict, rewards, flags, auto-reset and
    # final_info bookkeeping all inside the timed region
    venv._elapsed[:] = (state["t"] + phase) % EPISODE_LEN          # the wrapper's per-env step counters, in sync with the script
    for _ in range(min(args.warmup, 3)):
        venv.step(acts_np[advance()])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    chk = 0.0
    for _ in range(args.steps):
        obs, reward, terminated, truncated, info = venv.step(acts_np[advance()])
        chk += float(reward.sum()) + float(obs["agent_pos"][0, 0])
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
        if args.scaling == "strong":                                  # shards may differ by one environment: pad to the largest
            nmax = torch.tensor([B], device=dev)
            dist.all_reduce(nmax, op=dist.ReduceOp.MAX)
            succ = torch.nn.functional.pad(succ, (0, int(nmax.item()) - B))
        gathered = [torch.empty_like(succ) for _ in range(world)]
        dist.all_gather(gathered, succ)
        succ = torch.cat(gathered)
    n_succ = int(succ.sum().item())

    if rank == 0:
        value = total_envs * args.steps / (total_ms * 1e-3)
        e2e = total_envs * args.steps / (e2e_ms * 1e-3)
        peak, how = peaks()
        kern_avg_ms = total_ms / args.steps
        achieved = ALGO_BYTES_PER_ENV_STEP * B / (kern_avg_ms * 1e-3) / 1e9
        shape = batch.launch_shape
        split = bool(shape["split"])
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": kern_avg_ms, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"SlotInsertion-3Arms-v0 batch={B}/GPU ({total_envs} in total), no render, scripted grasp/lift/insert "
                                   f"policy (IK-solved joint targets + N(0,0.01) noise), 300-step episodes staggered over the batch, "
                                   f"auto-reset inside the step", "batch_per_gpu": B, "batch_total": total_envs, "nsubsteps": 20,
                       "solver": "Newton (primal), scaled-gradient tolerance 3e-7, + 3 noslip sweeps" if args.solver == "newton"
                                 else f"PGS {args.solver_iters} sweeps + 3 noslip, warm start {args.warmstart}",
                       "parallelism": f"env-sharded x{world}",
                       "kernel_shape": ("split pipeline per substep: avsim_substep_kernel (1 block/SM of 16 lockstep warps: 11 environment "
                                        "slices + 5 narrowphase helpers) -> avsim_solve_kernel (8 phase-locked warps per block, 1 environment each); "
                                        f"{shape['groups']} environment groups on separate streams; head records moved by cp.async.bulk") if split else
                                       f"fused step kernel, 1 block/SM of {shape['warps']} lockstep warps ({shape['env_warps']} environment slices)",
                       "launch_shape": shape,
                       "l2": "flushed between timed steps (256 MiB memset, outside the event pairs)",
                       "timing": "sum of per-step CUDA event pairs on the launching stream, max over ranks"},
            "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": int(B * model.njoints * 4),
                    "d2h_bytes_per_step": int(B * model.njoints * 4 + 2 * B * 4), "ms_per_step": e2e_ms / args.steps,
                    "api": "GuidedVisionVectorEnv.step(numpy actions) -> (obs dict, reward, terminated, truncated, info)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(B), "peak_source": how,
                         "kernel": "avsim_substep_kernel + avsim_solve_kernel (one env.step = 41 launches per environment group)" if split else "avsim_step_kernel",
                         "note": "instruction-issue / latency bound by construction (SURVEY.md 8d): 1032 algorithmic bytes per "
                                 "env-step; traffic = DRAM read+write of one env.step's kernels from the ncu capture of this command "
                                 "(profiles/r2_traffic.json), null when no capture matches this batch size"},
            "solver_residual": solver_residual,
            "clocks": clocks,
            "health": {"blown_up_envs": n_bad, "blowup_resets_e2e": venv.blowup_resets, "contact_overflow_envs": n_ovf, "ncon_mean": ncon_mean,
                       "reward_max": rew_max, "reward_mean": rew_mean, "successes": n_succ, "wall_s": wall,
                       "preroll_steps": args.preroll, "preroll_s": preroll_s, "e2e_checksum": chk,
                       "step_ms_min": float(min(kern_ms)), "step_ms_max": float(max(kern_ms)), "env_sm_cycles": cyc_stats},
        }
        if not args.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            n_envs, n_steps = 2 * max(cores, 4), 100
            _, line["cpu_baseline"] = cpu_baseline(args, n_envs, n_steps, cores)
    venv.close()
    del venv, batch
    # ---- BASELINE.json configs[4] as an extra object of the same line (all ranks take part): one policy rollout of 1024
    # environments in total, sharded over the ranks, 4 cameras at 480 x 640 rendered lazily, device-resident observations, one
    # all_gather of the episode results -- so the driver's 1..8-GPU runs carry the strong-scaling curve of the evaluation loop too
    c5 = None
    if not args.no_config5:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import eval_rollout
            c5 = eval_rollout.run(world, rank, dev, batch_total=1024, n_cameras=4, steps=EPISODE_LEN, rollouts=1, lazy=True,
                                  warm_rollout_steps=None)
        except Exception as e:   # the headline line must survive a failure of the extra leg
            c5 = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        if c5 is not None:
            line["config5"] = c5
        print(json.dumps(line))
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch environments per GPU; strong: --batch environments in total, sharded over the ranks")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-config5", action="store_true", dest="no_config5", help="skip the extra config-5 rollout leg")
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
