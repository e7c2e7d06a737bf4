"""Time block-shape (warps per lockstep block : blocks per SM) variants of the step kernel on the steady-state bench workload:
   python tools/variant_bench.py [B] [iters] [warps:blocks[:sync[:envwarps]],...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import steady
from av_aloha_b200 import capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
variants = [tuple(int(x) for x in v.split(":")) for v in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["16:1:2:11", "16:1:2:13", "13:1:2:13"])]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for v in variants:
    nw, nb = v[0], v[1]
    sync = v[2] if len(v) > 2 else 2
    envw = v[3] if len(v) > 3 else min(nw, 11)      # warps that own an environment slice; the rest only help the pooled narrowphase
    os.environ["AVSIM_WARPS"], os.environ["AVSIM_BLOCKS"], os.environ["AVSIM_SYNC"] = str(nw), str(nb), str(sync)
    os.environ["AVSIM_ENVW"] = str(envw)
    model, batch, acts, masks, mask_any, fp, t0 = steady.restore(B, iters)
    ts = []
    for k in range(5):
        e0.record(); steady.step(batch, acts, masks, mask_any, fp, t0 + k); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts[2:]) / 3
    print(f"warps/block={nw} (env {envw}) blocks/SM={nb} sync={sync}: {ms:7.1f} ms/step -> {B / ms * 1e3:8.0f} env-steps/s  (first {ts[0]:.0f} ms)  ncon {batch.get(capi.NCON).float().mean().item():.1f}", flush=True)
    cyc = batch.get(capi.ENV_CYCLES).double()
    print(f"      per-env SM cycles of the last step: mean {cyc.mean().item():.3e}  p50 {cyc.median().item():.3e}  p90 {cyc.quantile(0.9).item():.3e} "
          f"p99 {cyc.quantile(0.99).item():.3e}  max {cyc.max().item():.3e}  (kernel {ms * 1.965e6:.3e} cycles)", flush=True)
    batch.close()
