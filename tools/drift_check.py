import os, sys
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0,'/root/repo')
import torch, steady
from av_aloha_b200 import capi
model, batch, acts, masks, mask_any, fp, t0 = steady.restore(4096, 8)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
for k in range(260):
    e0.record(); steady.step(batch, acts, masks, mask_any, fp, t0 + k); e1.record(); torch.cuda.synchronize()
    if k%20==2:
        cyc=batch.get(capi.ENV_CYCLES).double(); nc=batch.get(capi.NCON).float(); rw=batch.get(capi.REWARD).float()
        print(f"k={k} {e0.elapsed_time(e1):.1f} ms  cyc mean {cyc.mean().item():.3e} max {cyc.max().item():.3e}  ncon mean {nc.mean().item():.2f} max {int(nc.max().item())} (>=24: {(nc>=24).sum().item()})  reward mean {rw.mean().item():.3f} (r>=1: {(rw>=1).sum().item()})", flush=True)
