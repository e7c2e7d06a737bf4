"""What makes the costliest environments costly: contacts of the top-k environments by intrinsic cycles."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch, steady
from av_aloha_b200 import capi, model_io
model, batch, acts, masks, mask_any, fp, t0 = steady.restore(4096, 8)
for k in range(6):
    steady.step(batch, acts, masks, mask_any, fp, t0 + k)
cyc = batch.get(capi.ENV_CYCLES).cpu().numpy().astype(float)
ncon = batch.get(capi.NCON).cpu().numpy()
con = batch.get(capi.CONTACTS).cpu().numpy()
names = model_io.load_names("slot_insertion", 3)["geom"]
avm = model_io.load_avm(model_io.model_path("slot_insertion", 3))
hn = avm["hull_num"]; gh = avm["geom_hull"]
print("cycles: mean %.2e p50 %.2e p90 %.2e p99 %.2e max %.2e" % (cyc.mean(), np.median(cyc), np.quantile(cyc, .9), np.quantile(cyc, .99), cyc.max()))
print("corr(cycles, ncon) = %.2f" % np.corrcoef(cyc, ncon)[0, 1])
top = np.argsort(-cyc)[:8]
for e in top:
    c = con[e][: ncon[e]]
    pairs = collections.Counter()
    for r in c:
        g1, g2 = int(r[7]), int(r[8])
        d = lambda g: f"{names[g] or 'hull'}#{g}({hn[gh[g]] if gh[g] >= 0 else '-'}v)"
        pairs[(d(g1), d(g2))] += 1
    print(f"env {e}: cycles {cyc[e]:.2e} ncon {ncon[e]} phase {(e*300)//4096}: " + ", ".join(f"{a}-{b} x{n}" for (a, b), n in pairs.most_common(8)))
