"""Config 3 (BASELINE.json): SewNeedle-3Arms, B = 1024, zed L/R + wrist L/R at 480x640 -- renderer throughput against the
HBM-write roofline, and a sample frame dump.  python tools/render_bench.py [B] [outdir]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from av_aloha_b200 import capi, model_io

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
out = sys.argv[2] if len(sys.argv) > 2 else None
task = "sew_needle"
model = capi.Model(model_io.model_path(task, 3), 0)
cams = model_io.load_names(task, 3)["camera"]
names = ["zed_cam_left", "zed_cam_right", "wrist_cam_left", "wrist_cam_right"]
ids = [cams.index(c) for c in names]
b = capi.Batch(model, B, seed=3)
b.reset()
HOME = np.array([0, -0.082, 1.06, 0, -0.953, 0, 1] * 2 + [0, -0.8, 0.8, 0, 0.5, 0, 0], np.float32)
act = torch.as_tensor(np.tile(HOME, (B, 1)), device="cuda")
for _ in range(3):
    b.step(act)
H, W = 480, 640
img = torch.empty((B, 4, H, W, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    b.render(ids, H, W, out=img)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    flush.zero_()
    e0.record(); b.render(ids, H, W, out=img); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
nbytes = img.numel()
print(f"render B={B} x 4 cams x {H}x{W}: {ms:.2f} ms -> {B / ms * 1e3:.0f} env-frames/s, {nbytes / ms / 1e6:.0f} GB/s of image writes "
      f"({nbytes / 1e9:.2f} GB per call; HBM copy peak 6650 GB/s fallback)")
if out:
    import cv2
    os.makedirs(out, exist_ok=True)
    frame = img[0].cpu().numpy()
    for k, n in enumerate(names):
        cv2.imwrite(os.path.join(out, f"render_{n}.png"), frame[k][:, :, ::-1])
    ov = b.render([cams.index("overhead_cam")], 225, 300)[0, 0].cpu().numpy()
    cv2.imwrite(os.path.join(out, "render_overhead.png"), ov[:, :, ::-1])
