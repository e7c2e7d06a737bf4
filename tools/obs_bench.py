"""Device-resident observation path: avsim_pixels_to_float against the HBM roofline, next to the reference's
host conversion (lerobot utils.py:37-50) timed on this box's cores, and the rendered rollout with / without lazy rendering.

    python tools/obs_bench.py [n_images=4096] [H=480] [W=640] [--rollout B]

n_images = 4096 is BASELINE.json config 3's observation (B = 1024 environments x 4 cameras).  Algorithmic bytes per image:
3*H*W read + 12*H*W written.  CUDA events on the launching stream; input + output (18.9 GB at the default) are far larger
than the 126 MB L2, so consecutive launches cannot hit in cache.
"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from av_aloha_b200 import capi

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if len(args) > 0 else 4096
H = int(args[1]) if len(args) > 1 else 480
W = int(args[2]) if len(args) > 2 else 640
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]); how = "measured"
except Exception:
    peak, how = 6650.0, "fallback"
img = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
out = torch.empty((n, 3, H, W), dtype=torch.float32, device="cuda")
for _ in range(3):
    capi.pixels_to_float(img, out=out)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
for a, b in ev:
    a.record(); capi.pixels_to_float(img, out=out); b.record()
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for a, b in ev)
ms = float(np.mean(ts))
nbytes = img.numel() * 5
print(f"avsim_pixels_to_float n={n} {H}x{W}: mean {ms:.3f} ms (min {ts[0]:.3f}, max {ts[-1]:.3f}) -> {nbytes / ms / 1e6:.0f} GB/s of "
      f"{nbytes / 1e9:.2f} GB algorithmic bytes = {nbytes / ms / 1e6 / peak * 100:.1f} % of the {how} copy peak {peak:.0f} GB/s; "
      f"{n / ms * 1e3:.0f} images/s")
# parity on the bench-sized buffer through a size-independent property: sum of planes == sum of bytes / 255 per channel (fp64)
s_out = out[:64].double().sum(dim=(0, 2, 3)).cpu().numpy()
s_in = img[:64].double().sum(dim=(0, 1, 2)).cpu().numpy() / 255.0
print(f"   channel checksums (first 64 images) rel err {np.abs(s_out - s_in).max() / s_in.max():.2e}; "
      f"bit-exact vs the reference's host arithmetic (torch CPU, 8 images): "
      f"{bool(torch.equal(out[:8].cpu(), torch.from_numpy(img[:8].cpu().numpy()).permute(0, 3, 1, 2).contiguous().float().div_(255)))} "
      f"(torch's CUDA div by a scalar multiplies by the reciprocal and is NOT bit-equal to its own CPU result)")
# torch's own elementwise path on the device, same arithmetic (library baseline, not the product)
for _ in range(2):
    ref = img.permute(0, 3, 1, 2).contiguous().float().div_(255)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ref = img.permute(0, 3, 1, 2).contiguous().float().div_(255); e1.record(); torch.cuda.synchronize()
print(f"   torch on the device (permute.contiguous.float.div_): {e0.elapsed_time(e1):.3f} ms")
del ref
# the reference's host conversion on a bounded sample (utils.py:37-50), all host threads torch uses
m = min(n, 64)
host = img[:m].cpu().numpy()
t0 = time.perf_counter()
t = torch.from_numpy(host).permute(0, 3, 1, 2).contiguous().type(torch.float32); t /= 255
dt = time.perf_counter() - t0
print(f"   reference host conversion ({m} images, {torch.get_num_threads()} threads): {dt * 1e3 / m:.3f} ms/image -> {m / dt:.0f} images/s "
      f"(+ the D2H of the frames and the H2D of 4x as many fp32 bytes, which the device path does not have)")
if "--rollout" in sys.argv:
    B = int(sys.argv[sys.argv.index("--rollout") + 1])
    from collections import deque
    from av_aloha_b200 import observation
    from av_aloha_b200.env import GuidedVisionVectorEnv

    class Chunk:   # ACT-shaped stand-in: reads the observation every 50 steps (n_action_steps), no network behind it
        def __init__(self, home): self.home, self._action_queue = home, deque([], maxlen=50)
        def reset(self): self._action_queue.clear()
        def select_action(self, batch):
            if not self._action_queue:
                lum = sum(v.mean(dim=(1, 2, 3)) for k, v in batch.items() if "images" in k)
                for k in range(50):
                    a = self.home[None].repeat(len(lum), 1); a[:, 1] += 0.001 * lum * k; self._action_queue.append(a)
            return self._action_queue.popleft()
    home = torch.tensor([0, -0.082, 1.06, 0, -0.953, 0, 1] * 2 + [0, -0.8, 0.8, 0, 0.5, 0, 0], dtype=torch.float32, device="cuda")
    cams = ["zed_cam_left", "zed_cam_right", "wrist_cam_left", "wrist_cam_right"]
    for lazy in (False, True):
        env = GuidedVisionVectorEnv("slot_insertion", B, cameras=cams, max_episode_steps=100, seed=1)
        pol = Chunk(home)
        observation.rollout(env, pol, lazy_render=lazy); torch.cuda.synchronize()
        t0 = time.perf_counter(); observation.rollout(env, pol, lazy_render=lazy); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"   device rollout B={B} x 100 steps, 4 cameras 480x640, lazy_render={lazy}: {dt:.2f} s -> {B * 100 / dt:.0f} env-steps/s")
        env.close()
