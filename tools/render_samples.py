"""Sample frames of the renderer for profiles/ (config 3's four cameras + the overhead view): python tools/render_samples.py outdir"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cv2
import numpy as np
import torch
from av_aloha_b200 import capi, model_io, workload
import bench

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
os.makedirs(out, exist_ok=True)
task = "slot_insertion"
model = capi.Model(model_io.model_path(task, 3), 0)
cams = model_io.load_names(task, 3)["camera"]
B = 4
obj = workload.sample_object_positions(B, 1234)
acts = workload.slot_insertion_script(bench.EPISODE_LEN, obj, 1234)
b = capi.Batch(model, B, seed=1234)
b.reset(free_pos=obj)
for t in range(190):                       # mid-grasp / lift
    b.step(torch.as_tensor(acts[t], device="cuda"))
for name in ("zed_cam_left", "wrist_cam_left", "overhead_cam", "worms_eye_cam"):
    img = b.render([cams.index(name)], 480, 640)[0, 0].cpu().numpy()
    cv2.imwrite(os.path.join(out, f"r2_render_{name}.png"), cv2.resize(img[:, :, ::-1], (320, 240), interpolation=cv2.INTER_AREA))
print("wrote", out)
