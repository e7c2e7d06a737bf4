#!/usr/bin/env python
"""Compile the ten AV-ALOHA models (5 tasks x {2,3} arms) from the reference's MJCF assets.

Usage:  python tools/compile_models.py [/path/to/gym_guided_vision/assets]

Runs in the build container (where /root/reference is mounted); the compiled ``.avm`` tables and the
``.json`` name sidecars under av_aloha_b200/models/ are committed because the GPU box has no copy of
the reference tree.  They hold derived constants only (kinematic tables, convex hulls), no XML.
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from av_aloha_b200 import mjcf_compile, model_io  # noqa: E402

DEFAULT_ASSETS = "/root/reference/gym_guided_vision/gym_guided_vision/assets"


def main():
    assets = sys.argv[1] if len(sys.argv) > 1 else DEFAULT_ASSETS
    os.makedirs(model_io.MODEL_DIR, exist_ok=True)
    for task in mjcf_compile.TASKS:
        for arms in (2, 3):
            m, names = mjcf_compile.compile_task(assets, task, arms)
            model_io.save_avm(model_io.model_path(task, arms), m)
            with open(os.path.join(model_io.MODEL_DIR, f"{task}_{arms}arms.json"), "w") as fh:
                json.dump(names, fh, indent=0)
            print(f"{task}-{arms}arms: nbody={len(m['body_parent'])} nq={len(m['qpos0'])} nv={len(m['dof_body'])} "
                  f"ngeom={len(m['geom_type'])} npair={len(m['pair_geom'])} nhullvert={len(m['hull_vert'])}")


if __name__ == "__main__":
    main()
