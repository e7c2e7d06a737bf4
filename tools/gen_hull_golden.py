"""Generate tests/golden/hull_support.json: support values h(d) = max_v d.v and h(-d) of every collision mesh for 48 fixed
directions, computed straight from the reference's STL files (all vertices, scaled as the <mesh> element says) by a reader written
for this script only.  The support function of a vertex cloud is that of its convex hull, so this pins the compiled hulls (STL
parsing, scale, hull construction, the bounded-error thinning of the two very fine hulls) without sharing code with the model
compiler.  The compiler stores hull vertices relative to the hull's volume centroid and moves the geom origin there (as MuJoCo
recentres meshes), so the test compares the translation-invariant WIDTHS h(d) + h(-d) for every hull, and the absolute support
through the compiled geom pose for geoms whose XML pose it restates.

    python tools/gen_hull_golden.py
"""
import json
import os
import struct
import sys
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from av_aloha_b200 import model_io  # noqa: E402  (only for the list of hull names to cover)

ASSETS = "/root/reference/gym_guided_vision/gym_guided_vision/assets"


def directions(n=48):
    """Fibonacci sphere; the test rebuilds the same directions"""
    k = np.arange(n) + 0.5
    phi, z = np.pi * (1 + 5 ** 0.5) * k, 1 - 2 * k / n
    r = np.sqrt(1 - z * z)
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)


def read_stl(path):
    raw = open(path, "rb").read()
    ntri = struct.unpack_from("<I", raw, 80)[0]
    if 84 + 50 * ntri == len(raw):                              # binary
        rec = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), offset=84, count=ntri)
        return rec["v"].reshape(-1, 3).astype(np.float64)
    verts = [[float(x) for x in ln.split()[1:4]] for ln in raw.decode("ascii", "ignore").splitlines() if ln.strip().startswith("vertex")]
    return np.array(verts, np.float64)


def main():
    meshes = {}
    for xml in ("aloha_sim.xml", "scene.xml"):
        for m in ET.parse(os.path.join(ASSETS, xml)).getroot().iter("mesh"):
            f = m.get("file")
            if f and f.endswith(".stl"):
                name = m.get("name") or os.path.splitext(os.path.basename(f))[0]
                scale = np.array([float(x) for x in m.get("scale", "1 1 1").split()])
                meshes[name] = (os.path.join(ASSETS, "meshes", f), scale)
    D = directions()
    out = {}
    for hull in model_io.load_names("slot_insertion", 3)["hull"]:
        path, scale = meshes[hull]
        v = read_stl(path) * scale
        out[hull] = {"nvert_stl": int(len(v)), "support": (v @ D.T).max(axis=0).tolist(), "support_neg": (-(v @ D.T).min(axis=0)).tolist()}
        print(f"{hull:32s} {len(v):7d} STL vertices, extent {np.ptp(v, axis=0).round(4)}")
    dst = os.path.join(ROOT, "tests", "golden", "hull_support.json")
    with open(dst, "w") as fh:
        json.dump(out, fh)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
