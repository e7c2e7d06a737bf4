"""Generate tests/golden/ik_geometry.json: the q = 0 joint axes, joint anchors and end-effector site poses of the three arms,
computed from the reference's MJCF (assets/aloha_sim.xml) by an INDEPENDENT walk of the XML -- xml.etree + scipy Rotation, written
for this script only, sharing no code with av_aloha_b200/mjcf_compile.py.

These arrays (ik_w0, ik_p0, ik_site0 of the compiled model) are what `create_fk_fn` captures from mujoco at q = 0 in the reference
(data_collection_scripts/kinematics.py:9-15) and what tools/gen_ik_golden.py feeds to the reference's own IK code, so this fixture
pins the inputs of the IK goldens against the XML itself.  MuJoCo conventions used: quat = (w, x, y, z); euler = intrinsic x-y-z
in radians (<compiler angle="radian">, default eulerseq "xyz"); a body's `childclass` sets the default class of its subtree;
nested <default> classes inherit from their parents; joint axis default (0, 0, 1), pos default (0, 0, 0).

    python tools/gen_ik_geometry_golden.py
"""
import json
import os
import xml.etree.ElementTree as ET

import numpy as np
from scipy.spatial.transform import Rotation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
XML = "/root/reference/gym_guided_vision/gym_guided_vision/assets/aloha_sim.xml"
ARMS = {   # reference gym_guided_vision/constants.py:29-55, 83-85 (the IK uses the first 6 / 7 joints)
    "left": (["left_waist", "left_shoulder", "left_elbow", "left_forearm_roll", "left_wrist_angle", "left_wrist_rotate"],
             "left_gripper_control"),
    "right": (["right_waist", "right_shoulder", "right_elbow", "right_forearm_roll", "right_wrist_angle", "right_wrist_rotate"],
              "right_gripper_control"),
    "middle": (["middle_waist", "middle_shoulder", "middle_elbow", "middle_forearm_roll", "middle_wrist_1_joint",
                "middle_wrist_2_joint", "middle_wrist_3_joint"], "middle_zed_camera_center"),
}


def vec(s, default):
    return np.array([float(x) for x in s.split()]) if s is not None else np.array(default, float)


def local_rotation(el):
    if el.get("quat") is not None:
        w, x, y, z = vec(el.get("quat"), None)
        return Rotation.from_quat([x, y, z, w]).as_matrix()
    if el.get("euler") is not None:
        return Rotation.from_euler("XYZ", vec(el.get("euler"), None)).as_matrix()      # intrinsic x-y-z
    assert not any(el.get(k) for k in ("axisangle", "xyaxes", "zaxis")), "orientation form not handled"
    return np.eye(3)


def joint_defaults(root):
    table = {}

    def walk(node, inherited):
        cur = dict(inherited)
        j = node.find("joint")
        if j is not None:
            cur.update({k: j.get(k) for k in ("axis", "pos") if j.get(k) is not None})
        if node.get("class"):
            table[node.get("class")] = cur
        for child in node.findall("default"):
            walk(child, cur)

    for top in root.findall("default"):
        walk(top, {})
    return table


def main():
    root = ET.parse(XML).getroot()
    defaults = joint_defaults(root)
    joints, sites = {}, {}

    def walk(body, R, p, childclass):
        for j in body.findall("joint"):
            d = defaults.get(j.get("class") or childclass, {})
            axis = vec(j.get("axis") or d.get("axis"), [0, 0, 1])
            pos = vec(j.get("pos") or d.get("pos"), [0, 0, 0])
            joints[j.get("name")] = (R @ (axis / np.linalg.norm(axis)), p + R @ pos)
        for s in body.findall("site"):
            T = np.eye(4)
            T[:3, :3] = R @ local_rotation(s)
            T[:3, 3] = p + R @ vec(s.get("pos"), [0, 0, 0])
            sites[s.get("name")] = T
        for b in body.findall("body"):
            walk(b, R @ local_rotation(b), p + R @ vec(b.get("pos"), [0, 0, 0]), b.get("childclass") or childclass)

    walk(root.find("worldbody"), np.eye(3), np.zeros(3), None)
    out = {}
    # explicit <inertial> elements (aloha_sim.xml:121 ff.): mass, centre of mass and the inertia tensor in the body frame
    # (diaginertia rotated by the inertial quat); non-normalised quats are normalised as MuJoCo does
    inertials = {}
    for b in root.iter("body"):
        el = b.find("inertial")
        if el is None:
            continue
        R = local_rotation(el)
        I = R @ np.diag(vec(el.get("diaginertia"), None)) @ R.T
        inertials[b.get("name")] = {"mass": float(el.get("mass")), "ipos": vec(el.get("pos"), [0, 0, 0]).tolist(),
                                    "inertia": [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]}
    out["inertials"] = inertials
    for arm, (names, site) in ARMS.items():
        out[arm] = {"joints": names, "site": site, "w0": [joints[n][0].tolist() for n in names],
                    "p0": [joints[n][1].tolist() for n in names], "site0": sites[site].tolist()}
        print(arm, "site position", np.round(sites[site][:3, 3], 5))
    dst = os.path.join(ROOT, "tests", "golden", "ik_geometry.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
