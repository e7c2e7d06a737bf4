"""MJCF-subset model compiler: XML + meshes -> flat constant tables.

The reference builds its model with ``mjcf.from_path`` + ``Physics.from_mjcf_model``
(reference gym_guided_vision/gym_guided_vision/env.py:53-56).  MuJoCo is not
installable here, so this module restates the part of the MJCF compile step
that the AV-ALOHA assets actually use (reference assets/aloha_sim.xml,
scene.xml, joint_position_actuators.xml, task_*.xml) and emits a dict of numpy
arrays which ``model_io`` packs into the ``.avm`` file the C-ABI library and
the CPU oracle both load.

Handled: <include>, <compiler angle=radian meshdir autolimits>, nested
<default class> + childclass, quat/euler(xyz)/xyaxes orientations, <inertial>,
geoms box/sphere/cylinder/mesh with mass|density, friction (1..3 values),
solref/solimp/condim/gap/group/contype/conaffinity, mesh scale, <exclude>,
<equality><joint polycoef>, <position kp kv ctrlrange>, joint armature /
frictionloss / damping / range / actuatorfrcrange, free joints, cameras,
sites.  Mesh geoms collide as the convex hull of their vertices
(scipy.spatial.ConvexHull; MuJoCo uses qhull as well [upstream]).
"""
from __future__ import annotations

import copy
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

# joint / geom type codes (follow the MuJoCo enum order so dumps read familiar)
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
GEOM_SPHERE, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = 2, 5, 6, 7
_GEOM_TYPES = {"sphere": GEOM_SPHERE, "cylinder": GEOM_CYLINDER, "box": GEOM_BOX, "mesh": GEOM_MESH}

TASKS = {
    # task name -> (xml file, task id, max_reward)   (reference env.py:412-423,503-511,592-602,693-702,782-790)
    "insert_peg": ("task_insert_peg.xml", 0, 4),
    "slot_insertion": ("task_slot_insertion.xml", 1, 4),
    "sew_needle": ("task_sew_needle.xml", 2, 5),
    "tube_transfer": ("task_tube_transfer.xml", 3, 3),
    "hook_package": ("task_hook_package.xml", 4, 4),
}

# built-in MuJoCo defaults [upstream], only the attributes the assets rely on
_BUILTIN = {
    "geom": dict(type="sphere", contype="1", conaffinity="1", condim="3", group="0",
                 friction="1 0.005 0.0001", solref="0.02 1", solimp="0.9 0.95 0.001 0.5 2",
                 margin="0", gap="0", density="1000", rgba="0.5 0.5 0.5 1"),
    "joint": dict(type="hinge", axis="0 0 1", pos="0 0 0", armature="0", damping="0", frictionloss="0",
                  solreflimit="0.02 1", solimplimit="0.9 0.95 0.001 0.5 2",
                  solreffriction="0.02 1", solimpfriction="0.9 0.95 0.001 0.5 2", margin="0"),
    "position": dict(kp="1", kv="0"),
    "site": dict(group="0"),
    "camera": dict(fovy="45"),
    "mesh": dict(scale="1 1 1"),
}


# ----------------------------------------------------------------------------- small math
def _f(s, n=None):
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None:
        assert v.size == n, (s, n)
    return v


def quat_mul(a, b):
    w1, x1, y1, z1 = a
    w2, x2, y2, z2 = b
    return np.array([
        w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
        w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
        w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
        w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
    ])


def quat2mat(q):
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def mat2quat(R):
    # Shepperd's method, w >= 0
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    if q[0] < 0:
        q = -q
    return q / np.linalg.norm(q)


def euler2quat(e):
    # intrinsic xyz (MuJoCo default eulerseq "xyz") [upstream]
    q = np.array([1.0, 0, 0, 0])
    for ax, ang in enumerate(e):
        r = np.zeros(4)
        r[0] = np.cos(ang / 2)
        r[1 + ax] = np.sin(ang / 2)
        q = quat_mul(q, r)
    return q


def _orientation(attrib):
    if "quat" in attrib:
        q = _f(attrib["quat"], 4)
        return q / np.linalg.norm(q)
    if "euler" in attrib:
        return euler2quat(_f(attrib["euler"], 3))
    if "xyaxes" in attrib:
        v = _f(attrib["xyaxes"], 6)
        x = v[:3] / np.linalg.norm(v[:3])
        y = v[3:] - x * np.dot(x, v[3:])
        y /= np.linalg.norm(y)
        z = np.cross(x, y)
        return mat2quat(np.stack([x, y, z], axis=1))
    return np.array([1.0, 0, 0, 0])


# ----------------------------------------------------------------------------- XML front end
def _load_xml(path, seen=None):
    """Parse ``path`` and splice <include> children in place (recursively)."""
    tree = ET.parse(path)
    root = tree.getroot()
    base = os.path.dirname(path)

    def expand(node):
        out = []
        for ch in list(node):
            if ch.tag == "include":
                sub = _load_xml(os.path.join(base, ch.attrib["file"]))
                out.extend(list(sub))
            else:
                expand(ch)
                out.append(ch)
        node[:] = out

    expand(root)
    return root


class _Defaults:
    """Nested <default class=...> tables; lookup(cls, tag) -> merged attribute dict."""

    def __init__(self):
        self.table = {"main": {}}
        self.parent = {"main": None}

    def ingest(self, node, cls="main", parent=None):
        if cls not in self.table:
            self.table[cls] = {}
            self.parent[cls] = parent
        for ch in node:
            if ch.tag == "default":
                self.ingest(ch, ch.attrib["class"], cls)
            else:
                self.table[cls].setdefault(ch.tag, {}).update(ch.attrib)

    def lookup(self, cls, tag):
        chain = []
        c = cls
        while c is not None:
            chain.append(c)
            c = self.parent[c]
        out = dict(_BUILTIN.get(tag, {}))
        for c in reversed(chain):
            out.update(self.table[c].get(tag, {}))
        return out


# ----------------------------------------------------------------------------- meshes
def _read_stl(path):
    with open(path, "rb") as fh:
        data = fh.read()
    ntri = struct.unpack_from("<I", data, 80)[0]
    assert len(data) == 84 + 50 * ntri, f"{path}: not a binary STL"
    rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", 9), ("a", "<u2")]), count=ntri, offset=84)
    return rec["v"].reshape(-1, 3).astype(np.float64)


def _read_obj(path):
    verts = []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                verts.append([float(x) for x in line.split()[1:4]])
    return np.array(verts, dtype=np.float64)


# the 13 slab directions of a 26-DOP (3 axes, 6 face diagonals, 4 body diagonals): what the renderer intersects a ray with for a
# mesh geom -- a convex polytope around the hull with 26 faces instead of the 6 of its bounding box
KDOP_DIRS = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, -1, 0], [1, 0, 1], [1, 0, -1], [0, 1, 1], [0, 1, -1],
                      [1, 1, 1], [1, 1, -1], [1, -1, 1], [1, -1, -1]], np.float64)
KDOP_DIRS = KDOP_DIRS / np.linalg.norm(KDOP_DIRS, axis=1, keepdims=True)


def _hull_of(verts):
    """Convex hull vertices + interior reference point (volume centroid of the hull)."""
    from scipy.spatial import ConvexHull

    uniq = np.unique(np.round(verts, 9), axis=0)
    hull = ConvexHull(uniq)
    hv = uniq[hull.vertices]
    # volume centroid via tetrahedra fan from an interior point
    c0 = hv.mean(axis=0)
    tri = uniq[hull.simplices] - c0
    vol = np.abs(np.einsum("ij,ij->i", tri[:, 0], np.cross(tri[:, 1], tri[:, 2]))) / 6.0
    cen = c0 + (vol[:, None] * tri.sum(axis=1) / 4.0).sum(axis=0) / vol.sum()
    if len(hv) > HULL_MAX_VERTS:
        hv = _thin_hull(hv, cen)
    return hv, cen


HULL_MAX_VERTS = 512


def _fibonacci_sphere(n):
    k = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * k / n)
    th = np.pi * (1 + 5 ** 0.5) * k
    return np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1)


def _thin_hull(hv, cen):
    """Bounded-error simplification of the two very fine hulls (wrist-mount 1 905 and D405 camera 7 003 vertices, SURVEY.md
    Appendix C): keep the vertices that are extreme in one of 642 evenly spread directions.  The support function of the
    kept set differs from the exact one by at most r * (1 - cos(direction spacing)) ~ 0.3 % of the hull radius
    (~ 0.15 mm here) while the narrowphase scan shrinks 3-12x; both the CUDA path and the oracle use the thinned hull."""
    d = _fibonacci_sphere(642)
    keep = np.unique(np.argmax((hv - cen) @ d.T, axis=0))
    return hv[np.sort(keep)]


# ----------------------------------------------------------------------------- inertia helpers
def _geom_mass_inertia(gtype, size, attrib, mesh_note=""):
    """(mass, inertia 3x3 in the geom frame about its centre) for a primitive geom."""
    if gtype == GEOM_BOX:
        vol = 8 * size[0] * size[1] * size[2]
        unit = np.array([size[1] ** 2 + size[2] ** 2, size[0] ** 2 + size[2] ** 2, size[0] ** 2 + size[1] ** 2]) / 3.0
    elif gtype == GEOM_SPHERE:
        vol = 4.0 / 3.0 * np.pi * size[0] ** 3
        unit = np.full(3, 0.4 * size[0] ** 2)
    elif gtype == GEOM_CYLINDER:
        r, h = size[0], size[1]
        vol = np.pi * r * r * 2 * h
        unit = np.array([(3 * r * r + (2 * h) ** 2) / 12.0, (3 * r * r + (2 * h) ** 2) / 12.0, r * r / 2.0])
    else:
        raise NotImplementedError("mass from mesh geoms is not needed by the AV-ALOHA assets " + mesh_note)
    mass = float(attrib["mass"]) if "mass" in attrib else float(attrib["density"]) * vol
    return mass, np.diag(unit * mass)


# ----------------------------------------------------------------------------- compiler
def compile_task(asset_dir, task, num_arms=3, timestep=0.002):
    """Compile one task XML into flat tables.  Returns dict[str, np.ndarray|scalar]."""
    xml, task_id, max_reward = TASKS[task]
    root = _load_xml(os.path.join(asset_dir, xml))

    meshdir = "meshes"
    for c in root.iter("compiler"):
        meshdir = c.attrib.get("meshdir", meshdir)
        assert c.attrib.get("angle", "radian") == "radian"
    opt = dict(noslip_iterations=0, impratio=1.0, cone="pyramidal", multiccd=0)
    for o in root.iter("option"):
        opt["noslip_iterations"] = int(o.attrib.get("noslip_iterations", opt["noslip_iterations"]))
        opt["impratio"] = float(o.attrib.get("impratio", opt["impratio"]))
        opt["cone"] = o.attrib.get("cone", opt["cone"])
        for fl in o.iter("flag"):
            opt["multiccd"] = int(fl.attrib.get("multiccd", "disable") == "enable")
    assert opt["cone"] == "elliptic"

    defaults = _Defaults()
    for d in root.findall("default"):
        defaults.ingest(d)

    # ---- materials (render colours; textures are not reproduced: the table's wood texture becomes one flat colour)
    materials = {}
    for a in root.findall("asset"):
        for el in a.findall("material"):
            materials[el.attrib["name"]] = _f(el.attrib["rgba"], 4) if "rgba" in el.attrib else np.array([0.62, 0.50, 0.38, 1.0])

    # ---- mesh assets
    mesh_assets = {}
    for a in root.findall("asset"):
        for m in a.findall("mesh"):
            at = defaults.lookup("main", "mesh")
            at.update(m.attrib)
            name = at.get("name", os.path.splitext(os.path.basename(at["file"]))[0])
            mesh_assets[name] = (os.path.join(asset_dir, meshdir, at["file"]), _f(at["scale"], 3))
    hull_cache = {}

    def hull(name):
        if name not in hull_cache:
            path, scale = mesh_assets[name]
            v = _read_stl(path) if path.lower().endswith(".stl") else _read_obj(path)
            hull_cache[name] = _hull_of(v * scale)
        return hull_cache[name]

    # ---- walk the body tree
    bodies = [dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), inertial=None,
                   joints=[], geoms=[], sites=[], cams=[])]
    joints, geoms, sites, cams = [], [], [], []

    def merged(el, tag, childclass):
        cls = el.attrib.get("class", childclass or "main")
        at = defaults.lookup(cls, tag)
        at.update({k: v for k, v in el.attrib.items() if k != "class"})
        return at

    def walk(node, bid, childclass):
        for el in node:
            if el.tag == "body":
                cc = el.attrib.get("childclass", childclass)
                b = dict(name=el.attrib.get("name", f"body{len(bodies)}"), parent=bid,
                         pos=_f(el.attrib.get("pos", "0 0 0"), 3), quat=_orientation(el.attrib),
                         inertial=None, joints=[], geoms=[], sites=[], cams=[])
                bodies.append(b)
                walk(el, len(bodies) - 1, cc)
            elif el.tag == "inertial":
                bodies[bid]["inertial"] = dict(pos=_f(el.attrib["pos"], 3), quat=_orientation(el.attrib),
                                               mass=float(el.attrib["mass"]), diag=_f(el.attrib["diaginertia"], 3))
            elif el.tag == "joint":
                at = merged(el, "joint", childclass)
                at["body"] = bid
                joints.append(at)
                bodies[bid]["joints"].append(len(joints) - 1)
            elif el.tag == "geom":
                at = merged(el, "geom", childclass)
                if "mesh" in at and "type" not in el.attrib and at.get("type") != "mesh":
                    at["type"] = "mesh"
                at["body"] = bid
                geoms.append(at)
                bodies[bid]["geoms"].append(len(geoms) - 1)
            elif el.tag == "site":
                at = merged(el, "site", childclass)
                at["body"] = bid
                sites.append(at)
            elif el.tag == "camera":
                at = merged(el, "camera", childclass)
                at["body"] = bid
                cams.append(at)

    for wb in root.findall("worldbody"):
        walk(wb, 0, None)

    nbody = len(bodies)
    body_id = {b["name"]: i for i, b in enumerate(bodies)}

    # 2-arm variant: the reference parks the middle arm out of view after compile (env.py:60-62,394-395)
    if num_arms == 2:
        bodies[body_id["middle_base_link"]]["pos"] = np.array([0.0, -2.4, -0.4])

    # ---- joints / dofs
    jnt_type, jnt_body, jnt_qposadr, jnt_dofadr = [], [], [], []
    jnt_axis, jnt_pos, jnt_range, jnt_limited, jnt_names = [], [], [], [], []
    jnt_solref_lim, jnt_solimp_lim = [], []
    dof_body, dof_jnt, dof_parent, dof_arm, dof_damp, dof_floss = [], [], [], [], [], []
    dof_frc_lo, dof_frc_hi, dof_frc_limited = [], [], []
    dof_solref_fr, dof_solimp_fr = [], []
    qpos0 = []
    body_dofadr = np.full(nbody, -1, np.int32)
    body_dofnum = np.zeros(nbody, np.int32)
    body_jntadr = np.full(nbody, -1, np.int32)
    body_jntnum = np.zeros(nbody, np.int32)
    body_lastdof = np.full(nbody, -1, np.int32)  # last dof on the path root..body (for dof_parent)
    for bi, b in enumerate(bodies):
        last = body_lastdof[b["parent"]] if bi else -1
        for k, ji in enumerate(b["joints"]):
            at = joints[ji]
            t = {"free": JNT_FREE, "slide": JNT_SLIDE, "hinge": JNT_HINGE}[at["type"]]
            if k == 0:
                body_jntadr[bi] = len(jnt_type)
                body_dofadr[bi] = len(dof_body)
            body_jntnum[bi] += 1
            jnt_names.append(at.get("name", f"joint{ji}"))
            jnt_type.append(t)
            jnt_body.append(bi)
            jnt_qposadr.append(len(qpos0))
            jnt_dofadr.append(len(dof_body))
            ax = _f(at["axis"], 3)
            jnt_axis.append(ax / np.linalg.norm(ax))
            jnt_pos.append(_f(at["pos"], 3))
            rng = _f(at["range"], 2) if "range" in at else np.zeros(2)
            jnt_range.append(rng)
            jnt_limited.append(int("range" in at))  # autolimits=true [upstream]
            jnt_solref_lim.append(_f(at["solreflimit"], 2))
            jnt_solimp_lim.append(_f(at["solimplimit"], 5))
            nd = 6 if t == JNT_FREE else 1
            if t == JNT_FREE:
                qpos0.extend(list(b["pos"]) + list(b["quat"]))
            else:
                qpos0.append(0.0)
            frc = _f(at["actuatorfrcrange"], 2) if "actuatorfrcrange" in at else np.zeros(2)
            for d in range(nd):
                dof_body.append(bi)
                dof_jnt.append(len(jnt_type) - 1)
                dof_parent.append(last)
                last = len(dof_body) - 1
                dof_arm.append(float(at["armature"]))
                dof_damp.append(float(at["damping"]))
                dof_floss.append(float(at["frictionloss"]))
                dof_frc_lo.append(frc[0])
                dof_frc_hi.append(frc[1])
                dof_frc_limited.append(int("actuatorfrcrange" in at))
                dof_solref_fr.append(_f(at["solreffriction"], 2))
                dof_solimp_fr.append(_f(at["solimpfriction"], 5))
            body_dofnum[bi] += nd
        body_lastdof[bi] = last
    nv, nq, njnt = len(dof_body), len(qpos0), len(jnt_type)
    jnt_id = {n: i for i, n in enumerate(jnt_names)}

    # weld ids: a jointless body is welded to its parent [upstream collision filtering]
    body_weld = np.zeros(nbody, np.int32)
    body_tree = np.full(nbody, -1, np.int32)  # kinematic tree id (root = first jointed ancestor below world)
    ntree = 0
    for bi in range(1, nbody):
        p = bodies[bi]["parent"]
        body_weld[bi] = bi if body_jntnum[bi] else body_weld[p]
        if body_jntnum[bi] and body_tree[p] < 0:
            body_tree[bi] = ntree
            ntree += 1
        else:
            body_tree[bi] = body_tree[p]

    # ---- body inertias (explicit <inertial> wins, else accumulate geoms) [upstream]
    body_mass = np.zeros(nbody)
    body_ipos = np.zeros((nbody, 3))
    body_inertia = np.zeros((nbody, 6))  # xx yy zz xy xz yz in the body frame about ipos

    def geom_shape(at):
        gt = _GEOM_TYPES[at["type"]]
        sz = np.zeros(3)
        if gt != GEOM_MESH:
            s = _f(at["size"])
            sz[: s.size] = s
        return gt, sz

    for bi, b in enumerate(bodies):
        if bi == 0:
            continue
        if b["inertial"] is not None:
            I = b["inertial"]
            R = quat2mat(I["quat"])
            Ib = R @ np.diag(I["diag"]) @ R.T
            body_mass[bi], body_ipos[bi] = I["mass"], I["pos"]
        elif body_jntnum[bi] or body_weld[bi] != 0:
            ms, cs, Is = [], [], []
            for gi in b["geoms"]:
                at = geoms[gi]
                gt, sz = geom_shape(at)
                m, Ig = _geom_mass_inertia(gt, sz, at, b["name"])
                R = quat2mat(_orientation(at))
                ms.append(m)
                cs.append(_f(at.get("pos", "0 0 0"), 3))
                Is.append(R @ Ig @ R.T)
            mtot = sum(ms)
            com = sum(m * c for m, c in zip(ms, cs)) / mtot
            Ib = np.zeros((3, 3))
            for m, c, Ig in zip(ms, cs, Is):
                d = c - com
                Ib += Ig + m * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
            body_mass[bi], body_ipos[bi] = mtot, com
        else:
            continue  # static body welded to the world: no inertia needed
        body_inertia[bi] = [Ib[0, 0], Ib[1, 1], Ib[2, 2], Ib[0, 1], Ib[0, 2], Ib[1, 2]]

    # ---- geoms: keep the collidable ones for physics
    g_keep = [i for i, at in enumerate(geoms) if int(at["contype"]) or int(at["conaffinity"])]
    hull_names = []
    G = dict(type=[], body=[], pos=[], quat=[], size=[], rbound=[], aabb=[], condim=[], friction=[], solref=[],
             solimp=[], gap=[], margin=[], hull=[], contype=[], conaffinity=[], name=[], rgba=[], visible=[])
    for gi in g_keep:
        at = geoms[gi]
        gt, sz = geom_shape(at)
        pos = _f(at.get("pos", "0 0 0"), 3)
        quat = _orientation(at)
        hid = -1
        if gt == GEOM_MESH:
            if at["mesh"] not in hull_names:
                hull_names.append(at["mesh"])
            hid = hull_names.index(at["mesh"])
            hv, cen = hull(at["mesh"])
            # like MuJoCo's mesh recentring [upstream]: the geom frame origin moves to an interior point of
            # the hull (its volume centroid) and the stored hull vertices are relative to it
            pos = pos + quat2mat(quat) @ cen
            lo, hi = (hv - cen).min(0), (hv - cen).max(0)
            aabb = np.maximum(np.abs(lo), np.abs(hi))  # half extents about the hull centroid
            rb = np.linalg.norm(hv - cen, axis=1).max()
        elif gt == GEOM_BOX:
            aabb, rb = sz.copy(), np.linalg.norm(sz)
        elif gt == GEOM_SPHERE:
            aabb, rb = np.full(3, sz[0]), sz[0]
        else:  # cylinder: radius, half height along local z
            aabb, rb = np.array([sz[0], sz[0], sz[1]]), np.hypot(sz[0], sz[1])
        fr = _f(at["friction"])
        fr3 = _f(_BUILTIN["geom"]["friction"], 3)
        fr3[: fr.size] = fr
        G["type"].append(gt); G["body"].append(at["body"]); G["pos"].append(pos); G["quat"].append(quat)
        G["size"].append(sz); G["rbound"].append(rb); G["aabb"].append(aabb); G["condim"].append(int(at["condim"]))
        G["friction"].append(fr3); G["solref"].append(_f(at["solref"], 2)); G["solimp"].append(_f(at["solimp"], 5))
        G["gap"].append(float(at["gap"])); G["margin"].append(float(at["margin"])); G["hull"].append(hid)
        G["contype"].append(int(at["contype"])); G["conaffinity"].append(int(at["conaffinity"]))
        G["name"].append(at.get("name", ""))
        # render colour / visibility of the physics geoms (K9 draws them as stand-ins for the visual meshes): group 3
        # 'collision' geoms of the robot take the black of the visual meshes they stand for; pin-* sensors and the
        # 0.6 mm pad spheres are not drawn
        rgba = materials[at["material"]] if at.get("material") in materials else _f(at.get("rgba", "0.5 0.5 0.5 1"), 4)
        is_robot_hull = gt == GEOM_MESH and int(at.get("group", 0)) == 3
        if is_robot_hull:
            rgba = np.array([0.15, 0.15, 0.15, 1.0])
        vis = not (at.get("name", "").startswith("pin-") or (gt == GEOM_SPHERE and sz[0] < 0.002))
        G["rgba"].append(rgba); G["visible"].append(int(vis))
    ngeom = len(g_keep)

    hull_adr, hull_num, hull_verts = [], [], []
    for hn in hull_names:
        hv, cen = hull(hn)
        hull_adr.append(sum(hull_num))
        hull_num.append(len(hv))
        hull_verts.append(hv - cen)
    hull_kdop = np.array([np.stack([(hv @ KDOP_DIRS.T).min(0), (hv @ KDOP_DIRS.T).max(0)], axis=1) for hv in hull_verts]) \
        if hull_verts else np.zeros((0, len(KDOP_DIRS), 2))          # [nhull, 13, 2]: the renderer's stand-in for the hull (K9)
    hull_verts = np.concatenate(hull_verts, axis=0) if hull_verts else np.zeros((0, 3))

    # ---- contact excludes + candidate pair list (static filters) [upstream collision filtering, SURVEY App. A]
    excl = set()
    for c in root.findall("contact"):
        for e in c.findall("exclude"):
            a, b_ = body_id[e.attrib["body1"]], body_id[e.attrib["body2"]]
            excl.add((min(a, b_), max(a, b_)))
    weld_parent = np.zeros(nbody, np.int32)
    for bi in range(1, nbody):
        w = body_weld[bi]
        weld_parent[bi] = body_weld[bodies[w]["parent"]] if w else 0
    pairs = []
    for i in range(ngeom):
        for j in range(i + 1, ngeom):
            b1, b2 = G["body"][i], G["body"][j]
            w1, w2 = body_weld[b1], body_weld[b2]
            if w1 == w2:
                continue  # same (weld) body; also both static
            if (min(b1, b2), max(b1, b2)) in excl:
                continue
            # parent-child filter on weld bodies, not applied when either weld body is the world
            if w1 != 0 and w2 != 0 and (weld_parent[b2] == w1 or weld_parent[b1] == w2):
                continue
            if not ((G["contype"][i] & G["conaffinity"][j]) or (G["contype"][j] & G["conaffinity"][i])):
                continue
            pairs.append((i, j))
    # primitive pairs (sphere / box on both sides) first, then the convex ones, each group in geom order: the CUDA
    # narrowphase drains a primitive and a convex candidate list and must emit contacts in this table's order
    prim = lambda g: G["type"][g] in (GEOM_SPHERE, GEOM_BOX)
    pairs = [p for p in pairs if prim(p[0]) and prim(p[1])] + [p for p in pairs if not (prim(p[0]) and prim(p[1]))]
    pairs = np.array(pairs, np.int32).reshape(-1, 2)

    # ---- equality, actuators
    eq_dof1, eq_dof2, eq_q1, eq_q2, eq_poly, eq_solref, eq_solimp = [], [], [], [], [], [], []
    for e in root.findall("equality"):
        for j in e.findall("joint"):
            j1, j2 = jnt_id[j.attrib["joint1"]], jnt_id[j.attrib["joint2"]]
            eq_dof1.append(jnt_dofadr[j1]); eq_dof2.append(jnt_dofadr[j2])
            eq_q1.append(jnt_qposadr[j1]); eq_q2.append(jnt_qposadr[j2])
            eq_poly.append(_f(j.attrib.get("polycoef", "0 1 0 0 0"), 5))
            eq_solref.append(_f(j.attrib.get("solref", "0.02 1"), 2))
            eq_solimp.append(_f(j.attrib.get("solimp", "0.9 0.95 0.001 0.5 2"), 5))
    act_names, act_dof, act_qadr, act_kp, act_kv, act_lo, act_hi = [], [], [], [], [], [], []
    for a in root.findall("actuator"):
        for p in a.findall("position"):
            at = defaults.lookup(p.attrib.get("class", "main"), "position")
            at.update(p.attrib)
            j = jnt_id[at["joint"]]
            cr = _f(at["ctrlrange"], 2)
            act_names.append(at["name"]); act_dof.append(jnt_dofadr[j]); act_qadr.append(jnt_qposadr[j])
            act_kp.append(float(at["kp"])); act_kv.append(float(at["kv"])); act_lo.append(cr[0]); act_hi.append(cr[1])

    m = dict(
        task_id=np.int32(task_id), max_reward=np.int32(max_reward), num_arms=np.int32(num_arms),
        timestep=np.float64(timestep), impratio=np.float64(opt["impratio"]),
        noslip_iterations=np.int32(opt["noslip_iterations"]), multiccd=np.int32(opt["multiccd"]),
        gravity=np.array([0, 0, -9.81]),
        body_parent=np.array([b["parent"] for b in bodies], np.int32),
        body_pos=np.array([b["pos"] for b in bodies]), body_quat=np.array([b["quat"] for b in bodies]),
        body_mass=body_mass, body_ipos=body_ipos, body_inertia=body_inertia,
        body_jntadr=body_jntadr, body_jntnum=body_jntnum, body_dofadr=body_dofadr, body_dofnum=body_dofnum,
        body_weld=body_weld, body_tree=body_tree,
        jnt_type=np.array(jnt_type, np.int32), jnt_body=np.array(jnt_body, np.int32),
        jnt_qposadr=np.array(jnt_qposadr, np.int32), jnt_dofadr=np.array(jnt_dofadr, np.int32),
        jnt_axis=np.array(jnt_axis), jnt_pos=np.array(jnt_pos), jnt_range=np.array(jnt_range),
        jnt_limited=np.array(jnt_limited, np.int32),
        jnt_solref=np.array(jnt_solref_lim), jnt_solimp=np.array(jnt_solimp_lim),
        dof_body=np.array(dof_body, np.int32), dof_jnt=np.array(dof_jnt, np.int32),
        dof_parent=np.array(dof_parent, np.int32), dof_armature=np.array(dof_arm), dof_damping=np.array(dof_damp),
        dof_frictionloss=np.array(dof_floss), dof_frc_lo=np.array(dof_frc_lo), dof_frc_hi=np.array(dof_frc_hi),
        dof_frc_limited=np.array(dof_frc_limited, np.int32),
        dof_solref=np.array(dof_solref_fr), dof_solimp=np.array(dof_solimp_fr),
        qpos0=np.array(qpos0),
        geom_type=np.array(G["type"], np.int32), geom_body=np.array(G["body"], np.int32),
        geom_pos=np.array(G["pos"]), geom_quat=np.array(G["quat"]), geom_size=np.array(G["size"]),
        geom_rbound=np.array(G["rbound"]), geom_aabb=np.array(G["aabb"]),
        geom_condim=np.array(G["condim"], np.int32), geom_friction=np.array(G["friction"]),
        geom_solref=np.array(G["solref"]), geom_solimp=np.array(G["solimp"]), geom_gap=np.array(G["gap"]),
        geom_margin=np.array(G["margin"]), geom_hull=np.array(G["hull"], np.int32),
        geom_rgba=np.array(G["rgba"]), geom_visible=np.array(G["visible"], np.int32),
        hull_adr=np.array(hull_adr, np.int32), hull_num=np.array(hull_num, np.int32), hull_vert=hull_verts,
        hull_kdop=hull_kdop,
        pair_geom=pairs,
        eq_dof1=np.array(eq_dof1, np.int32), eq_dof2=np.array(eq_dof2, np.int32),
        eq_qadr1=np.array(eq_q1, np.int32), eq_qadr2=np.array(eq_q2, np.int32),
        eq_polycoef=np.array(eq_poly), eq_solref=np.array(eq_solref), eq_solimp=np.array(eq_solimp),
        act_dof=np.array(act_dof, np.int32), act_qadr=np.array(act_qadr, np.int32),
        act_kp=np.array(act_kp), act_kv=np.array(act_kv), act_ctrl_lo=np.array(act_lo), act_ctrl_hi=np.array(act_hi),
    )
    names = dict(body=[b["name"] for b in bodies], joint=jnt_names, geom=G["name"], actuator=act_names,
                 site=[s.get("name", "") for s in sites], camera=[c.get("name", "") for c in cams],
                 hull=hull_names)

    # sites / cameras (body-relative frames; world poses are produced on demand)
    m["site_body"] = np.array([s["body"] for s in sites], np.int32)
    m["site_pos"] = np.array([_f(s.get("pos", "0 0 0"), 3) for s in sites])
    m["site_quat"] = np.array([_orientation(s) for s in sites])
    m["cam_body"] = np.array([c["body"] for c in cams], np.int32)
    m["cam_pos"] = np.array([_f(c.get("pos", "0 0 0"), 3) for c in cams])
    m["cam_quat"] = np.array([_orientation(c) for c in cams])
    m["cam_fovy"] = np.array([float(c["fovy"]) for c in cams])

    # ---- render-only constants (K9): the scene's directional light (scene.xml:50; MuJoCo's default light diffuse 0.7 [upstream]),
    # the headlight (scene.xml:9) and the table's diffuse texture (scene.xml:30,32: material "table" on the tabletop mesh), box
    # filtered down to 128 x 128 texels and drawn on the top face of the table box that stands in for it
    light = next(iter(root.iter("light")), None)
    ldir = _f(light.attrib.get("dir", "0 0 -1"), 3) if light is not None else np.array([0.0, 0.0, -1.0])
    head = next(iter(root.iter("headlight")), None)
    amb = float(_f(head.attrib.get("ambient", "0.1 0.1 0.1"))[0]) if head is not None else 0.1
    dif = float(_f(head.attrib.get("diffuse", "0.4 0.4 0.4"))[0]) if head is not None else 0.4
    m["light"] = np.array(list(ldir / np.linalg.norm(ldir)) + [0.7 if light is not None else 0.0, amb, dif])
    tex = np.full((128, 128, 3), 0.5)
    tex_path = os.path.join(asset_dir, "meshes", "small_meta_table_diffuse.png")
    if os.path.exists(tex_path):
        import cv2
        img = cv2.imread(tex_path, cv2.IMREAD_COLOR)
        tex = cv2.resize(img, (128, 128), interpolation=cv2.INTER_AREA)[:, :, ::-1].astype(np.float64) / 255.0
    m["table_tex"] = tex.reshape(-1)
    m["geom_tex"] = np.array([1 if n == "table" else 0 for n in G["name"]], np.int32)

    _finish(m, names)
    return m, names


# ----------------------------------------------------------------------------- derived constants
def fk_numpy(m, qpos):
    """Tree forward kinematics in numpy (compile-time only): world pose of every body + joint frames."""
    nb = len(m["body_parent"])
    xpos = np.zeros((nb, 3)); xquat = np.zeros((nb, 4)); xquat[0, 0] = 1
    xanchor = np.zeros((len(m["jnt_type"]), 3)); xaxis = np.zeros((len(m["jnt_type"]), 3))
    for b in range(1, nb):
        p = m["body_parent"][b]
        Rp = quat2mat(xquat[p])
        pos = xpos[p] + Rp @ m["body_pos"][b]
        quat = quat_mul(xquat[p], m["body_quat"][b])
        for k in range(m["body_jntnum"][b]):
            j = m["body_jntadr"][b] + k
            qa = m["jnt_qposadr"][j]
            t = m["jnt_type"][j]
            if t == JNT_FREE:
                pos = qpos[qa:qa + 3].copy()
                quat = qpos[qa + 3:qa + 7] / np.linalg.norm(qpos[qa + 3:qa + 7])
                xanchor[j] = pos; xaxis[j] = [0, 0, 1]
                continue
            R = quat2mat(quat)
            xaxis[j] = R @ m["jnt_axis"][j]
            xanchor[j] = pos + R @ m["jnt_pos"][j]
            dq = qpos[qa] - m["qpos0"][qa]
            if t == JNT_SLIDE:
                pos = pos + xaxis[j] * dq
            else:
                ax = m["jnt_axis"][j]
                quat = quat_mul(quat, np.concatenate([[np.cos(dq / 2)], np.sin(dq / 2) * ax]))
                pos = xanchor[j] - quat2mat(quat) @ m["jnt_pos"][j]
        xpos[b], xquat[b] = pos, quat / np.linalg.norm(quat)
    return xpos, xquat, xanchor, xaxis


def body_jacobian(m, xpos, xquat, xanchor, xaxis, b, point):
    """(jacp, jacr) 3 x nv of a world point attached to body b."""
    nv = len(m["dof_body"])
    jp = np.zeros((3, nv)); jr = np.zeros((3, nv))
    d = m["body_dofadr"][b] + m["body_dofnum"][b] - 1 if m["body_dofnum"][b] else -1
    if d < 0:  # climb to the first ancestor with dofs
        bb = b
        while bb and not m["body_dofnum"][bb]:
            bb = m["body_parent"][bb]
        d = m["body_dofadr"][bb] + m["body_dofnum"][bb] - 1 if bb else -1
    while d >= 0:
        j = m["dof_jnt"][d]
        t = m["jnt_type"][j]
        if t == JNT_FREE:
            k = d - m["jnt_dofadr"][j]
            if k < 3:
                jp[k, d] = 1.0
            else:
                ax = quat2mat(xquat[m["dof_body"][d]])[:, k - 3]
                jr[:, d] = ax
                jp[:, d] = np.cross(ax, point - xpos[m["dof_body"][d]])
        elif t == JNT_SLIDE:
            jp[:, d] = xaxis[j]
        else:
            jr[:, d] = xaxis[j]
            jp[:, d] = np.cross(xaxis[j], point - xanchor[j])
        d = m["dof_parent"][d]
    return jp, jr


def mass_matrix_numpy(m, qpos):
    """Joint-space inertia as sum_b J_b^T diag(m, I_b) J_b + armature (compile-time + test cross-check)."""
    xpos, xquat, xanchor, xaxis = fk_numpy(m, qpos)
    nv = len(m["dof_body"])
    M = np.diag(m["dof_armature"].astype(np.float64))
    for b in range(1, len(m["body_parent"])):
        if m["body_mass"][b] == 0 or m["body_weld"][b] == 0:
            continue
        R = quat2mat(xquat[b])
        c = xpos[b] + R @ m["body_ipos"][b]
        i6 = m["body_inertia"][b]
        Ib = np.array([[i6[0], i6[3], i6[4]], [i6[3], i6[1], i6[5]], [i6[4], i6[5], i6[2]]])
        Iw = R @ Ib @ R.T
        jp, jr = body_jacobian(m, xpos, xquat, xanchor, xaxis, b, c)
        M += m["body_mass"][b] * jp.T @ jp + jr.T @ Iw @ jr
    return M


def _finish(m, names):
    """invweight0 tables [upstream engine_setconst semantics, SURVEY App. A] and IK screw tables."""
    qpos0 = m["qpos0"]
    M = mass_matrix_numpy(m, qpos0)
    Minv = np.linalg.inv(M)
    nv = len(m["dof_body"])
    dof_inv = np.zeros(nv)
    for j, t in enumerate(m["jnt_type"]):
        a = m["jnt_dofadr"][j]
        if t == JNT_FREE:
            dof_inv[a:a + 3] = np.mean(np.diag(Minv)[a:a + 3])
            dof_inv[a + 3:a + 6] = np.mean(np.diag(Minv)[a + 3:a + 6])
        else:
            dof_inv[a] = Minv[a, a]
    m["dof_invweight0"] = dof_inv
    xpos, xquat, xanchor, xaxis = fk_numpy(m, qpos0)
    nb = len(m["body_parent"])
    binv = np.zeros((nb, 2))
    for b in range(1, nb):
        if m["body_weld"][b] == 0:
            continue
        c = xpos[b] + quat2mat(xquat[b]) @ m["body_ipos"][b]
        jp, jr = body_jacobian(m, xpos, xquat, xanchor, xaxis, b, c)
        binv[b, 0] = np.trace(jp @ Minv @ jp.T) / 3.0
        binv[b, 1] = np.trace(jr @ Minv @ jr.T) / 3.0
    m["body_invweight0"] = binv
    m["eq_invweight0"] = np.array([dof_inv[a] + dof_inv[b] for a, b in zip(m["eq_dof1"], m["eq_dof2"])])

    # IK tables: screw axes at q = 0, as create_fk_fn captures them (reference kinematics.py:7-15)
    q_zero = qpos0.copy()
    xpos, xquat, xanchor, xaxis = fk_numpy(m, q_zero)
    jid = {n: i for i, n in enumerate(names["joint"])}
    sid = {n: i for i, n in enumerate(names["site"])}
    arms = {
        "left": (["waist", "shoulder", "elbow", "forearm_roll", "wrist_angle", "wrist_rotate"], "left_gripper_control"),
        "right": (["waist", "shoulder", "elbow", "forearm_roll", "wrist_angle", "wrist_rotate"], "right_gripper_control"),
        "middle": (["waist", "shoulder", "elbow", "forearm_roll", "wrist_1_joint", "wrist_2_joint", "wrist_3_joint"],
                   "middle_zed_camera_center"),
    }
    w0 = np.zeros((3, 7, 3)); p0 = np.zeros((3, 7, 3)); site0 = np.zeros((3, 4, 4)); rng = np.zeros((3, 7, 2))
    ndof = np.zeros(3, np.int32); qadr = np.zeros((3, 7), np.int32)
    for a, (arm, (jn, site)) in enumerate(arms.items()):
        ndof[a] = len(jn)
        for k, n in enumerate(jn):
            j = jid[f"{arm}_{n}"]
            w0[a, k], p0[a, k], rng[a, k] = xaxis[j], xanchor[j], m["jnt_range"][j]
            qadr[a, k] = m["jnt_qposadr"][j]
        s = sid[site]
        b = m["site_body"][s]
        Rb = quat2mat(xquat[b])
        site0[a] = np.eye(4)
        site0[a, :3, :3] = Rb @ quat2mat(m["site_quat"][s])
        site0[a, :3, 3] = xpos[b] + Rb @ m["site_pos"][s]
    m["ik_ndof"], m["ik_w0"], m["ik_p0"], m["ik_site0"], m["ik_range"], m["ik_qadr"] = ndof, w0, p0, site0, rng, qadr

    # geom classes for the contact-name reward predicates (reference env.py:444-461,564-578,659-677,756-770,838-852)
    cls = np.zeros(len(names["geom"]), np.int32)
    for i, n in enumerate(names["geom"]):
        cls[i] = geom_class_mask(int(m["task_id"]), n)
    m["geom_class"] = cls

    # env-level index tables (reference constants.py:29-88, env.py:169-178,204-215,233-242)
    def qa(n):
        return m["jnt_qposadr"][jid[n]]
    arm_j = ["waist", "shoulder", "elbow", "forearm_roll", "wrist_angle", "wrist_rotate"]
    left = [qa(f"left_{n}") for n in arm_j] + [qa("left_left_finger")]
    right = [qa(f"right_{n}") for n in arm_j] + [qa("right_right_finger")]
    middle = [qa(f"middle_{n}") for n in arms["middle"][0]]
    m["obs_qadr"] = np.array(left + right + middle, np.int32)
    m["finger_qadr"] = np.array([qa("left_left_finger"), qa("left_right_finger"),
                                 qa("right_left_finger"), qa("right_right_finger")], np.int32)
    free = [j for j, t in enumerate(m["jnt_type"]) if t == JNT_FREE]
    # task reset ranges (reference env.py:478-493,517-533,608-624,709-723,796-808); uniform(lo, hi) per coordinate
    lo, hi, src = [], [], []
    fnames = [names["joint"][j] for j in free]
    for k, n in enumerate(fnames):
        rl, rh, sname = RESET_RANGES[int(m["task_id"])][n]
        lo.append(rl); hi.append(rh); src.append(fnames.index(sname) if sname else k)
    m["reset_lo"], m["reset_hi"], m["reset_draw"] = np.array(lo, np.float64), np.array(hi, np.float64), np.array(src, np.int32)
    m["free_qadr"] = np.array([m["jnt_qposadr"][j] for j in free], np.int32)
    m["free_names"] = names["free_joint"] = [names["joint"][j] for j in free]
    del m["free_names"]


# free joint -> (low[3], high[3], name of the free joint whose draw is re-used or None).  Note the reference passes
# low > high for the hole's x range (env.py:485): numpy's uniform then samples low + (high-low)*u, kept as is.
RESET_RANGES = {
    0: {"peg_joint": ([0.1, -0.1, 0.01], [0.2, 0.1, 0.01], None),
        "hole_joint": ([-0.1, -0.1, 0.021], [-0.2, 0.1, 0.021], None)},
    1: {"slot_joint": ([-0.05, 0.1, 0.0], [0.05, 0.15, 0.0], None),
        "stick_joint": ([-0.08, -0.1, 0.0], [0.08, 0.0, 0.0], None)},
    2: {"wall_joint": ([-0.025, -0.025, 0.0], [0.025, 0.1, 0.0], None),
        "needle_joint": ([0.15, -0.025, 0.0], [0.2, 0.1, 0.0], None)},
    3: {"ball_joint": ([0.05, -0.05, 0.0], [0.1, 0.05, 0.0], None),
        "tube1_joint": ([0.05, -0.05, 0.0], [0.1, 0.05, 0.0], "ball_joint"),
        "tube2_joint": ([-0.1, -0.05, 0.0], [-0.05, 0.05, 0.0], None)},
    4: {"hook_joint": ([-0.1, 0.3, 0.2], [0.1, 0.3, 0.3], None),
        "package_joint": ([-0.1, 0.0, 0.0], [0.1, 0.15, 0.0], None)},
}

# class bits shared by all tasks
CLS_LEFT, CLS_RIGHT, CLS_TABLE = 1, 2, 4
# task-specific bits start at 8
CLS_A, CLS_B, CLS_PIN_A, CLS_PIN_B, CLS_C = 8, 16, 32, 64, 128


def geom_class_mask(task_id, name):
    """Bitmask class of a geom name under the reward predicates of ``task_id``.

    The reference compares geom *names* (``==`` / ``startswith``); the kernel compares these bits.
    """
    c = 0
    if name.startswith("left"):
        c |= CLS_LEFT
    if name.startswith("right"):
        c |= CLS_RIGHT
    if name == "table":
        c |= CLS_TABLE
    if task_id == 0:  # insert peg: A=peg, B=hole-*, PIN_A=pin
        if name == "peg": c |= CLS_A
        if name.startswith("hole-"): c |= CLS_B
        if name == "pin": c |= CLS_PIN_A
    elif task_id == 1:  # slot insertion: A=stick, B=slot-*, pins
        if name == "stick": c |= CLS_A
        if name.startswith("slot-"): c |= CLS_B
        if name == "pin-stick": c |= CLS_PIN_A
        if name == "pin-slot": c |= CLS_PIN_B
    elif task_id == 2:  # sew needle: A=needle, B=wall-*, pins
        if name == "needle": c |= CLS_A
        if name.startswith("wall-"): c |= CLS_B
        if name == "pin-needle": c |= CLS_PIN_A
        if name == "pin-wall": c |= CLS_PIN_B
    elif task_id == 3:  # tube transfer: A=tube1-*, B=tube2-*, C=ball, PIN_A=pin
        if name.startswith("tube1-"): c |= CLS_A
        if name.startswith("tube2-"): c |= CLS_B
        if name == "ball": c |= CLS_C
        if name == "pin": c |= CLS_PIN_A
    elif task_id == 4:  # hook package: A=package-*, B=hook, pins
        if name.startswith("package-"): c |= CLS_A
        if name == "hook": c |= CLS_B
        if name == "pin-package": c |= CLS_PIN_A
        if name == "pin-hook": c |= CLS_PIN_B
    return c
