"""GPU tests of the camera renderer (K9).  There is no pixel oracle (no GL context, and the reference declares its
renders non-deterministic, gym_guided_vision/__init__.py:92-94), so these tests pin what CAN be pinned: the camera
model (a known world point lands on the pixel the MuJoCo pin-hole convention predicts: -Z forward, +Y up, vertical
fovy), determinism, the output layout, that body-mounted cameras follow their body, and -- through the id-buffer hook
avsim_render_ids -- that every mesh geom is drawn inside (and, where nothing occludes it, over) the projection of its convex
hull: silhouette IoU of the 26-DOP stand-ins against the hulls' projected vertices."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _quat2mat(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _project(p_world, cam_pos, cam_mat, fovy_deg, H, W):
    pc = cam_mat.T @ (p_world - cam_pos)
    th = np.tan(np.radians(fovy_deg) / 2)
    x, y = pc[0] / -pc[2], pc[1] / -pc[2]
    return (1 - y / th) * H / 2, (x / (th * W / H) + 1) * W / 2          # (row, col)


def test_static_camera_projection_and_layout():
    import torch
    from av_aloha_b200 import capi, model_io

    task = "slot_insertion"
    model = capi.Model(model_io.model_path(task, 3), 0)
    names = model_io.load_names(task, 3)
    cams = names["camera"]
    B, H, W = 3, 480, 640
    b = capi.Batch(model, B, seed=1)
    fp = np.array([[[0.0, 0.12, 0.0], [0.05 * e - 0.05, -0.05, 0.0]] for e in range(B)])     # stick at a different x per env
    b.reset(free_pos=fp)
    ids = [cams.index("overhead_cam"), cams.index("worms_eye_cam")]
    img = b.render(ids, H, W)
    assert img.shape == (B, 2, H, W, 3) and img.dtype == torch.uint8
    img2 = b.render(ids, H, W)
    assert torch.equal(img, img2)                                       # deterministic
    img = img.cpu().numpy()
    k = ids[0]
    assert model.table("cam_body")[k] == 0                              # world-fixed camera: pose straight from the model
    cpos, cmat, fovy = model.table("cam_pos")[k], _quat2mat(model.table("cam_quat")[k]), model.table("cam_fovy")[k]
    green = np.array([0.4, 0.8, 0.4])
    for e in range(B):
        top = np.array([fp[e, 1, 0], fp[e, 1, 1], 0.04])               # centre of the stick's top face
        r, c = _project(top, cpos, cmat, fovy, H, W)
        px = img[e, 0, int(r), int(c)].astype(float) / 255
        # stick colour (task_slot_insertion.xml:14) under some light level, channels clipped at 1 as GL does
        err = min(np.abs(px - np.minimum(1.0, green * lum)).max() for lum in np.linspace(0.3, 1.7, 141))
        assert err < 0.03, (e, r, c, px)
        off = img[e, 0, int(r), min(W - 1, int(c) + 200)].astype(float)                       # far to the side: not the stick
        assert np.abs(off / max(1.0, off.max()) - green / green.max()).max() > 0.1
    # the three environments differ only where the stick is
    assert not np.array_equal(img[0, 0], img[1, 0])
    assert len(np.unique(img[0, 0].reshape(-1, 3), axis=0)) > 20       # shaded scene, not a flat fill
    b.close()


def test_wrist_camera_follows_the_arm_and_env_pixels():
    from av_aloha_b200 import env

    e = env.SewNeedleEnv(num_arms=3, cameras=["zed_cam_left", "wrist_cam_left"], observation_height=120, observation_width=160)
    np.random.seed(0)
    obs, _ = e.reset()
    assert set(obs["pixels"]) == {"zed_cam_left", "wrist_cam_left"}
    assert obs["pixels"]["wrist_cam_left"].shape == (120, 160, 3) and obs["pixels"]["wrist_cam_left"].dtype == np.uint8
    a = np.array([0, -0.082, 1.06, 0, -0.953, 0, 1] * 2 + [0, -0.8, 0.8, 0, 0.5, 0, 0], np.float32)
    a[0] = 0.6                                                          # swing the left waist: the wrist camera view changes
    for _ in range(10):
        obs2, *_ = e.step(a)
    assert not np.array_equal(obs2["pixels"]["wrist_cam_left"], obs["pixels"]["wrist_cam_left"])
    e.close()
    v = env.GuidedVisionVectorEnv("sew_needle", 4, cameras=["zed_cam_left"], observation_height=60, observation_width=80,
                                  max_episode_steps=2)
    o, _ = v.reset()
    assert o["pixels"]["zed_cam_left"].shape == (4, 60, 80, 3)
    acts = np.tile(a, (4, 1))
    v.step(acts)
    o, r, term, trunc, info = v.step(acts)
    assert trunc.all() and info["final_observation"][0]["pixels"]["zed_cam_left"].shape == (60, 80, 3)
    assert len(v.call("render")) == 4 and v.call("render")[0].shape == (225, 300, 3)
    v.close()


def test_mesh_geoms_are_drawn_as_their_hulls_silhouette_iou():
    """Every mesh geom is drawn as the 26-DOP of its convex hull.  Check against the hull itself: for the overhead camera (a
    static camera that sees the whole workspace from above, everything in front of it) the set of pixels whose primary ray hits
    geom g -- read back with the id-buffer hook -- is compared with the projection of g's hull vertices (convex polygon, filled on
    the CPU).  Occlusion only removes pixels, so the check is: rendered pixels of g lie inside the dilated hull polygon
    (no pixel outside the k-DOP's slack), and the un-occluded geoms cover their hull polygon (IoU)."""
    import cv2
    from av_aloha_b200 import capi, model_io
    from oracle.oracle import OracleEnv, OracleModel

    task = "slot_insertion"
    path = model_io.model_path(task, 3)
    model = capi.Model(path, 0)
    avm = model_io.load_avm(path)
    cams = model_io.load_names(task, 3)["camera"]
    H, W = 480, 640
    b = capi.Batch(model, 1, seed=1)
    b.reset(free_pos=np.array([[[0.0, 0.12, 0.0], [0.03, -0.05, 0.0]]]))
    cam = cams.index("overhead_cam")
    ids = b.render([cam], H, W, ids=True)[0, 0, :, :, 0].cpu().numpy()
    rgb = b.render([cam], H, W)[0, 0].cpu().numpy()
    assert rgb.reshape(-1, 3).std(axis=0).min() > 5          # lit, textured picture (not a flat fill)
    o = OracleEnv(OracleModel(path))                          # world poses of the geoms at the same qpos (fp64 kinematics)
    o.qpos[:] = b.get(capi.QPOS).cpu().numpy()[0]
    o.forward()
    gpos, gmat = o.gpos, o.gmat
    cpos, cmat, fovy = avm["cam_pos"][cam], _quat2mat(avm["cam_quat"][cam]), avm["cam_fovy"][cam]
    ious, spill, checked = [], [], 0
    for g in np.nonzero(avm["geom_type"] == 7)[0]:
        if not avm["geom_visible"][g]:
            continue
        h = avm["geom_hull"][g]
        v = avm["hull_vert"][avm["hull_adr"][h]: avm["hull_adr"][h] + avm["hull_num"][h]]
        pw = gpos[g] + v @ gmat[g].T
        if ((cmat.T @ (pw - cpos).T)[2] > -0.05).any():
            continue                                          # partly behind the camera: not a plain polygon
        rc = np.array([_project(p, cpos, cmat, fovy, H, W) for p in pw])
        poly = cv2.convexHull(rc[:, ::-1].astype(np.float32))
        mask = np.zeros((H, W), np.uint8)
        cv2.fillConvexPoly(mask, np.round(poly).astype(np.int32), 1)
        if mask.sum() < 60:
            continue
        mine = ids == g
        if mine.sum() == 0:
            continue                                          # fully hidden behind something else
        grown = cv2.dilate(mask, np.ones((3, 3), np.uint8), iterations=1 + int(0.06 * np.sqrt(mask.sum())))
        spill.append(float((mine & (grown == 0)).sum()) / mine.sum())
        visible = (ids == g) | ((ids != g) & (mask == 0))     # pixels of the polygon not taken by another geom
        occluded = ((mask == 1) & (ids != g) & (ids != 255)).sum() / mask.sum()
        if occluded < 0.05:
            ious.append(float((mine & (mask == 1)).sum()) / float((mine | (mask == 1)).sum()))
        checked += 1
    assert checked >= 15 and len(ious) >= 6
    assert max(spill) <= 0.08 and np.median(spill) <= 0.02, (max(spill), np.median(spill))   # the 26-DOP's slack over the hull outline
    assert np.median(ious) >= 0.85 and min(ious) >= 0.7, (np.median(ious), min(ious))
    b.close()
