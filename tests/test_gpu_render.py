"""GPU tests of the camera renderer (K9).  There is no pixel oracle (no GL context, and the reference declares its
renders non-deterministic, gym_guided_vision/__init__.py:92-94), so these tests pin what CAN be pinned: the camera
model (a known world point lands on the pixel the MuJoCo pin-hole convention predicts: -Z forward, +Y up, vertical
fovy), determinism, the output layout, and that body-mounted cameras follow their body."""
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
        assert np.abs(px / px.max() - green / green.max()).max() < 0.05, (e, r, c, px)      # stick colour (task_slot_insertion.xml:14)
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
