"""Device-resident observation path and batched dataset replay (f2, f3).

The image conversion replaces lerobot's preprocess_observation (reference lerobot/lerobot/common/envs/utils.py:37-50);
its oracle is that function's own arithmetic -- rearrange, .type(float32), /= 255 -- restated with torch on the CPU.
The bar is bit-exact: fp32 division is correctly rounded on both sides.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_images(img_u8):
    """utils.py:37-50 on the host: b h w c -> b c h w, float32, /= 255"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(img_u8))
    t = t.permute(0, 3, 1, 2).contiguous().type(torch.float32)
    t /= 255
    return t


def test_u8_over_255_is_correctly_rounded_for_all_inputs():
    """av_u8_unit (csrc/avsim_obs.cuh: multiply by the rounded reciprocal + one fma correction) equals the IEEE quotient
    v / 255.0f for every uint8 v: the header's arithmetic is compiled for the host and compared exhaustively."""
    src = r'''
#include <stdio.h>
#include "avsim_obs.cuh"
int main() { int bad = 0; for (int v = 0; v < 256; v++) { volatile float a = (float)v, b = 255.0f; if (av_u8_unit((float)v) != a / b) bad++; }
  printf("%d\n", bad); return bad != 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "chk.cpp")
        open(c, "w").write(src)
        exe = os.path.join(d, "chk")
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "av_aloha_b200", "csrc"), "-o", exe, c])
        out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "0", out.stdout
    # and numpy agrees with what torch's fp32 division produces
    import torch
    v = np.arange(256, dtype=np.uint8)
    assert np.array_equal((v.astype(np.float32) / np.float32(255)), (torch.from_numpy(v).float() / 255).numpy())


def test_product_package_never_touches_the_oracle_and_has_no_cpu_fallback():
    """The oracle is test infrastructure: nothing under av_aloha_b200/ may import, load or execute it, and the C-ABI loader
    must fail loudly -- not fall back -- when the CUDA library is missing."""
    import re
    pkg = os.path.join(ROOT, "av_aloha_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "libavsim_oracle" not in text and "oracle/_ref" not in text, f
    from av_aloha_b200 import capi
    saved, capi._lib, capi.LIB_PATH = (capi._lib, capi.LIB_PATH), None, os.path.join(pkg, "csrc", "no_such_library.so")
    try:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            capi.load_library()
    finally:
        capi._lib, capi.LIB_PATH = saved


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 2, 48, 64, 3), (1, 480, 640, 3), (5, 15, 17, 3), (2, 225, 300, 3), (7, 1, 4, 3)])
def test_pixels_to_float_bit_exact(shape):
    import torch
    from av_aloha_b200 import capi
    rng = np.random.default_rng(sum(shape))
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    flat = img.reshape(-1)
    flat[: min(256, flat.size)] = np.arange(min(256, flat.size), dtype=np.uint8)      # every byte value where the image is large enough
    out = capi.pixels_to_float(torch.from_numpy(img).cuda())
    want = reference_images(img.reshape((-1,) + shape[-3:])).reshape(shape[:-3] + (3,) + shape[-3:-1])
    assert out.dtype == torch.float32 and tuple(out.shape) == tuple(want.shape)
    assert torch.equal(out.cpu(), want)


@pytest.mark.gpu
def test_pixels_to_float_edge_cases_and_errors():
    import torch
    from av_aloha_b200 import capi
    empty = capi.pixels_to_float(torch.empty((0, 8, 8, 3), dtype=torch.uint8, device="cuda"))
    assert tuple(empty.shape) == (0, 3, 8, 8)
    with pytest.raises(ValueError):
        capi.pixels_to_float(torch.zeros((2, 8, 8, 3), dtype=torch.float32, device="cuda"))
    with pytest.raises(ValueError):
        capi.pixels_to_float(torch.zeros((2, 8, 8, 4), dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        capi.pixels_to_float(torch.zeros((2, 8, 8, 3), dtype=torch.uint8))
    # a misaligned source view (offset by one byte) takes the general kernel and is still exact
    base = torch.randint(0, 256, (1 + 2 * 8 * 12 * 3,), dtype=torch.uint8, device="cuda")
    view = base[1:].view(2, 8, 12, 3)
    assert view.data_ptr() % 4 == 1
    got = capi.pixels_to_float(view)
    assert torch.equal(got.cpu(), reference_images(view.cpu().numpy()))
    lib = capi.load_library()
    assert lib.avsim_pixels_to_float(None, 1, 8, 8, None, 0, None) < 0 and b"pixels_to_float" in lib.avsim_last_error()


@pytest.mark.gpu
def test_preprocess_observation_matches_reference_contract():
    """numpy gym observations and device observations give the same tensors, equal to the reference's host conversion"""
    import torch
    from av_aloha_b200 import observation
    from av_aloha_b200.env import GuidedVisionVectorEnv
    cams = ["zed_cam_left", "wrist_cam_right"]
    env = GuidedVisionVectorEnv("slot_insertion", 3, cameras=cams, observation_height=48, observation_width=64)
    obs, _ = env.reset()
    got = observation.preprocess_observation(obs)
    assert set(got) == {"observation.images.zed_cam_left", "observation.images.wrist_cam_right", "observation.state"}
    for c in cams:
        assert obs["pixels"][c].shape == (3, 48, 64, 3) and int(obs["pixels"][c].max()) > 0
        assert torch.equal(got[f"observation.images.{c}"].cpu(), reference_images(obs["pixels"][c]))
    assert got["observation.state"].dtype == torch.float32 and got["observation.state"].is_cuda
    assert torch.equal(got["observation.state"].cpu(), torch.from_numpy(obs["agent_pos"]).float())
    dev_obs = env.observation_device()
    assert dev_obs["pixels"].is_cuda and tuple(dev_obs["pixels"].shape) == (3, 2, 48, 64, 3)
    got_dev = observation.preprocess_observation(dev_obs, cameras=cams)
    for k in got:
        assert torch.equal(got_dev[k], got[k]), k
    with pytest.raises(ValueError):
        observation.preprocess_observation(dev_obs)            # camera names missing
    env.close()


class _ChunkPolicy:
    """Stand-in with ACT's action-queue behaviour (modeling_act.py:123-131): reads the observation only when the queue is
    empty, then plans `n` actions from the image mean and the joint state."""

    def __init__(self, n, home):
        from collections import deque
        self.n, self.home, self._action_queue, self.calls = n, home, deque([], maxlen=n), 0

    def reset(self):
        self._action_queue.clear()

    def select_action(self, batch):
        import torch
        if len(self._action_queue) == 0:
            self.calls += 1
            lum = batch["observation.images.zed_cam_left"].mean(dim=(1, 2, 3))
            base = self.home[None, :] + 0.02 * torch.sin(batch["observation.state"] * 3.0)
            for k in range(self.n):
                a = base.clone()
                a[:, 1] += 0.05 * lum * (k + 1) / self.n
                a[:, 6] = a[:, 13] = 1.0
                self._action_queue.append(a)
        return self._action_queue.popleft()


@pytest.mark.gpu
def test_device_rollout_matches_reference_rollout_semantics_and_lazy_render_changes_nothing():
    import torch
    from av_aloha_b200 import observation
    from av_aloha_b200.env import GuidedVisionVectorEnv
    from oracle.oracle import HOME
    home = torch.as_tensor(HOME, dtype=torch.float32, device="cuda")
    res = {}
    for lazy in (False, True):
        env = GuidedVisionVectorEnv("slot_insertion", 4, cameras=["zed_cam_left"], observation_height=48, observation_width=64,
                                    max_episode_steps=12, seed=5)
        pol = _ChunkPolicy(5, home)
        launches0 = env._batch.launch_count
        res[lazy] = observation.rollout(env, pol, lazy_render=lazy, return_observations=not lazy)
        res[lazy]["launches"] = env._batch.launch_count - launches0
        assert pol.calls == 3                                  # steps 0, 5, 10
        env.close()
    r = res[False]
    assert tuple(r["action"].shape) == (4, 12, 21) and tuple(r["reward"].shape) == (4, 12)
    assert r["reward"].dtype == torch.float64 and r["done"].dtype == torch.bool and r["success"].dtype == torch.bool
    assert not bool(r["done"][:, :-1].any()) and bool(r["done"][:, -1].all())        # TimeLimit on the last step only
    assert tuple(r["observation"]["observation.images.zed_cam_left"].shape) == (4, 13, 3, 48, 64)
    assert tuple(r["observation"]["observation.state"].shape) == (4, 13, 21)
    for k in ("action", "reward", "done", "success"):
        assert torch.equal(res[True][k], r[k]), k             # skipping the unread frames changes nothing
    assert res[True]["launches"] < r["launches"]               # ... but renders 4 times instead of 13


@pytest.mark.gpu
def test_rerender_episode_equals_frame_by_frame_set_qpos():
    """f2: the batched re-render equals the reference's loop (set_qpos; get_obs) frame by frame, ragged last chunk included"""
    from av_aloha_b200 import capi, replay
    from av_aloha_b200.env import SlotInsertionEnv
    cams = ["overhead_cam", "wrist_cam_left"]
    env = SlotInsertionEnv(num_arms=3, cameras=cams, observation_height=48, observation_width=64)
    np.random.seed(3)
    env.reset()
    q0 = env._batch.get(capi.QPOS).cpu().numpy()[0]
    T = 5
    all_qpos = np.repeat(q0[None], T, axis=0)
    all_qpos[:, 0] += np.linspace(0.0, 0.6, T)             # left waist swings, wrist camera moves with it
    all_qpos[:, 8] += np.linspace(0.0, -0.4, T)
    out = replay.rerender_episode("SlotInsertion", all_qpos, cams, height=48, width=64, chunk=2)
    assert set(out) == set(cams)
    for t in range(T):
        env.set_qpos(all_qpos[t])
        px = env.get_obs()["pixels"]
        for c in cams:
            assert out[c].shape == (T, 48, 64, 3) and out[c].dtype == np.uint8
            assert np.array_equal(out[c][t], px[c]), (t, c)
    assert not np.array_equal(out["wrist_cam_left"][0], out["wrist_cam_left"][-1])
    assert replay.rerender_episode("SlotInsertion", all_qpos[:0], cams, height=48, width=64)["overhead_cam"].shape == (0, 48, 64, 3)
    with pytest.raises(ValueError):
        replay.rerender_episode("SlotInsertion", all_qpos[:, :10], cams)
    env.close()


@pytest.mark.gpu
def test_audit_rewards_equals_episode_by_episode_replay():
    """f3: all episodes at once == the reference's loop (set_qpos(all_qpos[0]); step_action; get_reward), one env at a time"""
    import torch
    from av_aloha_b200 import capi, model_io, replay, workload
    E, T = 5, 260                                            # the scripted grasp reaches reward >= 2 between steps 180 and 250
    obj = workload.sample_object_positions(E, 11)
    acts = workload.slot_insertion_script(300, obj, 11)[:T].transpose(1, 0, 2).copy()       # [E, T, 21]
    model = capi.Model(model_io.model_path("slot_insertion", 3), 0)
    first = np.empty((E, model.nq), np.float32)
    seq = np.zeros((E, T), np.int32)
    lengths = np.array([T, 100, T, 1, 0])
    for e in range(E):
        b = capi.Batch(model, 1, seed=0)
        b.reset(free_pos=obj[e][None])
        first[e] = b.get(capi.QPOS).cpu().numpy()[0]
        b.reset()
        b.set(capi.QPOS, first[e][None])
        b.forward()
        for t in range(int(lengths[e])):
            b.step(torch.as_tensor(acts[e, t][None], device="cuda"))
            seq[e, t] = int(b.get(capi.REWARD).item())
        b.close()
    out = replay.audit_rewards("slot_insertion", first, acts, lengths=lengths)
    for e in range(E):
        n = int(lengths[e])
        assert np.array_equal(out["rewards"][e, :n], seq[e, :n]), e
        assert int(out["episode_max"][e]) == (int(seq[e, :n].max()) if n else 0)
    assert out["episode_max"].max() >= 2                       # the scripted policy grasps the stick in episode 0 (step 198)
    # ... and against an independent implementation: the fp64 oracle replays episode 1 (100 steps) the way the reference's
    # check_dataset_reward.py does (set_qpos(all_qpos[0]); step_action; get_reward) -- same staged rewards, step by step
    from oracle.oracle import OracleEnv, OracleModel
    o = OracleEnv(OracleModel(model_io.model_path("slot_insertion", 3)))
    o.reset()
    o.qpos[:] = first[1]
    o.forward()
    ora = [o.step(acts[1, t].astype(np.float64)) for t in range(int(lengths[1]))]
    assert np.array_equal(out["rewards"][1, :len(ora)], np.array(ora, np.int32))
    assert out["not_max_reward_episodes"] == [int(i) for i in np.nonzero(~out["max_reward_reached"])[0]]
    assert 4 in out["not_max_reward_episodes"]                 # the empty episode can not have reached it
    with pytest.raises(ValueError):
        replay.audit_rewards("slot_insertion", first[:, :5], acts)
