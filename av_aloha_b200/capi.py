"""ctypes binding of libavsim.so (include/avsim.h) + thin torch-tensor helpers.

PyTorch is used for device memory and streams only; all arithmetic happens in the hand-written sm_100a kernels
behind the C-ABI.  There is no CPU fallback: importing this module without the built library, or creating a
batch without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AVSIM_LIB") or os.path.join(_HERE, "csrc", "libavsim.so")   # AVSIM_LIB: diagnostics build

# avsim_field
QPOS, QVEL, CTRL, WARMSTART, AGENT_POS, REWARD, SUCCESS, NCON, CONTACTS, STATUS, LATCH, QACC, XPOS, QFRC_BIAS, \
    QACC_SMOOTH, MASS_DIAG, ENV_CYCLES, FC_KEY, FC_N, FC_VAL, SOLVER_STAT = range(21)
MAX_CONTACTS = 64

SYMBOLS = [
    "avsim_model_load", "avsim_model_free", "avsim_model_dim", "avsim_create", "avsim_destroy", "avsim_set_options",
    "avsim_reset", "avsim_step", "avsim_forward", "avsim_get", "avsim_set", "avsim_step_host", "avsim_launch_count",
    "avsim_diffik", "avsim_gradik", "avsim_fk", "avsim_last_error", "avsim_stage_cycles", "avsim_render", "avsim_set_warmstart",
    "avsim_pixels_to_float", "avsim_jac", "avsim_set_solver", "avsim_transform", "avsim_render_ids", "avsim_launch_shape",
]
SOLVER_PGS, SOLVER_NEWTON = 0, 1


class DiffIKParams(C.Structure):
    _fields_ = [("k_pos", C.c_float), ("k_ori", C.c_float), ("damping", C.c_float), ("max_angvel", C.c_float),
                ("integration_dt", C.c_float), ("k_null", C.c_float * 7), ("q0", C.c_float * 7),
                ("iterations", C.c_int)]


class GradIKParams(C.Structure):
    _fields_ = [("step_size", C.c_float), ("min_cost_delta", C.c_float), ("position_weight", C.c_float),
                ("rotation_weight", C.c_float), ("position_threshold", C.c_float), ("rotation_threshold", C.c_float),
                ("max_pos_diff", C.c_float), ("max_rot_diff", C.c_float), ("joint_p", C.c_float),
                ("joint_center_weight", C.c_float * 7), ("joint_displacement_weight", C.c_float * 7),
                ("max_iterations", C.c_int)]


_lib = None


def load_library():
    """Load libavsim.so (fails loudly when it has not been built: run ``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: the CUDA extension has not been built (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    vp, cp, i32, u64 = C.c_void_p, C.c_char_p, C.c_int, C.c_uint64
    L.avsim_model_load.restype = vp; L.avsim_model_load.argtypes = [cp, i32]
    L.avsim_model_free.argtypes = [vp]
    L.avsim_model_dim.argtypes = [vp, cp]
    L.avsim_create.restype = vp; L.avsim_create.argtypes = [vp, i32, u64, vp]
    L.avsim_destroy.argtypes = [vp]
    L.avsim_set_options.argtypes = [vp, i32, i32, i32]
    L.avsim_set_warmstart.argtypes = [vp, i32]
    L.avsim_set_solver.argtypes = [vp, i32, i32, i32, C.c_float]
    L.avsim_reset.argtypes = [vp, vp, vp]
    L.avsim_step.argtypes = [vp, vp, i32]
    L.avsim_forward.argtypes = [vp]
    L.avsim_get.argtypes = [vp, i32, vp]
    L.avsim_set.argtypes = [vp, i32, vp]
    L.avsim_step_host.argtypes = [vp, vp, i32, vp, vp, vp]
    L.avsim_render.argtypes = [vp, C.POINTER(C.c_int), i32, i32, i32, vp]
    L.avsim_render_ids.argtypes = [vp, C.POINTER(C.c_int), i32, i32, i32, vp]
    L.avsim_pixels_to_float.argtypes = [vp, C.c_int64, i32, i32, vp, i32, vp]
    L.avsim_launch_count.restype = C.c_int64; L.avsim_launch_count.argtypes = [vp]
    L.avsim_launch_shape.argtypes = [vp, C.POINTER(C.c_int)]
    L.avsim_diffik.argtypes = [vp, i32, vp, vp, vp, i32, C.POINTER(DiffIKParams), vp, vp]
    L.avsim_gradik.argtypes = [vp, i32, vp, vp, vp, i32, C.POINTER(GradIKParams), vp, vp]
    L.avsim_fk.argtypes = [vp, i32, vp, i32, vp, vp]
    L.avsim_jac.argtypes = [vp, i32, vp, i32, vp, vp]
    L.avsim_transform.argtypes = [i32, vp, vp, i32, C.c_double, C.c_double, vp, i32, vp]
    L.avsim_last_error.restype = cp
    L.avsim_stage_cycles.argtypes = [vp, i32, i32]
    _lib = L
    return L


class AvsimError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise AvsimError(f"avsim error {rc}: {load_library().avsim_last_error().decode()}")


def pixels_to_float(img, out=None):
    """uint8 CUDA tensor [..., H, W, 3] -> float32 CUDA tensor [..., 3, H, W] in [0, 1] (avsim_pixels_to_float: the image half
    of lerobot's preprocess_observation, utils.py:37-50, without leaving the device).  Runs on torch's current stream."""
    import torch

    if not (torch.is_tensor(img) and img.is_cuda and img.dtype == torch.uint8 and img.is_contiguous()):
        raise ValueError("pixels_to_float: expected a contiguous uint8 CUDA tensor")
    if img.dim() < 3 or img.shape[-1] != 3:
        raise ValueError(f"pixels_to_float: expected channel-last images [..., H, W, 3], got {tuple(img.shape)}")
    lead, (H, W) = tuple(img.shape[:-3]), img.shape[-3:-1]
    if out is None:
        out = torch.empty(lead + (3, H, W), dtype=torch.float32, device=img.device)
    elif not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == lead + (3, H, W)):
        raise ValueError("pixels_to_float: bad output tensor")
    n = int(np.prod(lead)) if lead else 1
    stream = torch.cuda.current_stream(img.device).cuda_stream
    check(load_library().avsim_pixels_to_float(C.c_void_p(img.data_ptr()), C.c_int64(n), int(H), int(W), C.c_void_p(out.data_ptr()),
                                               img.device.index or 0, C.c_void_p(stream)))
    return out


class Model:
    """Compiled model resident on one CUDA device (replaces mjcf.from_path + Physics.from_mjcf_model, env.py:53-56)."""

    def __init__(self, avm_path, device=0):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("av_aloha_b200 needs a CUDA device; there is no CPU path")
        self.lib = load_library()
        self.device = int(device)
        self.avm_path = os.fspath(avm_path)
        self._tables = None
        self.ptr = self.lib.avsim_model_load(os.fsencode(avm_path), self.device)
        if not self.ptr:
            raise AvsimError(self.lib.avsim_last_error().decode())
        d = lambda n: self.lib.avsim_model_dim(self.ptr, n.encode())
        self.nq, self.nv, self.nu, self.nbody, self.ngeom = d("nq"), d("nv"), d("nu"), d("nbody"), d("ngeom")
        self.njoints, self.nfree, self.max_reward, self.task_id, self.num_arms = (
            d("njoints"), d("nfree"), d("max_reward"), d("task_id"), d("num_arms"))

    def table(self, name):
        """Host copy of a compiled-model array (names as in mjcf_compile.py), e.g. 'ik_range', 'reset_lo'."""
        if self._tables is None:
            from . import model_io
            self._tables = model_io.load_avm(self.avm_path)
        return self._tables[name]

    def ik_range(self, arm):
        n = int(self.table("ik_ndof")[arm])
        return self.table("ik_range")[arm, :n]

    def __del__(self):
        try:
            if self.ptr:
                self.lib.avsim_model_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Batch:
    """B environments in lockstep on one GPU (replaces B GuidedVisionEnv instances under SyncVectorEnv)."""

    def __init__(self, model: Model, num_envs: int, seed: int = 0, stream=None):
        import torch

        self.torch = torch
        self.model = model
        self.lib = model.lib
        self.num_envs = int(num_envs)
        self.dev = torch.device("cuda", model.device)
        self.stream = stream
        sp = C.c_void_p(stream.cuda_stream) if stream is not None else None
        self.ptr = self.lib.avsim_create(model.ptr, self.num_envs, C.c_uint64(seed), sp)
        if not self.ptr:
            raise AvsimError(self.lib.avsim_last_error().decode())
        m = model
        self._shape = {
            QPOS: (m.nq, torch.float32), QVEL: (m.nv, torch.float32), CTRL: (m.nu, torch.float32),
            WARMSTART: (m.nv, torch.float32), AGENT_POS: (m.njoints, torch.float32), REWARD: (None, torch.int32),
            SUCCESS: (None, torch.int32), NCON: (None, torch.int32), CONTACTS: ((MAX_CONTACTS, 16), torch.float32),
            STATUS: (None, torch.int32), LATCH: (None, torch.int32), QACC: (m.nv, torch.float32),
            XPOS: ((m.nbody, 3), torch.float32), QFRC_BIAS: (m.nv, torch.float32), QACC_SMOOTH: (m.nv, torch.float32),
            MASS_DIAG: (m.nv, torch.float32), ENV_CYCLES: (None, torch.int64),
            FC_KEY: (MAX_CONTACTS + 20, torch.int32), FC_N: (2, torch.int32), FC_VAL: (MAX_CONTACTS * 6 + 20, torch.float32),
            SOLVER_STAT: (4, torch.float32),
        }

    def close(self):
        if getattr(self, "ptr", None):
            self.lib.avsim_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, solver_iters=20, noslip_iters=-1, multiccd=-1):
        check(self.lib.avsim_set_options(self.ptr, solver_iters, noslip_iters, multiccd))

    def set_solver(self, solver="newton", max_iter=0, ls_iter=0, tol=0.0):
        """'newton' (default of a new batch): Newton on the primal, the reference's solver; 'pgs': fixed-sweep block Gauss-Seidel
        on the dual (sweep count from set_options).  max_iter / ls_iter / tol <= 0 keep the current Newton settings."""
        code = {"pgs": SOLVER_PGS, "newton": SOLVER_NEWTON}[solver] if isinstance(solver, str) else int(solver)
        check(self.lib.avsim_set_solver(self.ptr, code, int(max_iter), int(ls_iter), float(tol)))

    def set_warmstart(self, mode):
        """1: MuJoCo-style warm start from the previous qacc (default); 2: per-constraint force cache."""
        check(self.lib.avsim_set_warmstart(self.ptr, int(mode)))

    def reset(self, mask=None, free_pos=None):
        t = self.torch
        mp = fp = None
        if mask is not None:
            self._mask = t.as_tensor(mask, dtype=t.uint8, device=self.dev).contiguous()
            mp = C.c_void_p(self._mask.data_ptr())
        if free_pos is not None:
            if not t.is_tensor(free_pos):
                free_pos = np.asarray(free_pos, dtype=np.float32)
            self._fp = t.as_tensor(free_pos, dtype=t.float32, device=self.dev).contiguous()
            assert self._fp.shape == (self.num_envs, self.model.nfree, 3), self._fp.shape
            fp = C.c_void_p(self._fp.data_ptr())
        check(self.lib.avsim_reset(self.ptr, mp, fp))

    def step(self, action, nsubsteps=20):
        """action: float32 CUDA tensor [B, njoints] (stays on the device)."""
        assert action.is_cuda and action.dtype == self.torch.float32 and action.is_contiguous()
        assert tuple(action.shape) == (self.num_envs, self.model.njoints), action.shape
        check(self.lib.avsim_step(self.ptr, C.c_void_p(action.data_ptr()), nsubsteps))

    def step_host(self, action_np, nsubsteps=20, agent_pos_out=None, reward_out=None, status_out=None):
        """Host-buffer path: numpy float32 [B, njoints] in, numpy agent_pos / reward (/ status when status_out is given) out
        (copies inside the call)."""
        a = np.ascontiguousarray(action_np, dtype=np.float32)
        assert a.shape == (self.num_envs, self.model.njoints), a.shape
        if agent_pos_out is None:
            agent_pos_out = np.empty((self.num_envs, self.model.njoints), np.float32)
        if reward_out is None:
            reward_out = np.empty((self.num_envs,), np.int32)
        check(self.lib.avsim_step_host(self.ptr, a.ctypes.data_as(C.c_void_p), nsubsteps,
                                       agent_pos_out.ctypes.data_as(C.c_void_p), reward_out.ctypes.data_as(C.c_void_p),
                                       status_out.ctypes.data_as(C.c_void_p) if status_out is not None else None))
        return agent_pos_out, reward_out

    def forward(self):
        check(self.lib.avsim_forward(self.ptr))

    def render(self, cam_ids, height=480, width=640, out=None, ids=False):
        """uint8 CUDA tensor [B, ncam, H, W, 3] of the current state (cam_ids: indices into the model's camera list).  ids=True:
        the geom index each pixel sees instead of a colour (255 = background; test hook)."""
        t = self.torch
        cids = (C.c_int * len(cam_ids))(*[int(c) for c in cam_ids])
        if out is None:
            out = t.empty((self.num_envs, len(cam_ids), height, width, 3), dtype=t.uint8, device=self.dev)
        fn = self.lib.avsim_render_ids if ids else self.lib.avsim_render
        check(fn(self.ptr, cids, len(cam_ids), height, width, C.c_void_p(out.data_ptr())))
        return out

    def get(self, field, out=None):
        t = self.torch
        w, dt = self._shape[field]
        shape = (self.num_envs,) if w is None else (self.num_envs,) + (w if isinstance(w, tuple) else (w,))
        if out is None:
            out = t.empty(shape, dtype=dt, device=self.dev)
        check(self.lib.avsim_get(self.ptr, field, C.c_void_p(out.data_ptr())))
        return out

    def set(self, field, value):
        t = self.torch
        w, dt = self._shape[field]
        shape = (self.num_envs,) if w is None else (self.num_envs, w)
        v = t.as_tensor(value, dtype=dt, device=self.dev).contiguous()
        assert tuple(v.shape) == shape, (v.shape, shape)
        self._keep = v
        check(self.lib.avsim_set(self.ptr, field, C.c_void_p(v.data_ptr())))

    @property
    def launch_count(self):
        return int(self.lib.avsim_launch_count(self.ptr))

    @property
    def launch_shape(self):
        """{'split', 'groups', 'warps', 'env_warps', 'solve_warps', 'sms'}: how avsim_create laid this batch out on the GPU"""
        out = (C.c_int * 6)()
        check(self.lib.avsim_launch_shape(self.ptr, out))
        return dict(zip(("split", "groups", "warps", "env_warps", "solve_warps", "sms"), (int(v) for v in out)))
