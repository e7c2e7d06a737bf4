"""The reference's SO(3) / SE(3) helpers (data_collection_scripts/transform_utils.py) as batched device operators.

Same names, argument order and conventions as the reference -- quaternions are (x, y, z, w) inside this module, and
``xyzw_to_wxyz`` / ``wxyz_to_xyzw`` convert at the API boundary like the reference's callers do -- but every function also takes a
leading batch axis, and the arithmetic runs in `avsim_transform_kernel` (csrc/avsim_ik.cuh: the very device functions the DiffIK /
GradIK kernels call), fp64 like the reference.  Inputs may be numpy arrays (results come back as float64 numpy) or CUDA float64
tensors (results stay on the device).  There is no CPU path: without the CUDA library the calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

MAT2QUAT, QUAT2MAT, QUAT2AXISANGLE, AXISANGLE2QUAT, ANGULAR_ERROR, LIMIT_POSE, EXP2MAT, ADJOINT, WITHIN_POSE = range(9)


def _run(op, a, wa, wout, b=None, wb=0, p0=0.0, p1=0.0, device=0):
    import torch

    as_np = not hasattr(a, "is_cuda")
    dev = torch.device("cuda", device) if as_np else a.device
    ta = torch.as_tensor(np.asarray(a, np.float64) if as_np else a, dtype=torch.float64, device=dev)
    single = ta.numel() == wa
    ta = ta.reshape(-1, wa).contiguous()
    tb = None
    if b is not None:
        tb = torch.as_tensor(np.asarray(b, np.float64) if not hasattr(b, "is_cuda") else b, dtype=torch.float64, device=dev)
        tb = tb.reshape(-1, wb).contiguous()
        if len(tb) != len(ta):
            raise ValueError("operands must have the same leading dimension")
    out = torch.empty((len(ta), wout), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    capi.check(capi.load_library().avsim_transform(op, C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()) if tb is not None else None,
                                                   len(ta), float(p0), float(p1), C.c_void_p(out.data_ptr()), dev.index or 0,
                                                   C.c_void_p(stream)))
    if as_np:
        out = out.cpu().numpy()
    return (out[0] if single else out), single


def mat2quat(rmat):
    """3x3 rotation -> (x, y, z, w), w >= 0 (transform_utils.py:9-49)"""
    return _run(MAT2QUAT, rmat, 9, 4)[0]


def quat2mat(quaternion):
    """(x, y, z, w) -> 3x3, with the reference's float32 round trip (transform_utils.py:52-79)"""
    out, single = _run(QUAT2MAT, quaternion, 4, 9)
    return out.reshape(3, 3) if single else out.reshape(-1, 3, 3)


def quat2axisangle(quat):
    return _run(QUAT2AXISANGLE, quat, 4, 3)[0]


def axisangle2quat(vec):
    return _run(AXISANGLE2QUAT, vec, 3, 4)[0]


def angular_error(desired, current):
    """0.5 * sum_k current[:, k] x desired[:, k] (transform_utils.py:183-194)"""
    return _run(ANGULAR_ERROR, desired, 9, 3, b=current, wb=9)[0]


def _pose12(pos, mat):
    if hasattr(pos, "is_cuda"):
        import torch
        return torch.cat([pos.reshape(-1, 3), mat.reshape(-1, 9)], dim=1)
    return np.concatenate([np.asarray(pos, np.float64).reshape(-1, 3), np.asarray(mat, np.float64).reshape(-1, 9)], axis=1)


def limit_pose(current_xpos, current_xmat, target_xpos, target_xmat, max_pos_diff=0.1, max_rot_diff=0.1):
    """target pulled to within max_pos_diff / max_rot_diff of the current pose (transform_utils.py:263-287) -> (xpos, xmat)"""
    single = np.ndim(current_xpos) == 1
    out, _ = _run(LIMIT_POSE, _pose12(current_xpos, current_xmat), 12, 12, b=_pose12(target_xpos, target_xmat), wb=12,
                  p0=max_pos_diff, p1=max_rot_diff)
    out = out.reshape(-1, 12)
    pos, mat = out[:, :3], out[:, 3:].reshape(-1, 3, 3)
    return (pos[0], mat[0]) if single else (pos, mat)


def within_pose_threshold(current_xpos, current_xmat, target_xpos, target_xmat, position_threshold=0.001, rotation_threshold=0.001):
    single = np.ndim(current_xpos) == 1
    out, _ = _run(WITHIN_POSE, _pose12(current_xpos, current_xmat), 12, 1, b=_pose12(target_xpos, target_xmat), wb=12,
                  p0=position_threshold, p1=rotation_threshold)
    out = out.reshape(-1) > 0.5
    return bool(out[0]) if single else out


def exp2mat(w, v, theta):
    """matrix exponential of the screw (w, v) by theta -> 4x4 (transform_utils.py:239-261)"""
    single = np.ndim(w) == 1
    x = np.concatenate([np.asarray(w, np.float64).reshape(-1, 3), np.asarray(v, np.float64).reshape(-1, 3),
                        np.asarray(theta, np.float64).reshape(-1, 1)], axis=1)
    out, _ = _run(EXP2MAT, x, 7, 16)
    out = out.reshape(-1, 4, 4)
    return out[0] if single else out


def exp2rot(w, theta):
    """Rodrigues' formula (transform_utils.py:222-237): the rotation block of exp2mat"""
    return exp2mat(w, np.zeros_like(np.asarray(w, np.float64)), theta)[..., :3, :3]


def adjoint(T):
    single = np.ndim(T) == 2
    out, _ = _run(ADJOINT, np.asarray(T, np.float64).reshape(-1, 16), 16, 36)
    out = out.reshape(-1, 6, 6)
    return out[0] if single else out


def pose2mat(pos, quat):
    """(pos, (x, y, z, w)) -> 4x4 (transform_utils.py:136-151)"""
    single = np.ndim(pos) == 1
    R = np.asarray(quat2mat(quat)).reshape(-1, 3, 3)
    T = np.tile(np.eye(4), (len(R), 1, 1))
    T[:, :3, :3] = R
    T[:, :3, 3] = np.asarray(pos, np.float64).reshape(-1, 3)
    return T[0] if single else T


def mat2pose(homo_pose_mat):
    """4x4 -> (pos, (x, y, z, w)) (transform_utils.py:153-157)"""
    T = np.asarray(homo_pose_mat, np.float64)
    single = T.ndim == 2
    T = T.reshape(-1, 4, 4)
    q = np.asarray(mat2quat(np.ascontiguousarray(T[:, :3, :3]))).reshape(-1, 4)
    return (T[0, :3, 3], q[0]) if single else (T[:, :3, 3], q)


def xyzw_to_wxyz(quat):
    q = np.asarray(quat)
    return np.concatenate([q[..., 3:4], q[..., 0:3]], axis=-1)


def wxyz_to_xyzw(quat):
    q = np.asarray(quat)
    return np.concatenate([q[..., 1:4], q[..., 0:1]], axis=-1)


def skew_sym(x):
    x1, x2, x3 = np.asarray(x, np.float64).ravel()
    return np.array([[0, -x3, x2], [x3, 0, -x1], [-x2, x1, 0]])
