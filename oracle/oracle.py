"""ctypes front end of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see avsim_oracle.c header).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never
by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libavsim_oracle.so")

# reference constants.py:26-28
LEFT_ARM_POSE = [0, -0.082, 1.06, 0, -0.953, 0, 0.02239]
RIGHT_ARM_POSE = [0, -0.082, 1.06, 0, -0.953, 0, 0.02239]
MIDDLE_ARM_POSE = [0, -0.8, 0.8, 0, 0.5, 0, 0]
HOME = np.array(LEFT_ARM_POSE + RIGHT_ARM_POSE + MIDDLE_ARM_POSE, dtype=np.float64)


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("avsim_oracle.c", "avsim_oracle_collide.inc")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.ora_model_load.restype = C.c_void_p
        L.ora_model_load.argtypes = [C.c_char_p]
        L.ora_model_free.argtypes = [C.c_void_p]
        L.ora_data_new.restype = C.c_void_p
        L.ora_data_new.argtypes = [C.c_void_p]
        L.ora_data_free.argtypes = [C.c_void_p]
        L.ora_set_options.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
        L.ora_set_solver.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ora_inject_contacts.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        for fn in ("ora_forward", "ora_substep"):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.ora_position_pass.argtypes = [C.c_void_p]
        L.ora_env_step.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        L.ora_reset.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ora_agent_pos.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        for fn in ("ora_nq", "ora_nv", "ora_nu", "ora_nbody", "ora_ngeom", "ora_ncon", "ora_nefc", "ora_solver_iters",
                   "ora_reward"):
            getattr(L, fn).argtypes = [C.c_void_p]
        for fn in ("ora_qpos", "ora_qvel", "ora_ctrl", "ora_qacc_warmstart", "ora_qacc", "ora_qacc_smooth",
                   "ora_qfrc_bias", "ora_qfrc_actuator", "ora_qfrc_constraint", "ora_xpos", "ora_xquat", "ora_gpos",
                   "ora_gmat", "ora_M", "ora_efc_force", "ora_efc_aref", "ora_efc_R", "ora_efc_J"):
            getattr(L, fn).restype = C.POINTER(C.c_double)
            getattr(L, fn).argtypes = [C.c_void_p]
        L.ora_latch.restype = C.POINTER(C.c_int)
        L.ora_latch.argtypes = [C.c_void_p]
        L.ora_contact_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.ora_collide_pair.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                                       C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double),
                                       C.c_int]
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleModel:
    def __init__(self, path):
        self.path = path
        self.ptr = lib().ora_model_load(path.encode())
        if not self.ptr:
            raise FileNotFoundError(path)
        L = lib()
        self.nq, self.nv, self.nu = L.ora_nq(self.ptr), L.ora_nv(self.ptr), L.ora_nu(self.ptr)
        self.nbody, self.ngeom = L.ora_nbody(self.ptr), L.ora_ngeom(self.ptr)

    def collide_pair(self, g1, pos1, mat1, g2, pos2, mat2, multiccd=0):
        out = np.zeros((8, 13))
        args = [np.ascontiguousarray(x, dtype=np.float64) for x in (pos1, mat1, pos2, mat2)]
        n = lib().ora_collide_pair(self.ptr, g1, _dp(args[0]), _dp(args[1]), g2, _dp(args[2]), _dp(args[3]), multiccd,
                                   _dp(out), 8)
        return out[:n]


class OracleEnv:
    """One fp64 environment."""

    def __init__(self, model: OracleModel):
        self.model = model
        self.ptr = lib().ora_data_new(model.ptr)
        self.num_joints = 21 if model.nu == 21 and self._num_arms() == 3 else 14

    def _num_arms(self):
        from av_aloha_b200 import model_io  # data loader only
        return int(model_io.load_avm(self.model.path)["num_arms"][0])

    def __del__(self):
        try:
            lib().ora_data_free(self.ptr)
        except Exception:
            pass

    def _vec(self, name, n):
        p = getattr(lib(), name)(self.ptr)
        return np.ctypeslib.as_array(p, shape=(n,))

    qpos = property(lambda s: s._vec("ora_qpos", s.model.nq))
    qvel = property(lambda s: s._vec("ora_qvel", s.model.nv))
    ctrl = property(lambda s: s._vec("ora_ctrl", s.model.nu))
    qacc_warmstart = property(lambda s: s._vec("ora_qacc_warmstart", s.model.nv))
    qacc = property(lambda s: s._vec("ora_qacc", s.model.nv))
    qacc_smooth = property(lambda s: s._vec("ora_qacc_smooth", s.model.nv))
    qfrc_bias = property(lambda s: s._vec("ora_qfrc_bias", s.model.nv))
    qfrc_actuator = property(lambda s: s._vec("ora_qfrc_actuator", s.model.nv))
    qfrc_constraint = property(lambda s: s._vec("ora_qfrc_constraint", s.model.nv))
    xpos = property(lambda s: s._vec("ora_xpos", s.model.nbody * 3).reshape(-1, 3))
    xquat = property(lambda s: s._vec("ora_xquat", s.model.nbody * 4).reshape(-1, 4))
    gpos = property(lambda s: s._vec("ora_gpos", s.model.ngeom * 3).reshape(-1, 3))
    gmat = property(lambda s: s._vec("ora_gmat", s.model.ngeom * 9).reshape(-1, 3, 3))
    ncon = property(lambda s: lib().ora_ncon(s.ptr))
    nefc = property(lambda s: lib().ora_nefc(s.ptr))
    solver_iters = property(lambda s: lib().ora_solver_iters(s.ptr))
    reward = property(lambda s: lib().ora_reward(s.ptr))

    @property
    def M(self):
        stride = lib().ora_M_stride()
        return self._vec("ora_M", stride * stride).reshape(stride, stride)[: self.model.nv, : self.model.nv]

    @property
    def efc_force(self):
        return self._vec("ora_efc_force", self.nefc)

    @property
    def efc_J(self):
        stride = lib().ora_M_stride()
        return self._vec("ora_efc_J", self.nefc * stride).reshape(self.nefc, stride)[:, : self.model.nv]

    @property
    def efc_aref(self):
        return self._vec("ora_efc_aref", self.nefc)

    @property
    def efc_R(self):
        return self._vec("ora_efc_R", self.nefc)

    @property
    def latch(self):
        return lib().ora_latch(self.ptr)[0]

    def contacts(self):
        out = np.zeros((self.ncon, 22))
        for c in range(self.ncon):
            lib().ora_contact_get(self.ptr, c, _dp(out[c]))
        return out

    def reward_from_pairs(self, pairs, latch=0):
        """(reward, latch) of an explicit contact list [(geom1, geom2), ...] -- test hook for the reward tables"""
        arr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1))
        out = C.c_int(0)
        fn = lib().ora_reward_from_pairs
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int)]
        r = fn(self.ptr, arr.ctypes.data_as(C.POINTER(C.c_int)), len(arr) // 2, int(latch), C.byref(out))
        return int(r), int(out.value)

    def set_options(self, max_iter=3000, tol=1e-14, noslip_iter=-1, multiccd=-1, warmstart=1):
        lib().ora_set_options(self.ptr, max_iter, tol, noslip_iter, multiccd, warmstart)

    def set_solver(self, solver="pgs", newton_ls=0):
        """'pgs': block projected Gauss-Seidel on the dual; 'newton': Newton on the primal (the reference's solver)"""
        lib().ora_set_solver(self.ptr, {"pgs": 0, "newton": 1}[solver], int(newton_ls))

    def inject_contacts(self, records):
        """Use this contact list [n][9] = (dist, pos 3, normal 3, geom1, geom2) instead of the narrowphase (None: back to the
        narrowphase).  Lets a test check the constraint solve on exactly the contacts another implementation found."""
        if records is None:
            lib().ora_inject_contacts(self.ptr, -1, None)
            return
        r = np.ascontiguousarray(records, dtype=np.float64).reshape(-1, 9)
        lib().ora_inject_contacts(self.ptr, len(r), _dp(r))

    def reset(self, free_pos=None, arm_pose=HOME):
        ap = np.ascontiguousarray(arm_pose, dtype=np.float64)
        fp = None if free_pos is None else np.ascontiguousarray(free_pos, dtype=np.float64)
        lib().ora_reset(self.ptr, _dp(ap), _dp(fp) if fp is not None else None)

    def forward(self):
        return lib().ora_forward(self.ptr)

    def substep(self):
        return lib().ora_substep(self.ptr)

    def position_pass(self):
        lib().ora_position_pass(self.ptr)

    def step(self, action, nsub=20):
        a = np.ascontiguousarray(action, dtype=np.float64)
        return lib().ora_env_step(self.ptr, _dp(a), nsub)

    def agent_pos(self):
        out = np.zeros(21)
        lib().ora_agent_pos(self.ptr, _dp(out))
        return out
