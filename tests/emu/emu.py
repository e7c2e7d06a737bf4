"""ctypes front end of the warp-emulation debug harness (tests only; see cuda_emu.h)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "..", "..", "av_aloha_b200", "csrc")
_LIB = os.path.join(_HERE, "libavsim_emu.so")


def build():
    srcs = [os.path.join(_HERE, f) for f in ("avsim_emu.cpp", "cuda_emu.h")] + [
        os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-o", _LIB,
                               os.path.join(_HERE, "avsim_emu.cpp")])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emu_create.restype = C.c_void_p
        _lib.emu_create.argtypes = [C.c_char_p, C.c_int]
        _lib.emu_destroy.argtypes = [C.c_void_p]
        _lib.emu_set_options.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.emu_set_warmstart.argtypes = [C.c_void_p, C.c_int]
        _lib.emu_set_solver.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float]
        _lib.emu_dim.argtypes = [C.c_void_p, C.c_int]
        _lib.emu_f.restype = C.POINTER(C.c_float)
        _lib.emu_f.argtypes = [C.c_void_p, C.c_int]
        _lib.emu_i.restype = C.POINTER(C.c_int)
        _lib.emu_i.argtypes = [C.c_void_p, C.c_int]
        _lib.emu_forward.argtypes = [C.c_void_p]
        _lib.emu_step.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
        _lib.emu_reset.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        _lib.emu_reset_masked.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_float)]
    return _lib


_F = dict(qpos=0, qvel=1, ctrl=2, warm=3, agent_pos=4, contacts=5, qacc=6, xpos=7, qfrc_bias=8, qacc_smooth=9,
          mass_diag=10, nw_stat=11)
_I = dict(reward=0, status=1, latch=2, ncon=3, episode=4)


class EmuBatch:
    def __init__(self, path, num_envs=1):
        self.ptr = lib().emu_create(path.encode(), num_envs)
        assert self.ptr
        self.B = num_envs
        d = lambda k: lib().emu_dim(self.ptr, k)
        self.nq, self.nv, self.nu, self.nbody, self.nj, self.nfree = (d(k) for k in range(6))
        self._w = dict(qpos=self.nq, qvel=self.nv, ctrl=self.nu, warm=self.nv, agent_pos=self.nj, contacts=64 * 16,
                       qacc=self.nv, xpos=3 * self.nbody, qfrc_bias=self.nv, qacc_smooth=self.nv, mass_diag=self.nv, nw_stat=4)

    def __getattr__(self, name):
        if name in _F:
            p = lib().emu_f(self.ptr, _F[name])
            return np.ctypeslib.as_array(p, shape=(self.B, self._w[name]))
        if name in _I:
            p = lib().emu_i(self.ptr, _I[name])
            return np.ctypeslib.as_array(p, shape=(self.B,))
        raise AttributeError(name)

    def set_options(self, iters=50, noslip=-1, multiccd=-1):
        lib().emu_set_options(self.ptr, iters, noslip, multiccd)

    def set_warmstart(self, mode):
        lib().emu_set_warmstart(self.ptr, mode)

    def set_solver(self, solver="newton", max_iter=30, ls_iter=20, tol=1e-6):
        lib().emu_set_solver(self.ptr, {"pgs": 0, "newton": 1}[solver], max_iter, ls_iter, tol)

    def set_split(self, split):
        """Newton only: True = the split pipeline (substep + solve kernels), False = the fused step kernel"""
        lib().emu_set_split.argtypes = [C.c_void_p, C.c_int]
        lib().emu_set_split(self.ptr, int(bool(split)))

    def forward(self):
        lib().emu_forward(self.ptr)

    def step(self, action, nsub=20):
        a = np.ascontiguousarray(action, dtype=np.float32)
        lib().emu_step(self.ptr, a.ctypes.data_as(C.POINTER(C.c_float)), nsub)

    def reset(self, free_pos=None, mask=None):
        fp = mk = None
        if free_pos is not None:
            self._fp = np.ascontiguousarray(free_pos, dtype=np.float32)
            fp = self._fp.ctypes.data_as(C.POINTER(C.c_float))
        if mask is not None:
            self._mk = np.ascontiguousarray(mask, dtype=np.uint8)
            mk = self._mk.ctypes.data_as(C.POINTER(C.c_uint8))
        lib().emu_reset_masked(self.ptr, mk, fp)


def emu_reward_from_pairs(eb, pairs, latch=0):
    """(reward, latch) of stage_reward (CUDA source, emulated) on an explicit contact list [(geom1, geom2), ...]"""
    arr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1))
    out = C.c_int(0)
    fn = lib().emu_reward_from_pairs
    fn.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int)]
    r = fn(eb.ptr, arr.ctypes.data_as(C.POINTER(C.c_int)), len(arr) // 2, int(latch), C.byref(out))
    return int(r), int(out.value)


# ---- IK kernels through the emulator (same parameter structs as the C-ABI)
def _ik_setup():
    from av_aloha_b200.capi import DiffIKParams, GradIKParams
    L = lib()
    fp = C.POINTER(C.c_float)
    L.emu_fk.argtypes = [C.c_void_p, C.c_int, fp, C.c_int, fp]
    L.emu_jac.argtypes = [C.c_void_p, C.c_int, fp, C.c_int, fp]
    L.emu_diffik.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, C.c_int, C.POINTER(DiffIKParams), fp]
    L.emu_gradik.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, C.c_int, C.POINTER(GradIKParams), fp]
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def emu_fk(eb, arm, q):
    L = _ik_setup()
    q = np.ascontiguousarray(q, np.float32)
    out = np.zeros((len(q), 16), np.float32)
    L.emu_fk(eb.ptr, arm, _fp(q), len(q), _fp(out))
    return out.reshape(-1, 4, 4)


def emu_jac(eb, arm, q):
    L = _ik_setup()
    q = np.ascontiguousarray(q, np.float32)
    out = np.zeros((len(q), 6, q.shape[1]), np.float32)
    L.emu_jac(eb.ptr, arm, _fp(q), len(q), _fp(out))
    return out


def emu_diffik(eb, arm, q, pos, quat, params):
    L = _ik_setup()
    q, pos, quat = (np.ascontiguousarray(x, np.float32) for x in (q, pos, quat))
    out = np.zeros_like(q)
    L.emu_diffik(eb.ptr, arm, _fp(q), _fp(pos), _fp(quat), len(q), C.byref(params), _fp(out))
    return out


def emu_gradik(eb, arm, q, pos, quat, params):
    L = _ik_setup()
    q, pos, quat = (np.ascontiguousarray(x, np.float32) for x in (q, pos, quat))
    out = np.zeros_like(q)
    L.emu_gradik(eb.ptr, arm, _fp(q), _fp(pos), _fp(quat), len(q), C.byref(params), _fp(out))
    return out


def emu_transform(op, a, wa, wout, b=None, p0=0.0, p1=0.0):
    """avsim_transform_kernel (the CUDA source, emulated) on n items: a [n, wa] (b [n, wb]) -> [n, wout] float64"""
    L = lib()
    dp = C.POINTER(C.c_double)
    L.emu_transform.argtypes = [C.c_int, dp, dp, C.c_int, C.c_double, C.c_double, dp]
    a = np.ascontiguousarray(a, np.float64).reshape(-1, wa)
    out = np.zeros((len(a), wout))
    bp = None
    if b is not None:
        b = np.ascontiguousarray(b, np.float64).reshape(len(a), -1)
        bp = b.ctypes.data_as(dp)
    L.emu_transform(op, a.ctypes.data_as(dp), bp, len(a), p0, p1, out.ctypes.data_as(dp))
    return out
