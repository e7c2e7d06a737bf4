"""``.avm`` compiled-model container: a flat table-of-contents + raw arrays.

Layout (little endian):
  char[8]  magic "AVSIMMD1"
  u32      n_arrays
  n_arrays x { char[32] name; u32 dtype (0=f64,1=i32); u32 ndim; u32 shape[4]; u64 offset; u64 nbytes }
  raw data, every array 8-byte aligned, offsets from the start of the file

The C-ABI library (csrc/avsim_model.cpp) and the CPU oracle (oracle/avsim_oracle.c) both read this
file by array name; scalars are stored as 1-element arrays.
"""
from __future__ import annotations

import json
import os
import struct

import numpy as np

MAGIC = b"AVSIMMD1"
_ENTRY = struct.Struct("<32sII4IQQ")


def save_avm(path, m):
    names = sorted(m.keys())
    arrays = []
    for n in names:
        a = np.asarray(m[n])
        if a.dtype.kind in "iub":
            a = a.astype(np.int32)
            code = 1
        else:
            a = a.astype(np.float64)
            code = 0
        if a.ndim == 0:
            a = a.reshape(1)
        assert a.ndim <= 4 and len(n) < 32, n
        arrays.append((n, code, np.ascontiguousarray(a)))
    head = len(MAGIC) + 4 + _ENTRY.size * len(arrays)
    off = (head + 7) // 8 * 8
    toc, blobs = [], []
    for n, code, a in arrays:
        shape = list(a.shape) + [0] * (4 - a.ndim)
        toc.append(_ENTRY.pack(n.encode(), code, a.ndim, *shape, off, a.nbytes))
        pad = (-a.nbytes) % 8
        blobs.append(a.tobytes() + b"\0" * pad)
        off += a.nbytes + pad
    with open(path, "wb") as fh:
        fh.write(MAGIC + struct.pack("<I", len(arrays)) + b"".join(toc))
        fh.write(b"\0" * ((head + 7) // 8 * 8 - head))
        fh.write(b"".join(blobs))


def load_avm(path):
    with open(path, "rb") as fh:
        data = fh.read()
    assert data[:8] == MAGIC, f"{path}: bad magic"
    n = struct.unpack_from("<I", data, 8)[0]
    out = {}
    for i in range(n):
        name, code, ndim, s0, s1, s2, s3, off, nb = _ENTRY.unpack_from(data, 12 + i * _ENTRY.size)
        shape = (s0, s1, s2, s3)[:ndim]
        dt = np.float64 if code == 0 else np.int32
        out[name.rstrip(b"\0").decode()] = np.frombuffer(data, dtype=dt, count=nb // np.dtype(dt).itemsize,
                                                         offset=off).reshape(shape).copy()
    return out


MODEL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "models")


def model_path(task, num_arms):
    return os.path.join(MODEL_DIR, f"{task}_{num_arms}arms.avm")


def load_names(task, num_arms):
    with open(os.path.join(MODEL_DIR, f"{task}_{num_arms}arms.json")) as fh:
        return json.load(fh)
