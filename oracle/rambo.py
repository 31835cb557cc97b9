"""TEST INFRASTRUCTURE -- ctypes front-end of oracle/rambo_oracle.c (numpy in, numpy out).

The C file is the restatement of the reference's RAMBO map; this module only loads it.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librambo_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "rambo_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "memflow_b200.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(_SO) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def forward(desc, r):
    """r [N,3n-2] float64 -> momenta [N,n+2,4], weight, logw, x1, x2 (float64), flags int32."""
    r = np.ascontiguousarray(r, dtype=np.float64)
    N, n = r.shape[0], desc.n_final
    assert r.shape[1] == 3 * n - 2
    mom = np.empty((N, n + 2, 4)); w = np.empty(N); lw = np.empty(N)
    x1 = np.empty(N); x2 = np.empty(N); fl = np.zeros(N, dtype=np.int32)
    lib().oracle_rambo_forward(ctypes.byref(desc), _p(r), ctypes.c_int64(N), _p(mom), _p(w),
                               _p(lw), _p(x1), _p(x2), _p(fl))
    return mom, w, lw, x1, x2, fl


def inverse(desc, mom, x1, x2, perm=None):
    """mom [N,n,4], x1, x2 -> r [N,3n-2], prob [N], flags."""
    mom = np.ascontiguousarray(mom, dtype=np.float64)
    x1 = np.ascontiguousarray(x1, dtype=np.float64); x2 = np.ascontiguousarray(x2, dtype=np.float64)
    N, n = mom.shape[0], desc.n_final
    assert mom.shape[1:] == (n, 4)
    r = np.empty((N, 3 * n - 2)); prob = np.empty(N); fl = np.zeros(N, dtype=np.int32)
    pp = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
    lib().oracle_rambo_inverse(ctypes.byref(desc), _p(mom), _p(x1), _p(x2), _p(pp),
                               ctypes.c_int64(N), _p(r), _p(prob), _p(fl))
    return r, prob, fl


def get_ps(desc, mom, boost, perm=None):
    """Compute_ParticlesTensor.get_PS: mom [N,4,4], boost [N,4] -> ps [N,10], mask bool [N]."""
    mom = np.ascontiguousarray(mom, dtype=np.float64)
    boost = np.ascontiguousarray(boost, dtype=np.float64)
    N = mom.shape[0]
    ps = np.empty((N, 10)); mask = np.zeros(N, dtype=np.uint8)
    pp = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
    lib().oracle_rambo_get_ps(ctypes.byref(desc), _p(mom), _p(boost), _p(pp), ctypes.c_int64(N),
                              _p(ps), _p(mask))
    return ps, mask.astype(bool)


def forward_f32(desc, r32):
    """fp32-buffer variant used as the CPU baseline (OpenMP over events)."""
    r32 = np.ascontiguousarray(r32, dtype=np.float32)
    N, n = r32.shape[0], desc.n_final
    mom = np.empty((N, n + 2, 4), dtype=np.float32); w = np.empty(N, dtype=np.float32)
    x1 = np.empty(N, dtype=np.float32); x2 = np.empty(N, dtype=np.float32)
    lib().oracle_rambo_forward_f32(ctypes.byref(desc), _p(r32), ctypes.c_int64(N), _p(mom), _p(w),
                                   _p(x1), _p(x2))
    return mom, w, x1, x2
