"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes wrapper of oracle/conv_ref.c (the plain-C
restatement).  Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libconv_ref.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "libconv_ref.so"])
    return SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        _lib = C.CDLL(SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def conv2d_fwd(xp, w, b, stride):
    N, Cc, Hp, Wp = xp.shape
    K, _, R, S = w.shape
    y = np.empty((N, K, (Hp - R) // stride[0] + 1, (Wp - S) // stride[1] + 1), np.float32)
    lib().ref_conv2d_fwd(_p(np.ascontiguousarray(xp, np.float32)), _p(np.ascontiguousarray(w, np.float32)),
                         _p(None if b is None else np.ascontiguousarray(b, np.float32)), _p(y),
                         N, Cc, Hp, Wp, K, R, S, stride[0], stride[1])
    return y


def conv2d_bwd(xp, w, gy, stride, need_db=True):
    N, Cc, Hp, Wp = xp.shape
    K, _, R, S = w.shape
    xp = np.ascontiguousarray(xp, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    gy = np.ascontiguousarray(gy, np.float32)
    dxp = np.empty_like(xp)
    dw = np.empty_like(w)
    db = np.empty(K, np.float32) if need_db else None
    lib().ref_conv2d_dgrad(_p(gy), _p(w), _p(dxp), N, Cc, Hp, Wp, K, R, S, stride[0], stride[1])
    lib().ref_conv2d_wgrad(_p(xp), _p(gy), _p(dw), _p(db), N, Cc, Hp, Wp, K, R, S, stride[0], stride[1])
    return dxp, dw, db


def pool_fwd(xp, mode, k, stride):
    N, Cc, Hp, Wp = xp.shape
    xp = np.ascontiguousarray(xp, np.float32)
    y = np.empty((N, Cc, (Hp - k) // stride + 1, (Wp - k) // stride + 1), np.float32)
    lib().ref_pool2d_fwd(_p(xp), _p(y), N, Cc, Hp, Wp, k, stride, 0 if mode == "max" else 1)
    return y


def pool_bwd(xp, gy, mode, k, stride):
    N, Cc, Hp, Wp = xp.shape
    xp = np.ascontiguousarray(xp, np.float32)
    gy = np.ascontiguousarray(gy, np.float32)
    dxp = np.empty_like(xp)
    lib().ref_pool2d_bwd(_p(xp), _p(gy), _p(dxp), N, Cc, Hp, Wp, k, stride, 0 if mode == "max" else 1)
    return dxp
