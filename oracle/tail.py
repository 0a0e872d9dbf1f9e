"""numpy/ctypes front-end of oracle/salun_oracle.c (TEST INFRASTRUCTURE ONLY).

Each function restates reference arithmetic; see the C file for file:line citations.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsalun_oracle.so")
_lib = None


def build():
    src = os.path.join(_HERE, "salun_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_grad_norm.restype = C.c_double
        _lib.oracle_clip_coef.restype = C.c_float
        _lib.oracle_clip_coef.argtypes = [C.c_double, C.c_float]
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.c_void_p)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def saliency_accumulate(accum: np.ndarray, grad: np.ndarray):
    assert accum.dtype == np.float32 and accum.flags.c_contiguous
    g, gp = _f32(grad)
    lib().oracle_saliency_accumulate(_p(accum), gp, C.c_int64(accum.size))
    return accum


def abs_inplace(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    lib().oracle_abs_inplace(_p(a), C.c_int64(a.size))
    return a


def topk_mask(absg: np.ndarray, k: int):
    """returns (mask_i64, mask_bits(uint32), thr_key, n_gt, n_eq)"""
    a, ap = _f32(absg)
    n = a.size
    m64 = np.empty(n, dtype=np.int64)
    bits = np.zeros((n + 31) // 32, dtype=np.uint32)
    thr = C.c_uint32(); ngt = C.c_int64(); neq = C.c_int64()
    lib().oracle_topk_mask(ap, C.c_int64(n), C.c_int64(int(k)), _p(m64), _p(bits), C.byref(thr), C.byref(ngt),
                           C.byref(neq))
    return m64, bits, thr.value, ngt.value, neq.value


def topk_mask_argsort(absg: np.ndarray, k: int) -> np.ndarray:
    """The reference formula itself (generate_mask.py:57-80) with a STABLE argsort -- small inputs only."""
    neg = -np.asarray(absg, dtype=np.float32)
    neg = np.where(np.isnan(neg), np.float32(np.inf), neg)  # torch sorts NaN last
    positions = np.argsort(neg, kind="stable")
    ranks = np.argsort(positions, kind="stable")
    m = np.zeros(neg.size, dtype=np.int64)
    m[ranks < k] = 1
    return m


def pack_mask(m64: np.ndarray) -> np.ndarray:
    m64 = np.ascontiguousarray(m64, dtype=np.int64)
    bits = np.zeros((m64.size + 31) // 32, dtype=np.uint32)
    lib().oracle_pack_mask(_p(m64), C.c_int64(m64.size), _p(bits))
    return bits


def masked_sgd_step(p, g, v, bits, lr, momentum, wd):
    for a in (p, g, v):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    lib().oracle_masked_sgd_step(_p(p), _p(g), _p(v), _p(bits), C.c_int64(p.size), C.c_float(lr),
                                 C.c_float(momentum), C.c_float(wd))


def grad_norm(g) -> float:
    g, gp = _f32(g)
    return float(lib().oracle_grad_norm(gp, C.c_int64(g.size)))


def clip_coef(total_norm: float, max_norm: float) -> float:
    return float(lib().oracle_clip_coef(C.c_double(total_norm), C.c_float(max_norm)))


def masked_adam_step(p, g, m1, m2, bits, lr, b1, b2, eps, wd, step, clip):
    for a in (p, g, m1, m2):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    lib().oracle_masked_adam_step(_p(p), _p(g), _p(m1), _p(m2), _p(bits), C.c_int64(p.size), C.c_float(lr),
                                  C.c_float(b1), C.c_float(b2), C.c_float(eps), C.c_float(wd), C.c_int64(step),
                                  C.c_float(clip))


def apply_mask(g, bits):
    assert g.dtype == np.float32 and g.flags.c_contiguous
    lib().oracle_apply_mask(_p(g), _p(bits), C.c_int64(g.size))


def l1_penalty_grad(p: np.ndarray, g: np.ndarray, alpha: float) -> float:
    """g += alpha * sign(p) in place; returns sum |p| (FT.py:13-17,133-134)."""
    assert g.dtype == np.float32 and g.flags.c_contiguous
    pp, ppp = _f32(p)
    l = lib()
    l.oracle_l1_penalty_grad.restype = C.c_double
    return float(l.oracle_l1_penalty_grad(ppp, _p(g), C.c_int64(g.size), C.c_float(alpha)))
