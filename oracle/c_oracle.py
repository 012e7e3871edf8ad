"""ctypes loader for the C restatement ``oracle/floris_oracle.c`` (TEST INFRASTRUCTURE ONLY).

Builds ``oracle/_build/liboracle.so`` on demand with gcc.  Only tests/, ``__graft_entry__.smoke()`` and bench.py's
CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "floris_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        _lib.wf_oracle_solve_batch.restype = ctypes.c_int
        _lib.wf_oracle_solve_batch.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp, dp, ctypes.c_double,
                                               ctypes.c_int, ip, dp, dp, dp, dp, dp, dp, dp]
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def solve_batch(layout_x, layout_y, ws, wd, yaw_deg, cs=None, ti_ambient=0.06, nthreads=None):
    """Batched solve: ws, wd (B,), yaw_deg (B,T) degrees in original turbine order.  Returns a dict of (B,T) arrays
    (``power_W``, ``ws_local``, ``wd_local``, ``ti``, ``std_u``, ``std_v``, ``std_w``) plus ``order`` (int32)."""
    lx = np.ascontiguousarray(layout_x, dtype=np.float64)
    ly = np.ascontiguousarray(layout_y, dtype=np.float64)
    ws = np.ascontiguousarray(np.atleast_1d(ws), dtype=np.float64)
    wd = np.ascontiguousarray(np.atleast_1d(wd), dtype=np.float64)
    B, T = ws.shape[0], lx.shape[0]
    yaw = np.ascontiguousarray(np.asarray(yaw_deg, dtype=np.float64).reshape(B, T))
    csp = None
    if cs is not None:
        cs = np.ascontiguousarray(np.asarray(cs, dtype=np.float64).reshape(B, 2))
        csp = _dp(cs)
    out = {k: np.empty((B, T), dtype=np.float64)
           for k in ("power_W", "ws_local", "wd_local", "ti", "std_u", "std_v", "std_w")}
    order = np.empty((B, T), dtype=np.int32)
    if nthreads is None:
        nthreads = min(os.cpu_count() or 1, max(1, B))
    rc = lib().wf_oracle_solve_batch(B, T, _dp(lx), _dp(ly), _dp(ws), _dp(wd), _dp(yaw), csp, float(ti_ambient),
                                     int(nthreads), order.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                     _dp(out["power_W"]), _dp(out["ws_local"]), _dp(out["wd_local"]), _dp(out["ti"]),
                                     _dp(out["std_u"]), _dp(out["std_v"]), _dp(out["std_w"]))
    if rc != 0:
        raise RuntimeError(f"wf_oracle_solve_batch failed rc={rc}")
    out["order"] = order
    return out


class _Sol:
    __slots__ = ("order", "power_W", "ws_local", "wd_local", "ti", "std_u", "std_v", "std_w")


def solve(layout_x, layout_y, ws, wd, yaw_deg, *, cs=None, ti_ambient=0.06):
    """Single-env drop-in for ``floris_oracle.solve`` (same attribute names) backed by the C restatement.  The rotation
    uses numpy's cosd/sind (passed in) so that the turbine order and self-masks are those of the numpy oracle."""
    if cs is None:
        dev = ((np.float64(wd) - 270.0) % 360.0 + 360.0) % 360.0
        cs = (np.cos(np.radians(dev)), np.sin(np.radians(dev)))
    out = solve_batch(layout_x, layout_y, [ws], [wd], np.asarray(yaw_deg, dtype=np.float64)[None, :], cs=[cs],
                      ti_ambient=ti_ambient, nthreads=1)
    sol = _Sol()
    for k in _Sol.__slots__:
        setattr(sol, k, out[k][0])
    return sol
