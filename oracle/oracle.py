"""ctypes wrapper around the CPU oracle (oracle/mapleaf_oracle.c). TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_DIR, "_build", "libmapleaf_oracle.so")
TRACE_W = 16


def build(force=False):
    src = os.path.join(_DIR, "mapleaf_oracle.c")
    hdr = os.path.join(_DIR, "..", "include", "mlb_layout.h")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-s", "-C", _DIR])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_run.restype = ctypes.c_int
        _lib.oracle_force_eval.restype = ctypes.c_int
    return _lib


_OPLIB = os.path.join(_DIR, "_build", "libmapleaf_oracle_opcount.so")
_oplib = None


def opcountLib():
    """The oracle compiled with an op-counting scalar in place of double (oracle/opcount.hpp): exact algorithmic FLOP counts."""
    global _oplib
    if _oplib is None:
        srcs = [os.path.join(_DIR, f) for f in ("opcount_build.cpp", "opcount.hpp", "mapleaf_oracle.c")] + [os.path.join(_DIR, "..", "include", "mlb_layout.h")]
        if not os.path.exists(_OPLIB) or os.path.getmtime(_OPLIB) < max(os.path.getmtime(f) for f in srcs):
            os.makedirs(os.path.dirname(_OPLIB), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fpermissive", "-w", "-shared",
                                   "-o", _OPLIB, srcs[0], "-lm"])
        _oplib = ctypes.CDLL(_OPLIB)
        _oplib.oracle_run.restype = ctypes.c_int
    return _oplib


def runCounted(blob, nLanes, laneArrays=None):
    """oracle.run through the op-counting build. Returns (results dict, op counts dict)."""
    global _lib
    saved, _lib = _lib, opcountLib()
    try:
        _lib.oracle_opcount_reset()
        out = run(blob, nLanes, laneArrays)
        c = np.zeros(17, dtype=np.uint64)
        _lib.oracle_opcount_get(_p(c, ctypes.c_uint64))
    finally:
        _lib = saved
    names = ["eval_6dof", "eval_6dof_transonic", "eval_3dof", "step_rest_6dof", "step_rest_3dof"]
    ops = dict(add=int(c[0]), mul=int(c[1]), div=int(c[2]), sqrt=int(c[3]), compare=int(c[4]), transcendental=int(c[5]), uncounted=int(c[6]),
               buckets={n: dict(flops=int(c[7 + 2 * k]), count=int(c[8 + 2 * k])) for k, n in enumerate(names)})
    ops["flops"] = sum(int(x) for x in c[:6])
    return out, ops


def _p(a, t=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def run(blob, nLanes, laneArrays=None, traceLane=-1, traceMaxRows=0, dtReplay=None, separations=False):
    """laneArrays: dict {arrayId: ndarray [ncomp, nLanes]}. Returns dict of results. separations: also return the stage-separation rows
    [nLanes, 3, 16] (state, time, step size at each separation: what a dropped stage's flight starts from)."""
    blob = np.ascontiguousarray(blob, dtype=np.float64)
    laneArrays = laneArrays or {}
    ids = np.array(list(laneArrays.keys()), dtype=np.int32)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in laneArrays.values()]
    for a in arrs:
        assert a.shape[-1] == nLanes
    ptrs = (ctypes.POINTER(ctypes.c_double) * max(1, len(arrs)))(*[_p(a) for a in arrs])
    results = np.zeros((nLanes, 7))
    finals = np.zeros((nLanes, 14))
    status = np.zeros(nLanes, dtype=np.int32)
    counters = np.zeros((nLanes, 3), dtype=np.int64)
    trace = np.zeros((max(1, traceMaxRows), TRACE_W))
    rows = ctypes.c_int64(0)
    replay = np.ascontiguousarray(dtReplay, dtype=np.float64) if dtReplay is not None else None
    sep = np.zeros((nLanes, 3, 16)) if separations else None
    lib().oracle_set_separation_buffer(_p(sep))
    rc = lib().oracle_run(_p(blob), ctypes.c_int64(nLanes), ptrs, _p(ids, ctypes.c_int), ctypes.c_int(len(arrs)),
                          _p(results), _p(finals), _p(status, ctypes.c_int32), _p(counters, ctypes.c_int64),
                          _p(trace) if traceMaxRows > 0 else None, ctypes.c_int64(traceLane), ctypes.c_int64(traceMaxRows),
                          ctypes.byref(rows), _p(replay), ctypes.c_int64(0 if replay is None else replay.size))
    lib().oracle_set_separation_buffer(None)
    if rc != 0:
        raise ValueError("oracle_run failed: bad model block")
    out = dict(results=results, finals=finals, status=status, counters=counters, trace=trace[:rows.value])
    if separations:
        out["separations"] = sep
    return out


def forceEval(blob, state13, time=0.0, wind=None):
    blob = np.ascontiguousarray(blob, dtype=np.float64)
    s = np.ascontiguousarray(state13, dtype=np.float64)
    w = np.ascontiguousarray(wind, dtype=np.float64) if wind is not None else None
    out = np.zeros(19 + 9 * 30)
    lib().oracle_force_eval(_p(blob), _p(s), ctypes.c_double(time), _p(w), _p(out))
    return out


def gaussStream(seed, n, mu=0.0, sigma=1.0):
    out = np.zeros(n)
    lib().oracle_gauss_stream(ctypes.c_uint64(seed), ctypes.c_double(mu), ctypes.c_double(sigma), ctypes.c_int(n), _p(out))
    return out


def pinkNoise(seed, times):
    """times < 0 means 'next raw sample' (getValue() with no time)."""
    t = np.ascontiguousarray(times, dtype=np.float64)
    out = np.zeros(t.size)
    lib().oracle_pink_noise(ctypes.c_uint64(seed), _p(t), ctypes.c_int(t.size), _p(out))
    return out
