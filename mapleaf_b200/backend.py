"""
ctypes binding of libmapleaf_b200.so (C ABI: include/mapleaf_b200.h).

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a) into mapleaf_b200/lib/.
There is no CPU fallback: if the library is missing, or no CUDA device is usable, every call
raises (RuntimeError / ValueError, the reference's convention for bad input, e.g.
MAPLEAF/SimulationRunners/SingleSimulations.py:24, MAPLEAF/Motion/Integration.py:183).
"""
import ctypes
import hashlib
import math
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from .model import L

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MLB_LIB") or os.path.join(_PKG_DIR, "lib", "libmapleaf_b200.so")     # MLB_LIB: tuning variants
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")
INCLUDE_DIR = os.path.join(_PKG_DIR, "..", "include")
TRACE_W = 16

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]


class Counters(ctypes.Structure):
    _fields_ = [("lanes", ctypes.c_int64), ("steps", ctypes.c_int64), ("derivative_evals", ctypes.c_int64), ("rejected_steps", ctypes.c_int64),
                ("steps_3dof", ctypes.c_int64), ("derivative_evals_3dof", ctypes.c_int64), ("derivative_evals_transonic", ctypes.c_int64),
                ("lanes_ok", ctypes.c_int64), ("lanes_step_limit", ctypes.c_int64), ("lanes_nan", ctypes.c_int64),
                ("kernel_launches", ctypes.c_int64), ("last_run_ms", ctypes.c_double), ("dominant_kernel_ms", ctypes.c_double),
                ("chute_kernel_ms", ctypes.c_double), ("flight_block", ctypes.c_int64), ("flight_generations", ctypes.c_int64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _sources():
    out = [os.path.join(CSRC_DIR, f) for f in sorted(os.listdir(CSRC_DIR)) if f.endswith((".cu", ".cuh"))]
    out += [os.path.join(INCLUDE_DIR, "mlb_layout.h"), os.path.join(INCLUDE_DIR, "mapleaf_b200.h")]
    return out


def buildLibrary(force: bool = False, verbose: bool = False, extraFlags=()) -> str:
    """Compiles csrc/mlb_api.cu for sm_100a into mapleaf_b200/lib/libmapleaf_b200.so (nvcc cross-compiles without a GPU)."""
    newest = max(os.path.getmtime(p) for p in _sources())
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    cmd = ["nvcc"] + NVCC_FLAGS + list(extraFlags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(CSRC_DIR, "mlb_api.cu")]
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None
_specLibs = {}


def lib():
    """Loads the (generic) CUDA library. Raises RuntimeError when it has not been built: there is no other compute path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libmapleaf_b200.so not found at {}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). mapleaf_b200 has no CPU fallback.".format(LIB_PATH))
    _lib = _bind(LIB_PATH)
    return _lib


def _bind(path):
    l = ctypes.CDLL(path)
    vp, i64, dp, ip = ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    l.mlb_last_error.restype = ctypes.c_char_p
    l.mlb_version.restype = ctypes.c_char_p
    l.mlb_create.argtypes = [vp, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(vp)]
    l.mlb_destroy.argtypes = [vp]
    l.mlb_destroy.restype = None
    l.mlb_set_lanes.argtypes = [vp, i64, ctypes.POINTER(dp), ip, ctypes.c_int]
    l.mlb_set_lanes_device.argtypes = [vp, i64, ctypes.POINTER(vp), ip, ctypes.c_int]
    l.mlb_set_lane_models.argtypes = [vp, dp, i64]
    l.mlb_run.argtypes = [vp, vp]
    l.mlb_sync.argtypes = [vp]
    l.mlb_results.argtypes = [vp, dp, ctypes.POINTER(ctypes.c_int32)]
    l.mlb_finals.argtypes = [vp, dp]
    l.mlb_counters.argtypes = [vp, ctypes.POINTER(Counters)]
    l.mlb_lane_counters.argtypes = [vp, ctypes.POINTER(ctypes.c_int64)]
    l.mlb_stage_separations.argtypes = [vp, dp]
    l.mlb_device_results.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    l.mlb_set_result_buffers.argtypes = [vp, vp, vp]
    l.mlb_moments.argtypes = [vp, dp]
    l.mlb_set_flight_log.argtypes = [vp, i64, i64, i64, i64]
    l.mlb_get_flight_log.argtypes = [vp, dp, ctypes.POINTER(ctypes.c_int64)]
    l.mlb_set_trace.argtypes = [vp, i64, i64]
    l.mlb_get_trace.argtypes = [vp, dp, ctypes.POINTER(ctypes.c_int64)]
    l.mlb_set_dt_replay.argtypes = [vp, dp, i64]
    l.mlb_force_eval.argtypes = [vp, i64, dp, dp, dp, dp]
    l.mlb_host_gauss_stream.argtypes = [ctypes.POINTER(ctypes.c_uint32), ctypes.c_int, i64, dp, dp, ctypes.c_int, i64, dp]
    l.mlb_host_randrange_stream.argtypes = [ctypes.POINTER(ctypes.c_uint32), ctypes.c_int, i64, ctypes.c_uint32, i64, dp]
    ip = ctypes.POINTER(ctypes.c_int)
    l.mlb_host_gauss_from_state.argtypes = [ctypes.POINTER(ctypes.c_uint32), ip, ip, ctypes.POINTER(ctypes.c_double), dp, dp, ctypes.c_int, i64, dp]
    l.mlb_host_randrange_from_state.argtypes = [ctypes.POINTER(ctypes.c_uint32), ip, ctypes.c_uint32, i64, dp]
    l.mlb_fp64_peak.argtypes = [ctypes.c_int, dp]
    l.mlb_is_specialized.restype = ctypes.c_int
    return l


# ---------------------------------------------------------------------------------------------------------------------
# Model-specialised builds. The generic library INTERPRETS the model block: stage / component loops, a type switch per
# component, offsets and constants read from shared memory. A Monte Carlo campaign flies one model block millions of times,
# so the block can be compiled INTO the kernels instead: written out as a `static __device__ const double[]`, every
# constant-index read folds into the instruction stream, the loops unroll into the rocket's own straight-line force model
# and the options the definition does not use (other atmosphere / wind / earth / turbulence models, control system, staging)
# disappear. Measured on configs[1] at 10^6 lanes: 41 k -> 17 k SASS instructions in the free-flight kernel, 371 -> 513 M
# steps/s (gpurun_out r2t sweep, profiles/r2t_*). Cost: one nvcc run per distinct model block (~1 min), cached by content hash.
# ---------------------------------------------------------------------------------------------------------------------
SPEC_DIR = os.path.join(_PKG_DIR, "lib", "spec")
SPEC_FLAGS = ["-DMLB_COMP_NOINLINE=0", "-DMLB_COLD_MASK=0"]       # with the block folded in, inlined component / cold routines win (generic build: out of line)


def _specLiteral(v):
    if math.isnan(v):
        return "NAN"
    if math.isinf(v):
        return "INFINITY" if v > 0 else "-INFINITY"
    return float(v).hex()


def sharedPart(blob) -> np.ndarray:
    """The doubles of a model block the kernels keep on chip: [0, H_SHARED_LEN) (scattered-interpolation records behind it stay in HBM)."""
    b = np.ascontiguousarray(blob, dtype=np.float64)
    n = int(b[L.H_SHARED_LEN])
    return b[:n] if 0 < n <= b.size else b


def specializationKey(blob) -> str:
    """Content hash that names a specialised library: the shared part of the model block, the CUDA sources and the build flags."""
    h = hashlib.sha256()
    h.update(sharedPart(blob).tobytes())
    for p in _sources():
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + SPEC_FLAGS).encode())
    return h.hexdigest()[:16]


def writeSpecializationHeader(blob, path):
    """The model block as C source: MLB_SPEC, MLB_SPEC_LEN and the same values as a device array (folded by the compiler) and a host
    array (mlb_create checks the caller's block against it word by word)."""
    b = sharedPart(blob)
    rows = ["  " + ", ".join(_specLiteral(float(x)) for x in b[i:i + 4]) + "," for i in range(0, b.size, 4)]
    with open(path, "w") as f:
        f.write("/* generated by mapleaf_b200.backend.writeSpecializationHeader: one model block baked into the kernels */\n"
                "#include <math.h>\n#define MLB_SPEC 1\n#define MLB_SPEC_LEN %d\n#define MLB_SPEC_VALUES \\\n%s\n\n"
                "static __device__ const double kSpecBlob[MLB_SPEC_LEN] = { MLB_SPEC_VALUES };\n"
                "static const double kSpecBlobHost[MLB_SPEC_LEN] = { MLB_SPEC_VALUES };\n" % (b.size, " \\\n".join(rows)))


def specializedLibraryPath(blob) -> str:
    return os.path.join(SPEC_DIR, "libmapleaf_b200_" + specializationKey(blob) + ".so")


def buildSpecializedLibrary(blob, wait: bool = True, verbose: bool = False):
    """Compiles the library specialised for `blob` unless it is already cached. wait=False returns the running nvcc process (or None)."""
    out = specializedLibraryPath(blob)
    if os.path.exists(out):
        return out if wait else None
    os.makedirs(SPEC_DIR, exist_ok=True)
    hdr = out[:-3] + ".h"
    writeSpecializationHeader(blob, hdr)
    tmp = out + ".tmp%d" % os.getpid()
    cmd = ["nvcc"] + NVCC_FLAGS + SPEC_FLAGS + ["-include", hdr] + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, os.path.join(CSRC_DIR, "mlb_api.cu")]
    try:
        proc = subprocess.Popen(cmd + [], stdout=None if verbose else subprocess.DEVNULL, stderr=None if verbose else subprocess.PIPE)
    except FileNotFoundError:
        raise RuntimeError("a model-specialised library needs nvcc on PATH (build it where the toolkit is: backend.buildSpecializedLibrary)")
    proc._mlb_out = (tmp, out)
    if not wait:
        return proc
    finishSpecializedBuild(proc)
    return out


def finishSpecializedBuild(proc):
    if proc is None:
        return
    err = proc.communicate()[1]
    tmp, out = proc._mlb_out
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed for the specialised library:\n" + (err.decode()[-4000:] if err else ""))
    os.replace(tmp, out)


def specializedLibrary(blob, build: bool = True):
    """The loaded library specialised for this model block (compiled on first use when `build`, else it must be cached)."""
    path = specializedLibraryPath(blob)
    if path in _specLibs:
        return _specLibs[path]
    if not os.path.exists(path):
        if not build:
            raise RuntimeError("no specialised library cached for this model block ({}): call backend.buildSpecializedLibrary(blob) where nvcc is".format(path))
        buildSpecializedLibrary(blob)
    l = _bind(path)
    if l.mlb_is_specialized() != 1:
        raise RuntimeError("{} is not a model-specialised build".format(path))
    _specLibs[path] = l
    return l


EXPORTED_SYMBOLS = ["mlb_create", "mlb_destroy", "mlb_last_error", "mlb_version", "mlb_set_lanes", "mlb_set_lanes_device", "mlb_set_lane_models", "mlb_run", "mlb_sync",
                    "mlb_results", "mlb_finals", "mlb_counters", "mlb_lane_counters", "mlb_stage_separations", "mlb_device_results", "mlb_set_result_buffers", "mlb_moments", "mlb_set_flight_log", "mlb_get_flight_log", "mlb_set_trace",
                    "mlb_get_trace", "mlb_set_dt_replay", "mlb_force_eval", "mlb_host_gauss_stream", "mlb_host_randrange_stream", "mlb_host_gauss_from_state", "mlb_host_randrange_from_state", "mlb_fp64_peak", "mlb_is_specialized"]


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def _check(rc, what, l=None):
    if rc != 0:
        msg = (l or lib()).mlb_last_error().decode()
        if rc == 2:
            raise RuntimeError("{}: {}".format(what, msg))
        raise ValueError("{}: {}".format(what, msg))


_LANE_COMPONENTS = {L.LANE_WIND: 3, L.LANE_INIT_STATE: 13, L.LANE_PINK_SEED: 3, L.LANE_MOTOR_ADJUST: 2 * L.MLB_MAX_STAGES, L.LANE_START: 2}


class TrajectoryBatch:
    """One Monte Carlo case on one GPU: a model block plus per-lane sampled inputs."""

    def __init__(self, blob, device: int = 0, specialize=None):
        """specialize: True flies the batch on the library compiled for this very model block (backend.specializedLibrary: built with nvcc
        on first use, cached by content hash); "cached" uses it only if it is already built; False / None the generic library (None reads
        MLB_SPECIALIZE = 0 | 1 | cached from the environment). Results agree to rounding; per-lane model blocks need the generic one."""
        self._h = None
        self.blob = np.ascontiguousarray(blob, dtype=np.float64)
        self.nLanes = 0
        self._keep = []
        if specialize is None:
            specialize = {"1": True, "true": True, "cached": "cached"}.get(os.environ.get("MLB_SPECIALIZE", "0").lower(), False)
        if specialize == "cached":
            specialize = os.path.exists(specializedLibraryPath(self.blob))
        self._lib = specializedLibrary(self.blob) if specialize else lib()
        self.specialized = bool(specialize)
        self._h = ctypes.c_void_p()
        _check(self._lib.mlb_create(self.blob.ctypes.data_as(ctypes.c_void_p), self.blob.nbytes, device, ctypes.byref(self._h)), "mlb_create", self._lib)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mlb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def setLanes(self, nLanes: int, laneArrays: Optional[Dict[int, np.ndarray]] = None):
        """laneArrays: {LANE_* id: float64 array [components, nLanes]} in host memory."""
        laneArrays = laneArrays or {}
        ids = np.array(list(laneArrays.keys()), dtype=np.int32)
        arrs = []
        for k, a in laneArrays.items():
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != (_LANE_COMPONENTS[k], nLanes):
                raise ValueError("lane array {} must have shape ({}, {}), got {}".format(k, _LANE_COMPONENTS[k], nLanes, a.shape))
            arrs.append(a)
        ptrs = (ctypes.POINTER(ctypes.c_double) * max(1, len(arrs)))(*[_dp(a) for a in arrs])
        _check(self._lib.mlb_set_lanes(self._h, nLanes, ptrs, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(arrs)), "mlb_set_lanes", self._lib)
        self.nLanes = nLanes

    def setLanesDevice(self, nLanes: int, devicePointers: Dict[int, int], keepAlive=None):
        """devicePointers: {LANE_* id: device address of a float64 [components, nLanes] array on this GPU}."""
        ids = np.array(list(devicePointers.keys()), dtype=np.int32)
        ptrs = (ctypes.c_void_p * max(1, len(devicePointers)))(*[ctypes.c_void_p(p) for p in devicePointers.values()])
        _check(self._lib.mlb_set_lanes_device(self._h, nLanes, ptrs, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(devicePointers)), "mlb_set_lanes_device", self._lib)
        self.nLanes = nLanes
        self._keep = [keepAlive]

    def setLaneModels(self, models):
        """models: float64 [nLanes, >= H_SHARED_LEN]: each lane's own model block (see mlb_set_lane_models); None switches back."""
        if models is None:
            _check(self._lib.mlb_set_lane_models(self._h, None, 0), "mlb_set_lane_models", self._lib)
            return
        m = np.ascontiguousarray(models, dtype=np.float64)
        if m.ndim != 2 or m.shape[0] != self.nLanes:
            raise ValueError("lane models must have shape (nLanes, words), got {}".format(m.shape))
        _check(self._lib.mlb_set_lane_models(self._h, _dp(m), m.shape[1]), "mlb_set_lane_models", self._lib)

    def run(self, stream: int = 0, sync: bool = False):
        _check(self._lib.mlb_run(self._h, ctypes.c_void_p(stream)), "mlb_run", self._lib)
        if sync:
            self.sync()

    def sync(self):
        _check(self._lib.mlb_sync(self._h), "mlb_sync", self._lib)

    def results(self):
        res = np.empty((self.nLanes, L.RES_SIZE))
        st = np.empty(self.nLanes, dtype=np.int32)
        _check(self._lib.mlb_results(self._h, _dp(res), st.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))), "mlb_results", self._lib)
        return res, st

    def finals(self):
        out = np.empty((self.nLanes, L.FIN_SIZE))
        _check(self._lib.mlb_finals(self._h, _dp(out)), "mlb_finals", self._lib)
        return out

    def counters(self) -> dict:
        c = Counters()
        _check(self._lib.mlb_counters(self._h, ctypes.byref(c)), "mlb_counters", self._lib)
        return c.asdict()

    def laneCounters(self):
        out = np.empty((self.nLanes, 6), dtype=np.int64)
        _check(self._lib.mlb_lane_counters(self._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))), "mlb_lane_counters", self._lib)
        return out

    def stageSeparations(self):
        """[nLanes, MLB_MAX_STAGES - 1, SEP_W]: state, time and step size of the rocket at each stage separation (SEP_VALID = 1 where one happened)."""
        out = np.empty((self.nLanes, L.MLB_MAX_STAGES - 1, L.SEP_W))
        _check(self._lib.mlb_stage_separations(self._h, _dp(out)), "mlb_stage_separations", self._lib)
        return out

    def deviceResults(self):
        r, s = ctypes.c_void_p(), ctypes.c_void_p()
        _check(self._lib.mlb_device_results(self._h, ctypes.byref(r), ctypes.byref(s)), "mlb_device_results", self._lib)
        return r.value, s.value

    def setResultBuffers(self, resultsPtr: int, statusPtr: int, keepAlive=None):
        """Device addresses of caller-owned [nLanes, 7] float64 and [nLanes] int32 buffers the next runs write into."""
        _check(self._lib.mlb_set_result_buffers(self._h, ctypes.c_void_p(resultsPtr), ctypes.c_void_p(statusPtr)), "mlb_set_result_buffers", self._lib)
        self._keepOut = keepAlive

    def moments(self):
        """count, mean[7], M2[7] of the result columns over lanes that ended normally (computed on the device)."""
        out = np.empty(15)
        _check(self._lib.mlb_moments(self._h, _dp(out)), "mlb_moments", self._lib)
        return out[0], out[1::2].copy(), out[2::2].copy()

    def setFlightLog(self, firstLane: int, nLanes: int, every: int = 1, maxRowsPerLane: int = 2048):
        """Record the flight paths of lanes [firstLane, firstLane + nLanes) during the next run (see mlb_set_flight_log)."""
        _check(self._lib.mlb_set_flight_log(self._h, firstLane, nLanes, every, maxRowsPerLane), "mlb_set_flight_log", self._lib)
        self._flightLog = (nLanes, maxRowsPerLane)

    def flightLog(self):
        """List of [rows_i, 16] arrays, one per logged lane: time, pos, vel, quaternion, angular velocity, dt, error estimate."""
        n, m = self._flightLog
        rows = np.empty((n, m, TRACE_W))
        counts = np.zeros(n, dtype=np.int64)
        _check(self._lib.mlb_get_flight_log(self._h, _dp(rows), counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))), "mlb_get_flight_log", self._lib)
        return [rows[i, :counts[i]].copy() for i in range(n)]

    def flightLogRowCounts(self):
        """Rows kept per logged lane (without copying the rows)."""
        n, _ = self._flightLog
        counts = np.zeros(n, dtype=np.int64)
        _check(self._lib.mlb_get_flight_log(self._h, None, counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))), "mlb_get_flight_log", self._lib)
        return counts

    def setTrace(self, lane: int, maxRows: int):
        _check(self._lib.mlb_set_trace(self._h, lane, maxRows), "mlb_set_trace", self._lib)
        self._traceMax = maxRows

    def trace(self):
        rows = np.empty((max(1, getattr(self, "_traceMax", 0)), TRACE_W))
        n = ctypes.c_int64(0)
        _check(self._lib.mlb_get_trace(self._h, _dp(rows), ctypes.byref(n)), "mlb_get_trace", self._lib)
        return rows[:n.value]

    def setDtReplay(self, dts):
        a = np.ascontiguousarray(dts, dtype=np.float64) if dts is not None else None
        _check(self._lib.mlb_set_dt_replay(self._h, _dp(a), 0 if a is None else a.size), "mlb_set_dt_replay", self._lib)

    def forceEval(self, states, times=None, winds=None):
        s = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, 13)
        n = s.shape[0]
        t = np.zeros(n) if times is None else np.ascontiguousarray(times, dtype=np.float64)
        w = None if winds is None else np.ascontiguousarray(winds, dtype=np.float64).reshape(n, 3)
        out = np.empty((n, 18))
        _check(self._lib.mlb_force_eval(self._h, n, _dp(s), _dp(t), _dp(w), _dp(out)), "mlb_force_eval", self._lib)
        return out


def seedKey(seed) -> np.ndarray:
    """The 32-bit key words CPython's random.seed() feeds to init_by_array: |int| in little-endian 32-bit chunks; a str
    seed becomes int.from_bytes(s.encode() + sha512(s.encode()).digest(), 'big') (Lib/random.py, version 2)."""
    if isinstance(seed, str):
        b = seed.encode()
        n = int.from_bytes(b + hashlib.sha512(b).digest(), "big")
    else:
        n = abs(int(seed))
    words = []
    while True:
        words.append(n & 0xFFFFFFFF)
        n >>= 32
        if n == 0:
            break
    return np.array(words, dtype=np.uint32)


def hostGaussStream(seed, mu, sigma, nSamples: int, skipDraws: int = 0) -> np.ndarray:
    """nSamples rows of gauss(mu[j], sigma[j]) drawn from random.Random(seed) in CPython's order (C++ sampler)."""
    mu = np.ascontiguousarray(np.atleast_1d(mu), dtype=np.float64)
    sigma = np.ascontiguousarray(np.atleast_1d(sigma), dtype=np.float64)
    key = seedKey(seed)
    out = np.empty((nSamples, mu.size))
    _check(lib().mlb_host_gauss_stream(key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), key.size, skipDraws, _dp(mu), _dp(sigma), mu.size,
                                      nSamples * mu.size, _dp(out)), "mlb_host_gauss_stream")
    return out


def hostRandrangeStream(seed, n: int, nDraws: int, skipDraws: int = 0) -> np.ndarray:
    """nDraws values of random.Random(seed).randrange(n) in CPython's order (C++ sampler), as float64."""
    key = seedKey(seed)
    out = np.empty(nDraws)
    _check(lib().mlb_host_randrange_stream(key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), key.size, skipDraws, n, nDraws, _dp(out)),
           "mlb_host_randrange_stream")
    return out


def _pyRandomState(rng):
    version, internal, gaussNext = rng.getstate()
    words = np.array(internal[:624], dtype=np.uint32)
    return version, words, ctypes.c_int(internal[624]), gaussNext


def gaussFromGenerator(rng, mu, sigma, nSamples: int) -> np.ndarray:
    """nSamples rows of rng.gauss(mu[j], sigma[j]) for a random.Random (or the `random` module), drawn by the C++ sampler; the generator
    is left exactly where nSamples * len(mu) Python-level gauss() calls would have left it."""
    mu = np.ascontiguousarray(np.atleast_1d(mu), dtype=np.float64)
    sigma = np.ascontiguousarray(np.atleast_1d(sigma), dtype=np.float64)
    version, words, pos, gaussNext = _pyRandomState(rng)
    hasNext = ctypes.c_int(0 if gaussNext is None else 1)
    nxt = ctypes.c_double(0.0 if gaussNext is None else gaussNext)
    out = np.empty((nSamples, mu.size))
    _check(lib().mlb_host_gauss_from_state(words.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.byref(pos), ctypes.byref(hasNext), ctypes.byref(nxt),
                                          _dp(mu), _dp(sigma), mu.size, nSamples * mu.size, _dp(out)), "mlb_host_gauss_from_state")
    rng.setstate((version, tuple(int(w) for w in words) + (pos.value,), nxt.value if hasNext.value else None))
    return out


def randrangeFromGenerator(rng, n: int, nDraws: int) -> np.ndarray:
    """nDraws values of rng.randrange(n), drawn by the C++ sampler; the generator is advanced as Python would have advanced it."""
    version, words, pos, gaussNext = _pyRandomState(rng)
    out = np.empty(nDraws)
    _check(lib().mlb_host_randrange_from_state(words.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.byref(pos), n, nDraws, _dp(out)),
           "mlb_host_randrange_from_state")
    rng.setstate((version, tuple(int(w) for w in words) + (pos.value,), gaussNext))
    return out


def fp64Peak(device: int = 0) -> float:
    """Measured FP64 FMA throughput of the device in TFLOP/s."""
    out = ctypes.c_double(0)
    _check(lib().mlb_fp64_peak(device, ctypes.byref(out)), "mlb_fp64_peak")
    return out.value
