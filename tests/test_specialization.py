"""
Host side of the model-specialised builds (mapleaf_b200.backend.specializedLibrary), no GPU needed: the generated header carries the
model block bit for bit, the cache key follows content, and the compiled library exports the whole C ABI, says what it is, and
refuses any other model block before it touches a device. (GPU parity of these libraries: tests/test_gpu_specialized.py.)
"""
import ctypes
import math
import os
import re
import struct

import numpy as np

from conftest import DATA
from mapleaf_b200 import L, SimDefinition, backend, buildModel


def _model(name="MonteCarloAdaptive.mapleaf", overrides=None):
    sd = SimDefinition(os.path.join(DATA, name), silent=True, disableDistributionSampling=True)
    for k, v in (overrides or {}).items():
        sd.setValue(k, v)
    return buildModel(sd)


def test_header_carries_the_shared_part_bit_for_bit(tmp_path):
    m = _model()
    blob = m.blob.copy()
    blob[L.H_TURB_INTENSITY] = float("nan")            # a NaN marks "not given" in the block: it must survive as a NaN
    path = str(tmp_path / "spec.h")
    backend.writeSpecializationHeader(blob, path)
    text = open(path).read()
    n = int(re.search(r"#define MLB_SPEC_LEN (\d+)", text).group(1))
    shared = backend.sharedPart(blob)
    assert n == shared.size == int(blob[L.H_SHARED_LEN])
    body = text.split("#define MLB_SPEC_VALUES")[1].split("\n\nstatic")[0].replace("\\", " ")
    toks = [t.strip() for t in body.split(",") if t.strip()]
    assert len(toks) == n
    for t, v in zip(toks, shared):
        if t == "NAN":
            assert math.isnan(v)
        elif t.endswith("INFINITY"):
            assert math.isinf(v) and (v < 0) == t.startswith("-")
        else:
            assert struct.pack("<d", float.fromhex(t)) == struct.pack("<d", v)
    assert "static __device__ const double kSpecBlob[MLB_SPEC_LEN]" in text and "kSpecBlobHost" in text


def test_cache_key_follows_the_model_block():
    a, b = _model(), _model()
    assert backend.specializationKey(a.blob) == backend.specializationKey(b.blob)
    c = a.blob.copy()
    c[L.H_DT0] = np.nextafter(c[L.H_DT0], 1.0)       # one ulp in one word is another rocket
    assert backend.specializationKey(c) != backend.specializationKey(a.blob)
    # words behind the shared part (scattered-interpolation records in HBM) are not compiled in
    d = np.concatenate([a.blob, [1.0, 2.0]])
    assert backend.specializationKey(d) == backend.specializationKey(a.blob)


def test_specialised_library_exports_the_abi_and_serves_one_block_only():
    m = _model()                                       # bench configs[1]: compiled ahead of time by __graft_entry__.build()
    lib = backend.specializedLibrary(m.blob)
    raw = ctypes.CDLL(backend.specializedLibraryPath(m.blob))
    for sym in backend.EXPORTED_SYMBOLS:
        getattr(raw, sym)
    assert lib.mlb_is_specialized() == 1 and b"model-specialised" in lib.mlb_version()
    assert backend.lib().mlb_is_specialized() == 0
    other = m.blob.copy()
    other[L.H_DT0] *= 2
    h = ctypes.c_void_p()
    rc = lib.mlb_create(other.ctypes.data_as(ctypes.c_void_p), other.nbytes, 0, ctypes.byref(h))
    assert rc == 1 and b"specialised for another model block (word %d differs)" % L.H_DT0 in lib.mlb_last_error()
    # its own block passes that check; what happens next depends on the box (no device here: MLB_ERR_CUDA, never a CPU path)
    rc = lib.mlb_create(m.blob.ctypes.data_as(ctypes.c_void_p), m.blob.nbytes, 0, ctypes.byref(h))
    if rc == 0:
        lib.mlb_destroy(h)
    else:
        assert rc == 2 and b"no CPU path" in lib.mlb_last_error()


def test_trajectory_batch_library_choice(monkeypatch):
    m = _model()
    seen = []
    monkeypatch.setattr(backend, "specializedLibrary", lambda blob, build=True: seen.append("spec") or (_ for _ in ()).throw(RuntimeError("stop")))
    monkeypatch.setattr(backend, "lib", lambda: seen.append("generic") or (_ for _ in ()).throw(RuntimeError("stop")))
    for spec, env, want in ((True, None, "spec"), (False, "1", "generic"), (None, "1", "spec"), (None, "0", "generic"), (None, None, "generic"),
                            (None, "cached", "spec")):            # configs[1]'s library is cached by build()
        seen.clear()
        if env is None:
            monkeypatch.delenv("MLB_SPECIALIZE", raising=False)
        else:
            monkeypatch.setenv("MLB_SPECIALIZE", env)
        try:
            backend.TrajectoryBatch(m.blob, specialize=spec)
        except RuntimeError:
            pass
        assert seen[:1] == [want], (spec, env, seen)
    # "cached": the specialised library only if it has been built already
    monkeypatch.setattr(backend.os.path, "exists", lambda p: False)
    seen.clear()
    try:
        backend.TrajectoryBatch(m.blob, specialize="cached")
    except RuntimeError:
        pass
    assert seen[:1] == ["generic"]


def test_bench_reads_the_metrics_pass_of_the_library_and_shape_it_flew():
    """roofline.traffic comes from the committed ncu pass of the SAME library (generic / specialised) and CTA shape; a pass over the
    generations of one flight adds up its .sum metrics."""
    import bench
    generic, gsrc = bench.measuredKernelMetrics(1, specialised=False)
    spec, ssrc = bench.measuredKernelMetrics(1, specialised=True, block=1024)
    spec512, s512src = bench.measuredKernelMetrics(1, specialised=True, block=512)
    assert "_spec_" not in gsrc and "_spec_" in ssrc and "_block" not in ssrc and "_block512_" in s512src
    assert generic["dram_bytes"] > spec["dram_bytes"] > spec512["dram_bytes"] > 1e6
    # the same flights execute the same FP64 work in both shapes
    for k in ("dfma", "dadd", "dmul"):
        assert abs(spec[k] - spec512[k]) <= 1e-3 * spec[k]
    assert bench.measuredKernelMetrics(1, specialised=True, block=768) == (None, None)      # no pass of that shape: traffic stays null
    for c in (0, 2, 3, 4):
        per, src = bench.measuredKernelMetrics(c, specialised=True)
        assert per is not None and "config%d" % c in src


def test_the_model_block_really_folds_into_the_kernel():
    """Static check on the compiled code (cuobjdump, no GPU): the free-flight kernel of the library compiled for configs[1] is less than half
    the generic kernel, carries no shared-memory staging loop's loads of the block, and keeps its launch bounds (64 registers at 1024 lanes)."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("CUDA binary utilities not installed")

    def kernelSass(path):
        txt = subprocess.run(["cuobjdump", "-sass", "-fun", "_Z11flightLanesILi1024ELi1ELb0EEv7RunArgs", path], capture_output=True, text=True).stdout
        return [l for l in txt.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l)]
    m = _model()
    generic = kernelSass(backend.buildLibrary())
    spec = kernelSass(backend.buildSpecializedLibrary(m.blob))
    assert len(generic) > 30000 and 5000 < len(spec) < 0.5 * len(generic), (len(generic), len(spec))
    usage = subprocess.run(["cuobjdump", "-res-usage", backend.specializedLibraryPath(m.blob)], capture_output=True, text=True).stdout
    row = usage[usage.index("_Z11flightLanesILi1024ELi1ELb0EEv7RunArgs"):].splitlines()[1]
    assert "REG:64 " in row, row
