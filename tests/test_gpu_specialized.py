"""
The MODEL-SPECIALISED build (mapleaf_b200.backend.specializedLibrary: the model block compiled into the kernels as constants)
against the same fixtures and gates as the generic library. Each case below IS a test of tests/test_gpu_parity.py, re-run with
MLB_SPECIALIZE=1 so that every TrajectoryBatch it creates flies on the library compiled for its own model block; the reference
fixtures (tests/golden/*.json, flown by the unmodified reference) and the oracle stay the checkers.

The libraries are compiled ahead of time where nvcc is (`__graft_entry__.build()` collects the model blocks of these cases with
`collectModelBlocks` below) and travel in-tree (mapleaf_b200/lib/spec/); a block that has no cached library is compiled on the spot.
"""
import inspect

import numpy as np
import pytest

import test_gpu_parity as T
from mapleaf_b200 import L

# (test function of test_gpu_parity.py, extra positional arguments)
CASES = [
    (T.test_rk4_to_apogee_vs_reference, ()),                       # fixed step, 6-DoF, trace rows: <= 1e-9, same step counts
    (T.test_rk4_to_ground_vs_reference, ()),                       # drogue / main chute, 3-DoF switch, landing: <= 1e-9
    (T.test_rk4_matches_oracle_on_seeded_lanes, ()),               # 256 seeded lanes vs the oracle lane by lane
    (T.test_rk45_lockstep_replay, ()),                             # adaptive arithmetic on the reference's accepted step sizes
    (T.test_rk45_pink_noise_vs_reference, ()),                     # free-running adaptive lanes inside the self-sensitivity band
    (T.test_rk45_ensemble_statistics_vs_oracle, ()),               # configs[1]: 1024-lane dispersion statistics
    (T.test_two_stage_fixed_step_vs_reference, ()),                # stage separation (the stage loop is NOT folded for two stages)
    (T.test_nasa_two_stage_orbital_rocket_vs_reference, ()),       # configs[3]: WGS84 + J2, tabulated aero, staging
    (T.test_canard_control_system_vs_reference, ()),               # configs[2]: GNC PID, deflection table, forced RK4
    (T.test_motor_adjust_factors_vs_reference, ("motoradj_rk4.json",)),        # per-lane motor factors on a specialised block
    (T.test_option_fixed_step_vs_reference, ("opt_atm_tabulated.json",)),      # tabulated atmosphere + options folded away
    (T.test_windtunnel_sweep, ()),                                 # force-model known-answer test, 100 points
]


def _call(fn, extra, backend, golden):
    it = iter(extra)
    kw = {}
    for n in inspect.signature(fn).parameters:
        kw[n] = backend if n == "backend" else (golden if n == "golden" else next(it))
    return fn(**kw)


def collectModelBlocks():
    """The model blocks the cases above fly, found by running them with a TrajectoryBatch that records its block and stops the case.
    No GPU needed: this is how build() knows which specialised libraries to compile."""
    import json, os
    from conftest import GOLDEN
    from mapleaf_b200 import backend
    blobs = {}

    class _Stop(Exception):
        pass

    def record(blob, build=True):
        b = np.ascontiguousarray(blob, dtype=np.float64)
        blobs[backend.specializationKey(b)] = b
        raise _Stop()

    def init(self, blob, device=0, specialize=None):
        self._h = None
        record(blob)
    specLib = backend.specializedLibrary
    backend.specializedLibrary = record

    def golden(name):
        with open(os.path.join(GOLDEN, name)) as f:
            return json.load(f)
    saved = backend.TrajectoryBatch.__init__
    backend.TrajectoryBatch.__init__ = init
    try:
        for fn, extra in CASES + [(test_specialised_and_generic_library_agree_lane_by_lane, ()), (test_specialised_library_serves_one_model_block_only, ())]:
            try:
                _call(fn, extra, backend, golden)
            except _Stop:
                pass
    finally:
        backend.TrajectoryBatch.__init__ = saved
        backend.specializedLibrary = specLib
    return list(blobs.values())


pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend():
    from mapleaf_b200 import backend
    backend.lib()
    return backend


@pytest.mark.parametrize("case", CASES, ids=[c[0].__name__[5:] + ("-" + c[1][0] if c[1] else "") for c in CASES])
def test_specialised_library_passes_the_generic_gates(case, backend, golden, monkeypatch):
    monkeypatch.setenv("MLB_SPECIALIZE", "1")
    before = len(backend._specLibs)
    created = []
    init = backend.TrajectoryBatch.__init__

    def spy(self, *a, **k):
        init(self, *a, **k)
        created.append(self.specialized and self._lib.mlb_is_specialized() == 1)
    monkeypatch.setattr(backend.TrajectoryBatch, "__init__", spy)
    _call(case[0], case[1], backend, golden)
    assert created and all(created), "the case did not fly on a specialised library"
    assert len(backend._specLibs) >= max(before, 1)


def test_specialised_and_generic_library_agree_lane_by_lane(backend):
    """Same 512 seeded lanes, fixed-step flights to the ground: the two builds differ only in the rounding of folded constants."""
    m = T.buildFrom("MonteCarlo.mapleaf", {"SimControl.EndCondition": "Altitude", "SimControl.EndConditionValue": "0"})
    rng = np.random.default_rng(11)
    n = 512
    w = rng.normal(size=(3, n)) * np.array([[5.0], [5.0], [0.5]]) + np.array([[2.0], [0.0], [0.0]])
    out = []
    for spec in (False, True):
        with backend.TrajectoryBatch(m.blob, specialize=spec) as b:
            assert b._lib.mlb_is_specialized() == int(spec)
            b.setLanes(n, {L.LANE_WIND: w})
            b.run()
            res, st = b.results()
            out.append((res, st, b.laneCounters()))
    assert (out[0][1] == L.ST_OK).all() and (out[1][1] == L.ST_OK).all()
    assert (out[0][2][:, 0] == out[1][2][:, 0]).all()                        # same step counts
    for i in range(n):
        assert T.relErr(out[1][0][i], out[0][0][i]) < 1e-9


def test_specialised_library_serves_one_model_block_only(backend):
    m = T.buildFrom("MonteCarlo.mapleaf", {"SimControl.EndCondition": "Altitude", "SimControl.EndConditionValue": "0"})
    lib = backend.specializedLibrary(m.blob)
    other = m.blob.copy()
    other[L.H_DT0] *= 2
    import ctypes
    h = ctypes.c_void_p()
    rc = lib.mlb_create(other.ctypes.data_as(ctypes.c_void_p), other.nbytes, 0, ctypes.byref(h))
    assert rc == 1 and b"specialised for another model block" in lib.mlb_last_error()
    with backend.TrajectoryBatch(m.blob, specialize=True) as b:
        b.setLanes(4, {})
        with pytest.raises(ValueError, match="per-lane model blocks need the generic library"):
            b.setLaneModels(np.tile(m.blob[:int(m.blob[L.H_SHARED_LEN])], (4, 1)))


def test_shape_tuning_measures_both_shapes_and_changes_no_result(backend):
    """A specialised library times the 1024- and the 512-lane CTA shape on the first two multi-wave runs of a batch size and keeps the
    faster (mlb_run, tunedShape). The shape changes where lanes run, not what they compute: every run returns the same results (the
    10^6-lane adaptive sweeps of profiles/r3l count identical steps in both shapes; gated here at 1e-12)."""
    m = T.buildFrom("MonteCarlo.mapleaf", {"SimControl.EndCondition": "Altitude", "SimControl.EndConditionValue": "0"})
    n = 2 * 148 * 1024 + 17 * 32              # two waves of the chip and a bit: a batch size no other test uses
    rng = np.random.default_rng(12)
    w = rng.normal(size=(3, n)) * np.array([[5.0], [5.0], [0.5]]) + np.array([[2.0], [0.0], [0.0]])
    shapes, outs = [], []
    with backend.TrajectoryBatch(m.blob, specialize=True) as b:
        b.setLanes(n, {L.LANE_WIND: w})
        for _ in range(3):
            b.run()
            res, st = b.results()                      # synchronises: the next run can read this one's time
            shapes.append(b.counters()["flight_block"])
            outs.append((res.copy(), st.copy()))
    assert shapes[0] == 1024 and shapes[1] == 512 and shapes[2] in (512, 1024)
    assert (outs[0][1] == L.ST_OK).all()
    for res, st in outs[1:]:
        assert np.allclose(res, outs[0][0], rtol=1e-12, atol=1e-9) and np.array_equal(st, outs[0][1])
    # a second handle of the same library and batch size starts with the settled shape
    with backend.TrajectoryBatch(m.blob, specialize=True) as b:
        b.setLanes(n, {L.LANE_WIND: w})
        b.run()
        assert b.counters()["flight_block"] == shapes[2]
