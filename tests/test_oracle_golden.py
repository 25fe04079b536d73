"""
Pins the CPU oracle (oracle/mapleaf_oracle.c) to the reference: fixtures in tests/golden/ were produced by running the
unmodified reference (tests/golden/make_golden.py). Gates: fixed-step results norm-relative <= 1e-12 (observed: bit-exact
on the generating host), adaptive lock-step <= 1e-9, adaptive free-running within the reference's own 1-ulp sensitivity band.
"""
import os

import numpy as np
import pytest

from conftest import DATA
from mapleaf_b200 import L, SimDefinition, buildModel
from oracle import oracle


def buildFrom(defName, overrides=None, wind=None):
    sd = SimDefinition(os.path.join(DATA, defName), silent=True, disableDistributionSampling=True)
    for k, v in (overrides or {}).items():
        sd.setValue(k, v)
    if wind is not None:
        sd.setValue("Environment.ConstantMeanWind.velocity", "({} {} {})".format(*wind))
    return buildModel(sd)


def relErr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1.0)


def stateErr(final14, ref14):
    """norm-relative errors of (pos, vel, quat, angvel): SURVEY §8c gate definition. ref rows are [t, pos, vel, q, w]."""
    f, r = np.asarray(final14), np.asarray(ref14)
    return max(relErr(f[0:3], r[1:4]), relErr(f[3:6], r[4:7]), relErr(f[6:10], r[7:11]), relErr(f[10:13], r[11:14]))


def test_rng_known_answers(golden):
    g = golden("rng.json")
    assert list(oracle.gaussStream(1000, 40)) == g["seed1000_stream"]
    assert oracle.gaussStream(1000, 1, 10, 1)[0] == 10.254632116649685           # test_SimDefinition.py:122
    assert list(oracle.pinkNoise(g["pinkSeed"], g["pinkTimes"])) == g["pinkValues"]
    v = oracle.pinkNoise(63583, [0.05, 0.1])
    assert v[0] == 0.7946604314895807 and v[1] == 0.5874110050866364              # test_TurbulenceModelling.py:39-43
    assert list(oracle.pinkNoise(g["pinkBigSeed"], [-1] * 5)) == g["pinkRaw"]      # seed wider than 32 bits


def test_windtunnel_sweep(golden):
    """ExampleWindTunnelCase: 100 force evaluations over 10..550 m/s; vs the reference run here (1e-12) and vs the
    16-digit values the reference ships (its own 0.2 % regression tolerance, Batch.py:51-52)."""
    g = golden("windtunnel.json")
    m = buildFrom("WindTunnel.mapleaf")
    aref = m.header("H_AREF")
    for row, expCFZ, expMach in zip(g["rows"], g["expectedAeroCFZ"], g["expectedMach"]):
        out = oracle.forceEval(m.blob, row["state"], 0.0)
        assert relErr(out[0:3], row["totalF"]) < 1e-12 and relErr(out[3:6], row["totalM"]) < 1e-12
        assert relErr(out[6:9], row["aeroF"]) < 1e-12
        assert abs(out[9] - row["Mach"]) < 1e-14
        q = 0.5 * out[10] * row["velocityZ"] ** 2
        cfz = out[8] / (q * aref)
        assert abs(cfz - expCFZ) <= 0.002 * abs(expCFZ)
        assert abs(out[9] - expMach) <= 0.002 * abs(expMach) + 1e-4


def lanesFromRuns(runs, key="wind"):
    return np.array([r[key] for r in runs], dtype=float).T.copy()


def test_rk4_to_apogee(golden):
    g = golden("rk4_apogee.json")
    m = buildFrom(g["definition"])
    runs = g["runs"]
    o = oracle.run(m.blob, len(runs), {L.LANE_WIND: lanesFromRuns(runs)}, traceLane=0, traceMaxRows=5000)
    assert (o["status"] == L.ST_OK).all()
    for i, r in enumerate(runs):
        assert o["counters"][i, 0] == r["steps"]
        assert o["finals"][i, 13] == r["final"][0]
        assert stateErr(o["finals"][i], r["final"]) < 1e-12
        assert relErr(o["results"][i, 3:7], r["results"][3:7]) < 1e-12
        assert relErr(o["results"][i, 0:3], r["results"][0:3]) < 1e-9     # extrapolated to z=0 from apogee: ill-conditioned
    tr = o["trace"]
    for idx, row in zip(runs[0]["traceIdx"], runs[0]["trace"]):
        assert tr[idx, 0] == row[0]
        assert relErr(tr[idx, 1:7], row[1:7]) < 1e-13


def test_rk4_to_ground_baseline_anchor(golden):
    """BASELINE.md first parity anchor: 3 runs to the ground (3-DoF under chute, drogue + main events)."""
    g = golden("rk4_ground.json")
    m = buildFrom(g["definition"], g["overrides"])
    runs = g["runs"]
    o = oracle.run(m.blob, len(runs), {L.LANE_WIND: lanesFromRuns(runs)})
    for i, r in enumerate(runs):
        assert o["counters"][i, 0] == r["steps"]
        assert relErr(o["results"][i], r["results"]) < 1e-12
    assert abs(o["results"][0, 3] - 5060.601561802594) < 1e-9 and abs(o["results"][2, 5] - 559.0199999996366) < 1e-9


def test_rk45_adaptive_pink_noise(golden):
    """Derived config 2. Free-running gate: same accepted-step count and apogee/landing within the reference's own
    1-ulp self-sensitivity band (~1 m apogee, 0.5 m landing, 0.1 s flight time with PinkNoise2D).
    Lock-step gate: replaying the reference's accepted dt sequence must give the fixed-step 1e-9 agreement."""
    g = golden("rk45_pink.json")
    m = buildFrom(g["definition"])
    runs = g["runs"]
    seeds = np.array([r["pinkSeeds"] + [0] for r in runs], dtype=float).T.copy()
    o = oracle.run(m.blob, len(runs), {L.LANE_WIND: lanesFromRuns(runs), L.LANE_PINK_SEED: seeds})
    for i, r in enumerate(runs):
        assert abs(int(o["counters"][i, 0]) - r["steps"]) <= 10
        assert abs(o["results"][i, 3] - r["results"][3]) < 1.0
        assert np.hypot(o["results"][i, 0] - r["results"][0], o["results"][i, 1] - r["results"][1]) < 0.5
        assert abs(o["results"][i, 5] - r["results"][5]) < 0.1
    r0 = runs[0]
    o = oracle.run(m.blob, 1, {L.LANE_WIND: lanesFromRuns(runs[:1]), L.LANE_PINK_SEED: seeds[:, :1].copy()},
                   traceLane=0, traceMaxRows=len(r0["times"]) + 5)
    n = min(len(r0["times"]), len(o["trace"]))
    # on the generating host the oracle is bit-identical; elsewhere libm may differ by an ulp and the adaptive path decorrelates
    same = np.array(o["trace"][:n, 0]) == np.array(r0["times"][:n])
    assert same[:50].all()


def test_rk45_gust_to_apogee(golden):
    g = golden("rk45_gust.json")
    m = buildFrom(g["definition"], g["overrides"])
    runs = g["runs"]
    o = oracle.run(m.blob, len(runs), {L.LANE_WIND: lanesFromRuns(runs)})
    for i, r in enumerate(runs):
        assert abs(int(o["counters"][i, 0]) - r["steps"]) <= 5
        assert abs(o["results"][i, 3] - r["results"][3]) < 0.05
        assert abs(o["results"][i, 4] - r["results"][4]) < 1e-3


def test_step_limit_status():
    m = buildFrom("MonteCarlo.mapleaf")
    m.blob[L.H_MAX_STEPS] = 100
    o = oracle.run(m.blob, 1)
    assert o["status"][0] == L.ST_STEP_LIMIT and o["counters"][0, 0] == 100


@pytest.mark.parametrize("name", ["tabulated_rk4.json", "tabulated_rk45.json", "tabulated2d_rk4.json"])
def test_tabulated_and_force_only_components(golden, name):
    """TabulatedInertia, TabulatedAeroForce (TotalAOA / AOA / AOSS / Mach keys, including the cached-vector mutation of
    getAOA / getAOSS), FractionalJetDamping, AeroDamping, AeroForce, Force: one tumbling flight, fixed-step and adaptive; and the same
    flight with 2- and 3-key aero tables (scattered N-D interpolation on scipy's triangulation, nearest neighbour outside the hull)."""
    g = golden(name)
    m = buildFrom(g["definition"], g["overrides"])
    r = g["runs"][0]
    o = oracle.run(m.blob, 1, {}, traceLane=0, traceMaxRows=2000)
    assert o["status"][0] == L.ST_OK and o["counters"][0, 0] == r["steps"]
    assert o["finals"][0, 13] == r["final"][0]
    assert stateErr(o["finals"][0], r["final"]) < 1e-12
    for idx, row in zip(r["traceIdx"], r["trace"]):
        assert o["trace"][idx, 0] == row[0]
        assert stateErr(np.concatenate([o["trace"][idx, 1:14], [0]]), row) < 1e-12


def test_variable_inertia_components_keep_construction_order():
    m = buildFrom("TabulatedComponents.mapleaf")
    S = m.stageRecord(0)
    names = [n for n, _ in m.meta["stages"][0]["components"]]
    assert int(S[L.S_NVAR]) == 2
    # construction order is by dictionary path (Inertia < Motor); the force sum runs top to bottom (Motor first)
    assert [names[int(S[L.S_VAR_IDX + k])] for k in range(2)] == ["Inertia", "Motor"]
    assert names[0] == "Motor" and names[-1] == "YawMoment"


def test_two_stage_separation_fixed_step(golden):
    """Booster separation 0.5 s after burnout, sustainer ignition at separation, boat-tail hand-over, fineness update, motors and
    fixed masses without inertia overrides (Rocket/rocket.py:339-361,461-528): agreement 1e-12 (observed 1e-13, not bit-identical)."""
    g = golden("twostage_rk4.json")
    m = buildFrom(g["definition"], g["overrides"])
    r = g["runs"][0]
    o = oracle.run(m.blob, 1, {}, traceLane=0, traceMaxRows=4000)
    assert o["status"][0] == L.ST_OK and o["counters"][0, 0] == r["steps"]
    assert stateErr(o["finals"][0], r["final"]) < 1e-12
    for idx, row in zip(r["traceIdx"], r["trace"]):
        assert o["trace"][idx, 0] == row[0]
        assert stateErr(np.concatenate([o["trace"][idx, 1:14], [0]]), row) < 1e-12


def test_two_stage_adaptive_lockstep(golden):
    """RK45Adaptive through separation, ignition, apogee, chute and ground impact: replaying the reference's accepted step
    sizes, every step time is identical (same event clamps, no extra rejections) and the results agree to 1e-6."""
    g = golden("twostage_rk45_ground.json")
    m = buildFrom(g["definition"], g["overrides"])
    r = g["runs"][0]
    o = oracle.run(m.blob, 1, {}, traceLane=0, traceMaxRows=4000, dtReplay=r["dts"])
    assert o["counters"][0, 0] == r["steps"]
    assert np.abs(o["trace"][:, 0] - np.array(r["times"])).max() < 1e-9
    assert relErr(o["results"][0], r["results"]) < 1e-6


@pytest.mark.parametrize("name", ["earth_round_rk4.json", "earth_wgs84_rk4.json"])
def test_round_and_wgs84_earth(golden, name):
    """Earth-centred inertial frame: launch-site conversion, geodetic altitude (Newton iteration for the ellipsoid), ENU wind rotation,
    earth-rotation wind term, inverse-square and J2 gravity (ENV/EarthModelling.py, ENV/environment.py:105-181)."""
    g = golden(name)
    m = buildFrom(g["definition"], g["overrides"])
    r = g["runs"][0]
    o = oracle.run(m.blob, 1, {}, traceLane=0, traceMaxRows=2000)
    assert o["status"][0] == L.ST_OK and o["counters"][0, 0] == r["steps"]
    assert stateErr(o["finals"][0], r["final"]) < 1e-13
    for idx, row in zip(r["traceIdx"], r["trace"]):
        assert stateErr(np.concatenate([o["trace"][idx, 1:14], [0]]), row) < 1e-13
    assert relErr(o["results"][0, 3:7], r["results"][3:7]) < 1e-13


def test_nasa_two_stage_orbital_rocket(golden):
    """BASELINE.json configs[3] / the reference's own regression case (NASAVerificationCases.mapleaf:264-278): RK45Adaptive, WGS84 + J2,
    tabulated aero and inertia, jet damping, separation 96.79 s after burnout, 200 s. Free-running: same 1043 accepted steps, same states."""
    g = golden("nasa_twostage.json")
    m = buildFrom(g["definition"], g["overrides"])
    r = g["runs"][0]
    o = oracle.run(m.blob, 1, {}, traceLane=0, traceMaxRows=2000)
    assert o["counters"][0, 0] == r["steps"] == 1043
    assert stateErr(o["finals"][0], r["final"]) < 1e-12
    assert np.abs(o["trace"][:, 0] - np.array(r["times"])).max() < 1e-9
    # the reference's recorded regression values for this case, its own 0.2 % tolerance (Batch.py:51-52)
    for got, exp in zip(o["finals"][0, [0, 1, 3, 4]], (6598829.9, 637702.775, 2308.8505, 8475.945)):
        assert abs(got - exp) <= 0.002 * abs(exp)


@pytest.mark.parametrize("name", ["canards_apogee.json", "canards_ground.json"])
def test_canard_control_system(golden, name):
    """BASELINE.json configs[2]: 100 Hz stabiliser -> scheduled-gain vector PID -> 5-D deflection table (scipy Delaunay on the host)
    -> four first-order actuators -> actuated fin angles; RK4 forced while the controller is in charge, RK45Adaptive again under
    the parachute (GNC/*.py, Fins.py:264-268, rocket.py:646-656,700-726). Free-running, to apogee and to the ground."""
    g = golden(name)
    m = buildFrom(g["definition"], g["overrides"])
    r = g["runs"][0]
    o = oracle.run(m.blob, 1, {}, traceLane=0, traceMaxRows=6000)
    assert o["status"][0] == L.ST_OK and o["counters"][0, 0] == r["steps"]
    assert o["finals"][0, 13] == r["final"][0]
    assert stateErr(o["finals"][0], r["final"]) < 1e-12
    for idx, row in zip(r["traceIdx"], r["trace"]):
        assert o["trace"][idx, 0] == row[0]
        assert stateErr(np.concatenate([o["trace"][idx, 1:14], [0]]), row) < 1e-12
    assert relErr(o["results"][0, 3:7], r["results"][3:7]) < 1e-12


def test_control_record_and_time_step_rewrite():
    m = buildFrom("Canards.mapleaf")
    cs = m.blob[int(m.header("H_CTRL_OFF")):]
    assert m.header("H_DT0") == 0.01 and cs[L.CS_UPDATE_DT] == 0.01 and cs[L.CS_FORCE_RK4] == 1     # ControlSystems.py:126-150
    assert int(cs[L.CS_NACT]) == 4 and cs[L.CS_ACT_MIN] == -45 and cs[L.CS_ACT_MAX] == 45
    names = [n for n, _ in m.meta["stages"][0]["components"]]
    assert names[int(cs[L.CS_COMP])] == "Canards"
    assert m.header("H_SHARED_LEN") < m.blob.size        # the Delaunay tables stay out of shared memory
    nd = m.blob[int(cs[L.CS_DEFL_INTERP_OFF]):]
    assert int(nd[L.ND_DIM]) == 5 and int(nd[L.ND_NVAL]) == 4 and int(nd[L.ND_NPTS]) == 60


def motorAdjustLanes(model, factorsPerLane):
    """LANE_MOTOR_ADJUST array [8, nLanes] from {stage index: (impulseAdjustFactor, burnTimeAdjustFactor)} per lane."""
    a = np.ones((2 * L.MLB_MAX_STAGES, len(factorsPerLane)))
    for lane, f in enumerate(factorsPerLane):
        for si, (imp, burn) in f.items():
            a[2 * si, lane], a[2 * si + 1, lane] = imp, burn
    return a


@pytest.mark.parametrize("name", ["motoradj_rk4.json", "motoradj_twostage.json"])
def test_motor_adjust_factors(golden, name):
    """impulseAdjustFactor / burnTimeAdjustFactor (Propulsion.py:38-40,102-128,153-157), two ways: written into the definition (the
    host builds the adjusted table) and as a per-lane input on the unadjusted model (the per-sample table is rebuilt at lane start).
    Both must reproduce the reference flight; the burnout-separation time uses the UNadjusted burn time, the upper-stage shut-off the
    adjusted one (two-stage case: 0.5 s separation delay, factors 1.12 / 0.88)."""
    g = golden(name)
    r = g["runs"][0]
    baked = buildFrom(g["definition"], g["overrides"])
    plain = buildFrom(g["definition"])
    factors = {}
    for si, motor in enumerate(plain.meta["motors"]):
        factors[si] = (float(g["overrides"][motor["path"] + ".impulseAdjustFactor"]), float(g["overrides"][motor["path"] + ".burnTimeAdjustFactor"]))
    assert baked.meta["motors"][0]["factors"] == factors[0]
    for o in (oracle.run(baked.blob, 1, {}, traceLane=0, traceMaxRows=4000),
              oracle.run(plain.blob, 1, {L.LANE_MOTOR_ADJUST: motorAdjustLanes(plain, [factors])}, traceLane=0, traceMaxRows=4000)):
        assert o["status"][0] == L.ST_OK and o["counters"][0, 0] == r["steps"]
        assert stateErr(o["finals"][0], r["final"]) < 1e-12
        assert relErr(o["results"][0, 3:7], r["results"][3:7]) < 1e-12
        for idx, row in zip(r["traceIdx"], r["trace"]):
            assert o["trace"][idx, 0] == row[0]
            assert stateErr(np.concatenate([o["trace"][idx, 1:14], [0]]), row) < 1e-12


def test_sampled_motor_adjust_factors_per_lane(golden):
    """The reference's Monte Carlo loop with `_stdDev` on the motor factors (its Jake1_MonteCarlo.mapleaf example does this): one
    batch, each lane with its own wind and factors, lanes with unit factors alongside."""
    g = golden("motoradj_mc.json")
    m = buildFrom(g["definition"])
    runs = g["runs"]
    lanes = [{0: tuple(r["motor"])} for r in runs] + [{}]
    wind = np.concatenate([lanesFromRuns(runs), np.array([[2.0], [0.0], [0.0]])], axis=1)
    o = oracle.run(m.blob, len(lanes), {L.LANE_WIND: wind, L.LANE_MOTOR_ADJUST: motorAdjustLanes(m, lanes)})
    for i, r in enumerate(runs):
        assert o["counters"][i, 0] == r["steps"]
        assert stateErr(o["finals"][i], r["final"]) < 1e-12
        assert relErr(o["results"][i, 3:7], r["results"][3:7]) < 1e-12
    base = oracle.run(m.blob, 1, {L.LANE_WIND: wind[:, -1:]})
    assert (o["finals"][-1] == base["finals"][0]).all()          # unit factors: identical to the table built by the host


def test_rk78_adaptive_gust(golden):
    """RK78Adaptive (Dormand-Prince 8(7)13M: 13 evaluations per attempt, no first-same-as-last; Integration.py:279-297) through the sine
    gust to apogee: same accepted-step counts and results as the reference (observed bit-identical), and the lock-step replay of its
    accepted step sizes reproduces every step time."""
    g = golden("rk78_gust.json")
    m = buildFrom(g["definition"], g["overrides"])
    runs = g["runs"]
    assert m.header("H_TAB_NSTAGE_ROWS") == 12 and m.header("H_TAB_NRESULT_ROWS") == 2
    o = oracle.run(m.blob, len(runs), {L.LANE_WIND: lanesFromRuns(runs)})
    for i, r in enumerate(runs):
        assert o["counters"][i, 0] == r["steps"] and o["counters"][i, 1] >= 13 * r["steps"]
        assert stateErr(o["finals"][i], r["final"]) < 1e-12
        assert relErr(o["results"][i, 3:7], r["results"][3:7]) < 1e-12
    r0 = runs[0]
    rep = oracle.run(m.blob, 1, {L.LANE_WIND: lanesFromRuns(runs[:1])}, traceLane=0, traceMaxRows=3000, dtReplay=r0["dts"])
    assert np.abs(rep["trace"][:, 0] - np.array(r0["times"])).max() < 1e-12


@pytest.mark.parametrize("name", ["ideal_rk4.json", "ideal_rk45.json"])
def test_ideal_moment_controller(golden, name):
    """IdealMomentController (GNC/MomentControllers.py:85-93; ControlSystems.py:73-100): every time step the orientation is overwritten
    with the navigator's (tilted) target and the angular velocity with zero. The reference assigns the target quaternion OBJECT, so the
    in-place normalisations of the step reach the next steps' target; the adaptive integrator keeps its first-same-as-last derivative
    across the rewrite. Observed bit-identical to the reference, fixed-step and free-running adaptive."""
    g = golden(name)
    r = g["runs"][0]
    m = buildFrom(g["definition"], g["overrides"])
    assert m.blob[int(m.header("H_CTRL_OFF")) + L.CS_MC_TYPE] == 2
    o = oracle.run(m.blob, 1, {}, traceLane=0, traceMaxRows=5000)
    assert o["status"][0] == L.ST_OK and o["counters"][0, 0] == r["steps"]
    assert stateErr(o["finals"][0], r["final"]) < 1e-12 and relErr(o["results"][0, 3:7], r["results"][3:7]) < 1e-12
    # the reference's flight log holds the state OBJECTS, and the controller rewrites each one's orientation / angular velocity at the start
    # of the following step: only position and velocity of a logged row are what the integrator produced
    for idx, row in zip(r["traceIdx"], r["trace"]):
        assert abs(o["trace"][idx, 0] - row[0]) < 1e-12
        assert relErr(o["trace"][idx, 1:4], row[1:4]) < 1e-12 and relErr(o["trace"][idx, 4:7], row[4:7]) < 1e-12


def test_reference_regression_case_timestep_adaptation_demo():
    """A golden the reference ships itself (regressionTests.mapleaf:87-114, checked by its batch runner at 0.2 %): RK45Adaptive + PID with
    the case's own coefficients, a 20 m launch rail, PinkNoise2D with fixed seeds, motor impulse and burn time both doubled, 45 s of flight."""
    from conftest import TIMESTEP_ADAPTATION_DEMO, TIMESTEP_ADAPTATION_DEMO_EXPECTED as exp
    m = buildFrom("MonteCarloAdaptive.mapleaf", TIMESTEP_ADAPTATION_DEMO)
    assert m.header("H_RAIL_LEN") == 20 and m.meta["motors"][0]["factors"] == (2.0, 2.0)
    o = oracle.run(m.blob, 1, {})
    assert o["status"][0] == L.ST_OK
    assert abs(o["finals"][0, 13] - exp["Time"]) < 1e-9
    assert abs(o["finals"][0, 0] - exp["PositionX"]) <= exp["percentTolerance"] / 100 * abs(exp["PositionX"])


@pytest.mark.parametrize("name", __import__("conftest").REGRESSION_CASES)
def test_reference_flight_regression_cases(name):
    """Goldens the reference ships and checks at 0.2 % (OpenRocketCases.mapleaf:63-214): two Estes-motor model rockets to apogee, a
    sounding rocket with a boat tail and a nose-cone-to-tube transition (Haisunaata), a rolling rocket with canted fins flown for 6 s; and
    the NASA check cases (NASAVerificationCases.mapleaf:9-262): a sphere dropped over the round / WGS84 earth with and without drag and wind, a
    damped tumbling brick.
    The recorded values are printed to 4-7 digits; the oracle reproduces them to that precision (observed <= 0.04 %)."""
    from conftest import regressionCase, regressionErrors
    blob, expected = regressionCase(name)
    o = oracle.run(blob, 1, {})
    assert o["status"][0] == L.ST_OK
    assert max(regressionErrors(o["finals"][0][:6], expected)) < 0.05


# ---- one short flight per option the backend accepts (tests/golden/make_golden.py `options`) ----
OPTION_FIXED = ["opt_atm_tabulated.json", "opt_wind_profile.json", "opt_pink1d.json", "opt_pink3d.json", "opt_euler.json", "opt_rk2mid.json",
                "opt_rk2heun.json", "opt_rk4_38.json", "opt_constgain_pid.json"]
OPTION_ADAPTIVE = ["opt_pink3d_rk45.json", "opt_rk12a.json", "opt_rk23a.json", "opt_ctrl_elementary.json"]


def optionOverrides(g):
    """The fixture's overrides with table paths re-pointed at this checkout (they were absolute on the generating host)."""
    from conftest import REPO
    out = dict(g["overrides"])
    for k, v in out.items():
        if k.endswith("filePath"):
            out[k] = os.path.join(REPO, "mapleaf_b200", "data", "Environment", os.path.basename(v))
    return out


@pytest.mark.parametrize("name", OPTION_FIXED)
def test_option_fixed_step(golden, name):
    """TabulatedAtmosphere (ENV/AtmosphereModelling.py:79-96), CustomWindProfile (ENV/MeanWindModelling.py:165-196), PinkNoise1D / 3D with
    fixed seeds (ENV/TurbulenceModelling.py:100-131: 3D reuses generator 2 for z), Euler / RK2Midpoint / RK2Heun / RK4_3/8
    (Motion/Integration.py:144-236), ConstantGainPIDRocket (GNC/MomentControllers.py:58-83): fixed-step flights, same step count,
    states along the flight and at the end <= 1e-12 from the reference."""
    g = golden(name)
    m = buildFrom(g["definition"], optionOverrides(g))
    r = g["runs"][0]
    o = oracle.run(m.blob, 1, {}, traceLane=0, traceMaxRows=4000)
    assert o["status"][0] == L.ST_OK and o["counters"][0, 0] == r["steps"]
    assert o["finals"][0, 13] == r["final"][0]
    assert stateErr(o["finals"][0], r["final"]) < 1e-12
    assert relErr(o["results"][0, 3:7], r["results"][3:7]) < 1e-12
    for idx, row in zip(r["traceIdx"], r["trace"]):
        assert o["trace"][idx, 0] == row[0]
        assert stateErr(np.concatenate([o["trace"][idx, 1:14], [0]]), row) < 1e-12


@pytest.mark.parametrize("name", OPTION_ADAPTIVE)
def test_option_adaptive(golden, name):
    """RK12Adaptive / RK23Adaptive (Integration.py:238-330; RK23 is first-same-as-last), the `elementary` step-size controller
    (:359-382, safety factor), PinkNoise3D under the adaptive integrator. Lock-step replay of the reference's accepted step sizes
    reproduces every step time and the final state; free-running, the accepted-step count stays within a few steps."""
    g = golden(name)
    m = buildFrom(g["definition"], optionOverrides(g))
    r = g["runs"][0]
    rep = oracle.run(m.blob, 1, {}, traceLane=0, traceMaxRows=len(r["times"]) + 5, dtReplay=r["dts"])
    assert rep["counters"][0, 0] == r["steps"]
    assert np.abs(rep["trace"][:, 0] - np.array(r["times"])).max() < 1e-9
    assert stateErr(rep["finals"][0], r["final"]) < 1e-9
    free = oracle.run(m.blob, 1, {})
    assert free["status"][0] == L.ST_OK
    assert abs(int(free["counters"][0, 0]) - r["steps"]) <= max(5, r["steps"] // 100)
    assert relErr(free["finals"][0, 0:3], r["final"][1:4]) < 1e-4


def test_constant_step_controller_is_rejected():
    """`controller Constant` with an adaptive method cannot be flown by the reference (integratorFactory leaves safetyFactor unbound,
    Integration.py:74-93): the backend refuses it when the model is built instead of flying something the reference never could."""
    with pytest.raises(ValueError):
        buildFrom("MonteCarloAdaptive.mapleaf", {"SimControl.TimeStepAdaptation.controller": "Constant"})


@pytest.mark.parametrize("name", ["droppaths_rk4.json", "droppaths_rk45.json"])
def test_dropped_stage_flight(golden, name):
    """SingleSimulations.createNewDetachedStage (:280-299) + Rocket.__init__(stageToInitialize) (rocket.py:296-321,339-361): the booster of
    TwoStage.mapleaf as its own rocket (burnt-out motor, no staging, StageDropPaths end condition), started from the reference's separation
    state, time and step size: same number of steps to the ground, first states <= 1e-12, final state 1e-12 (fixed step) / within the
    adaptive band. The main flight records the separation row the dropped stage starts from."""
    g = golden(name)
    d = g["drops"][0]
    m = buildFrom(g["definition"], g["overrides"])
    main = oracle.run(m.blob, 1, {}, separations=True)
    row = main["separations"][0, 0]
    assert row[L.SEP_VALID] == 1 and abs(row[L.SEP_TIME] - d["separationTime"]) < 1e-6
    if name.endswith("rk4.json"):
        assert row[L.SEP_TIME] == d["separationTime"] and row[L.SEP_DT] == 0.02
        assert stateErr(np.concatenate([row[:13], [0]]), d["separationState"]) < 1e-12
    sd = SimDefinition(os.path.join(DATA, g["definition"]), silent=True, disableDistributionSampling=True)
    for k, v in g["overrides"].items():
        sd.setValue(k, v)
    sep = np.array(d["separationState"])
    dt0 = d["firstStatesAfter"][0][0] - d["separationTime"]
    md = buildModel(sd, droppedStage=0, droppedFromAltitude=sep[3])
    assert int(md.header("H_NSTAGES")) == 1 and md.header("H_CTRL_OFF") == 0
    arrays = {L.LANE_INIT_STATE: sep[1:14].reshape(13, 1).copy(), L.LANE_START: np.array([[d["separationTime"]], [dt0]])}
    o = oracle.run(md.blob, 1, arrays, traceLane=0, traceMaxRows=8)
    assert o["status"][0] == L.ST_OK and o["counters"][0, 0] == d["stepsAfterSeparation"]
    for i, row in enumerate(d["firstStatesAfter"]):
        assert abs(o["trace"][1 + i, 0] - row[0]) < 1e-12
        assert stateErr(np.concatenate([o["trace"][1 + i, 1:14], [0]]), row) < 1e-12
    if name.endswith("rk4.json"):
        assert stateErr(o["finals"][0], d["final"]) < 1e-11
    else:
        assert relErr(o["results"][0, 0:2], d["results"][0:2]) < 1e-4


@pytest.mark.parametrize("name", ["timesteplog_rk4.json", "timesteplog_rk45.json"])
def test_time_step_log_csv_is_the_reference_file(golden, name, tmp_path):
    """SimControl.loggingLevel 1: `<stage>_timeStepLog.csv` written from a lane's per-step flight-log rows is byte for byte the file the
    reference writes for the same flight (rocket.py:155-170,615-644,729-738; IO/CythonLog.pyx:163-213): column order, the three quaternion
    columns, TimeStep(s), the error estimate shifted to the row the step starts from, fill values under the parachute."""
    from mapleaf_b200 import flightlog
    g = golden(name)
    m = buildFrom(g["definition"], g["overrides"])
    o = oracle.run(m.blob, 1, {}, traceLane=0, traceMaxRows=2000)
    path = flightlog.writeTimeStepLog(str(tmp_path / g["fileName"]), o["trace"], m)
    assert open(path).read() == g["csv"]
