"""
CPU tests of the host side: the Monte Carlo runner's sampling order and log format, the C-ABI library's exported symbols,
the bit-exact host samplers and the multi-GPU gather logic (gloo, world size 2). No CUDA device is used; where a run needs
trajectories, the test substitutes the CPU oracle for the device (tests are the only place allowed to do that).
"""
import ctypes
import os
import random
import re
import socket
import sys

import numpy as np
import pytest

from conftest import DATA, REPO
from mapleaf_b200 import L, SimDefinition, buildModel, montecarlo
from mapleaf_b200 import backend, distributed


class OracleBatch:
    """Stands in for backend.TrajectoryBatch in host-logic tests: same methods, trajectories from the CPU oracle."""

    def __init__(self, blob, device=0, specialize=None):
        self.blob, self.device = np.asarray(blob, dtype=np.float64), device

    def setLanes(self, n, arrays=None):
        self.n, self.arrays = n, arrays or {}

    def setLaneModels(self, models):
        self.models = None if models is None else np.asarray(models, dtype=np.float64)

    def run(self, stream=0, sync=False):
        from oracle import oracle
        if getattr(self, "models", None) is None:
            self.out = oracle.run(self.blob, self.n, self.arrays, separations=True)
            return
        # per-lane model blocks: one oracle flight per lane with its own block (shared part from the lane, tables from the handle's block)
        outs = []
        for i in range(self.n):
            blob = np.concatenate([self.models[i], self.blob[self.models.shape[1]:]])
            outs.append(oracle.run(blob, 1, {k: np.ascontiguousarray(a[:, i:i + 1]) for k, a in self.arrays.items()}))
        self.out = {k: np.concatenate([o[k] for o in outs]) for k in ("results", "finals", "status", "counters")}

    def results(self):
        return self.out["results"], self.out["status"]

    def finals(self):
        return self.out["finals"]

    def laneCounters(self):
        c = self.out["counters"]
        return np.concatenate([c, np.zeros((c.shape[0], 3), dtype=c.dtype)], axis=1)

    def stageSeparations(self):
        return self.out["separations"]

    def setFlightLog(self, firstLane, nLanes, every=1, maxRowsPerLane=2048):
        self.logged = (firstLane, nLanes, maxRowsPerLane)

    def flightLog(self):
        from oracle import oracle
        first, n, maxRows = self.logged
        return [oracle.run(self.blob, self.n, self.arrays, traceLane=first + i, traceMaxRows=maxRows)["trace"] for i in range(n)]

    def counters(self):
        c = self.out["counters"]
        return dict(lanes=self.n, steps=int(c[:, 0].sum()), derivative_evals=int(c[:, 1].sum()), rejected_steps=int(c[:, 2].sum()), last_run_ms=1000.0)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def montecarloDefinition(tmp_path, nRuns=3, **overrides):
    sd = SimDefinition(os.path.join(DATA, "MonteCarlo.mapleaf"), silent=True, disableDistributionSampling=True)
    sd.setValue("MonteCarlo.randomSeed", "2394781")
    sd.setValue("MonteCarlo.numberRuns", str(nRuns))
    for k, v in overrides.items():
        sd.setValue(k, v)
    sd = SimDefinition(dictionary=dict(sd.dict), silent=True)           # seeds the rng with the *string* seed, consumes draw set 1
    sd.fileName = str(tmp_path / "MonteCarlo.mapleaf")
    return sd


def test_library_exports_every_declared_symbol():
    path = backend.buildLibrary()
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(REPO, "include", "mapleaf_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mlb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(backend.EXPORTED_SYMBOLS) == sorted(n for n in declared)
    lib.mlb_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.mlb_version()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(backend, "_lib", None)
    monkeypatch.setattr(backend, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        backend.lib()


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    m = buildModel(SimDefinition(os.path.join(DATA, "MonteCarlo.mapleaf"), silent=True, disableDistributionSampling=True))
    with pytest.raises(RuntimeError, match="no CPU path|no CUDA"):
        backend.TrajectoryBatch(m.blob)


def test_host_samplers_are_cpython_exact():
    r = random.Random("2394781")
    ref = [r.gauss(m, s) for _ in range(50) for m, s in ((2, 5), (0, 5), (0, 0.5))]
    got = backend.hostGaussStream("2394781", [2, 0, 0], [5, 5, 0.5], 50)
    assert list(got.reshape(-1)) == ref
    assert list(backend.hostGaussStream("2394781", [2, 0, 0], [5, 5, 0.5], 10, skipDraws=120).reshape(-1)) == ref[120:150]
    r = random.Random(987654321012)
    ref = [float(r.randrange(1000000)) for _ in range(200)]
    assert list(backend.hostRandrangeStream(987654321012, 1000000, 150, skipDraws=50)) == ref[50:]
    assert list(backend.seedKey(2 ** 40 + 5)) == [5, 256]


def test_runner_sampling_order_and_log(monkeypatch, tmp_path, golden, capsys):
    """Run #i of the runner uses draw set i+1 of random.Random("2394781") (set 0 is consumed at construction) and the
    log has the reference's lines in the reference's order (MonteCarlo.py:67, simDefinition.py:508-532, Plotting.py:671-741)."""
    g = golden("rk4_apogee.json")
    monkeypatch.setattr(backend, "TrajectoryBatch", OracleBatch)
    sd = montecarloDefinition(tmp_path, nRuns=3, **{"MonteCarlo.output": "landingLocations apogees"})
    out = montecarlo.runMonteCarloSimulation(simDefinition=sd, silent=True)
    for i in range(3):
        ref = g["runs"][i]
        assert np.allclose(out.results[i, 3:7], ref["results"][3:7], rtol=1e-12)
        assert out.apogees[i] == out.results[i, L.RES_APOGEE]
        assert tuple(out.landingLocations[i]) == tuple(out.results[i, 0:3])
    log = open(out.logPath).read().splitlines()
    body = [l for l in log if not l.startswith("#")]
    assert "Running Monte Carlo Simulation: 3 runs" in body
    i1 = body.index("Monte Carlo Run #1")
    assert body[i1 + 1] == "Sampling vector parameter: Environment.ConstantMeanWind.velocity, value: ({:1.3f} {:1.3f} {:1.3f})".format(*g["runs"][0]["wind"])
    assert body.index("Monte Carlo Run #2") > i1 and "Monte Carlo results:" in body
    assert "Landing locations:   X       Y        Z" in body
    assert any(l.startswith("Mean Landing location: (") for l in body) and any(l.startswith("Standard deviation:    (") for l in body)
    assert "Simulation #1: {}".format(out.apogees[0]) in body
    from statistics import mean, stdev
    assert "Mean Apogee:        {:>8.1f}".format(mean(out.apogees)) in body
    assert "Standard Deviation: {:>8.1f}".format(stdev(out.apogees)) in body
    assert out.logPath.endswith("MonteCarlo_monteCarloLog_run1.txt")


def test_runner_pink_noise_seeds_follow_the_global_stream(monkeypatch, tmp_path, golden):
    g = golden("rk45_pink.json")
    sd = SimDefinition(os.path.join(DATA, "MonteCarloAdaptive.mapleaf"), silent=True, disableDistributionSampling=True)
    sd.setValue("MonteCarlo.randomSeed", g["randomSeed"])
    sd = SimDefinition(dictionary=dict(sd.dict), silent=True)
    model = buildModel(sd)
    random.seed(g["globalSeed"])
    arrays = montecarlo.sampleRuns(sd, len(g["runs"]), montecarlo.MonteCarloLogger(echo=False), model)
    for i, r in enumerate(g["runs"]):
        assert list(arrays[L.LANE_WIND][:, i]) == r["wind"]
        assert list(arrays[L.LANE_PINK_SEED][:2, i]) == [float(s) for s in r["pinkSeeds"]]


def test_runner_routes_sampled_keys(tmp_path):
    """Wind, initial state and motor factors are per-lane arrays; any other sampled key, and a sampled launch position / direction under
    a launch rail (the rail origin and direction follow it, ENV/environment.py:57-91), gives every run its own model block."""
    sd = montecarloDefinition(tmp_path, nRuns=2)
    assert montecarlo._checkSampledKeys(sd, buildModel(sd)) == (["Environment.ConstantMeanWind.velocity"], False)
    sd = montecarloDefinition(tmp_path, nRuns=2, **{"Rocket.Sustainer.UpperBodyTube.mass_stdDev": "0.02"})
    keys, laneModels = montecarlo._checkSampledKeys(sd, buildModel(sd))
    assert "Rocket.Sustainer.UpperBodyTube.mass" in keys and laneModels
    sd = montecarloDefinition(tmp_path, nRuns=2, **{"Rocket.position_stdDev": "(1 1 0)"})
    assert montecarlo._checkSampledKeys(sd, buildModel(sd))[1] is False
    sd = montecarloDefinition(tmp_path, nRuns=2, **{"Rocket.position_stdDev": "(1 1 0)", "Environment.LaunchSite.railLength": "5"})
    assert montecarlo._checkSampledKeys(sd, buildModel(sd))[1] is True
    sd = montecarloDefinition(tmp_path, nRuns=2, **{"Rocket.rotationAngle": "3", "Rocket.rotationAngle_stdDev": "1", "Rocket.rotationAxis": "(0 1 0)",
                                                    "Environment.LaunchSite.railLength": "5"})
    assert montecarlo._checkSampledKeys(sd, buildModel(sd))[1] is True


def test_runner_samples_motor_adjust_factors(monkeypatch, tmp_path, golden):
    """`_stdDev` on a motor's impulseAdjustFactor / burnTimeAdjustFactor (the reference's Jake1_MonteCarlo.mapleaf): drawn after the
    wind, in dictionary order, logged like the reference, and flown per lane."""
    g = golden("motoradj_mc.json")
    monkeypatch.setattr(backend, "TrajectoryBatch", OracleBatch)
    sd = montecarloDefinition(tmp_path, nRuns=3)
    for k, v in g["overrides"].items():          # after construction, as tests/golden/make_golden.py does: draw set 1 sampled the wind only
        sd.setValue(k, v)
    out = montecarlo.runMonteCarloSimulation(simDefinition=sd, silent=True)
    for i, ref in enumerate(g["runs"]):
        assert np.allclose(out.results[i, 3:7], ref["results"][3:7], rtol=1e-12)
    body = open(out.logPath).read().splitlines()
    i1 = body.index("Monte Carlo Run #1")
    assert body[i1 + 2] == "Sampling scalar parameter: Rocket.Sustainer.Motor.impulseAdjustFactor, value: {:1.3f}".format(g["runs"][0]["motor"][0])
    assert body[i1 + 3] == "Sampling scalar parameter: Rocket.Sustainer.Motor.burnTimeAdjustFactor, value: {:1.3f}".format(g["runs"][0]["motor"][1])


def test_sampled_initial_state_matches_model_builder(tmp_path):
    sd = montecarloDefinition(tmp_path, nRuns=2)
    sd.setValue("Rocket.initialDirection", "(0 0.1 1)")
    m = buildModel(sd)
    assert montecarlo.laneInitialState(sd, m) == list(m.blob[L.H_INIT_STATE:L.H_INIT_STATE + 13])


def test_moments_combination_matches_numpy():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(1000, 7)) * np.arange(1, 8) + 100
    parts = [distributed.momentsOf(x[a:b]) for a, b in ((0, 10), (10, 500), (500, 1000))]
    n, mean, std = distributed.combineMoments(parts)
    assert n == 1000 and np.allclose(mean, x.mean(axis=0), rtol=1e-13) and np.allclose(std, x.std(axis=0, ddof=1), rtol=1e-12)
    assert [distributed.shardRange(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]


def _gatherWorker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, REPO)
    from mapleaf_b200 import L, SimDefinition, buildModel, distributed
    from oracle import oracle
    m = buildModel(SimDefinition(os.path.join(REPO, "mapleaf_b200", "data", "Simulations", "MonteCarlo.mapleaf"), silent=True, disableDistributionSampling=True))
    nTotal = 7                                                            # not divisible by the world size: ranks hold 3 and 4 samples
    w = np.random.default_rng(11).normal(size=(3, nTotal)) * 5            # every rank derives the same full sample stream ...
    lo, hi = distributed.shardRange(nTotal, rank, world)                  # ... and flies its own contiguous range
    o = oracle.run(m.blob, hi - lo, {L.LANE_WIND: np.ascontiguousarray(w[:, lo:hi])})
    res, st = distributed.gatherResults(torch.from_numpy(o["results"]), torch.from_numpy(o["status"]))
    n, mean, std = distributed.gatherMoments(*distributed.momentsOf(o["results"], o["status"]))
    np.save(os.path.join(tmp, "res{}.npy".format(rank)), res.numpy())
    np.save(os.path.join(tmp, "mom{}.npy".format(rank)), np.concatenate([[n], mean, std]))
    dist.destroy_process_group()


def test_two_rank_gather_restores_sample_order(tmp_path):
    """world_size 2 over gloo: contiguous shards of unequal size (7 samples), one all-gather of results + status padded to the larger
    shard and trimmed, combined moments on every rank."""
    import torch.multiprocessing as mp
    from oracle import oracle
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_gatherWorker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    m = buildModel(SimDefinition(os.path.join(DATA, "MonteCarlo.mapleaf"), silent=True, disableDistributionSampling=True))
    w = np.random.default_rng(11).normal(size=(3, 7)) * 5
    full = oracle.run(m.blob, 7, {L.LANE_WIND: w})["results"]
    for rank in range(2):
        assert np.array_equal(np.load(tmp_path / "res{}.npy".format(rank)), full)
        mom = np.load(tmp_path / "mom{}.npy".format(rank))
        assert mom[0] == 7 and np.allclose(mom[1:8], full.mean(axis=0), rtol=1e-12) and np.allclose(mom[8:], full.std(axis=0, ddof=1), rtol=1e-10)


def test_bench_flop_constants_are_the_oracle_op_counts():
    """bench.py's algorithmic FLOP figures (the roofline numerator) are exact op counts of the oracle (oracle/opcount.hpp): re-count on two
    lanes of the bench workload; the op-counting build must also reproduce the plain oracle bit for bit."""
    import bench
    from oracle import oracle
    cfg = bench.CONFIGS[1]
    m = bench.buildConfigModel(cfg)
    wind, seeds = np.array([[3.0, -1.0], [1.5, 4.0], [0.2, -0.3]]), np.array([[11.0, 5.0], [7.0, 123456.0], [0.0, 0.0]])
    arrays = {L.LANE_WIND: wind, L.LANE_PINK_SEED: seeds}
    out, ops = oracle.runCounted(m.blob, 2, arrays)
    plain = oracle.run(m.blob, 2, arrays)
    assert (out["finals"] == plain["finals"]).all() and (out["counters"] == plain["counters"]).all()
    per = {k: v["flops"] / v["count"] for k, v in ops["buckets"].items()}
    f = cfg["flops"]
    for key, const in (("eval_6dof", f["eval6"]), ("eval_6dof_transonic", f["eval6t"]), ("eval_3dof", f["eval3"]),
                       ("step_rest_6dof", f["step6"]), ("step_rest_3dof", f["step3"])):
        assert abs(per[key] - const) <= 0.01 * const, (key, per[key], const)
    c = dict(derivative_evals=int(out["counters"][:, 1].sum()), derivative_evals_3dof=ops["buckets"]["eval_3dof"]["count"],
             derivative_evals_transonic=ops["buckets"]["eval_6dof_transonic"]["count"], steps=int(out["counters"][:, 0].sum()),
             steps_3dof=ops["buckets"]["step_rest_3dof"]["count"], rejected_steps=int(out["counters"][:, 2].sum()))
    assert abs(sum(bench.algorithmicFlops(c, f)) - ops["flops"]) <= 0.01 * ops["flops"]          # the bench formula reproduces the total


def test_bench_reference_arm_inputs_match_the_product_sampler():
    """The reference arm draws its inputs with CPython's `random` alone (no product library on that arm); the product arm draws the same
    samples with the C++ sampler."""
    import bench
    w1, s1 = bench.sampleInputs(5, 7)
    w2, s2 = bench.sampleInputsPython(5, 7)
    assert (w1 == w2).all() and (s1 == s2).all()
    assert (bench.motorFactors(3, 5) == bench.motorFactors(3, 5, python=True)).all()


def test_runner_flight_paths_output(monkeypatch, tmp_path, golden):
    """`MonteCarlo.output flightPaths` (MonteCarlo.py:60-62): every run's path, 900 frames evenly spaced in time, ending where the run ended."""
    g = golden("rk4_apogee.json")
    monkeypatch.setattr(backend, "TrajectoryBatch", OracleBatch)
    monkeypatch.setattr(montecarlo, "_FLIGHT_PATH_ROWS", 4000)
    sd = montecarloDefinition(tmp_path, nRuns=2, **{"MonteCarlo.output": "flightPaths apogees"})
    out = montecarlo.runMonteCarloSimulation(simDefinition=sd, silent=True)
    assert len(out.flights) == 2
    for i, f in enumerate(out.flights):
        ref = g["runs"][i]
        assert f.times.shape == (900,) and f.positions.shape == (900, 3) and f.orientations.shape == (900, 4)
        assert f.times[0] == 0 and abs(f.times[-1] - ref["final"][0]) < 1e-12
        assert np.allclose(f.positions[-1], ref["final"][1:4], rtol=1e-12) and np.allclose(f.velocities[-1], ref["final"][4:7], rtol=1e-9, atol=1e-9)
        assert abs(f.positions[:, 2].max() - ref["results"][3]) < 1e-3 * ref["results"][3]          # apogee is the last frame here
        assert np.allclose(np.linalg.norm(f.orientations, axis=1), 1.0)


def test_vectorized_sampling_is_the_per_run_loop(monkeypatch, tmp_path):
    """The C++ sampler continuing the SimDefinition's own generator and the global `random` module (mlb_host_*_from_state) reproduces the
    per-run Python loop: same per-lane arrays, same log, same dictionary, and both generators left in the same state."""
    def prepared():
        sd = SimDefinition(os.path.join(DATA, "MonteCarloAdaptive.mapleaf"), silent=True, disableDistributionSampling=True)
        sd.setValue("MonteCarlo.randomSeed", "2394781")
        sd.setValue("Rocket.Sustainer.Motor.impulseAdjustFactor", "1.0")
        sd.setValue("Rocket.Sustainer.Motor.impulseAdjustFactor_stdDev", "0.02331")
        sd = SimDefinition(dictionary=dict(sd.dict), silent=True)
        return sd, buildModel(sd)
    n = 300
    out = []
    for threshold in (10 ** 9, 1):
        monkeypatch.setattr(montecarlo, "_VECTORIZED_SAMPLING_FROM", threshold)
        sd, model = prepared()
        random.seed(99)
        log = montecarlo.MonteCarloLogger(echo=False)
        sd.monteCarloLogger = log          # as _prepSim does
        arrays = montecarlo.sampleRuns(sd, n, log, model)
        sd.monteCarloLogger = None
        out.append((arrays, list(log.monteCarloLog[4:]), dict(sd.dict), sd.rng.getstate(), random.getstate(), sd.rng.gauss(0, 1), random.random()))
    slow, fast = out
    assert sorted(slow[0]) == sorted(fast[0]) == [L.LANE_WIND, L.LANE_PINK_SEED, L.LANE_MOTOR_ADJUST]
    for k in slow[0]:
        assert np.array_equal(slow[0][k], fast[0][k])
    assert slow[1] == fast[1] and slow[2] == fast[2]
    assert slow[3] == fast[3] and slow[4] == fast[4] and slow[5:] == fast[5:]


@pytest.mark.parametrize("name,keys", [("perlane_mass_span.json", ("Rocket.Sustainer.GeneralMass.mass", "Rocket.Sustainer.TailFins.span")),
                                       ("perlane_rail_position.json", ("Rocket.position", "Rocket.rotationAngle"))])
def test_runner_samples_any_key_through_per_lane_models(monkeypatch, tmp_path, golden, name, keys):
    """`_stdDev` on keys that change what the constructors precompute (a mass and a fin span; the launch position and tilt under a launch
    rail): the runner draws them in the reference's order (IO/simDefinition.py:471-532), builds one model block per run and flies them as
    one batch; every run reproduces the reference's flight of the same sample (tests/golden/make_golden.py `perlane`)."""
    g = golden(name)
    monkeypatch.setattr(backend, "TrajectoryBatch", OracleBatch)
    sd = montecarloDefinition(tmp_path, nRuns=3, **{"MonteCarlo.output": "landingLocations apogees"})
    for k, v in g["overrides"].items():          # as the fixture's generator: added after the definition's construction draw
        sd.setValue(k, v)
    out = montecarlo.runMonteCarloSimulation(simDefinition=sd, silent=True)
    assert (out.status == L.ST_OK).all()
    for i, ref in enumerate(g["runs"]):
        assert np.allclose(out.results[i, 3:7], ref["results"][3:7], rtol=1e-12, atol=1e-12), (i, out.results[i], ref["results"])
    log = open(out.logPath).read()
    for k in keys:
        assert "parameter: " + k in log


@pytest.mark.parametrize("name", ["droppaths_rk4.json", "droppaths_rk45.json"])
def test_runner_flies_dropped_stages(monkeypatch, tmp_path, golden, name):
    """SimControl.StageDropPaths.compute (SingleSimulations.py:91-145,280-299): the booster of TwoStage.mapleaf flown on from the separation
    as a lane of a second batch (model block of the dropped stage alone, burnt-out motor, separation state / time / step size from the main
    batch). Fixed step: same number of steps after the separation and the same final state as the reference's dropped-stage flight."""
    g = golden(name)
    monkeypatch.setattr(backend, "TrajectoryBatch", OracleBatch)
    sd = SimDefinition(os.path.join(DATA, g["definition"]), silent=True, disableDistributionSampling=True)
    for k, v in {**g["overrides"], "MonteCarlo.numberRuns": "2", "MonteCarlo.output": "landingLocations"}.items():
        sd.setValue(k, v)
    sd.fileName = str(tmp_path / "TwoStage.mapleaf")
    out = montecarlo.runMonteCarloSimulation(simDefinition=sd, silent=True)
    assert len(out.droppedStages) == 1
    d, ref = out.droppedStages[0], g["drops"][0]
    assert list(d["lanes"]) == [0, 1] and (d["status"] == L.ST_OK).all()
    fixed = name.endswith("rk4.json")
    for lane in range(2):            # nothing is sampled in this definition: both runs are the reference's flight
        if fixed:
            assert abs(d["finals"][lane, 13] - ref["final"][0]) < 1e-9
            assert np.linalg.norm(d["finals"][lane, 0:3] - np.array(ref["final"][1:4])) <= 1e-11 * np.linalg.norm(ref["final"][1:4])
        else:
            # free-running adaptive: the main flight's step sequence decorrelates from the reference's near the separation (a step that just
            # clears or just misses the event clamp), so the dropped stage starts 6e-7 s apart; where it comes down is what is compared
            assert abs(d["finals"][lane, 13] - ref["final"][0]) < 0.5
            assert np.linalg.norm(d["results"][lane, 0:2] - np.array(ref["results"][0:2])) <= 1e-3 * np.linalg.norm(ref["results"][0:2])
    assert np.allclose(out.results[0, 3:7], g["runs"][0]["results"][3:7], rtol=1e-12 if fixed else 3e-3)
    assert "Dropped stage 1: 2 of 2 runs separated" in open(out.logPath).read()


def test_runner_writes_time_step_logs(monkeypatch, tmp_path, golden):
    """SimControl.loggingLevel 1 in a Monte Carlo run: every run gets `<definition>_RunN/<top stage>_timeStepLog.csv` like the reference's
    runs do (SingleSimulations.py:351-392); an un-sampled definition reproduces the reference's file byte for byte."""
    g = golden("timesteplog_rk4.json")
    monkeypatch.setattr(backend, "TrajectoryBatch", OracleBatch)
    sd = SimDefinition(os.path.join(DATA, g["definition"]), silent=True, disableDistributionSampling=True)
    for k, v in {**g["overrides"], "MonteCarlo.numberRuns": "2", "MonteCarlo.output": "apogees", "SimControl.loggingLevel": "1"}.items():
        sd.setValue(k, v)
    sd.fileName = str(tmp_path / "MonteCarlo.mapleaf")
    out = montecarlo.runMonteCarloSimulation(simDefinition=sd, silent=True)
    assert [os.path.relpath(p, tmp_path) for p in out.timeStepLogPaths] == ["MonteCarlo_Run1/Sustainer_timeStepLog.csv", "MonteCarlo_Run2/Sustainer_timeStepLog.csv"]
    for p in out.timeStepLogPaths:
        assert open(p).read() == g["csv"]


def test_wind_rose_sampler_matches_the_reference(golden, monkeypatch):
    """WindRoseDataSampler.sampleWindRoses (ENV/MeanWindModelling.py:198-329): averaged rose, speed histogram with the extreme-wind tail,
    heading histogram interpolated at the drawn speed, two draws of the global generator per wind. Same winds as the reference drew from
    the same seed, on this repository's synthetic tables (and on the reference's own tables where the built reference is present)."""
    from mapleaf_b200.windsampling import WindRoseSampler
    g = golden("windrose.json")
    refWind = os.path.join(REPO, "baseline", "_ref", "MAPLEAF", "Examples", "Wind")
    for label, d in g.items():
        if label == "reference_tables":
            if not os.path.isdir(refWind):
                continue
            monkeypatch.setenv("MAPLEAF_WIND_DATA", refWind)
        random.seed(d["seed"])
        s = WindRoseSampler(d["locations"], d["weights"], d["month"])
        mine = np.array([s.sample() for _ in range(len(d["winds"]))])
        assert np.allclose(mine, np.array(d["winds"]), rtol=1e-12, atol=1e-12), label


@pytest.mark.parametrize("name", ["windrose_mc.json", "windrose_hellman_mc.json"])
def test_runner_samples_ground_winds_from_wind_roses(monkeypatch, tmp_path, golden, name):
    """MeanWindModel SampledGroundWindData and Hellman.groundWindModel SampledGroundWindData in a Monte Carlo loop: every run's
    Environment draws its ground wind from the wind roses with the GLOBAL generator, before the run's pink-noise seeds
    (ENV/environment.py:47-48): the runner draws the same winds in the same order and flies them per lane."""
    g = golden(name)
    monkeypatch.setattr(backend, "TrajectoryBatch", OracleBatch)
    sd = SimDefinition(os.path.join(DATA, g["definition"]), silent=True, disableDistributionSampling=True)
    sd.setValue("MonteCarlo.randomSeed", g["randomSeed"])
    sd = SimDefinition(dictionary=dict(sd.dict), silent=True)
    sd.fileName = str(tmp_path / g["definition"])
    for k, v in {**g["overrides"], "MonteCarlo.numberRuns": str(len(g["runs"])), "MonteCarlo.output": "apogees"}.items():
        sd.setValue(k, v)
    random.seed(g["globalSeed"])
    captured = {}
    realSample = montecarlo.sampleRuns

    def spy(*a, **k):
        captured["arrays"] = realSample(*a, **k)
        return captured["arrays"]
    monkeypatch.setattr(montecarlo, "sampleRuns", spy)
    out = montecarlo.runMonteCarloSimulation(simDefinition=sd, silent=True)
    adaptive = "hellman" in name
    for i, ref in enumerate(g["runs"]):
        assert np.allclose(captured["arrays"][L.LANE_WIND][:, i], ref["groundWind"], rtol=1e-12, atol=1e-12)
        if adaptive:
            # (the fixture's "pinkSeeds" field assumes the seeds are the run's first global draws, which the wind draws are here; the flight
            # itself shows that the seeds drawn after the wind are the reference's: free-running adaptive, apogee within the 1-2 m band of such flights; other seeds move it by tens of metres)
            assert abs(out.results[i, L.RES_APOGEE] - ref["results"][3]) < 2.0 and abs(out.results[i, L.RES_MAX_SPEED] - ref["results"][4]) < 0.2
        else:
            assert np.allclose(out.results[i, 3:7], ref["results"][3:7], rtol=1e-10)


def test_runner_parallel_branch_seeding(monkeypatch, tmp_path):
    """nCores > 1: the seeding of the reference's ray branch (MonteCarlo.py:110-136): master = Random(MonteCarlo.randomSeed); run i draws
    its `_stdDev` keys from Random(master.randrange(10^7)). The expected winds are drawn here with CPython's random, as that branch does."""
    monkeypatch.setattr(backend, "TrajectoryBatch", OracleBatch)
    sd = montecarloDefinition(tmp_path, nRuns=3, **{"MonteCarlo.output": "apogees", "SimControl.EndCondition": "Time", "SimControl.EndConditionValue": "0.5"})
    captured = {}
    realSample = montecarlo.sampleRuns

    def spy(*a, **k):
        captured["arrays"] = realSample(*a, **k)
        return captured["arrays"]
    monkeypatch.setattr(montecarlo, "sampleRuns", spy)
    montecarlo.runMonteCarloSimulation(simDefinition=sd, silent=True, nCores=4)
    master = random.Random("2394781")
    for i in range(3):
        r = random.Random(master.randrange(10000000))
        expected = [r.gauss(2, 5), r.gauss(0, 5), r.gauss(0, 5)]
        assert list(captured["arrays"][L.LANE_WIND][:, i]) == expected


def test_radio_sonde_sampler_and_runner_match_the_reference(monkeypatch, tmp_path, golden):
    """MeanWindModel SampledRadioSondeData (ENV/MeanWindModelling.py:96-118,331-485): the sampler re-seeds the global generator, picks a
    location and one of its soundings of the month and turns it into an altitude / wind-vector profile; every run flies its own profile
    as a per-lane model block whose wind table is padded to the longest sounding of the archives. Same profiles as the reference picked
    for the same seeds, and the reference's Monte Carlo flights on the sampled profile."""
    from mapleaf_b200.windsampling import RadioSondeSampler
    g = golden("radiosonde_mc.json")
    sm = g["sampler"]
    for pick in g["picks"]:
        s = RadioSondeSampler(sm["locations"], sm["weights"], sm["aslAltitudes"], pick["month"], pick["seed"])
        alt, vec = s.sample()
        assert list(alt) == pick["altitudes"] and np.allclose(vec, np.array(pick["winds"]), rtol=1e-15, atol=0)
        assert s.maxRows >= len(alt)
    monkeypatch.setattr(backend, "TrajectoryBatch", OracleBatch)
    sd = montecarloDefinition(tmp_path, nRuns=2, **{"MonteCarlo.output": "apogees"})
    for k, v in g["overrides"].items():
        sd.setValue(k, v)
    out = montecarlo.runMonteCarloSimulation(simDefinition=sd, silent=True)
    for i, ref in enumerate(g["runs"]):
        assert np.allclose(out.results[i, 3:7], ref["results"][3:7], rtol=1e-12), (out.results[i], ref["results"])


@pytest.mark.parametrize("label", ["rk4", "rk45", "rk4_apogee"])
def test_convergence_runner_matches_the_reference(monkeypatch, golden, label, capsys):
    """ConvergenceSimRunner.convergeSimEndPosition (SimulationRunners/Convergence.py:15-149, IO/gridConvergenceFunctions.py): the refinement
    levels fly as lanes of one batch (per-lane first step size for fixed-step methods, per-lane model blocks for the adaptive target
    error); the final positions, the convergence checks and the printed lines are the reference's. `rk4_apogee`: a definition that does
    not end on Time is flown once to find its end time first (:300-317)."""
    from mapleaf_b200 import ConvergenceSimRunner
    from mapleaf_b200.convergence import checkConvergence
    g = golden("convergence.json")[label]
    monkeypatch.setattr(backend, "TrajectoryBatch", OracleBatch)
    sd = SimDefinition(os.path.join(DATA, g["definition"]), silent=True, disableDistributionSampling=True)
    for k, v in g["overrides"].items():
        sd.setValue(k, v)
    runner = ConvergenceSimRunner(simDefinition=sd, silent=False)
    steps, positions, times = runner.convergeSimEndPosition(refinementRatio=2, simLimit=g["simLimit"], plot=False, showPlot=False)
    assert steps == g["steps"] and len(times) == g["simLimit"]
    tol = 1e-12 if label != "rk45" else 1e-6
    for mine, ref in zip(positions, g["positions"]):
        assert np.allclose(list(mine), ref, rtol=tol, atol=tol)
    assert [sd.getValue("SimControl.EndCondition"), float(sd.getValue("SimControl.EndConditionValue"))] == [g["endCondition"][0], float(g["endCondition"][1])]
    for i, ref in enumerate(g["checks"]):
        mine = checkConvergence(g["positions"][i], g["positions"][i + 1], g["positions"][i + 2], 2)
        assert np.allclose(np.array(mine), np.array(ref), rtol=1e-12, atol=0)
    if label != "rk45":
        # the lines the reference prints (Convergence.py:100-131), rebuilt from its recorded numbers with its format strings (the fixture
        # itself caught only the first one: the reference's run() re-routes sys.stdout)
        printed = [l for l in capsys.readouterr().out.splitlines() if l.startswith(("Simulation ", "Final Position", "X-Direction", "Y-Direction", "Z-Direction"))]
        expected = []
        for i, (dt, pos) in enumerate(zip(g["steps"], g["positions"])):
            expected.append("Simulation {}, Time step: {}".format(i + 1, dt))
            expected.append("Final Position: {:1.3f} {:1.3f} {:1.3f}".format(*pos))
            if i >= 2:
                orders, _, _, asym, rich, unc = g["checks"][i - 2]
                for d, name in enumerate("XYZ"):
                    expected.append("{}-Direction: Order: {:>4.3f}, Asymptotic Check: {:>6.3f}, RichardsonExtrap: {:>7.3f}, Estimated Uncertainty: {:>6.3f}".format(
                        name, orders[d], asym[d], rich[d], unc[d]))
        assert printed == expected and (not g["lines"] or g["lines"][0] in printed)
