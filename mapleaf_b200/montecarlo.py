"""
Monte Carlo runner with the reference's entry point, running every sampled rocket as one lane of the
B200 trajectory kernels.

Mirrors MAPLEAF/SimulationRunners/MonteCarlo.py:
    runMonteCarloSimulation(simDefinitionFilePath, simDefinition, silent, nCores)      :10-20
    _prepSim ....................................................................... :22-46
    _runSimulations_B200  (replaces _runSimulations_SingleThreaded :48-76 / _Parallel :78-146, same fill contract)
    _showResults ................................................................... :148-174
and the text half of IO/Plotting.plotAndSummarize{Vector,Scalar}Result (:671-752) and IO/Logging.MonteCarloLogger (:61-89).

Sampling is the reference's: one `random.Random` owned by the SimDefinition, re-drawn once per run in dictionary order
(IO/simDefinition.py:471-532), pink-noise seeds from the global `random` module, two (1D: one, 3D: three) per run
(ENV/TurbulenceModelling.py:191-195). The Monte Carlo log has the same lines in the same order.
What differs: all runs are sampled first and then flown together. Wind, initial state and motor adjust factors are per-lane input
arrays of the kernels; any other `_stdDev` key gives every run its own model block (mlb_set_lane_models). There is no CPU fallback.
"""
import os
import random
from statistics import mean, stdev
from typing import List, Optional

import numpy as np

from . import hostmath as hm
from .model import L, buildModel, initialStateGlobalFrame
from .simdef import SimDefinition, SubDictReader, formatVector, parseVector

__all__ = ["runMonteCarloSimulation", "MonteCarloLogger", "Vector", "MonteCarloResults", "laneInitialState", "flyDroppedStages"]

# sampled keys the kernels accept per lane (everything else would change the model block itself)
_WIND_KEY = "Environment.ConstantMeanWind.velocity"
_INIT_STATE_KEYS = {"Rocket.position", "Rocket.velocity", "Rocket.initialDirection", "Rocket.angularVelocity", "Rocket.rotationAngle"}
_EXACT_STATISTICS_LIMIT = 100000      # statistics.mean/stdev (exact rational arithmetic, as the reference) up to here


class Vector:
    """Minimal stand-in for MAPLEAF.Motion.Vector in result lists: X/Y/Z, iteration, str and format like the reference
    (Motion/CythonVector.pyx:65-72)."""
    __slots__ = ("X", "Y", "Z")

    def __init__(self, X, Y, Z):
        self.X, self.Y, self.Z = float(X), float(Y), float(Z)

    def __iter__(self):
        return iter((self.X, self.Y, self.Z))

    def __getitem__(self, i):
        return (self.X, self.Y, self.Z)[i]

    def __format__(self, fmt_spec=""):
        return "{} {} {}".format(*(format(c, fmt_spec) for c in (self.X, self.Y, self.Z)))

    def __str__(self):
        return "({} {} {})".format(self.X, self.Y, self.Z)

    def __eq__(self, other):
        return tuple(self) == tuple(other)


class MonteCarloLogger:
    """IO/Logging.py:61-89: lines go to the console and to the log that is written next to the definition file."""

    def __init__(self, monteCarloLog=None, echo=True):
        self.monteCarloLog = [] if monteCarloLog is None else monteCarloLog
        self.echo = echo
        from datetime import datetime
        from platform import platform
        self.monteCarloLog += ["# mapleaf_b200 (B200 batched trajectory kernels behind MAPLEAF's Monte Carlo interface)",
                               "# Date: {}".format(datetime.now()), "# Platform: {}".format(platform()), ""]

    def log(self, string):
        self.monteCarloLog.append(string)
        if self.echo:
            print(string)

    def writeToFile(self, fileBaseName="monteCarloLog", filePath=None):
        if filePath is None:
            n = 0
            while filePath is None or os.path.exists(filePath):
                n += 1
                filePath = fileBaseName + str(n) + ".txt"
        with open(filePath, "w+") as f:
            f.write("".join(line if (len(line) > 0 and line[-1] == "\n") else line + "\n" for line in self.monteCarloLog))
        return filePath


class MonteCarloResults:
    """What a run returns in addition to the reference's side effects (the reference returns None)."""

    def __init__(self, results, status, counters, logPath, outputLists):
        self.results, self.status, self.counters, self.logPath = results, status, counters, logPath
        self.landingLocations, self.apogees, self.maxSpeeds, self.flightTimes, self.maxHorizontalVels, self.flights = outputLists
        self.droppedStages = []      # SimControl.StageDropPaths.compute: one dict per dropped stage (flyDroppedStages)
        self.timeStepLogPaths = []   # SimControl.loggingLevel >= 1: the `<stage>_timeStepLog.csv` files written


_TIME_STEP_LOG_LIMIT = 100     # runs whose per-step logs are written with SimControl.loggingLevel >= 1 (the reference writes every run's)
_TIME_STEP_LOG_ROWS = 65536    # per-step rows kept per logged run (8 MB of device memory each); longer flights are decimated, with a note
_FLIGHT_PATH_LIMIT = 2000      # runs whose paths are kept (256 KB of device memory each while the batch runs)
_FLIGHT_PATH_ROWS = 2048       # strided states kept per flight on the device: 1024-2048 of them, whatever the flight's length
_FLIGHT_PATH_FRAMES = 900      # MonteCarlo.py:60-62: Plotting._keepNTimeSteps(flight, 900)


class FlightPath:
    """What the reference keeps of a flight for `MonteCarlo.output flightPaths` (RocketFlight after Plotting._keepNTimeSteps(flight, 900),
    IO/Plotting.py:320-375): 900 frames evenly spaced in time. Positions, velocities and angular velocities are interpolated linearly
    like the reference's; the orientation is the normalised linear blend of the neighbouring quaternions (the reference uses slerp) and the
    source is the device's strided flight log (every 1st, 2nd, 4th ... accepted step) rather than every step."""

    def __init__(self, times, positions, velocities, orientations, angularVelocities):
        self.times, self.positions, self.velocities = times, positions, velocities
        self.orientations, self.angularVelocities = orientations, angularVelocities

    @classmethod
    def fromLog(cls, rows, nFrames=_FLIGHT_PATH_FRAMES):
        t = rows[:, 0]
        newTimes = np.linspace(0, t[-1], num=nFrames)
        cols = np.stack([np.interp(newTimes, t, rows[:, k]) for k in range(1, 14)], axis=1)
        q = cols[:, 6:10]
        n = np.linalg.norm(q, axis=1, keepdims=True)
        q = np.where(n > 0, q / np.where(n > 0, n, 1), np.array([1.0, 0, 0, 0]))
        return cls(newTimes, cols[:, 0:3], cols[:, 3:6], q, cols[:, 10:13])


def laneInitialState(simDefinition: SimDefinition, model) -> List[float]:
    """Initial inertial-frame state of the CG for the current (sampled) Rocket.* values: rocket.py:251-291,416-421."""
    rk = SubDictReader("Rocket", simDefinition)
    initPos, initVel, q0, w0 = initialStateGlobalFrame(rk, model.meta["earth"], model.header("H_ELEV"), model.header("H_LAT"),
                                                       model.header("H_LON"), model.header("H_RAIL_LEN") > 0)
    cgGlobal, q0n = hm.q_rotate(q0, model.meta["initialInertia"].CG)
    return list(hm.v_add(initPos, cgGlobal)) + list(initVel) + list(q0n) + list(w0)


def loadSimDefinition(simDefinitionFilePath=None, simDefinition=None, silent=False) -> SimDefinition:
    """SingleSimulations.py:15-26"""
    if simDefinition is None and simDefinitionFilePath is not None:
        return SimDefinition(simDefinitionFilePath, silent=silent)
    if simDefinition is not None:
        return simDefinition
    raise ValueError("No definition file path or SimDefinition object provided to the Monte Carlo runner")


def _prepSim(simDefinition, echo=True):
    simDefinition.setValue("SimControl.plot", "None")
    simDefinition.setValue("SimControl.RocketPlot", "Off")
    nRuns = int(simDefinition.getValue("MonteCarlo.numberRuns"))
    mCLogger = MonteCarloLogger(echo=echo)
    mCLogger.log("")
    mCLogger.log("Running Monte Carlo Simulation: {} runs".format(nRuns))
    simDefinition.monteCarloLogger = mCLogger
    outputLists = [[], [], [], [], [], []]
    return nRuns, mCLogger, outputLists


def _motorKeys(model):
    """{definition key: row of LANE_MOTOR_ADJUST} for the impulse / burn-time adjust factors of every stage's motor."""
    keys = {}
    for si, motor in enumerate(model.meta["motors"] if model is not None else []):
        if motor is not None:
            keys[motor["path"] + ".impulseAdjustFactor"] = 2 * si
            keys[motor["path"] + ".burnTimeAdjustFactor"] = 2 * si + 1
    return keys


_RAIL_STATE_KEYS = {"Rocket.position", "Rocket.rotationAngle", "Rocket.rotationAxis", "Rocket.initialDirection"}
LANE_MODELS = "models"      # key of the per-lane model blocks in the dictionary sampleRuns returns


def _checkSampledKeys(simDefinition, model=None):
    """(sampled keys in draw order, needLaneModels). Wind, initial state and motor adjust factors are per-lane input arrays of the kernels;
    any other `_stdDev` key (the reference re-draws whatever has one, IO/simDefinition.py:471-532) changes what the constructors
    precompute, so every run gets its own model block. So does a sampled launch position / direction when there is a launch rail: the rail's
    origin and direction and the choice of the Altitude end condition follow the sample (ENV/environment.py:82-91, SingleSimulations.py:214-231)."""
    keys = simDefinition.probabilisticKeys()
    motorKeys = _motorKeys(model)
    other = [k for k in keys if k != _WIND_KEY and k not in _INIT_STATE_KEYS and k not in motorKeys]
    onRail = float(simDefinition.getValue("Environment.LaunchSite.railLength")) > 0 and any(k in _RAIL_STATE_KEYS for k in keys)
    return keys, bool(other) or onRail


def _sampleRunsLaneModels(simDefinition, nRuns, mCLogger, model, nGen, hasSeed, windSampler=None):
    """One model block per run: resample, run the constructors' precomputation on the sampled definition (buildModel), keep the shared part.
    ~0.1 s of host time per run for a Barrowman rocket (fin MAC integral, ogive area integral): the general path, not the fast one."""
    shared = int(model.header("H_SHARED_LEN")) or model.blob.size
    models = np.empty((nRuns, shared))
    seeds = np.zeros((3, nRuns)) if nGen else None
    for i in range(nRuns):
        mCLogger.log("\nMonte Carlo Run #{}".format(i + 1))
        simDefinition.resampleProbabilisticValues()
        if windSampler is not None and windSampler.kind == "profile":
            mi = buildModel(simDefinition, windProfile=windSampler.sample(), windProfileRows=windSampler.maxRows)
        else:
            mi = buildModel(simDefinition, groundWind=windSampler.sample() if windSampler is not None else None)
        if mi.blob.size != model.blob.size or int(mi.header("H_SHARED_LEN")) != int(model.header("H_SHARED_LEN")) \
                or not np.array_equal(mi.blob[shared:], model.blob[shared:], equal_nan=True):
            raise NotImplementedError("Monte Carlo sampling changed the structure of the model (component list, table sizes or interpolation tables) "
                                      "in run {}: per-lane model blocks must share one layout".format(i + 1))
        models[i] = mi.blob[:shared]
        for g in range(nGen):
            seeds[g, i] = mi.blob[L.H_PINK_SEED + g] if (hasSeed >> g) & 1 else random.randrange(1000000)
    arrays = {LANE_MODELS: models}
    if nGen:
        arrays[L.LANE_PINK_SEED] = seeds
    return arrays


def sampleRuns(simDefinition, nRuns, mCLogger, model, windSampler=None, parallelSeeding=False):
    """Draws every run's inputs in the reference's order (MonteCarlo.py:64-72) and returns the per-lane arrays. windSampler: the ground
    wind of every run comes from wind roses (two draws of the global generator per run, made when the run's Environment is constructed:
    after the run's resample, before its pink-noise seeds, ENV/environment.py:47-48).
    parallelSeeding: the seeding of the reference's ray branch (nCores > 1, MonteCarlo.py:110-136): a master generator seeded with
    MonteCarlo.randomSeed hands every run its own seed, master.randrange(10^7), and the run's `_stdDev` keys are drawn from
    random.Random(that seed) - a different, equally repeatable stream than the single-threaded one."""
    if parallelSeeding:
        try:
            masterSeed = simDefinition.getValue("MonteCarlo.randomSeed")
        except KeyError:
            masterSeed = random.randrange(10000000)
        master = random.Random(masterSeed)
    keys, needLaneModels = _checkSampledKeys(simDefinition, model)
    motorKeys = _motorKeys(model)
    sampleMotor = any(k in motorKeys for k in keys)
    sampleWind = _WIND_KEY in keys
    sampleState = any(k in _INIT_STATE_KEYS for k in keys)
    turb = int(model.header("H_TURB"))
    nGen = {L.TURB_PINK1D: 1, L.TURB_PINK2D: 2, L.TURB_PINK3D: 3}.get(turb, 0)
    hasSeed = int(model.header("H_PINK_HAS_SEED"))
    if windSampler is not None and windSampler.kind == "profile":
        needLaneModels = True              # a wind table per run
    if needLaneModels:
        if parallelSeeding:
            raise NotImplementedError("nCores > 1 seeding together with Monte Carlo keys that need per-lane model blocks")
        return _sampleRunsLaneModels(simDefinition, nRuns, mCLogger, model, nGen, hasSeed, windSampler)
    if windSampler is not None:
        sampleWind = True
    if windSampler is None and not parallelSeeding and not sampleState and not simDefinition.disableDistributionSampling and nRuns >= _VECTORIZED_SAMPLING_FROM:
        return _sampleRunsVectorized(simDefinition, nRuns, mCLogger, model, keys, motorKeys, nGen, hasSeed)
    wind = np.empty((3, nRuns)) if sampleWind else None
    state = np.empty((13, nRuns)) if sampleState else None
    seeds = np.zeros((3, nRuns)) if nGen else None
    motor = np.ones((2 * L.MLB_MAX_STAGES, nRuns)) if sampleMotor else None
    for i in range(nRuns):
        mCLogger.log("\nMonte Carlo Run #{}".format(i + 1))
        if parallelSeeding:
            simDefinition.rng = random.Random(master.randrange(10000000))      # randrange(1e7) of the reference, an int since Python 3.12
        simDefinition.resampleProbabilisticValues()
        if windSampler is not None:
            wind[:, i] = windSampler.sample()
        elif sampleWind:
            wind[:, i] = parseVector(simDefinition.getValue(_WIND_KEY))
        if sampleState:
            state[:, i] = laneInitialState(simDefinition, model)
        if sampleMotor:
            for key, row in motorKeys.items():
                motor[row, i] = float(simDefinition.getValue(key))
        for g in range(nGen):
            # PinkNoiseGenerator.__init__ draws from the *global* random module when no seed key is given
            seeds[g, i] = model.blob[L.H_PINK_SEED + g] if (hasSeed >> g) & 1 else random.randrange(1000000)
    arrays = {}
    if sampleWind:
        arrays[L.LANE_WIND] = wind
    if sampleState:
        arrays[L.LANE_INIT_STATE] = state
    if nGen:
        arrays[L.LANE_PINK_SEED] = seeds
    if sampleMotor:
        arrays[L.LANE_MOTOR_ADJUST] = motor
    return arrays


_VECTORIZED_SAMPLING_FROM = 64      # runs; below this the per-run Python loop is as fast


def _sampleRunsVectorized(simDefinition, nRuns, mCLogger, model, keys, motorKeys, nGen, hasSeed):
    """sampleRuns for 10^5+ runs: the same draws from the same two generators (the SimDefinition's own `rng` for the `_stdDev` keys, the
    global `random` module for the pink-noise seeds), made by the C++ CPython-exact sampler (mlb_host_gauss_from_state /
    mlb_host_randrange_from_state) instead of nRuns passes of resampleProbabilisticValues; the generators, the dictionary and the Monte
    Carlo log end up as the per-run loop leaves them (IO/simDefinition.py:471-532, MonteCarlo.py:64-72)."""
    from . import backend
    D = simDefinition.dict
    mus, sigmas, spans = [], [], []            # one (mu, sigma) per drawn component, in draw order; spans: (key, first column, width)
    for key in keys:
        meanString = D.setdefault(key + "_mean", D[key])
        try:
            m, sd = [float(meanString)], [float(D[key + "_stdDev"])]
        except ValueError:
            m, sd = list(parseVector(meanString)), list(parseVector(D[key + "_stdDev"]))
        spans.append((key, len(mus), len(m)))
        mus += m
        sigmas += sd
    draws = backend.gaussFromGenerator(simDefinition.rng, mus, sigmas, nRuns) if mus else np.empty((nRuns, 0))
    arrays = {}
    perRunLines = []
    for key, c0, w in spans:
        v = draws[:, c0:c0 + w]
        if w == 1:
            D[key] = str(float(v[-1, 0]))
            perRunLines.append(["Sampling scalar parameter: {}, value: {:1.3f}".format(key, x) for x in v[:, 0].tolist()])
        else:
            D[key] = formatVector(v[-1].tolist())
            perRunLines.append(["Sampling vector parameter: {}, value: ({:1.3f} {:1.3f} {:1.3f})".format(key, *x) for x in v.tolist()])
        if key == _WIND_KEY:
            arrays[L.LANE_WIND] = np.ascontiguousarray(v.T)
        else:
            motor = arrays.setdefault(L.LANE_MOTOR_ADJUST, np.ones((2 * L.MLB_MAX_STAGES, nRuns)))
            motor[motorKeys[key]] = v[:, 0]
    if L.LANE_MOTOR_ADJUST in arrays:              # motors whose factors are not sampled keep the definition's values
        for key, row in motorKeys.items():
            if key not in keys:
                arrays[L.LANE_MOTOR_ADJUST][row] = float(simDefinition.getValue(key))
    if nGen:
        seeds = np.zeros((3, nRuns))
        free = [g for g in range(nGen) if not (hasSeed >> g) & 1]
        if free:
            drawn = backend.randrangeFromGenerator(random, 1000000, len(free) * nRuns).reshape(nRuns, len(free))
            for j, g in enumerate(free):
                seeds[g] = drawn[:, j]
        for g in range(nGen):
            if (hasSeed >> g) & 1:
                seeds[g] = model.blob[L.H_PINK_SEED + g]
        arrays[L.LANE_PINK_SEED] = seeds
    log = mCLogger.log
    # resampleProbabilisticValues reports to the definition's own logger, else to the console unless silent (simDefinition.py:528-532)
    sampleLog = simDefinition.monteCarloLogger.log if simDefinition.monteCarloLogger is not None else (None if simDefinition.silent else print)
    for i in range(nRuns):
        log("\nMonte Carlo Run #{}".format(i + 1))
        if sampleLog is not None:
            for lines in perRunLines:
                sampleLog(lines[i])
    return arrays


def _runSimulations_B200(simDefinition, nRuns, outputLists, mCLogger, silent=False, nGPUs=1, device=0, parallelSeeding=False, specialize="cached"):
    """Same fill contract as _runSimulations_SingleThreaded: outputLists filled in sample order."""
    from . import backend
    landingLocations, apogees, maxSpeeds, flightTimes, maxHorizontalVels, flights = outputLists
    resultsToOutput = simDefinition.getValue("MonteCarlo.output")
    wantPaths = "flightPaths" in resultsToOutput
    if wantPaths and nRuns > _FLIGHT_PATH_LIMIT:
        mCLogger.log("NOTE: MonteCarlo.output flightPaths: flight paths are kept for the first {} of {} runs".format(_FLIGHT_PATH_LIMIT, nRuns))
    from .windsampling import samplerFromDefinition
    windSampler = samplerFromDefinition(SubDictReader("Environment", simDefinition))
    # the template model is built without a draw from the global generator: placeholders stand in for the sampled wind
    if windSampler is not None and windSampler.kind == "profile":
        model = buildModel(simDefinition, windProfile=(np.arange(windSampler.maxRows, dtype=float), np.zeros((windSampler.maxRows, 3))))
    else:
        model = buildModel(simDefinition, groundWind=(0.0, 0.0, 0.0) if windSampler is not None else None)
    # SimControl.loggingLevel >= 1: the reference writes `<definition>_RunN/<top stage>_timeStepLog.csv` for every run
    # (SingleSimulations.py:351-392, rocket.py:729-738); here for the first _TIME_STEP_LOG_LIMIT runs, from the lanes' per-step flight log
    wantStepLogs = int(simDefinition.getValue("SimControl.loggingLevel")) >= 1 and bool(simDefinition.fileName)
    if wantStepLogs and nRuns > _TIME_STEP_LOG_LIMIT:
        mCLogger.log("NOTE: SimControl.loggingLevel: time-step logs are written for the first {} of {} runs".format(_TIME_STEP_LOG_LIMIT, nRuns))
    arrays = sampleRuns(simDefinition, nRuns, mCLogger, model, windSampler, parallelSeeding=parallelSeeding)
    laneModels = arrays.pop(LANE_MODELS, None)

    # contiguous sample ranges per GPU: gathered results are already in sample order
    nGPUs = max(1, min(int(nGPUs), nRuns))
    bounds = [nRuns * g // nGPUs for g in range(nGPUs + 1)]
    batches = []
    for g in range(nGPUs):
        lo, hi = bounds[g], bounds[g + 1]
        # the library compiled for this very model block when there is one (or when asked for); per-lane model blocks need the generic one
        b = backend.TrajectoryBatch(model.blob, device=device + g, specialize=False if laneModels is not None else specialize)
        b.setLanes(hi - lo, {k: np.ascontiguousarray(a[:, lo:hi]) for k, a in arrays.items()})
        if laneModels is not None:
            b.setLaneModels(laneModels[lo:hi])
        nLogged = max(0, min(hi, _FLIGHT_PATH_LIMIT) - lo) if wantPaths else 0
        nStepLogged = max(0, min(hi, _TIME_STEP_LOG_LIMIT) - lo) if wantStepLogs else 0
        if nStepLogged:           # every accepted step (the path frames below are interpolated from the same rows)
            b.setFlightLog(0, max(nLogged, nStepLogged), 1, _TIME_STEP_LOG_ROWS)
        elif nLogged:
            b.setFlightLog(0, nLogged, 1, _FLIGHT_PATH_ROWS)
        b.run()                                   # asynchronous: all GPUs integrate concurrently
        batches.append(b)
    results = np.empty((nRuns, L.RES_SIZE))
    status = np.empty(nRuns, dtype=np.int32)
    counters = []
    stepLogPaths = []
    wantDrops = int(model.header("H_NSTAGES")) > 1 and str(simDefinition.getValue("SimControl.StageDropPaths.compute")).lower() == "true" and laneModels is None
    separations = np.zeros((nRuns, L.MLB_MAX_STAGES - 1, L.SEP_W)) if wantDrops else None
    for g, b in enumerate(batches):
        r, s = b.results()
        results[bounds[g]:bounds[g + 1]], status[bounds[g]:bounds[g + 1]] = r, s
        counters.append(b.counters())
        if wantDrops:
            separations[bounds[g]:bounds[g + 1]] = b.stageSeparations()
        logged = b.flightLog() if ((wantPaths and bounds[g] < _FLIGHT_PATH_LIMIT) or (wantStepLogs and bounds[g] < _TIME_STEP_LOG_LIMIT)) else []
        if wantPaths and bounds[g] < _FLIGHT_PATH_LIMIT:
            flights.extend(FlightPath.fromLog(rows) for rows in logged[:max(0, min(bounds[g + 1], _FLIGHT_PATH_LIMIT) - bounds[g])])
        if wantStepLogs and bounds[g] < _TIME_STEP_LOG_LIMIT:
            from .flightlog import writeTimeStepLog
            base = simDefinition.fileName[:simDefinition.fileName.rfind(".")] + "_Run"
            laneSteps = b.laneCounters() if hasattr(b, "laneCounters") else None
            for i, rows in enumerate(logged[:max(0, min(bounds[g + 1], _TIME_STEP_LOG_LIMIT) - bounds[g])]):
                folder = _nextNumberedName(base)
                os.mkdir(folder)
                if laneSteps is not None and len(rows) - 1 != int(laneSteps[i, 0]):
                    mCLogger.log("NOTE: run {}: flight longer than {} steps, time-step log decimated".format(bounds[g] + i + 1, _TIME_STEP_LOG_ROWS))
                path = writeTimeStepLog(os.path.join(folder, "{}_timeStepLog.csv".format(model.meta["stages"][0]["name"])), rows, model)
                stepLogPaths.append(path)
        b.close()

    failed = np.nonzero(status != L.ST_OK)[0]
    if failed.size:
        names = {L.ST_STEP_LIMIT: "step limit", L.ST_NAN: "NaN state", L.ST_UNSUPPORTED: "unsupported"}
        mCLogger.log("WARNING: {} of {} runs did not end normally: {}".format(
            failed.size, nRuns, ", ".join("#{} ({})".format(i + 1, names.get(int(status[i]), status[i])) for i in failed[:20])))
    for i in range(nRuns):
        landingLocations.append(Vector(results[i, L.RES_LAND_X], results[i, L.RES_LAND_Y], results[i, L.RES_LAND_Z]))
    apogees.extend(results[:, L.RES_APOGEE].tolist())
    maxSpeeds.extend(results[:, L.RES_MAX_SPEED].tolist())
    flightTimes.extend(results[:, L.RES_FLIGHT_TIME].tolist())
    maxHorizontalVels.extend(results[:, L.RES_MAX_HVEL].tolist())
    droppedStages = flyDroppedStages(simDefinition, model, arrays, separations, device=device) if wantDrops else []
    for k, d in enumerate(droppedStages):
        mCLogger.log("Dropped stage {}: {} of {} runs separated; flown from the separation to SimControl.StageDropPaths.endCondition".format(k + 1, d["lanes"].size, nRuns))
    return results, status, counters, droppedStages, stepLogPaths


def _nextNumberedName(base):
    """Logging.findNextAvailableNumberedFileName (IO/Logging.py): base + the first integer >= 1 that is not taken."""
    n = 1
    while os.path.exists(base + str(n)):
        n += 1
    return base + str(n)


def flyDroppedStages(simDefinition, model, arrays, separations, device=0, batchClass=None):
    """The dropped stages' own flights (SimControl.StageDropPaths.compute; SingleSimulations.py:91-145,280-299): for every separation the
    reference creates a new rocket made of the dropped stage alone (burnt-out motor, its own recovery system if it has one), gives it the
    whole rocket's state and time at the separation and flies it to SimControl.StageDropPaths.endCondition. Here every dropped stage of
    every run is one lane of a second batch: model block of the dropped stage, LANE_INIT_STATE / LANE_START from the separation rows the
    main batch recorded (mlb_stage_separations), the run's own sampled wind.
    Returns one dict per dropped stage (first dropped first): lanes (indices of the runs in which it separated), results, status, finals."""
    from . import backend
    batchClass = batchClass or backend.TrajectoryBatch
    if int(model.header("H_TURB")) in (L.TURB_PINK1D, L.TURB_PINK2D, L.TURB_PINK3D):
        raise NotImplementedError("StageDropPaths with a pink-noise turbulence model: the reference flies the dropped stage through the generator "
                                  "state the main flight left behind, backwards in time (ENV/TurbulenceModelling.py:197-247)")
    out = []
    nStages = int(model.header("H_NSTAGES"))
    for k in range(nStages - 1):
        rows = separations[:, k]
        lanes = np.nonzero(rows[:, L.SEP_VALID] == 1)[0]
        if lanes.size == 0:
            out.append(dict(lanes=lanes, results=np.empty((0, L.RES_SIZE)), status=np.empty(0, dtype=np.int32), finals=np.empty((0, L.FIN_SIZE))))
            continue
        z = rows[lanes, L.SEP_STATE + 2]
        endValue = float(simDefinition.getValue("SimControl.StageDropPaths.endConditionValue"))
        if simDefinition.getValue("SimControl.StageDropPaths.endCondition") == "Altitude" and (z < endValue).any() and (z >= endValue).any():
            raise NotImplementedError("StageDropPaths.endCondition Altitude with separations both below and above the end altitude")
        dropModel = buildModel(simDefinition, droppedStage=k, droppedFromAltitude=float(z[0]))
        laneArrays = {L.LANE_INIT_STATE: np.ascontiguousarray(rows[lanes, L.SEP_STATE:L.SEP_STATE + 13].T),
                      L.LANE_START: np.ascontiguousarray(rows[lanes][:, [L.SEP_TIME, L.SEP_DT]].T)}
        if L.LANE_WIND in arrays:
            laneArrays[L.LANE_WIND] = np.ascontiguousarray(arrays[L.LANE_WIND][:, lanes])
        if L.LANE_MOTOR_ADJUST in arrays:
            # the dropped stage is stage 0 of its own model: its motor's factors move to rows 0 / 1 (the k-th dropped stage is the
            # bottom stage of what is left of the rocket: stage nStages - 1 - k of the full model)
            full = arrays[L.LANE_MOTOR_ADJUST][:, lanes]
            adj = np.ones_like(full)
            si = nStages - 1 - k
            adj[0:2] = full[2 * si:2 * si + 2]
            laneArrays[L.LANE_MOTOR_ADJUST] = np.ascontiguousarray(adj)
        b = batchClass(dropModel.blob, device=device)
        try:
            b.setLanes(int(lanes.size), laneArrays)
            b.run()
            res, st = b.results()
            fin = b.finals()
        finally:
            b.close()
        out.append(dict(lanes=lanes, results=res, status=st, finals=fin))
    return out


def _mean(x):
    return mean(x) if len(x) <= _EXACT_STATISTICS_LIMIT else float(np.mean(np.asarray(x, dtype=np.float64)))


def _stdev(x):
    if len(x) < 2:
        return float("nan")
    return stdev(x) if len(x) <= _EXACT_STATISTICS_LIMIT else float(np.std(np.asarray(x, dtype=np.float64), ddof=1))


def summarizeVectorResult(vectorList, name="Landing location", monteCarloLogger=None):
    """Text of Plotting.plotAndSummarizeVectorResult (IO/Plotting.py:671-699)."""
    out = ["{}s:   X       Y        Z".format(name)]
    x, y, z = [], [], []
    for i, loc in enumerate(vectorList):
        out.append("Simulation #{}:   {:>7.1f} {:>7.1f} {:>7.1f}".format(i + 1, loc.X, loc.Y, loc.Z))       # "{:>7.1f}".format(Vector)
        x.append(loc.X), y.append(loc.Y), z.append(loc.Z)
    out.append("")
    out.append("Mean {}: ({:>8.1f} {:>8.1f} {:>8.1f})".format(name, _mean(x), _mean(y), _mean(z)))
    out.append("Standard deviation:    ({:>8.1f} {:>8.1f} {:>8.1f})".format(_stdev(x), _stdev(y), _stdev(z)))
    out.append("")
    for line in out:
        monteCarloLogger.log(line) if monteCarloLogger is not None else print(line)


def summarizeScalarResult(scalarList, name="Apogee", monteCarloLogger=None):
    """Text of Plotting.plotAndSummarizeScalarResult (IO/Plotting.py:723-741)."""
    out = ["{}s:".format(name)]
    for i, v in enumerate(scalarList):
        out.append("Simulation #{}: {}".format(i + 1, v))
    out.append("")
    out.append("Mean {}:        {:>8.1f}".format(name, _mean(scalarList)))
    out.append("Standard Deviation: {:>8.1f}".format(_stdev(scalarList)))
    out.append("")
    for line in out:
        monteCarloLogger.log(line) if monteCarloLogger is not None else print(line)


def _showResults(simDefinition, outputLists, mCLogger):
    landingLocations, apogees, maxSpeeds, flightTimes, maxHorizontalVels, flights = outputLists
    resultsToOutput = simDefinition.getValue("MonteCarlo.output")
    mCLogger.log("")
    mCLogger.log("Monte Carlo results:")
    if "landingLocations" in resultsToOutput:
        summarizeVectorResult(landingLocations, name="Landing location", monteCarloLogger=mCLogger)
    if "apogees" in resultsToOutput:
        summarizeScalarResult(apogees, name="Apogee", monteCarloLogger=mCLogger)
    if "maxSpeeds" in resultsToOutput:
        summarizeScalarResult(maxSpeeds, name="Max speed", monteCarloLogger=mCLogger)
    if "flightTimes" in resultsToOutput:
        summarizeScalarResult(flightTimes, name="Flight time", monteCarloLogger=mCLogger)
    if "maxHorizontalVels" in resultsToOutput:
        summarizeScalarResult(maxHorizontalVels, name="Max horizontal speed", monteCarloLogger=mCLogger)
    logPath = None
    if resultsToOutput != "None" and len(resultsToOutput) > 0 and simDefinition.fileName:
        dotIndex = simDefinition.fileName.rfind(".")
        logPath = mCLogger.writeToFile(fileBaseName=simDefinition.fileName[0:dotIndex] + "_monteCarloLog_run")
        print("Wrote Monte Carlo Log to: {}".format(logPath))
    return logPath


def runMonteCarloSimulation(simDefinitionFilePath=None, simDefinition=None, silent=False, nCores=1, nGPUs: Optional[int] = None, device: int = 0, specialize="cached"):
    """Drop-in for MAPLEAF.SimulationRunners.runMonteCarloSimulation. `nGPUs` spreads contiguous sample ranges over that many GPUs of
    this box. `nCores` > 1 selects the SEEDING of the reference's ray branch (every run draws from its own generator, seeded by a
    master generator: MonteCarlo.py:110-136), so that a definition run with `--parallel` keeps drawing the samples it drew there; the
    work itself is not fanned out over processes. Two things of that branch are not reproducible even in the reference and are not
    imitated: its results arrive in completion order (here: sample order), and its pink-noise seeds come from each worker process's own,
    OS-seeded global generator (here: this process's global generator, as in the single-threaded branch).
    `specialize`: "cached" (default) flies on the library compiled for this definition's model block if one is cached
    (backend.buildSpecializedLibrary, ~1 min of nvcc per distinct rocket), True compiles it on first use, False always uses the generic
    library. Returns MonteCarloResults."""
    simDefinition = loadSimDefinition(simDefinitionFilePath, simDefinition, silent)
    nRuns, mCLogger, outputLists = _prepSim(simDefinition, echo=not silent)
    results, status, counters, droppedStages, stepLogPaths = _runSimulations_B200(simDefinition, nRuns, outputLists, mCLogger, silent, nGPUs=nGPUs or 1, device=device,
                                                                                  parallelSeeding=int(nCores or 1) > 1, specialize=specialize)
    logPath = _showResults(simDefinition, outputLists, mCLogger)
    out = MonteCarloResults(results, status, counters, logPath, outputLists)
    out.droppedStages = droppedStages
    out.timeStepLogPaths = stepLogPaths
    return out
