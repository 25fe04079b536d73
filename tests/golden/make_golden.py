#!/usr/bin/env python
"""
Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(henrystoldt/MAPLEAF, built under baseline/_ref by oracle/build_ref.sh) in this container.

    python tests/golden/make_golden.py            # all fixtures (several minutes: the reference runs ~500 steps/s)

The fixtures pin the CPU oracle (tests/test_oracle_golden.py) and, through it and directly, the
CUDA path (tests/test_gpu_parity.py). /root/reference is never read at test time.

What is recorded, and which reference code produced it:
  model_montecarlo.json ... attributes of the reference's constructed Rocket / Environment
                            (Rocket/rocket.py:34-219 and component constructors)
  rk4_apogee.json ......... `runMonteCarloSimulation`-style loop (SimulationRunners/MonteCarlo.py:64-76),
                            seed "2394781", 8 runs, RK4 dt 0.01, EndCondition Apogee
  rk4_ground.json ......... same, 3 runs, EndCondition Altitude 0 (BASELINE.md first parity anchor)
  ideal_rk4/rk45.json ..... IdealMomentController (GNC/MomentControllers.py:85-93): orientation forced to a tilted target every step, zero angular
                            velocity; fixed-step and adaptive (the adaptive integrator keeps its first-same-as-last derivative across the rewrite)
  rk78_gust.json .......... RK78Adaptive (Dormand-Prince 8(7)13M, no first-same-as-last) with the sine gust, to apogee, 2 runs
  rk45_pink.json .......... RK45Adaptive + PID, Hellman wind, PinkNoise2D (derived config 2), 4 runs
  windtunnel.json ......... WindTunnelSimulation.runSweep force evaluations (SingleSimulations.py:444-475)
                            + the reference's recorded regression values (BatchSims/regressionTests.mapleaf:1-27)
  tabulated_rk4/rk45.json . TabulatedComponents.mapleaf: TabulatedInertia, TabulatedAeroForce (TotalAOA / AOA / AOSS / Mach keys),
                            FractionalJetDamping, AeroDamping, AeroForce, FixedForce (Rocket/RocketComponents.py:183-413)
  tabulated2d_rk4.json .... TabulatedComponents2D.mapleaf: TabulatedAeroForce tables with 2 and 3 key columns (scipy LinearNDInterpolator with
                            nearest-neighbour fallback behind NoNaNLinearNDInterpolator, Motion/Interpolation.py:109-124)
  twostage_*.json ......... TwoStage.mapleaf: motor-burnout separation with delay, upper-stage ignition, automatic boat tail and
                            fineness update (Rocket/rocket.py:339-361,461-528), RK4 to apogee and RK45Adaptive to the ground
  earth_*.json ............ TabulatedComponents.mapleaf flown over the Round and the WGS84 (J2) earth from 47.5 N, 112.25 W
                            (ENV/EarthModelling.py:107-369, ENV/environment.py:105-181)
  nasa_twostage.json ...... NASATwoStageOrbitalRocket.mapleaf, 200 s (BASELINE.json configs[3]; the reference's own regression case,
                            BatchSims/NASAVerificationCases.mapleaf:264-278)
  canards_*.json .......... Canards.mapleaf (BASELINE.json configs[2]): 100 Hz stabiliser + scheduled-gain PID + deflection table +
                            first-order actuators (GNC/*.py), RK4 forced in the ascent, adaptive again under the parachute
  motoradj_*.json ......... motor impulseAdjustFactor / burnTimeAdjustFactor (Rocket/Propulsion.py:38-40,102-128,153-157): set in the
                            definition for a single-stage and a two-stage flight, and sampled (`_stdDev`) in a 3-run Monte Carlo loop
                            like the reference's Jake1_MonteCarlo.mapleaf example
  regression_blobs.npz .... (tests/golden/make_regression_blobs.py) the reference's own flight regression cases as flattened model blocks +
                            the final values the reference records for them
  opt_*.json .............. one short flight per option the backend accepts and no other fixture exercises: TabulatedAtmosphere
                            (ENV/AtmosphereModelling.py:79-96), CustomWindProfile / InterpolatedWind (ENV/MeanWindModelling.py:165-196),
                            PinkNoise1D / 3D with fixed seeds (ENV/TurbulenceModelling.py:100-131), Euler / RK2Midpoint / RK2Heun / RK4_3/8
                            (Motion/Integration.py:144-187,189-236), RK12Adaptive / RK23Adaptive (:238-330), the `elementary` step-size controller (:359-382), ConstantGainPIDRocket (GNC/MomentControllers.py:58-83)
  perlane_*.json .......... Monte Carlo loops with `_stdDev` on keys that change the model itself (a mass and a fin span; the launch position and
                            tilt with a launch rail): what mlb_set_lane_models / per-lane model blocks reproduce
  droppaths_*.json ........ TwoStage.mapleaf with SimControl.StageDropPaths.compute: the booster's own flight from the separation to the ground
                            (SingleSimulations.py:91-145,280-299), fixed step and adaptive
  timesteplog_*.json ...... the reference's own `<stage>_timeStepLog.csv` (SimControl.loggingLevel 1) of two short flights, as text
  windrose.json ........... ground winds drawn by the reference's WindRoseDataSampler (seeded global random)
  radiosonde_mc.json ...... wind profiles picked by the reference's RadioSondeDataSampler and a Monte Carlo loop flown on one
  convergence.json ........ ConvergenceSimRunner.convergeSimEndPosition: final positions at successively halved time steps / target errors and
                            the convergence checks computed from them
  optimization.json ....... ScipyMinimizeRunner (SimulationRunners/Optimization.py:507-614, Nelder-Mead, two independent and one dependent variable):
                            every (trial point, cost) the reference evaluated, its optimum and the continuation definition it wrote
  rng.json ................ random.Random / PinkNoiseGenerator known answers (test_SimDefinition.py:118-151,
                            test_TurbulenceModelling.py:37-60)
"""
import json
import os
import random
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle.ref_loader import REF_DIR, load_reference  # noqa: E402

load_reference()
os.chdir(REF_DIR)

import numpy as np  # noqa: E402
from MAPLEAF.ENV import PinkNoiseGenerator  # noqa: E402
from MAPLEAF.IO import SimDefinition  # noqa: E402
from MAPLEAF.Motion import Vector  # noqa: E402
from MAPLEAF.SimulationRunners import Simulation  # noqa: E402

MY_DATA = os.path.join(REPO, "mapleaf_b200", "data", "Simulations")
QUIET = {"SimControl.loggingLevel": "0", "SimControl.RocketPlot": "Off", "SimControl.plot": "None"}


def stateRow(t, s):
    try:
        return [t, *s.position, *s.velocity, *s.orientation, *s.angularVelocity]
    except AttributeError:      # 3-DoF state after chute deployment
        return [t, *s.position, *s.velocity, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]


def runOne(sd):
    sim = Simulation(simDefinition=sd, silent=True)
    flights, _ = sim.run()
    sys.stdout = sys.__stdout__
    runOne.lastSimulation = sim
    return flights[0]


def flightRecord(f, traceEvery=0, keepDts=False):
    try:
        land = f.getLandingLocation()
        land = [land.X, land.Y, land.Z]
    except ZeroDivisionError:      # Time end condition: the last step is ~1e-14 s long, dz == 0 (rocketFlight.py:65 divides by it)
        land = [None, None, None]
    rec = dict(steps=len(f.times) - 1, final=stateRow(f.times[-1], f.rigidBodyStates[-1]),
               results=[*land, f.getApogee(), f.getMaxSpeed(), f.getFlightTime(), f.getMaxHorizontalVel()])
    if traceEvery:
        idx = list(range(0, len(f.times), traceEvery))
        rec["traceIdx"] = idx
        rec["trace"] = [stateRow(f.times[i], f.rigidBodyStates[i]) for i in idx]
    if keepDts:
        rec["dts"] = [f.times[i + 1] - f.times[i] for i in range(len(f.times) - 1)]
        rec["times"] = list(f.times)
    return rec


def monteCarlo(defFile, overrides, seed, nRuns, globalSeed=None, traceEvery=0, keepDts=False, motorKeys=(), sampledKeys=()):
    """Same sampling order as Main.main -> runMonteCarloSimulation single-threaded (MonteCarlo.py:64-76)."""
    if globalSeed is not None:
        random.seed(globalSeed)          # the pink-noise seeds come from the *global* random module
    sd = SimDefinition(defFile, silent=True, disableDistributionSampling=True)      # parse only
    sd.setValue("MonteCarlo.randomSeed", seed)
    sd = SimDefinition(dictionary=dict(sd.dict), silent=True)      # seeds rng with the *string* seed, consumes draw set 1
    sd.fileName = defFile
    for k, v in {**QUIET, **overrides}.items():
        sd.setValue(k, v)
    runs = []
    for i in range(nRuns):
        sd.resampleProbabilisticValues()
        wind = [*Vector(sd.getValue("Environment.ConstantMeanWind.velocity"))]
        state = random.getstate()
        f = runOne(sd)
        # recover the pink-noise seeds this run drew from the global stream
        after = random.getstate()
        random.setstate(state)
        seeds = [random.randrange(1000000), random.randrange(1000000)]
        random.setstate(after)
        rec = flightRecord(f, traceEvery if i == 0 else 0, keepDts and i == 0)
        rec["wind"] = wind
        rec["motor"] = [float(sd.getValue(k)) for k in motorKeys]
        rec["sampled"] = {k: sd.getValue(k) for k in sampledKeys}
        rec["groundWind"] = [*runOne.lastSimulation.environment.meanWindModel.getMeanWind(10.0)]      # at the 10 m reference height of the Hellman law
        rec["pinkSeeds"] = seeds
        runs.append(rec)
        print("  run", i + 1, "steps", rec["steps"], "apogee", rec["results"][3], flush=True)
    return dict(definition=os.path.basename(defFile), overrides=overrides, randomSeed=seed, globalSeed=globalSeed, runs=runs)


def modelFields():
    sd = SimDefinition(os.path.join(MY_DATA, "MonteCarlo.mapleaf"), silent=True)
    for k, v in QUIET.items():
        sd.setValue(k, v)
    sd.setValue("Environment.ConstantMeanWind.velocity", "(2 0 0)")
    sim = Simulation(simDefinition=sd, silent=True)
    rocket = sim.createRocket()
    sys.stdout = sys.__stdout__
    st = rocket.stages[0]
    s = rocket.rigidBody.state
    inert = rocket.getInertia(0, s)
    out = dict(Aref=rocket.Aref, fineness=rocket.finenessRatio, init=[*s.position, *s.velocity, *s.orientation, *s.angularVelocity],
               fixedMOI=[*st.fixedMassInertiaComponent.MOI], fixedCG=[*st.fixedMassInertiaComponent.CG], fixedMass=st.fixedMassInertiaComponent.mass,
               inertia0=[*inert.MOI, *inert.CG, inert.mass], usstdT=rocket.environment.atmosphericModel.baseTemps[:8],
               usstdP=rocket.environment.atmosphericModel.basePressures[:8], components=[])
    for c in st.components:
        d = dict(cls=type(c).__name__, position=[*c.position])
        n = type(c).__name__
        if n == "NoseCone":
            d.update(baseArea=c.baseArea, length=c.length, wettedArea=c.wettedArea, CP=[*c.CPLocation], coneHalfAngle=c.coneHalfAngle,
                     sub=list(c.SubsonicCdPolyCoeffs), trans=[float(x) for x in np.array(c.TransonicCdPolyCoeffs).reshape(-1)], roughness=c.surfaceRoughness)
        elif n == "BodyTube":
            d.update(length=c.length, outerDiameter=c.outerDiameter, planformArea=c.planformArea, wettedArea=c.wettedArea, CP=[*c.CPLocation])
        elif n == "BoatTail":
            d.update(topArea=c.topArea, bottomArea=c.bottomArea, frontalArea=c.frontalArea, length=c.length, wettedArea=c.wettedArea,
                     CdAdjustmentFactor=c.CdAdjustmentFactor, CP=[*c.CPLocation])
        elif n == "FinSet":
            for a in ("numFins", "span", "rootChord", "tipChord", "thickness", "surfaceRoughness", "sweepAngle", "midChordSweep", "planformArea",
                      "wettedArea", "aspectRatio", "stallAngle", "bodyRadius", "bodyOnFinInterferenceFactor", "finNumberInterferenceFactor",
                      "MACLength", "MACYPos", "XMACLeadingEdge", "finAngle", "subsonicFinThicknessK"):
                d[a] = getattr(c, a)
            d.update(poly=[float(x) for x in c.x], sliceR=list(c.spanSliceRadii), sliceA=list(c.spanSliceAreas), sliceLE=list(c.sliceLEPositions),
                     sliceLen=list(c.sliceLengths), tipPosition=c.finList[0].tipPosition,
                     spandir=[v for f in c.finList for v in (f.spanwiseDirection.X, f.spanwiseDirection.Y)])
        elif n == "TabulatedMotor":
            d.update(times=list(c.times), thrust=list(c.thrustLevels), oxW=list(c.oxWeights), fuelW=list(c.fuelWeights),
                     oxMOI=[[*v] for v in c.oxMOIs], fuelMOI=[[*v] for v in c.fuelMOIs],
                     cgs=[c.initOxCG_Z, c.finalOxCG_Z, c.initFuelCG_Z, c.finalFuelCG_Z], w0=[c.initialOxidizerWeight, c.initialFuelWeight])
        elif n == "RecoverySystem":
            d.update(triggers=c.stageTriggers[1:], values=c.stageTriggerValues[1:], areas=c.chuteAreas[1:], cds=c.chuteCds[1:], delays=c.delayTimes[1:])
        out["components"].append(d)
    return out


def windTunnel():
    from MAPLEAF.SimulationRunners import WindTunnelSimulation  # noqa: F401  (documented harness; evaluated inline below)
    sd = SimDefinition(os.path.join(MY_DATA, "WindTunnel.mapleaf"), silent=True)
    for k, v in QUIET.items():
        sd.setValue(k, v)
    vels = [np.linspace(10, 550, 100)[i] for i in range(100)]
    rows = []
    for vz in vels:
        sd.setValue("Rocket.velocity", "(0 0 {})".format(vz))     # Batch.py:374-409 builds the same linspace strings
        sim = Simulation(simDefinition=sd, silent=True)
        rocket = sim.createRocket()
        sys.stdout = sys.__stdout__
        s = rocket.rigidBody.state
        from MAPLEAF.Motion import AeroParameters
        env = rocket._getEnvironmentalConditions(0.0, s)
        inertia = rocket.getInertia(0.0, s)
        comp = rocket.getAppliedForce(s, 0.0, env, inertia.CG).getAt(inertia.CG)
        total = rocket._getAppliedForce(0.0, s)
        air2 = sum((a - b) ** 2 for a, b in zip(s.velocity, env.Wind))
        q = air2 * env.Density * 0.5
        rows.append(dict(velocityZ=float(vz), state=[*s.position, *s.velocity, *s.orientation, *s.angularVelocity],
                         Mach=AeroParameters.getMachNumber(s, env), aeroF=[*comp.force], aeroM=[*comp.moment],
                         totalF=[*total.force], totalM=[*total.moment], AeroCFZ=comp.force.Z / (q * rocket.Aref)))
    # the reference's own recorded regression values for this sweep
    with open("MAPLEAF/Examples/BatchSims/regressionTests.mapleaf") as f:
        text = f.read()
    case = text[text.index("ExampleWindTunnelCase"):text.index("PlotsToGenerate")]
    exp = [[float(x) for x in m.split(",")] for m in re.findall(r"expectedValues\s+(\S+)", case)]
    return dict(rows=rows, expectedAeroCFZ=exp[0], expectedMach=exp[1])


def rng():
    r = random.Random()
    r.seed(1000)
    g = [r.gauss(10, 1), r.gauss(11, 2), r.gauss(12, 3)]
    r2 = random.Random(1000)
    stream = [r2.gauss(0, 1) for _ in range(40)]
    png = PinkNoiseGenerator(alpha=5 / 3, nPoles=2, seed=63583, simulatedSamplingFreqHz=20)
    times = [0.05, 0.075, 0.09, 0.1, 0.1, 0.03, 0.32, 0.4, 2.0]
    pink = [png.getValue(t) for t in times]
    png2 = PinkNoiseGenerator(alpha=5 / 3, nPoles=2, seed=987654321012, simulatedSamplingFreqHz=20)
    raw = [png2.getValue() for _ in range(5)]
    return dict(seed1000_vector=g, seed1000_stream=stream, pinkSeed=63583, pinkTimes=times, pinkValues=pink,
                pinkBigSeed=987654321012, pinkRaw=raw, strSeedDraws=[random.Random("2394781").gauss(2, 5) for _ in range(1)])


def singleFlight(defFile, overrides, traceEvery=25, keepDts=False):
    """One un-sampled flight of a definition through the reference's Simulation.run()."""
    sd = SimDefinition(defFile, silent=True, disableDistributionSampling=True)
    for k, v in {**QUIET, **overrides}.items():
        sd.setValue(k, v)
    f = runOne(sd)
    rec = flightRecord(f, traceEvery, keepDts)
    print("  steps", rec["steps"], "final", rec["final"][:4], flush=True)
    return dict(definition=os.path.basename(defFile), overrides=overrides, runs=[rec])


def stageDropFlight(defFile, overrides):
    """Main flight + dropped-stage flights of one un-sampled definition (SingleSimulations.py:91-145,280-299)."""
    sd = SimDefinition(defFile, silent=True, disableDistributionSampling=True)
    for k, v in {**QUIET, **overrides}.items():
        sd.setValue(k, v)
    sim = Simulation(simDefinition=sd, silent=True)
    flights, _ = sim.run()
    sys.stdout = sys.__stdout__
    main = flightRecord(flights[0], 50)
    drops = []
    for f in flights[1:]:
        # a dropped stage's flight object starts as a copy of the main flight up to the separation (createNewDetachedStage)
        k = 0
        while k < min(len(f.times), len(flights[0].times)) and f.times[k] == flights[0].times[k] and \
                [*f.rigidBodyStates[k].position] == [*flights[0].rigidBodyStates[k].position]:
            k += 1
        rec = flightRecord(f)
        rec.update(separationIndex=k - 1, separationTime=f.times[k - 1], separationState=stateRow(f.times[k - 1], f.rigidBodyStates[k - 1]),
                   stepsAfterSeparation=len(f.times) - k, firstStatesAfter=[stateRow(f.times[i], f.rigidBodyStates[i]) for i in range(k, min(k + 3, len(f.times)))])
        drops.append(rec)
        print("  dropped stage: separation at t =", rec["separationTime"], "steps after", rec["stepsAfterSeparation"], "final", rec["final"][:4], flush=True)
    return dict(definition=os.path.basename(defFile), overrides=overrides, runs=[main], drops=drops)


def timeStepLogText(defFile, overrides):
    """The reference's own `<stage>_timeStepLog.csv` (SimControl.loggingLevel 1: rocket.py:155-170,615-644,729-738, IO/CythonLog.pyx:163-213) of
    one un-sampled flight, as text."""
    import shutil, tempfile
    tmp = tempfile.mkdtemp()
    sd = SimDefinition(defFile, silent=True, disableDistributionSampling=True)
    for k, v in {**QUIET, **overrides, "SimControl.loggingLevel": "1"}.items():
        sd.setValue(k, v)
    sd.fileName = os.path.join(tmp, os.path.basename(defFile))
    sim = Simulation(simDefinition=sd, silent=True)
    flights, logPaths = sim.run()
    sys.stdout = sys.__stdout__
    path = [p for p in logPaths if p.endswith("_timeStepLog.csv")][0]
    with open(path) as f:
        text = f.read()
    shutil.rmtree(tmp)
    print("  time-step log:", os.path.basename(path), len(text.splitlines()), "lines", flush=True)
    return dict(definition=os.path.basename(defFile), overrides=overrides, fileName=os.path.basename(path), csv=text,
                runs=[flightRecord(flights[0])])


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f)
    print("wrote", name, os.path.getsize(os.path.join(HERE, name)), "bytes", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["model", "rng", "windtunnel", "rk4_apogee", "rk4_ground", "rk45_pink", "rk45_gust", "tabulated", "twostage", "earth", "canards", "motoradj", "rk78", "ideal"]
    mc = os.path.join(MY_DATA, "MonteCarlo.mapleaf")
    if "model" in which:
        dump("model_montecarlo.json", modelFields())
    if "rng" in which:
        dump("rng.json", rng())
    if "windtunnel" in which:
        dump("windtunnel.json", windTunnel())
    if "rk4_apogee" in which:
        dump("rk4_apogee.json", monteCarlo(mc, {}, "2394781", 8, traceEvery=200))
    if "rk4_ground" in which:
        dump("rk4_ground.json", monteCarlo(mc, {"SimControl.EndCondition": "Altitude", "SimControl.EndConditionValue": "0"}, "2394781", 3))
    if "tabulated" in which:
        tab = os.path.join(MY_DATA, "TabulatedComponents.mapleaf")
        dump("tabulated_rk4.json", singleFlight(tab, {}))
        dump("tabulated2d_rk4.json", singleFlight(os.path.join(MY_DATA, "TabulatedComponents2D.mapleaf"), {}))
        dump("tabulated_rk45.json", singleFlight(tab, {"SimControl.timeDiscretization": "RK45Adaptive", "SimControl.EndConditionValue": "6"}, traceEvery=10, keepDts=True))
    if "twostage" in which:
        two = os.path.join(MY_DATA, "TwoStage.mapleaf")
        dump("twostage_rk4.json", singleFlight(two, {}, traceEvery=50))
        dump("twostage_rk45_ground.json", singleFlight(two, {"SimControl.timeDiscretization": "RK45Adaptive", "SimControl.EndCondition": "Altitude"}, traceEvery=40, keepDts=True))
    if "earth" in which:
        tab = os.path.join(MY_DATA, "TabulatedComponents.mapleaf")
        site = {"Environment.LaunchSite.latitude": "47.5", "Environment.LaunchSite.longitude": "-112.25", "Environment.LaunchSite.elevation": "350"}
        dump("earth_round_rk4.json", singleFlight(tab, {"Environment.EarthModel": "Round", **site}))
        dump("earth_wgs84_rk4.json", singleFlight(tab, {"Environment.EarthModel": "WGS84", **site}))
        dump("nasa_twostage.json", singleFlight(os.path.join(MY_DATA, "NASATwoStageOrbitalRocket.mapleaf"), {}, traceEvery=40, keepDts=True))
    if "canards" in which:
        can = os.path.join(MY_DATA, "Canards.mapleaf")
        dump("canards_apogee.json", singleFlight(can, {}, traceEvery=50))
        dump("canards_ground.json", singleFlight(can, {"SimControl.EndCondition": "Altitude"}, traceEvery=100, keepDts=True))
    if "motoradj" in which:
        m = "Rocket.Sustainer.Motor."
        dump("motoradj_rk4.json", singleFlight(mc, {m + "impulseAdjustFactor": "1.0421", m + "burnTimeAdjustFactor": "0.9137"}, traceEvery=200))
        two = os.path.join(MY_DATA, "TwoStage.mapleaf")
        dump("motoradj_twostage.json", singleFlight(two, {"Rocket.Booster.Motor.impulseAdjustFactor": "0.97", "Rocket.Booster.Motor.burnTimeAdjustFactor": "1.12",
                                                          "Rocket.Sustainer.Motor.impulseAdjustFactor": "1.05", "Rocket.Sustainer.Motor.burnTimeAdjustFactor": "0.88"}, traceEvery=50))
        dump("motoradj_mc.json", monteCarlo(mc, {m + "impulseAdjustFactor": "1.0", m + "impulseAdjustFactor_stdDev": "0.02331",
                                                     m + "burnTimeAdjustFactor": "1.0", m + "burnTimeAdjustFactor_stdDev": "0.10679"}, "2394781", 3,
                                            motorKeys=(m + "impulseAdjustFactor", m + "burnTimeAdjustFactor")))
    adaptive = os.path.join(MY_DATA, "MonteCarloAdaptive.mapleaf")
    if "options" in which:
        envDir = os.path.join(REPO, "mapleaf_b200", "data", "Environment")
        wind = {"Environment.ConstantMeanWind.velocity": "(4 -2.5 0.3)"}
        # The reference's TabulatedAtmosphere hands numpy scalars to the Cython vector algebra, which numpy >= 2 turns into ndarrays
        # (TypeError in forceMomentSystem.getAt on the first step). Shim for this one fixture: the same four interpolated values as
        # Python floats (np.float64 -> float is exact), nothing else touched.
        from MAPLEAF.ENV import AtmosphereModelling
        _orig = AtmosphereModelling.TabulatedAtmosphere.getAirProperties
        AtmosphereModelling.TabulatedAtmosphere.getAirProperties = lambda self, h, _=None: [float(x) for x in _orig(self, h)]
        dump("opt_atm_tabulated.json", singleFlight(mc, {**wind, "Environment.AtmosphericPropertiesModel": "TabulatedAtmosphere",
                                                        "Environment.TabulatedAtmosphere.filePath": os.path.join(envDir, "atmosphereTable.txt")}, traceEvery=200))
        dump("opt_wind_profile.json", singleFlight(mc, {"Environment.MeanWindModel": "CustomWindProfile",
                                                       "Environment.CustomWindProfile.filePath": os.path.join(envDir, "windProfile.txt")}, traceEvery=200))
        pink = {**wind, "Environment.MeanWindModel": "Constant", "SimControl.timeDiscretization": "RK4", "SimControl.timeStep": "0.02", "SimControl.EndCondition": "Apogee",
                "Environment.PinkNoiseModel.randomSeed1": "63583", "Environment.PinkNoiseModel.randomSeed2": "4021", "Environment.PinkNoiseModel.randomSeed3": "77"}
        dump("opt_pink1d.json", singleFlight(adaptive, {**pink, "Environment.TurbulenceModel": "PinkNoise1D"}, traceEvery=100))
        dump("opt_pink3d.json", singleFlight(adaptive, {**pink, "Environment.TurbulenceModel": "PinkNoise3D"}, traceEvery=100))
        dump("opt_pink3d_rk45.json", singleFlight(adaptive, {**pink, "Environment.TurbulenceModel": "PinkNoise3D", "SimControl.timeDiscretization": "RK45Adaptive",
                                                            "SimControl.timeStep": "0.01", "SimControl.EndCondition": "Time", "SimControl.EndConditionValue": "8"}, traceEvery=20, keepDts=True))
        short = {**wind, "SimControl.EndCondition": "Time", "SimControl.EndConditionValue": "4", "SimControl.timeStep": "0.01"}
        for name, method in (("euler", "Euler"), ("rk2mid", "RK2Midpoint"), ("rk2heun", "RK2Heun"), ("rk4_38", "RK4_3/8")):
            dump("opt_%s.json" % name, singleFlight(mc, {**short, "SimControl.timeDiscretization": method}, traceEvery=100))
        gust = {"Environment.TurbulenceModel": "customSineGust", "Environment.MeanWindModel": "Constant", "SimControl.EndCondition": "Apogee", **wind}
        dump("opt_rk12a.json", singleFlight(adaptive, {**gust, "SimControl.timeDiscretization": "RK12Adaptive", "SimControl.EndCondition": "Time", "SimControl.EndConditionValue": "3"},
                                            traceEvery=50, keepDts=True))
        dump("opt_rk23a.json", singleFlight(adaptive, {**gust, "SimControl.timeDiscretization": "RK23Adaptive", "SimControl.EndCondition": "Time", "SimControl.EndConditionValue": "12"},
                                            traceEvery=50, keepDts=True))
        dump("opt_ctrl_elementary.json", singleFlight(adaptive, {**gust, "SimControl.TimeStepAdaptation.controller": "elementary"}, traceEvery=20, keepDts=True))
        # controller "Constant" cannot be flown: integratorFactory leaves safetyFactor unbound for it (Integration.py:74-93, UnboundLocalError
        # while the rocket is constructed), so the backend rejects it as well
        can = os.path.join(MY_DATA, "Canards.mapleaf")
        mcK = "Rocket.ControlSystem.MomentController."
        dump("opt_constgain_pid.json", singleFlight(can, {mcK + "Type": "ConstantGainPIDRocket", mcK + "Pxy": "4", mcK + "Ixy": "0.5", mcK + "Dxy": "1.2",
                                                         mcK + "Pz": "0.5", mcK + "Iz": "0.1", mcK + "Dz": "0.2"}, traceEvery=50))
    if "perlane" in which:
        # `_stdDev` on keys that change what the constructors precompute: a mass (inertia, initial CG), a fin span (planform, MAC, slices, CP
        # polynomial, interference factors) ...
        keys = {"Rocket.Sustainer.GeneralMass.mass": "50", "Rocket.Sustainer.GeneralMass.mass_stdDev": "2.5",
                "Rocket.Sustainer.TailFins.span": "0.1397", "Rocket.Sustainer.TailFins.span_stdDev": "0.006"}
        dump("perlane_mass_span.json", monteCarlo(mc, keys, "2394781", 3, sampledKeys=("Rocket.Sustainer.GeneralMass.mass", "Rocket.Sustainer.TailFins.span")))
        # ... and a sampled launch position with a launch rail (rail origin, direction and the end-condition choice follow the sample)
        keys = {"Environment.LaunchSite.railLength": "6", "Rocket.position": "(0 0 15)", "Rocket.position_stdDev": "(3 3 1)", "Rocket.velocity": "(0 0 0)",
                "Rocket.rotationAngle": "4", "Rocket.rotationAngle_stdDev": "1.5", "Rocket.rotationAxis": "(0 1 0)"}
        dump("perlane_rail_position.json", monteCarlo(mc, keys, "2394781", 3, sampledKeys=("Rocket.position", "Rocket.rotationAngle")))
    if "droppaths" in which:
        two = os.path.join(MY_DATA, "TwoStage.mapleaf")
        drop = {"SimControl.StageDropPaths.compute": "true", "SimControl.StageDropPaths.endCondition": "Altitude", "SimControl.StageDropPaths.endConditionValue": "0"}
        dump("droppaths_rk4.json", stageDropFlight(two, drop))
        dump("droppaths_rk45.json", stageDropFlight(two, {**drop, "SimControl.timeDiscretization": "RK45Adaptive", "SimControl.EndCondition": "Altitude"}))
    if "timesteplog" in which:
        # 1.2 s of RK4 flight with the drogue opened at 0.5 s (6-DoF rows, then 3-DoF rows), and 2 s of the adaptive flight (error column)
        early = {"Rocket.Sustainer.RecoverySystem.stage1Trigger": "Time", "Rocket.Sustainer.RecoverySystem.stage1TriggerValue": "0.5",
                 "Rocket.Sustainer.RecoverySystem.stage1DelayTime": "0", "SimControl.EndCondition": "Time", "SimControl.EndConditionValue": "1.2",
                 "SimControl.timeStep": "0.02", "Environment.ConstantMeanWind.velocity": "(4 -2.5 0.3)"}
        dump("timesteplog_rk4.json", timeStepLogText(mc, early))
        dump("timesteplog_rk45.json", timeStepLogText(adaptive, {"Environment.TurbulenceModel": "None", "SimControl.EndCondition": "Time", "SimControl.EndConditionValue": "2",
                                                                 "Environment.ConstantMeanWind.velocity": "(4 -2.5 0.3)"}))
    if "windrose" in which:
        # WindRoseDataSampler (ENV/MeanWindModelling.py:198-329) on this repository's synthetic tables (the reference reads them from its
        # own Examples/Wind folder, so they are copied into the built copy under baseline/_ref for the run) and, separately, on two of the
        # reference's own tables: seeded global random, ten ground winds each
        import shutil
        from MAPLEAF.ENV.MeanWindModelling import WindRoseDataSampler
        mine = os.path.join(REPO, "mapleaf_b200", "data", "Wind")
        for f in os.listdir(mine):
            shutil.copy(os.path.join(mine, f), os.path.join(REF_DIR, "MAPLEAF", "Examples", "Wind", f))
        out = {}
        for label, locs, weights in (("synthetic", ["TestFieldA", "TestFieldB"], [0.35, 0.65]), ("synthetic_single", ["TestFieldB"], [1.0]),
                                     ("reference_tables", ["Suffield", "MedecineHat"], [0.4, 0.6])):
            random.seed(20240611)
            smp = WindRoseDataSampler(silent=True)
            out[label] = dict(locations=locs, weights=weights, month="Apr", seed=20240611,
                              winds=[[*smp.sampleWindRoses(locs, weights, "Apr")] for _ in range(10)])
        dump("windrose.json", out)
        rose = {"Environment.SampledGroundWindData.locationsToSample": "TestFieldA 0.35 TestFieldB 0.65", "Environment.SampledGroundWindData.launchMonth": "Apr"}
        dump("windrose_mc.json", monteCarlo(mc, {**rose, "Environment.MeanWindModel": "SampledGroundWindData"}, "2394781", 3, globalSeed=777))
        dump("windrose_hellman_mc.json", monteCarlo(adaptive, {**rose, "Environment.Hellman.groundWindModel": "SampledGroundWindData", "SimControl.EndCondition": "Apogee"},
                                                    "2394781", 2, globalSeed=778, keepDts=True))
    if "radiosonde" in which:
        # RadioSondeDataSampler (ENV/MeanWindModelling.py:331-485) on this repository's synthetic archives (copied into the built reference's
        # Examples/Wind for the run): the profiles it picks for a few seeds, and a 2-run Monte Carlo loop flown on a sampled profile
        import shutil
        from MAPLEAF.ENV.MeanWindModelling import RadioSondeDataSampler
        mine = os.path.join(REPO, "mapleaf_b200", "data", "Wind")
        for f in os.listdir(mine):
            shutil.copy(os.path.join(mine, f), os.path.join(REF_DIR, "MAPLEAF", "Examples", "Wind", f))
        locs, weights, asl = ["TestFieldA", "TestFieldB"], [0.48, 0.52], [710.0, 638.0]
        picks = []
        for seed, month in (("11", "Apr"), ("12", "Apr"), ("13", "Jul"), ("14", "Yearly"), ("15", "Yearly")):
            alt, vec = RadioSondeDataSampler(silent=True).getRadioSondeWindProfile(locs, weights, asl, month, seed)
            picks.append(dict(seed=seed, month=month, altitudes=[float(x) for x in alt], winds=[[float(c) for c in v] for v in vec]))
        sonde = {"Environment.MeanWindModel": "SampledRadioSondeData", "Environment.SampledRadioSondeData.locationsToSample": "TestFieldA 0.48 TestFieldB 0.52",
                 "Environment.SampledRadioSondeData.locationASLAltitudes": "710 638", "Environment.SampledRadioSondeData.launchMonth": "Apr",
                 "Environment.SampledRadioSondeData.randomSeed": "12"}
        out = monteCarlo(mc, sonde, "2394781", 2)
        out["picks"] = picks
        out["sampler"] = dict(locations=locs, weights=weights, aslAltitudes=asl)
        dump("radiosonde_mc.json", out)
    if "convergence" in which:
        # ConvergenceSimRunner.convergeSimEndPosition (SimulationRunners/Convergence.py:15-149): the same flight at time steps (or target errors)
        # divided by the refinement ratio each time, Richardson-style convergence check of the final position (IO/gridConvergenceFunctions.py)
        import contextlib, io
        from MAPLEAF.SimulationRunners.Convergence import ConvergenceSimRunner
        from MAPLEAF.IO.gridConvergenceFunctions import checkConvergence
        out = {}
        for label, defFile, ov, limit in (("rk4", mc, {"SimControl.EndCondition": "Time", "SimControl.EndConditionValue": "3", "SimControl.timeStep": "0.04",
                                                       "Environment.ConstantMeanWind.velocity": "(4 -2.5 0.3)"}, 4),
                                          ("rk45", adaptive, {"SimControl.EndCondition": "Time", "SimControl.EndConditionValue": "3", "Environment.TurbulenceModel": "None",
                                                              "SimControl.TimeStepAdaptation.targetError": "0.01", "Environment.ConstantMeanWind.velocity": "(4 -2.5 0.3)"}, 3),
                                          ("rk4_apogee", mc, {"SimControl.timeStep": "0.04", "Environment.ConstantMeanWind.velocity": "(4 -2.5 0.3)"}, 3)):
            sd = SimDefinition(defFile, silent=True, disableDistributionSampling=True)
            for k, v in {**QUIET, **ov}.items():
                sd.setValue(k, v)
            runner = ConvergenceSimRunner(simDefinition=sd, silent=True)
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                steps, positions, _ = runner.convergeSimEndPosition(refinementRatio=2, simLimit=limit, plot=False, showPlot=False)
            sys.stdout = sys.__stdout__
            pos = [[*p] for p in positions]
            checks = [checkConvergence(positions[i], positions[i + 1], positions[i + 2], 2) for i in range(len(positions) - 2)]
            out[label] = dict(definition=os.path.basename(defFile), overrides=ov, simLimit=limit, steps=[float(x) for x in steps], positions=pos,
                              checks=[[[float(x) for x in part] for part in c] for c in checks],
                              endCondition=[sd.getValue("SimControl.EndCondition"), sd.getValue("SimControl.EndConditionValue")],
                              lines=[l for l in buf.getvalue().splitlines() if l.startswith(("Simulation ", "Final Position", "X-Direction", "Y-Direction", "Z-Direction", "Setting EndCondition"))])
            print(" ", label, steps, pos[-1], flush=True)
        dump("convergence.json", out)
    if "optimization" in which:
        # ScipyMinimizeRunner of the reference (SimulationRunners/Optimization.py:507-614) on the single-stage rocket: every cost evaluation it
        # made (trial point, cost), the optimum it returned, and the continuation definition it wrote
        import contextlib, io, tempfile
        from MAPLEAF.SimulationRunners.Optimization import optimizationRunnerFactory
        opt = {"Optimization.method": "scipy.optimize.minimize Nelder-Mead",
               "Optimization.costFunction": "1/apogee if (apogee > 1) else max(1, -1*apogee)",
               "Optimization.IndependentVariables.finSpan": "0.08 < Rocket.Sustainer.TailFins.span < 0.25",
               "Optimization.IndependentVariables.noseAR": "2 < Rocket.Sustainer.Nosecone.aspectRatio < 8",
               "Optimization.DependentVariables.Rocket.Sustainer.TailFins.tipChord": "!0.1 + 0.4*finSpan!",
               "Optimization.ScipyMinimize.tolerance": "1e-6", "Optimization.ScipyMinimize.maxIterations": "8",
               "Optimization.showConvergencePlot": "false", "Environment.ConstantMeanWind.velocity": "(4 -2.5 0.3)"}
        sd = SimDefinition(mc, silent=True, disableDistributionSampling=True)
        for k, v in {**QUIET, **opt}.items():
            sd.setValue(k, v)
        tmp = tempfile.mkdtemp()
        # the continuation file is written next to the definition file
        sd.fileName = os.path.join(tmp, "Optimization.mapleaf")
        runner = optimizationRunnerFactory(simDefinition=sd, silent=True)
        evaluations = []
        inner = runner._evaluateCostFunction

        def record(x):
            c = inner(x)
            evaluations.append(dict(x=[float(v) for v in x], cost=float(c)))
            return c
        runner._evaluateCostFunction = record
        with contextlib.redirect_stdout(io.StringIO()):
            cost, pos = runner.runOptimization()
        sys.stdout = sys.__stdout__
        cont = SimDefinition(os.path.join(tmp, "Optimization_continue.mapleaf"), silent=True, disableDistributionSampling=True)
        dump("optimization.json", dict(definition="MonteCarlo.mapleaf", overrides=opt, 
                                       evaluations=evaluations, bestCost=float(cost), bestPosition=[float(v) for v in pos],
                                       continuation={k: cont.getValue(k) for k in cont.dict if k.startswith("Optimization.IndependentVariables.")}))
        print("  optimisation:", len(evaluations), "evaluations, best", float(cost), [float(v) for v in pos], flush=True)
    if "rk45_pink" in which:
        dump("rk45_pink.json", monteCarlo(adaptive, {}, "2394781", 4, globalSeed=2394781, keepDts=True))
    if "ideal" in which:
        ideal = {"Rocket.ControlSystem.MomentController.Type": "IdealMomentController", "Rocket.ControlSystem.desiredFlightDirection": "(0.2 0.1 1)",
                 "Environment.ConstantMeanWind.velocity": "(6 -3 0)"}
        dump("ideal_rk4.json", singleFlight(mc, ideal, traceEvery=200))
        dump("ideal_rk45.json", singleFlight(mc, {**ideal, "SimControl.timeDiscretization": "RK45Adaptive"}, traceEvery=20, keepDts=True))
    if "rk78" in which:
        dump("rk78_gust.json", monteCarlo(adaptive, {"Environment.TurbulenceModel": "customSineGust", "Environment.MeanWindModel": "Constant",
                                                     "SimControl.EndCondition": "Apogee", "SimControl.timeDiscretization": "RK78Adaptive"}, "2394781", 2, keepDts=True))
    if "rk45_gust" in which:
        dump("rk45_gust.json", monteCarlo(adaptive, {"Environment.TurbulenceModel": "customSineGust", "Environment.MeanWindModel": "Constant",
                                                     "SimControl.EndCondition": "Apogee"}, "2394781", 2, keepDts=True))
