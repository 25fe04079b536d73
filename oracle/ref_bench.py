"""
TEST / BENCH INFRASTRUCTURE ONLY — CPU baselines for bench.py (the product path never imports this).

Two CPU arms, timed on the host cores of the box bench.py runs on:
  "reference": the UNMODIFIED reference (henrystoldt/MAPLEAF built under baseline/_ref by oracle/build_ref.sh) flying
               the same simulation definition through its own stock code path, one
               `Simulation(simDefinition).run()` per sample (MAPLEAF/SimulationRunners/MonteCarlo.py:64-76),
               fanned out over worker processes the way the reference's ray branch does (MonteCarlo.py:78-146;
               ray itself is not installed and that branch is broken on Python 3.12).
  "port":      the C restatement (oracle/mapleaf_oracle.c), one lane after another per worker process.

Every worker returns (accepted steps, trajectories, seconds of pure simulation time).
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
if REPO not in sys.path:
    sys.path.insert(0, REPO)

_state = {}


def _initReference(defFile, overrides):
    import warnings
    warnings.simplefilter("ignore")          # the reference's regex literals raise SyntaxWarning on Python 3.12
    from oracle.ref_loader import REF_DIR, load_reference
    load_reference()
    os.chdir(REF_DIR)
    from MAPLEAF.IO import SimDefinition
    from MAPLEAF.SimulationRunners import Simulation
    sd = SimDefinition(defFile, silent=True, disableDistributionSampling=True)
    for k, v in {"SimControl.loggingLevel": "0", "SimControl.RocketPlot": "Off", "SimControl.plot": "None", **(overrides or {})}.items():
        sd.setValue(k, v)
    _state["sd"] = sd
    _state["Simulation"] = Simulation


def _runReference(job):
    """job = {definition key: value} of one sample (sampled wind, pink-noise seeds, motor factors). One trajectory through the reference's
    own Simulation.run()."""
    sd = _state["sd"]
    for k, v in job.items():
        sd.setValue(k, v)
    t0 = time.perf_counter()
    sim = _state["Simulation"](simDefinition=sd, silent=True)
    flights, _ = sim.run()
    dt = time.perf_counter() - t0
    sys.stdout = sys.__stdout__
    return len(flights[0].times) - 1, 1, dt


def _initPort(defFile, overrides):
    import numpy as np  # noqa: F401
    from mapleaf_b200 import SimDefinition, buildModel
    from oracle import oracle
    sd = SimDefinition(defFile, silent=True, disableDistributionSampling=True)
    for k, v in (overrides or {}).items():
        sd.setValue(k, v)
    _state["model"] = buildModel(sd)
    oracle.lib()


def _runPort(job):
    """job = {LANE_* id: [components, n] array}."""
    import numpy as np
    from oracle import oracle
    arrays = {k: np.asarray(a, dtype=float) for k, a in job.items()}
    n = next(iter(arrays.values())).shape[1]
    t0 = time.perf_counter()
    o = oracle.run(_state["model"].blob, n, arrays)
    dt = time.perf_counter() - t0
    return int(o["counters"][:, 0].sum()), n, dt


class CpuArm:
    """A pool of worker processes that each fly trajectories of `defFile` on one host core."""

    def __init__(self, kind, defFile, overrides, cores):
        import multiprocessing as mp
        self.kind, self.cores = kind, cores
        ctx = mp.get_context("spawn")
        init = _initReference if kind == "reference" else _initPort
        self.pool = ctx.Pool(cores, initializer=init, initargs=(defFile, overrides))
        self.fn = _runReference if kind == "reference" else _runPort

    def step(self, jobs):
        """Runs the jobs over the pool; returns (accepted steps, trajectories, wall seconds)."""
        t0 = time.perf_counter()
        out = self.pool.map(self.fn, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        return sum(o[0] for o in out), sum(o[1] for o in out), wall

    def close(self):
        self.pool.close()
        self.pool.join()


def referenceAvailable():
    from oracle.ref_loader import reference_available
    return reference_available()
