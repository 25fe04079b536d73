"""
Optimisation runners (mapleaf_b200/optimization.py) against the reference's own runner (MAPLEAF/SimulationRunners/Optimization.py):
tests/golden/optimization.json holds every (trial point, cost) the unmodified reference evaluated in a Nelder-Mead run, its optimum and the
continuation definition it wrote (tests/golden/make_golden.py optimization). Host logic on the CPU oracle here; the same run on the GPU
in tests/test_gpu_parity.py.
"""
import os

import numpy as np
import pytest

from conftest import DATA
from mapleaf_b200 import SimDefinition
from mapleaf_b200.optimization import PSORunner, ScipyMinimizeRunner, evalExpression, optimizationRunnerFactory
from test_host_runner import OracleBatch


def referenceCase(golden, tmp_path, extra=None):
    g = golden("optimization.json")
    sd = SimDefinition(os.path.join(DATA, g["definition"]), silent=True, disableDistributionSampling=True)
    for k, v in {**g["overrides"], **(extra or {})}.items():
        sd.setValue(k, v)
    sd.fileName = str(tmp_path / "Optimization.mapleaf")
    return g, sd


def checkAgainstReference(g, runner, cost, pos, tmp_path, tol):
    assert len(runner.costHistory) == len(g["evaluations"])
    for mine, x, ref in zip(runner.costHistory, runner.positionHistory, g["evaluations"]):
        assert x == ref["x"]                                   # scipy asked for the very same points
        assert abs(mine - ref["cost"]) <= tol * abs(ref["cost"])
    assert abs(cost - g["bestCost"]) <= tol * abs(g["bestCost"]) and list(pos) == g["bestPosition"]
    cont = SimDefinition(str(tmp_path / "Optimization_continue.mapleaf"), silent=True, disableDistributionSampling=True)
    for k, v in g["continuation"].items():
        assert cont.getValue(k) == v
    # the rest of the definition survives the round trip through the writer
    assert cont.getValue("Rocket.Sustainer.TailFins.rootChord") == "0.3048" and cont.getValue("Optimization.method") == g["overrides"]["Optimization.method"]


def test_scipy_minimize_runner_follows_the_reference(golden, tmp_path):
    g, sd = referenceCase(golden, tmp_path)
    runner = optimizationRunnerFactory(simDefinition=sd, silent=True, batchClass=OracleBatch)
    assert isinstance(runner, ScipyMinimizeRunner) and runner.minimizeMethod == "Nelder-Mead"
    assert runner.varNames == ["finSpan", "noseAR"] and runner.varKeys == ["Rocket.Sustainer.TailFins.span", "Rocket.Sustainer.Nosecone.aspectRatio"]
    assert runner.minVals == [0.08, 2.0] and runner.maxVals == [0.25, 8.0]
    assert runner.dependentVars == ["Rocket.Sustainer.TailFins.tipChord"]
    cost, pos = runner.runOptimization()
    checkAgainstReference(g, runner, cost, pos, tmp_path, 1e-11)
    assert runner.nSimulations == len(g["evaluations"]) == runner.nBatches


def test_dependent_variables_and_cost_expression(golden, tmp_path):
    g, sd = referenceCase(golden, tmp_path)
    runner = optimizationRunnerFactory(simDefinition=sd, silent=True, batchClass=OracleBatch)
    trial = SimDefinition(os.path.join(DATA, g["definition"]), silent=True, disableDistributionSampling=True)
    varDict = runner._updateIndependentVariableValues(trial, [0.2, 3.0])
    runner._updateDependentVariableValues(trial, varDict)
    assert trial.getValue("Rocket.Sustainer.TailFins.span") == "0.2" and trial.getValue("Rocket.Sustainer.Nosecone.aspectRatio") == "3.0"
    assert trial.getValue("Rocket.Sustainer.TailFins.tipChord") == str(0.1 + 0.4 * 0.2)
    assert evalExpression("1/apogee if (apogee > 1) else max(1, -1*apogee)", {"apogee": 4.0}) == 0.25
    assert evalExpression("math.sqrt(flightTime)", {"flightTime": 9.0}) == 3.0
    runner.costFunctionDefinition = "someModule:someFunction"
    with pytest.raises(NotImplementedError, match="log files"):
        runner._computeCostFunctionValues([[0.2, 3.0]])
    sd.setValue("Optimization.IndependentVariables.finSpan", "0.08 < Rocket.Sustainer.TailFins.span")
    with pytest.raises(ValueError, match="should be scalars and conform"):
        optimizationRunnerFactory(simDefinition=sd, silent=True, batchClass=OracleBatch)
    sd.setValue("Optimization.method", "Annealing")
    sd.setValue("Optimization.IndependentVariables.finSpan", "0.08 < Rocket.Sustainer.TailFins.span < 0.25")
    with pytest.raises(ValueError, match="not implemented, try PSO"):
        optimizationRunnerFactory(simDefinition=sd, silent=True, batchClass=OracleBatch)


def test_particle_swarm_flies_every_iteration_as_one_batch(golden, tmp_path):
    """Four particles (two given, two drawn), three iterations: 3 batches of 4 lanes; the best cost never rises, every position stays in
    the bounds, the given positions are flown first, and the continuation file restarts from the swarm's last positions."""
    g, sd = referenceCase(golden, tmp_path, {
        "Optimization.method": "PSO", "Optimization.ParticleSwarm.nParticles": "4", "Optimization.ParticleSwarm.nIterations": "3",
        "Optimization.IndependentVariables.InitialParticlePositions.p0.finSpan": "0.165", "Optimization.IndependentVariables.InitialParticlePositions.p0.noseAR": "5.0",
        "Optimization.IndependentVariables.InitialParticlePositions.p1.finSpan": "0.0886875"})
    np.random.seed(3)
    runner = optimizationRunnerFactory(simDefinition=sd, silent=True, batchClass=OracleBatch)
    assert isinstance(runner, PSORunner) and (runner.c1, runner.c2, runner.w) == (0.5, 0.6, 0.9)
    assert list(runner.initPositions[0]) == [0.165, 5.0] and runner.initPositions[1, 0] == 0.0886875
    flown = []
    inner = runner._computeCostFunctionValues
    runner._computeCostFunctionValues = lambda xs: (flown.append(np.array(xs)), inner(xs))[1]
    cost, pos = runner.runOptimization()
    assert runner.nBatches == 3 and runner.nSimulations == 12 and len(flown) == 3
    assert np.array_equal(flown[0], runner.initPositions)
    ref = {tuple(e["x"]): e["cost"] for e in g["evaluations"]}
    firstCosts = inner([[0.165, 5.0]])
    assert abs(firstCosts[0] - ref[(0.165, 5.0)]) <= 1e-11 * ref[(0.165, 5.0)]          # same flight, same cost as the reference's first evaluation
    for xs in flown + [runner.finalPositions]:
        assert (xs >= np.array(runner.minVals)).all() and (xs <= np.array(runner.maxVals)).all()
    assert all(b <= a for a, b in zip(runner.costHistory, runner.costHistory[1:])) and cost == runner.costHistory[-1] <= ref[(0.165, 5.0)] * (1 + 1e-11)
    cont = SimDefinition(str(tmp_path / "Optimization_continue.mapleaf"), silent=True, disableDistributionSampling=True)
    assert float(cont.getValue("Optimization.IndependentVariables.InitialParticlePositions.bestCost")) == cost
    assert [float(x) for x in cont.getValue("Optimization.IndependentVariables.InitialParticlePositions.bestPosition").split(" ")] == list(pos)
    for i in range(4):
        assert float(cont.getValue("Optimization.IndependentVariables.InitialParticlePositions.p{}.noseAR".format(i))) == runner.finalPositions[i, 1]
    # and the continuation is a valid start: positions and best point are read back
    again = optimizationRunnerFactory(simDefinition=cont, silent=True, batchClass=OracleBatch)
    assert np.array_equal(again.initPositions, runner.finalPositions) and again.bestCost == cost and list(again.bestPosition) == list(pos)


def test_written_definition_reads_back_identically(tmp_path):
    sd = SimDefinition(os.path.join(DATA, "TwoStage.mapleaf"), silent=True, disableDistributionSampling=True)
    sd.setValue("Optimization.IndependentVariables.a", "1 < Rocket.position < 2")
    path = str(tmp_path / "copy.mapleaf")
    before = dict(sd.dict)
    sd.writeToFile(path)
    back = SimDefinition(path, silent=True, disableDistributionSampling=True)
    assert back.dict == before
    text = open(path).read()
    assert text.startswith("# MAPLEAF\n# File: " + path) and "\nRocket{\n" in text and text.rstrip().endswith("}")


def test_batch_of_cases_and_inner_optimisation(golden, tmp_path):
    """Optimization.batchDefinition: the cost of a trial solution is the mean over the cases (Optimization.py:201-233,268-298), all cases of all
    trial solutions in one batch. InnerOptimization: every outer trial first optimises the inner variables (:240-247,286-290)."""
    g = golden("optimization.json")
    deck = os.path.join(DATA, g["definition"])
    batchFile = tmp_path / "cases.mapleaf"
    batchFile.write_text("calm{\n    simDefinitionFile " + deck + "\n    ParameterOverrides{\n        Environment.ConstantMeanWind.velocity (0 0 0)\n    }\n}\n"
                         "windy{\n    simDefinitionFile " + deck + "\n    ParameterOverrides{\n        Environment.ConstantMeanWind.velocity (9 -4 0.3)\n    }\n}\n")
    base = {k: v for k, v in g["overrides"].items() if "noseAR" not in k}
    single = {}
    for name, wind in (("calm", "(0 0 0)"), ("windy", "(9 -4 0.3)")):
        sd = SimDefinition(deck, silent=True, disableDistributionSampling=True)
        for k, v in {**base, "Environment.ConstantMeanWind.velocity": wind}.items():
            sd.setValue(k, v)
        single[name] = optimizationRunnerFactory(simDefinition=sd, silent=True, batchClass=OracleBatch)._computeCostFunctionValues([[0.1], [0.2]])
    sd = SimDefinition(deck, silent=True, disableDistributionSampling=True)
    for k, v in {**base, "Optimization.batchDefinition": str(batchFile)}.items():
        sd.setValue(k, v)
    runner = optimizationRunnerFactory(simDefinition=sd, silent=True, batchClass=OracleBatch)
    assert len(runner.simDefinitions) == 2
    both = runner._computeCostFunctionValues([[0.1], [0.2]])
    assert runner.nBatches == 1 and runner.nSimulations == 4                       # 2 trial solutions x 2 cases, one kernel run
    for i in range(2):
        assert abs(both[i] - (single["calm"][i] + single["windy"][i]) / 2) <= 1e-15
    assert single["calm"][0] != single["windy"][0]

    # nested: the outer variable is the fin span, the inner optimisation (Nelder-Mead, 2 iterations) picks the nose aspect ratio for each trial
    sd = SimDefinition(deck, silent=True, disableDistributionSampling=True)
    nested = {**base, "Optimization.InnerOptimization.method": "scipy.optimize.minimize Nelder-Mead",
              "Optimization.InnerOptimization.costFunction": g["overrides"]["Optimization.costFunction"],
              "Optimization.InnerOptimization.IndependentVariables.noseAR": g["overrides"]["Optimization.IndependentVariables.noseAR"],
              "Optimization.InnerOptimization.ScipyMinimize.maxIterations": "2", "Optimization.InnerOptimization.showConvergencePlot": "false"}
    for k, v in nested.items():
        sd.setValue(k, v)
    sd.fileName = str(tmp_path / "Nested.mapleaf")
    outer = optimizationRunnerFactory(simDefinition=sd, silent=True, batchClass=OracleBatch)
    assert outer.varNames == ["finSpan"] and outer._hasInnerOptimization()
    cost = outer._computeCostFunctionValues([[0.165]])[0]
    plain = single["calm"]          # (other wind; only used for the type) the nested cost is a flight's cost at the inner optimum:
    inner = outer._createNestedOptimization(sd)
    assert inner.varNames == ["noseAR"] and isinstance(inner, ScipyMinimizeRunner)
    assert isinstance(cost, float) and isinstance(plain[0], float) and outer.nSimulations > 3       # inner evaluations are counted
    # the inner optimum is at least as good as the inner start point (centre of its bounds, noseAR = 5) at the same fin span
    ref = {tuple(e["x"]): e["cost"] for e in g["evaluations"]}
    assert cost <= ref[(0.165, 5.0)] * (1 + 1e-11)


def test_scipy_runner_gives_a_crashed_trial_a_high_cost(golden, tmp_path):
    """scipy's optimisers may leave the bounds; a trial the model builder (or the flight) rejects gets max(costHistory) x 1.1 while fewer than six
    costs are known, max(costHistory) after that, and a crash of the very first trial propagates (Optimization.py:529-546)."""
    g, sd = referenceCase(golden, tmp_path)
    runner = optimizationRunnerFactory(simDefinition=sd, silent=True, batchClass=OracleBatch)
    costs = iter([2.0, 3.0, ValueError("boom"), 1.0, 1.5, 2.5, 0.5, ValueError("boom")])

    def fake(trials):
        c = next(costs)
        if isinstance(c, Exception):
            raise c
        return [c]
    runner._computeCostFunctionValues = fake
    out = [runner._evaluateCostFunction([0.1, 3.0]) for _ in range(8)]
    assert out == [2.0, 3.0, 3.0 * 1.1, 1.0, 1.5, 2.5, 0.5, 3.0 * 1.1]
    fresh = optimizationRunnerFactory(simDefinition=sd, silent=True, batchClass=OracleBatch)
    fresh._computeCostFunctionValues = lambda trials: (_ for _ in ()).throw(ValueError("first trial"))
    with pytest.raises(ValueError, match="first trial"):
        fresh._evaluateCostFunction([0.1, 3.0])
