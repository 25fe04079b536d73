"""
Design optimisation on the batched kernels.

Mirrors MAPLEAF/SimulationRunners/Optimization.py: the `Optimization` dictionary of a .mapleaf file (independent variables
`min < path.To.Key < max`, dependent variables with `!expression!` parts, a cost-function expression in flightTime / apogee / maxSpeed /
maxHorizontalVel, optional batch of cases, optional InnerOptimization), `optimizationRunnerFactory`, `PSORunner`, `ScipyMinimizeRunner`
(:22-56 cost function, :59-343 the shared runner, :345-505 particle swarm, :507-614 scipy.optimize.minimize).

What is different, and why. The reference flies one `Simulation` per trial solution (and per case), one after the other or one ray task
each. Here ALL trial solutions an optimiser asks for at once are lanes of ONE batch: each trial's definition is flattened into its own
model block (`buildModel`) and handed over as a per-lane model block (`mlb_set_lane_models`), so a swarm iteration of N particles x C
cases is one kernel run of N x C lanes. scipy.optimize.minimize asks for one point at a time, so there a batch is a single lane (plus
the finite-difference points when the method takes the gradient from `jac`-less differences, which scipy also requests one by one).

* `ScipyMinimizeRunner` calls the same scipy routine with the same arguments as the reference (:575-600), so the iterate sequence is the
  reference's as long as the cost values agree (pinned by tests/golden/optimization.json, produced by the reference's own runner).
* `PSORunner`: the reference delegates to `pyswarms.single.GlobalBestPSO` (:419). pyswarms is not a dependency of this package; the swarm
  update here is the published global-best rule that class implements (Kennedy & Eberhart 1995 with inertia weight, Shi & Eberhart 1998:
  v <- w v + c1 r1 (pbest - x) + c2 r2 (gbest - x); x <- x + v, positions folded back into the bounds periodically, which is pyswarms'
  default `BoundaryHandler(strategy="periodic")`), with numpy's global generator for r1, r2 and the initial velocities. Random draws are
  therefore NOT draw-for-draw those of a pyswarms run: parity for this runner is the update rule and the reference's parsing, initial
  positions, continuation file and return values, not a recorded trajectory (parity unpinned for the swarm's random sequence).
* Cost functions given as `module:function` receive the paths of the reference's log files (:36-45); this backend writes time-step logs
  only through the Monte Carlo runner, so that form raises NotImplementedError.
* Plots (`showConvergencePlot`) stay with MAPLEAF: accepted and ignored.
"""
import math
import re
from copy import deepcopy
from statistics import mean

import numpy as np

from .model import L, buildModel
from .montecarlo import loadSimDefinition
from .simdef import SimDefinition, SubDictReader

__all__ = ["optimizationRunnerFactory", "OptimizingSimRunner", "PSORunner", "ScipyMinimizeRunner", "evalExpression"]


def evalExpression(statement: str, additionalVars=None):
    """MAPLEAF.Utilities.evalExpression (Utilities.py:25-30): Python expression with `math` available."""
    return eval(statement, {"math": math}, dict(additionalVars or {}))


class OptimizingSimRunner:
    """Reads the `Optimization` dictionary; child classes implement the search (Optimization.py:59-343)."""

    def __init__(self, optimizationReader, silent=False, parallel=False, device=0, batchClass=None):
        self.silent, self.parallel, self.device = silent, parallel, device
        self.optimizationReader = optimizationReader
        self._batchClass = batchClass
        self.costFunctionDefinition = optimizationReader.getString("costFunction")
        self.varKeys, self.varNames, self.minVals, self.maxVals = self._loadIndependentVariables()
        self.dependentVars, self.dependentVarDefinitions = self._loadDependentVariables()
        self.bestPosition, self.bestCost = self._loadBestPosition()
        self.simDefinitions = [optimizationReader.simDefinition]
        self.nSimulations = 0           # trajectories flown so far
        self.nBatches = 0               # kernel runs they took
        if optimizationReader.tryGetString("batchDefinition", defaultValue="notPresent.fakeFile") != "notPresent.fakeFile":
            self._loadBatchOptimization()

    def _print(self, *a, **k):
        if not self.silent:
            print(*a, **k)

    # ---- parsing (Optimization.py:91-199) ----------------------------------------------------------------------------
    def _loadIndependentVariables(self):
        varKeys, varNames, minVals, maxVals = [], [], [], []
        for key in self.optimizationReader.getImmediateSubKeys("IndependentVariables"):
            strings = self.optimizationReader.getString(key).split("<")
            if len(strings) != 3:
                raise ValueError("Parameters in the Optimization.IndependentVariables dictionary should be scalars and conform to the following format: "
                                 "'[minValue] < [path.To.Parameter] < [maxValue]' Problem key is {}, which has a value of {}".format(key, " ".join(strings)))
            minVal, keyPath, maxVal = [s.strip() for s in strings]
            varKeys.append(keyPath)
            varNames.append(key.split(".")[-1])
            minVals.append(float(minVal))
            maxVals.append(float(maxVal))
        self._print("Independent Variables: ")
        for name, lo, hi in zip(varNames, minVals, maxVals):
            self._print("    {} < {} < {}".format(lo, name, hi))
        self._print("")
        return varKeys, varNames, minVals, maxVals

    def _loadDependentVariables(self):
        names, definitions = [], []
        for depVar in self.optimizationReader.getSubKeys("DependentVariables"):
            names.append(re.sub("Optimization.*DependentVariables.", "", depVar))
            definitions.append(self.optimizationReader.getString(depVar))
        if names:
            self._print("Dependent variables:")
            for n, d in zip(names, definitions):
                self._print("    {} = {}".format(n, d))
            self._print("")
        return names, definitions

    def _loadBestPosition(self):
        bestPosition = self.optimizationReader.tryGetString("IndependentVariables.InitialParticlePositions.bestPosition", defaultValue=None)
        bestCost = self.optimizationReader.tryGetFloat("IndependentVariables.InitialParticlePositions.bestCost", defaultValue=None)
        if bestPosition is not None and bestCost is not None:
            bestPosition = np.array([float(x) for x in bestPosition.split(" ")], np.float64)
            self._print("Loaded previous best position: {}".format(bestPosition))
            self._print("Loaded previous best cost: {}\n".format(bestCost))
            return bestPosition, np.float64(bestCost)
        if bestPosition is None and bestCost is None:
            return None, None
        raise ValueError("Please specify bestPosition and bestCost or neither, current values: {} and {}, respectively".format(bestPosition, bestCost))

    def _loadBatchOptimization(self):
        """Optimization.batchDefinition: every root-level dictionary of that file is a case (simDefinitionFile + ParameterOverrides); the
        cost of a trial solution is the mean over the cases (Optimization.py:201-233, Batch.py `_implementParameterOverrides`)."""
        import os
        self._print("Batch Optimization")
        path = self.optimizationReader.getString("batchDefinition")
        if not os.path.isabs(path):
            base = os.path.dirname(os.path.abspath(self.optimizationReader.simDefinition.fileName or "."))
            path = os.path.join(base, path) if os.path.exists(os.path.join(base, path)) else os.path.abspath(path)
        batchDefinition = SimDefinition(path, silent=True, disableDistributionSampling=True)
        self.simDefinitions = []
        self._print("Found the following cases:")
        for case in sorted({k.split(".")[0] for k in batchDefinition.dict if "." in k}):          # every root-level dictionary is a case
            self._print(case)
            simDefPath = batchDefinition.getValue("{}.simDefinitionFile".format(case))
            if not os.path.isabs(simDefPath) and not os.path.exists(simDefPath):
                simDefPath = os.path.join(os.path.dirname(path), simDefPath)
            caseDefinition = SimDefinition(simDefPath, silent=True, disableDistributionSampling=True)
            prefix = case + ".ParameterOverrides."
            for k in batchDefinition.getSubKeys(case + ".ParameterOverrides"):
                caseDefinition.setValue(k[len(prefix):], batchDefinition.getValue(k))
            self.simDefinitions.append(caseDefinition)
        self._print("")

    # ---- trial solutions -> definitions (Optimization.py:300-336) ------------------------------------------------------
    def _updateIndependentVariableValues(self, simDefinition, trialSolution):
        indVarValueDict = {}
        for i in range(len(trialSolution)):
            simDefinition.setValue(self.varKeys[i], str(trialSolution[i]))
            indVarValueDict[self.varNames[i]] = trialSolution[i]
        return indVarValueDict

    def _updateDependentVariableValues(self, simDefinition, indVarValueDict):
        for name, definition in zip(self.dependentVars, self.dependentVarDefinitions):
            parts = definition.split("!")          # "(0 0 !a+b!)" -> ["(0 0 ", "a+b", ")"]: odd parts are expressions
            for j in range(1, len(parts), 2):
                parts[j] = str(evalExpression(parts[j], indVarValueDict))
            simDefinition.setValue(name, "".join(parts))

    def _createNestedOptimization(self, simDef):
        reader = SubDictReader(self.optimizationReader.simDefDictPathToReadFrom + ".InnerOptimization", simDef)
        return optimizationRunnerFactory(optimizationReader=reader, silent=self.silent, parallel=self.parallel, device=self.device, batchClass=self._batchClass)

    def _hasInnerOptimization(self):
        return self.optimizationReader.simDefDictPathToReadFrom + ".InnerOptimization" in self.optimizationReader.getImmediateSubDicts()

    # ---- cost of a set of trial solutions: ONE batch (replaces Optimization.py:240-298) --------------------------------
    def _costOfFlight(self, results):
        if ":" in self.costFunctionDefinition:
            raise NotImplementedError("Optimization.costFunction '{}': cost functions in an importable module read the reference's log files "
                                      "(Optimization.py:36-45); give an expression in flightTime, apogee, maxSpeed, maxHorizontalVel".format(self.costFunctionDefinition))
        varVals = {"flightTime": float(results[L.RES_FLIGHT_TIME]), "apogee": float(results[L.RES_APOGEE]), "maxSpeed": float(results[L.RES_MAX_SPEED]),
                   "maxHorizontalVel": float(results[L.RES_MAX_HVEL])}
        return evalExpression(self.costFunctionDefinition, varVals)

    def _fly(self, models):
        """One kernel run: lane i flies models[i] (its own model block). Returns the [n, 7] result block and the status words."""
        from . import backend
        cls = self._batchClass or backend.TrajectoryBatch
        shared = int(models[0].header("H_SHARED_LEN")) or models[0].blob.size
        b = cls(models[0].blob, device=self.device, specialize=False)          # every lane has its own model block: the generic library
        try:
            b.setLanes(len(models), {})
            if len(models) > 1:
                b.setLaneModels(np.stack([m.blob[:shared] for m in models]))
            b.run()
            res, status = b.results()
        finally:
            if hasattr(b, "close"):
                b.close()
        self.nSimulations += len(models)
        self.nBatches += 1
        return res, status

    def _computeCostFunctionValues(self, trialSolutions):
        """Cost of every trial solution (mean over the cases of a batch optimisation), all flown as lanes of one batch."""
        definitions = []
        for independentVariableValues in trialSolutions:
            for case in self.simDefinitions:
                simDef = deepcopy(case)
                simDef.disableDistributionSampling = True
                varDict = self._updateIndependentVariableValues(simDef, independentVariableValues)
                self._updateDependentVariableValues(simDef, varDict)
                if self._hasInnerOptimization():
                    inner = self._createNestedOptimization(simDef)
                    _, pos = inner.runOptimization()
                    varDict = inner._updateIndependentVariableValues(simDef, pos)
                    inner._updateDependentVariableValues(simDef, varDict)
                    self.nSimulations += inner.nSimulations
                    self.nBatches += inner.nBatches
                definitions.append(simDef)
        models = [buildModel(sd) for sd in definitions]          # a definition the constructors reject raises here, as Simulation() would
        res, status = self._fly(models)
        bad = np.nonzero(np.asarray(status) != L.ST_OK)[0]
        if bad.size:
            raise ValueError("optimisation: trial flight {} ended with status {}".format(int(bad[0]), int(status[bad[0]])))
        nCases = len(self.simDefinitions)
        costs = []
        for t in range(len(trialSolutions)):
            caseCosts = [self._costOfFlight(res[t * nCases + c]) for c in range(nCases)]
            costs.append(mean(caseCosts) if nCases > 1 else caseCosts[0])
        return costs

    # the reference's two entry points, kept by name: both are the batched evaluation
    _computeCostFunctionValues_SingleThreaded = _computeCostFunctionValues
    _computeCostFunctionValues_Parallel = _computeCostFunctionValues

    def runOptimization(self):
        raise NotImplementedError


def optimizationRunnerFactory(simDefinitionFilePath=None, simDefinition=None, optimizationReader=None, silent=False, parallel=False, device=0, batchClass=None) -> OptimizingSimRunner:
    """Optimization.py:338-361."""
    if simDefinition is not None or simDefinitionFilePath is not None:
        simDefinition = loadSimDefinition(simDefinitionFilePath, simDefinition, silent)
        optimizationReader = SubDictReader("Optimization", simDefinition)
        simDefinition.setValue("SimControl.plot", "None")
        simDefinition.setValue("SimControl.RocketPlot", "Off")
    elif optimizationReader is None:
        raise ValueError("subDictReader not initialized for a nested optimization")
    method = optimizationReader.tryGetString("method", defaultValue="PSO")
    if method == "PSO":
        return PSORunner(optimizationReader, silent=silent, parallel=parallel, device=device, batchClass=batchClass)
    if "scipy.optimize.minimize" in method:
        return ScipyMinimizeRunner(optimizationReader, silent=silent, parallel=parallel, device=device, batchClass=batchClass)
    raise ValueError("Optimization method: {} not implemented, try PSO".format(method))


class PSORunner(OptimizingSimRunner):
    """Global-best particle swarm (Optimization.py:345-505); every iteration's particles are one batch."""

    def __init__(self, optimizationReader, silent=False, parallel=False, device=0, batchClass=None):
        if not silent:
            print("Particle Swarm Optimization")
        super().__init__(optimizationReader, silent, parallel, device, batchClass)
        self.initPositions = self._loadInitialPositions()
        reader = SubDictReader(optimizationReader.simDefDictPathToReadFrom + ".ParticleSwarm", optimizationReader.simDefinition)
        self.nParticles = reader.tryGetInt("nParticles", defaultValue=20)
        self.nIterations = reader.tryGetInt("nIterations", defaultValue=50)
        self.c1 = reader.tryGetFloat("cognitiveParam", defaultValue=0.5)
        self.c2 = reader.tryGetFloat("socialParam", defaultValue=0.6)
        self.w = reader.tryGetFloat("inertiaParam", defaultValue=0.9)
        self.showConvergence = optimizationReader.tryGetBool("showConvergencePlot", defaultValue=False)
        self.costHistory = []
        self._print("Optimization Parameters:")
        self._print("{} Particles".format(self.nParticles))
        self._print("{} Iterations".format(self.nIterations))
        self._print("c1 = {}, c2 = {}, w = {}\n".format(self.c1, self.c2, self.w))
        self._print("Cost Function:")
        self._print(self.costFunctionDefinition + "\n")

    def _loadInitialPositions(self):
        """Optimization.py:354-413: given positions first, the rest uniform in the bounds (numpy's global generator, variable by variable)."""
        nParticles = self.optimizationReader.getInt("ParticleSwarm.nParticles")
        nVars = len(self.varNames)
        dicts = self.optimizationReader.getImmediateSubDicts("IndependentVariables.InitialParticlePositions")
        if len(dicts) > nParticles:
            raise ValueError("Number of initial particle positions: {}({}) cannot exceed the total number of particles: {}".format(len(dicts), dicts, nParticles))
        positions = np.empty((nParticles, nVars), dtype=np.float64)
        for i in range(nParticles):
            paths, names = [], []
            if i < len(dicts):
                if i == 0:
                    self._print("Loading specified initial particle positions")
                self._print("    " + dicts[i] + ":")
                paths = self.optimizationReader.getImmediateSubKeys(dicts[i])
                for v in paths:
                    name = v.split(".")[-1]
                    names.append(name)
                    if name not in self.varNames:
                        raise ValueError("Variable {} not expected in {}. Only expecting independent variables: {}".format(v, dicts[i], self.varNames))
            for j, var in enumerate(self.varNames):
                if var in names:
                    value = self.optimizationReader.getFloat(paths[names.index(var)])
                    self._print("        {} = {}".format(var, value))
                    if value < self.minVals[j] or value > self.maxVals[j]:
                        raise ValueError("{} value in initial particle position: {} ({}) outside of expected range ({}-{}".format(var, dicts[i], value, self.minVals[j], self.maxVals[j]))
                else:
                    value = np.random.uniform(self.minVals[j], self.maxVals[j])
                positions[i, j] = value
        return positions

    def _createContinuationFile(self, particlePositions, bestPosition, bestCost):
        """Optimization.py:446-478: the definition with the swarm's last positions and best point as InitialParticlePositions."""
        restart = deepcopy(self.optimizationReader.simDefinition)
        for p in self.optimizationReader.getSubKeys("IndependentVariables.InitialParticlePositions"):
            restart.removeKey(p)
        path = self.optimizationReader.simDefDictPathToReadFrom
        restart.setValue(path + ".IndependentVariables.InitialParticlePositions.bestPosition", " ".join(str(x) for x in bestPosition))
        restart.setValue(path + ".IndependentVariables.InitialParticlePositions.bestCost", str(bestCost))
        for i, position in enumerate(particlePositions):
            for j, value in enumerate(position):
                restart.setValue(path + ".IndependentVariables.InitialParticlePositions.p{}.{}".format(i, self.varNames[j]), str(value))
        if restart.fileName:
            newFileName = restart.fileName.replace(".mapleaf", "_continue.mapleaf")
            self._print("Creating continuation file: {}".format(newFileName))
            restart.writeToFile(newFileName)
        return restart

    def runOptimization(self):
        lo, hi = np.array(self.minVals, dtype=np.float64), np.array(self.maxVals, dtype=np.float64)
        x = self.initPositions.copy()
        n, d = x.shape
        v = np.random.uniform(0.0, 1.0, size=(n, d))                       # pyswarms' generate_velocity default clamp (0, 1)
        pbest, pbestCost = x.copy(), np.full(n, np.inf)
        if self.bestPosition is not None and self.bestCost is not None:
            gbest, gbestCost = np.array(self.bestPosition, dtype=np.float64), float(self.bestCost)
        else:
            gbest, gbestCost = x[0].copy(), np.inf
        for _ in range(self.nIterations):
            cost = np.asarray(self._computeCostFunctionValues(x), dtype=np.float64)       # one batch of n x cases lanes
            better = cost < pbestCost
            pbest[better], pbestCost[better] = x[better], cost[better]
            if pbestCost.min() < gbestCost:
                k = int(np.argmin(pbestCost))
                gbest, gbestCost = pbest[k].copy(), float(pbestCost[k])
            self.costHistory.append(gbestCost)
            r1, r2 = np.random.uniform(0, 1, size=(n, d)), np.random.uniform(0, 1, size=(n, d))
            v = self.w * v + self.c1 * r1 * (pbest - x) + self.c2 * r2 * (gbest - x)
            x = x + v
            span = hi - lo                                                    # periodic boundary: fold back into [lo, hi]
            x = np.where(x > hi, lo + np.mod(x - hi, span), x)
            x = np.where(x < lo, hi - np.mod(lo - x, span), x)
        self.finalPositions = x
        self._createContinuationFile(x, gbest, gbestCost)
        self._print("\nBest position: {}\nBest cost: {}\n".format(gbest, gbestCost))
        return gbestCost, gbest


class ScipyMinimizeRunner(OptimizingSimRunner):
    """scipy.optimize.minimize with the reference's arguments (Optimization.py:507-614)."""

    def __init__(self, optimizationReader, silent=False, parallel=False, device=0, batchClass=None):
        method = optimizationReader.getString("method")
        splitMethod = method.split(" ")
        if len(splitMethod) > 2:
            raise ValueError("Unexpected format for optimization method. Expected scipy.optimize.minimize [method], got: {}".format(method))
        self.minimizeMethod = splitMethod[1]
        if not silent:
            print("Optimization using scipy.optimize.minimize, method={}".format(self.minimizeMethod))
        super().__init__(optimizationReader, silent, parallel, device, batchClass)
        self.showConvergence = optimizationReader.tryGetBool("showConvergencePlot", defaultValue=True)
        self.costHistory = []
        self.positionHistory = []

    def _evaluateCostFunction(self, independentVariableValues):
        try:
            result = self._computeCostFunctionValues([independentVariableValues])[0]
        except NotImplementedError:
            raise
        except Exception:
            # scipy's optimisers may leave the bounds: a definition the model builder rejects gets a high cost (Optimization.py:529-546)
            self._print("Simulation crash, assigning high cost: ", end="")
            if len(self.costHistory) > 5:
                result = max(self.costHistory)
            elif len(self.costHistory) > 0:
                result = max(self.costHistory) * 1.1
            else:
                raise
        self.costHistory.append(result)
        self.positionHistory.append([float(x) for x in independentVariableValues])
        self._print(result)
        return result

    def _createContinuationFile(self, currentSolution):
        restart = deepcopy(self.optimizationReader.simDefinition)
        for i, name in enumerate(self.varNames):
            path = self.optimizationReader.simDefDictPathToReadFrom + ".IndependentVariables." + name
            eps = 1e-12
            restart.setValue(path, "{} < {} < {}".format(currentSolution[i] - eps, self.varKeys[i], currentSolution[i] + eps))
        if restart.fileName:
            newFileName = restart.fileName.replace(".mapleaf", "_continue.mapleaf")
            self._print("Creating continuation file: {}".format(newFileName))
            restart.writeToFile(newFileName)
        return restart

    def runOptimization(self):
        from scipy.optimize import minimize
        tolerance = self.optimizationReader.tryGetFloat("ScipyMinimize.tolerance", defaultValue=0.01)
        maxIterations = self.optimizationReader.tryGetInt("ScipyMinimize.maxIterations", defaultValue=50)
        printConvergence = self.optimizationReader.tryGetBool("ScipyMinimize.printStatus", defaultValue=True) and not self.silent
        self._print("Options:")
        self._print("    Tolerance: {}".format(tolerance))
        self._print("    Max Iterations: {}\n".format(maxIterations))
        self._print("Running optimization:")
        initialPosition = np.array([(lo + hi) / 2 for lo, hi in zip(self.minVals, self.maxVals)])     # centre of the bounds
        result = minimize(self._evaluateCostFunction, initialPosition, method=self.minimizeMethod, options={"maxiter": maxIterations, "disp": printConvergence}, tol=tolerance)
        self._createContinuationFile(result.x)
        self._print("\nBest position: {}\nBest cost: {}\n".format(result.x, result.fun))
        return result.fun, result.x
