"""
Time-step / target-error convergence studies on the batched kernels.

Mirrors MAPLEAF/SimulationRunners/Convergence.py (ConvergenceSimRunner.convergeSimEndPosition, :15-149, and its set-up step :300-317) and
the grid-convergence arithmetic of MAPLEAF/IO/gridConvergenceFunctions.py (:11-45,71-73,75-172). The reference flies the same definition
again and again, dividing the time step (fixed-step methods) or the adaptive integrator's target error by the refinement ratio each time;
here every refinement level is ONE LANE of a single batch: fixed-step levels differ in their first step size (LANE_START), adaptive
levels in the target error of their own model block (mlb_set_lane_models). The printed lines and the returned histories are the
reference's; `simTimeHistory` holds each level's share of the batch's device time (by accepted steps) instead of a wall time per run.
Plotting stays with MAPLEAF (`plot` / `showPlot` / axes arguments are accepted and ignored).
"""
from math import log
from statistics import mean

import numpy as np

from .model import L, buildModel
from .montecarlo import Vector, loadSimDefinition

__all__ = ["ConvergenceSimRunner", "checkConvergence"]


def _orderOfConvergence(coarseVal, medVal, fineVal, ratio, minOrder=0.5, maxOrder=2):
    logBody = (coarseVal - medVal) / (medVal - fineVal)
    if logBody <= 0:
        return minOrder
    return min(maxOrder, max(minOrder, log(logBody) / log(ratio)))


def _GCI(coarserVal, finerVal, ratio, order, factorOfSafety=3):
    return abs(factorOfSafety * abs((coarserVal - finerVal) / finerVal) / (ratio ** order - 1)) * 100


def checkConvergence(coarseVals, medVals, fineVals, gridRefinementRatio, minConvergOrder=0.5, maxConvergOrder=2, orderTolerance=0.2,
                     asympTolerance=0.05, theoreticalOrderOfConvergence=2):
    """[orders, GCI12s, GCI23s, asymptotic checks, Richardson extrapolations, uncertainties] per component
    (gridConvergenceFunctions.checkConvergence with its defaults: local orders, GCI2g uncertainty)."""
    out = [[], [], [], [], [], []]
    for cV, mV, fV in zip(coarseVals, medVals, fineVals):
        p = _orderOfConvergence(cV, mV, fV, gridRefinementRatio, minConvergOrder, maxConvergOrder)
        GCI12, GCI23 = _GCI(mV, fV, gridRefinementRatio, p), _GCI(cV, mV, gridRefinementRatio, p)
        asymp = GCI23 / (GCI12 * gridRefinementRatio ** p)
        if (theoreticalOrderOfConvergence - orderTolerance < p < theoreticalOrderOfConvergence + orderTolerance) and abs(asymp - 1) < asympTolerance:
            GCI12, GCI23 = _GCI(mV, fV, gridRefinementRatio, p, 1.25), _GCI(cV, mV, gridRefinementRatio, p, 1.25)
        richardson = fV + (fV - mV) / (gridRefinementRatio ** p - 1)
        uncertainty = 3 * (abs(fV - mV) / (gridRefinementRatio ** theoreticalOrderOfConvergence - 1))
        for lst, v in zip(out, (p, GCI12, GCI23, asymp, richardson, uncertainty)):
            lst.append(v)
    return out


class ConvergenceSimRunner:
    """Drop-in for MAPLEAF.SimulationRunners.ConvergenceSimRunner (convergeSimEndPosition)."""

    def __init__(self, simDefinitionFilePath=None, simDefinition=None, silent=False, device=0):
        self.simDefinition = loadSimDefinition(simDefinitionFilePath, simDefinition, silent)
        self.silent, self.device = silent, device

    def _print(self, *a):
        if not self.silent:
            print(*a)

    def _fly(self, model, nLanes, laneArrays=None, laneModels=None):
        from . import backend
        with backend.TrajectoryBatch(model.blob, device=self.device) as b:
            b.setLanes(nLanes, laneArrays or {})
            if laneModels is not None:
                b.setLaneModels(laneModels)
            b.run()
            _, status = b.results()
            finals = b.finals()
            steps = b.laneCounters()[:, 0] if hasattr(b, "laneCounters") else np.ones(nLanes)
            ms = b.counters().get("last_run_ms", 0.0)
        return finals, status, steps, ms

    def _setUpConfigFileForConvergenceRun(self):
        sd = self.simDefinition
        self._print("Will attempt to converge final rocket position of simulation: {}".format(sd.fileName))
        sd.disableDistributionSampling = True
        sd.setValue("SimControl.plot", "None")
        if sd.getValue("SimControl.EndCondition") != "Time":
            self._print("Running simulation to determine end time")
            finals, _, _, _ = self._fly(buildModel(sd), 1)
            endTime = float(finals[0, 13])
            self._print("Setting EndCondition = Time, EndConditionValue = {}".format(endTime))
            sd.setValue("SimControl.EndCondition", "Time")
            sd.setValue("SimControl.EndConditionValue", str(endTime))

    def convergeSimEndPosition(self, refinementRatio=2, simLimit=10, plot=True, stopAtConvergence=False, showPlot=True, plotLineLabel="Simulations",
                               ax1=None, ax2=None):
        sd = self.simDefinition
        self._setUpConfigFileForConvergenceRun()
        method = sd.getValue("SimControl.timeDiscretization")
        adaptive = "Adaptive" in method
        timeStepKey, targetErrorKey = "SimControl.timeStep", "SimControl.TimeStepAdaptation.targetError"
        self._print("Starting convergence simulations")
        value = float(sd.getValue(targetErrorKey if adaptive else timeStepKey)) * refinementRatio
        levels = []
        for _ in range(simLimit):
            value /= refinementRatio
            levels.append(value)
        # every level is a lane of one batch
        if not adaptive:
            model = buildModel(sd)
            start = np.array([[model.header("H_START_TIME")] * simLimit, levels], dtype=np.float64)
            finals, status, steps, ms = self._fly(model, simLimit, {L.LANE_START: start})
        else:
            models = []
            for v in levels:
                sd.setValue(targetErrorKey, str(v))
                models.append(buildModel(sd))
            shared = int(models[0].header("H_SHARED_LEN")) or models[0].blob.size
            finals, status, steps, ms = self._fly(models[0], simLimit, None, np.stack([m.blob[:shared] for m in models]))
        sd.setValue(targetErrorKey if adaptive else timeStepKey, str(levels[-1]))          # where the reference's loop leaves the definition
        if (np.asarray(status) != L.ST_OK).any():
            raise ValueError("convergence study: {} of {} refinement levels did not end normally".format(int((np.asarray(status) != L.ST_OK).sum()), simLimit))
        share = np.asarray(steps, dtype=float) / max(float(np.sum(steps)), 1.0)

        timeStepHistory, finalPositionHistory, simTimeHistory, convergenceHistory = [], [], [], []
        for i, v in enumerate(levels):
            timeStepHistory.append(v)
            self._print("Simulation {}, Time step: {}".format(i + 1, v))
            simTimeHistory.append(ms * 1e-3 * share[i])
            finalPositionHistory.append(Vector(*finals[i, 0:3]))
            self._print("Final Position: {:1.3f}".format(finalPositionHistory[-1]))
            if len(finalPositionHistory) >= 3:
                cV, mV, fV = finalPositionHistory[-3:]
                self._print("Checking convergence")
                result = checkConvergence(cV, mV, fV, refinementRatio)
                orders, GCI12s, GCI23s, asymptoticChecks, richardsonVals, uncertainties = result
                convergenceHistory.append(result)
                for d, name in enumerate("XYZ"):
                    self._print("{}-Direction: Order: {:>4.3f}, Asymptotic Check: {:>6.3f}, RichardsonExtrap: {:>7.3f}, Estimated Uncertainty: {:>6.3f}".format(
                        name, orders[d], asymptoticChecks[d], richardsonVals[d], uncertainties[d]))
                if stopAtConvergence and abs(sum(asymptoticChecks) / len(asymptoticChecks) - 1) < 0.1 and max(asymptoticChecks) - min(asymptoticChecks) < 0.2:
                    self._print("Simulation Converging Asymptotically")
                    self._printConvergenceHistory(method, finalPositionHistory, simTimeHistory, convergenceHistory)
                    self.convergenceHistory = convergenceHistory
                    return timeStepHistory, finalPositionHistory, finals[i]
        if simLimit >= 3:
            self._print("Asymptotic convergence not reached within {} simulations".format(simLimit))
        else:
            self._print("Asymptotic convergence impossible to reach with less than 3 iterations (performed {}). Adjust the parameter 'simLimit' to perform more iterations".format(simLimit))
        self._printConvergenceHistory(method, finalPositionHistory, simTimeHistory, convergenceHistory)
        self.convergenceHistory = convergenceHistory
        return timeStepHistory, finalPositionHistory, simTimeHistory

    def _printConvergenceHistory(self, method, finalPositionHistory, simTimeHistory, convergenceHistory):
        self._print("")
        self._print("Convergence History:")
        self._print("Integration Method: {}".format(method))
        for i, finalPos in enumerate(finalPositionHistory):
            line = "FinalPosition(m): {:>7.3f} WallTime(s): {:>7.3f} ".format(finalPos, simTimeHistory[i])
            if i > 1:
                orders, _, _, asymptoticChecks, _, _ = convergenceHistory[i - 2]
                line += " Avg Order: {:>4.2f}, Avg Asymptotic Check: {:>6.3f}".format(mean(orders), mean(asymptoticChecks))
            self._print(line)
