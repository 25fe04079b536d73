import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")
DATA = os.path.join(REPO, "mapleaf_b200", "data", "Simulations")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json

    def load(name):
        with open(os.path.join(GOLDEN, name)) as f:
            return json.load(f)
    return load


# The reference's own regression case `TimeStepAdaptationDemo` (MAPLEAF/Examples/BatchSims/regressionTests.mapleaf:87-114): its
# AdaptTimeStep.mapleaf rocket is the rocket of MonteCarloAdaptive.mapleaf with a different stage CG / MOI override; the radiosonde wind of
# that file is replaced by the case's own ParameterOverrides. Expected by the reference: PositionX -141.033 at Time 45 s, within 0.2 %
# (SimulationRunners/Batch.py:51-52).
TIMESTEP_ADAPTATION_DEMO = {
    "Rocket.Sustainer.constCG": "(0 0 -2.8)", "Rocket.Sustainer.constMOI": "(85 85 0.3)",
    "Environment.LaunchSite.elevation": "755", "Environment.AtmosphericPropertiesModel": "Constant",
    "Environment.ConstantAtmosphere.temp": "15", "Environment.ConstantAtmosphere.pressure": "101325",
    "Environment.ConstantAtmosphere.density": "1.225", "Environment.ConstantAtmosphere.viscosity": "1.789e-5",
    "Environment.MeanWindModel": "Constant", "Environment.ConstantMeanWind.velocity": "(1 0 0)",
    "Environment.TurbulenceModel": "PinkNoise2D", "Environment.PinkNoiseModel.turbulenceIntensity": "2",
    "Environment.PinkNoiseModel.velocityStdDeviation": "1",
    "Environment.PinkNoiseModel.randomSeed1": "63583", "Environment.PinkNoiseModel.randomSeed2": "63583",
    "Environment.LaunchSite.railLength": "20",
    "SimControl.EndCondition": "Time", "SimControl.EndConditionValue": "45", "SimControl.timeStep": "0.1",
    "SimControl.TimeStepAdaptation.targetError": "0.002", "SimControl.TimeStepAdaptation.minFactor": "0.1",
    "SimControl.TimeStepAdaptation.maxFactor": "10", "SimControl.TimeStepAdaptation.minTimeStep": "0.001",
    "SimControl.TimeStepAdaptation.maxTimeStep": "5", "SimControl.TimeStepAdaptation.PID.coefficients": "-0.016 -0.004 -0.0012",
    "Rocket.Sustainer.Motor.impulseAdjustFactor": "2", "Rocket.Sustainer.Motor.burnTimeAdjustFactor": "2",
}
TIMESTEP_ADAPTATION_DEMO_EXPECTED = {"PositionX": -141.033, "Time": 45.0, "percentTolerance": 0.2}


REGRESSION_CASES = ["B4-4SimpleRocketCase", "C6-3SimpleRocketCase", "HaisunaataCase", "RollingRocketCase", "NASADraglessSphere", "NASACdSphere_RoundEarth",
                    "NASACdSphere_WGS84", "NASACdSphere_WGS84Wind", "NASADampedBrick"]


def regressionCase(name):
    """(model block, expected [x, y, z, vx, vy, vz]) of one of the reference's own flight regression cases
    (MAPLEAF/Examples/BatchSims/OpenRocketCases.mapleaf:63-214, NASAVerificationCases.mapleaf:9-262), flattened by tests/golden/make_regression_blobs.py."""
    import numpy as np
    from mapleaf_b200 import L
    z = np.load(os.path.join(GOLDEN, "regression_blobs.npz"))
    if list(z["layoutVersion"]) != [L.MLB_VERSION, L.H_SIZE, L.FS_SIZE, L.S_SIZE, L.TA_TABLE, L.CS_SIZE]:
        pytest.fail("regression_blobs.npz was flattened with another model-block layout: re-run tests/golden/make_regression_blobs.py")
    return z[name + "/blob"], z[name + "/expected"]


def regressionErrors(final, expected):
    """The reference's batch-runner check (SimulationRunners/Batch.py:488-520): percent error, absolute where the expectation is 0."""
    return [abs(a - b) / abs(b) * 100 if b != 0 else abs(a) for a, b in zip(final, expected)]
