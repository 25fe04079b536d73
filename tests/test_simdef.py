"""Host configuration layer: parser, defaults, Monte Carlo sampling (reference: IO/simDefinition.py, IO/subDictReader.py;
reference tests mirrored: test/test_IO/test_SimDefinition.py:112-151)."""
import os
import random

import pytest

from conftest import DATA
from mapleaf_b200.simdef import SimDefinition, SubDictReader, formatVector, parseVector


def loadMC(**kw):
    return SimDefinition(os.path.join(DATA, "MonteCarlo.mapleaf"), silent=True, **kw)


def test_parse_nested_keys_and_comments():
    sd = loadMC(disableDistributionSampling=True)
    assert sd.getValue("Rocket.Sustainer.TailFins.numFins") == "4"
    assert sd.getValue("Environment.ConstantMeanWind.velocity") == "( 2 0 0 )"
    assert sd.getValue("Rocket.Sustainer.Motor.path").endswith(os.path.join("Motors", "test.txt"))
    assert os.path.isabs(sd.getValue("Rocket.Sustainer.Motor.path"))


def test_defaults_and_class_based_defaults():
    sd = loadMC(disableDistributionSampling=True)
    assert sd.getValue("Environment.EarthModel") == "Flat"                       # plain default
    assert sd.getValue("Rocket.Sustainer.TailFins.finCantAngle") == "0"          # FinSet.finCantAngle
    assert sd.getValue("Rocket.Sustainer.TailFins.LeadingEdge.shape") == "Round"
    assert sd.getValue("Rocket.Sustainer.stageNumber") == "0"                    # Stage.stageNumber
    with pytest.raises(KeyError):
        sd.getValue("Rocket.Sustainer.TailFins.noSuchKey")


def test_subdictreader_relative_then_absolute():
    sd = loadMC(disableDistributionSampling=True)
    r = SubDictReader("Rocket", sd)
    assert r.getString("SimControl.timeDiscretization") == "RK4"
    assert r.getVector("position") == (0.0, 0.0, 15.0)
    assert r.tryGetVector("rotationAxis", defaultValue=None) is None
    assert r.getBool("Aero.fullyTurbulentBL") is True
    assert r.getImmediateSubDicts() == ["Rocket.Sustainer"]


def test_duplicate_key_raises(tmp_path):
    p = tmp_path / "dup.mapleaf"
    p.write_text("A{\n  x 1\n  x 2\n}\n")
    with pytest.raises(ValueError):
        SimDefinition(str(p), silent=True)


def test_vector_parse_formats():
    for s in ("(1 2 3)", "[1, 2, 3]", "{1;2;3}", "1; 2; 3", "1,2,3"):
        assert parseVector(s) == (1.0, 2.0, 3.0)
    with pytest.raises(ValueError):
        parseVector("(1 2)")
    assert formatVector((1, 2.5, -3)) == "(1.0 2.5 -3.0)"


def test_sampling_seed1000_reference_vectors(golden):
    """test_SimDefinition.py:118-151: rng.seed(1000) -> 10.2546..., (10.2546..., 8.0664..., 14.2736...)."""
    g = golden("rng.json")
    sd = SimDefinition(dictionary={"scalar1": "10", "scalar1_stdDev": "1"}, silent=True)
    sd.rng.seed(1000)
    sd.resampleProbabilisticValues()
    assert float(sd.getValue("scalar1")) == 10.254632116649685
    sd = SimDefinition(dictionary={"vector1": "(10, 11, 12)", "vector1_stdDev": "(1, 2, 3)"}, silent=True)
    sd.rng.seed(1000)
    sd.resampleProbabilisticValues()
    assert list(parseVector(sd.getValue("vector1"))) == [10.254632116649685, 8.066448705920658, 14.27360999981685]
    assert list(parseVector(sd.getValue("vector1"))) == g["seed1000_vector"]
    assert sd.getValue("vector1_mean") == "(10, 11, 12)"


def test_string_seed_stream_matches_reference_run(golden):
    """MonteCarlo.randomSeed is a *string* seed; draw set 1 is consumed by the constructor (App. A #10)."""
    g = golden("rk4_apogee.json")
    sd = loadMC(disableDistributionSampling=True)
    sd.setValue("MonteCarlo.randomSeed", g["randomSeed"])
    sd = SimDefinition(dictionary=dict(sd.dict), silent=True)
    for run in g["runs"]:
        sd.resampleProbabilisticValues()
        assert list(parseVector(sd.getValue("Environment.ConstantMeanWind.velocity"))) == run["wind"]


def test_sampling_log_lines():
    class Log:
        lines = []

        def log(self, s):
            self.lines.append(s)
    sd = SimDefinition(dictionary={"a.v": "(1 2 3)", "a.v_stdDev": "(1 1 1)", "b": "5", "b_stdDev": "0.1", "MonteCarlo.randomSeed": "7"}, silent=True)
    sd.monteCarloLogger = Log()
    sd.resampleProbabilisticValues()
    assert sd.monteCarloLogger.lines[0].startswith("Sampling vector parameter: a.v, value: (")
    assert sd.monteCarloLogger.lines[1].startswith("Sampling scalar parameter: b, value: ")
    r = random.Random("7")
    [r.gauss(0, 1) for _ in range(4)]       # constructor draw set
    expect = [r.gauss(m, 1) for m in (1, 2, 3)]
    assert list(parseVector(sd.getValue("a.v"))) == expect
