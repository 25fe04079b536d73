"""
Simulation-definition (.mapleaf) reader and Monte Carlo sampler for the B200 backend.

Host-side mirror of the reference's configuration layer for the Monte Carlo hot path
(reference: MAPLEAF/IO/simDefinition.py:143-235 parse/init, :237-305 dictionary grammar,
:471-532 resampleProbabilisticValues, :534-561 getValue, :740-780 class-based defaults;
MAPLEAF/IO/subDictReader.py:12-121). Same names, argument meaning and error behaviour
(KeyError for a missing key, ValueError for unparsable vectors / duplicate keys), written
from scratch so that the GPU box needs no MAPLEAF install.

Sampling contract: ONE `random.Random` owned by the definition,
seeded with the *string* value of MonteCarlo.randomSeed when present, normal draws in
dictionary insertion order, first draw set consumed at construction.
"""
import os
import random
import re
from typing import Dict, List, Optional

__all__ = ["SimDefinition", "SubDictReader", "defaultConfigValues", "parseVector", "formatVector", "strtobool"]

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
DATA_DIR = os.path.join(_PKG_DIR, "data")

# Default values used when a key is absent from the file. Values and key names are the
# reference's configuration contract (MAPLEAF/IO/simDefinition.py:21-129); only the keys the
# Monte Carlo trajectory path reads are listed.
defaultConfigValues: Dict[str, str] = {
    "MonteCarlo.output": "landingLocations",
    "SimControl.plot": "Position FlightAnimation",
    "SimControl.loggingLevel": "2",
    "SimControl.EndCondition": "Altitude",
    "SimControl.EndConditionValue": "-1",
    "SimControl.StageDropPaths.compute": "true",
    "SimControl.StageDropPaths.endCondition": "Altitude",
    "SimControl.StageDropPaths.endConditionValue": "0",
    "SimControl.timeDiscretization": "RK45Adaptive",
    "SimControl.timeStep": "0.01",
    "SimControl.TimeStepAdaptation.controller": "PID",
    "SimControl.TimeStepAdaptation.targetError": "0.001",
    "SimControl.TimeStepAdaptation.minFactor": "0.3",
    "SimControl.TimeStepAdaptation.maxFactor": "1.5",
    "SimControl.TimeStepAdaptation.Elementary.safetyFactor": "0.9",
    "SimControl.TimeStepAdaptation.maxTimeStep": "30",
    "SimControl.TimeStepAdaptation.minTimeStep": "0.0001",
    "SimControl.TimeStepAdaptation.PID.coefficients": "-0.01 -0.001 0",
    "SimControl.TimeStepAdaptation.eventTimingAccuracy": "0.001",
    "SimControl.RocketPlot": "Off",
    "Environment.EarthModel": "Flat",
    "Environment.AtmosphericPropertiesModel": "USStandardAtmosphere",
    "Environment.LaunchSite.elevation": "0",
    "Environment.LaunchSite.railLength": "0",
    "Environment.LaunchSite.latitude": "0",
    "Environment.LaunchSite.longitude": "0",
    "Environment.MeanWindModel": "Constant",
    "Environment.ConstantMeanWind.velocity": "(0 0 0)",
    "Environment.Hellman.alphaCoeff": "0.1429",
    "Environment.Hellman.altitudeLimit": "1000",
    "Environment.TurbulenceModel": "None",
    "Environment.turbulenceOffWhenUnderChute": "True",
    "Environment.ConstantAtmosphere.temp": "15",
    "Environment.ConstantAtmosphere.pressure": "101325",
    "Environment.ConstantAtmosphere.density": "1.225",
    "Environment.ConstantAtmosphere.viscosity": "1.789e-5",
    "Rocket.ControlSystem.desiredFlightDirection": "(0 0 1)",
    "Rocket.ControlSystem.MomentController.Type": "ScheduledGainPIDRocket",
    "Rocket.ControlSystem.updateRate": "0",
    "Rocket.name": "Rocket",
    "Rocket.position": "(0 0 10)",
    "Rocket.initialDirection": "(0 0 1)",
    "Rocket.velocity": "(0 0 0)",
    "Rocket.angularVelocity": "(0 0 0)",
    "Rocket.Aero.fullyTurbulentBL": "true",
    "Rocket.Aero.addZeroLengthBoatTailsToAccountForBaseDrag": "true",
    "Rocket.Aero.surfaceRoughness": "0.000005",
    "Stage.stageNumber": "0",
    "Stage.separationTriggerType": "None",
    "Stage.separationTriggerValue": "0",
    "Stage.separationDelay": "0",
    "Stage.position": "(0 0 0)",
    "AeroForce.Lref": "0",
    "AeroForce.Cd": "0",
    "AeroForce.Cl": "0",
    "AeroForce.momentCoeffs": "(0 0 0)",
    "AeroDamping.zDampingCoeffs": "(0 0 0)",
    "AeroDamping.yDampingCoeffs": "(0 0 0)",
    "AeroDamping.xDampingCoeffs": "(0 0 0)",
    "FinSet.finCantAngle": "0",
    "FinSet.firstFinAngle": "0",
    "FinSet.LeadingEdge.shape": "Round",
    "FinSet.TrailingEdge.shape": "Tapered",
    "FinSet.numFinSpanSlicesForIntegration": "10",
    "Nosecone.shape": "tangentOgive",
    "BoatTail.shape": "cone",
    "Mass.cg": "(0 0 0)",
    "Motor.impulseAdjustFactor": "1.0",
    "Motor.burnTimeAdjustFactor": "1.0",
    "Actuator.controller": "TableInterpolating",
    "Actuator.responseModel": "FirstOrder",
    "Actuator.responseTime": "0.1",
    "RecoverySystem.cg": "(0 0 0)",
    "TabulatedAeroForce.Lref": "0",
}


def strtobool(val: str) -> bool:
    """distutils.util.strtobool semantics (reference: subDictReader.py:4,34)."""
    v = str(val).lower()
    if v in ("y", "yes", "t", "true", "on", "1"):
        return True
    if v in ("n", "no", "f", "false", "off", "0"):
        return False
    raise ValueError("invalid truth value %r" % (val,))


def parseVector(text: str):
    """'(x y z)' / '[x, y, z]' / 'x; y; z' -> (x, y, z) floats. Reference: Motion/CythonVector.pyx:26-49."""
    orig = text
    X = text.strip()
    if X and X[0] in "[({":
        X = X[1:-1].strip()
    for sep in (", ", "; ", None, ",", ";"):
        parts = X.split(sep) if sep is not None else X.split()
        if len(parts) != 1:
            break
    if len(parts) == 3:
        return (float(parts[0]), float(parts[1]), float(parts[2]))
    raise ValueError("'" + orig + "' Was not successfully parsed into Vector format")


def formatVector(v) -> str:
    """str(Vector): '({} {} {})' with float repr (CythonVector.pyx:66-68)."""
    return "({} {} {})".format(float(v[0]), float(v[1]), float(v[2]))


def _looksLikeFileName(val: str) -> bool:
    return bool(re.search(r"\.[A-Za-z0-9]{1,8}$", val)) and (" " not in val.strip())


class SimDefinition:
    """Flat `dict[str,str]` view of a .mapleaf file with defaults and Monte Carlo sampling."""

    def __init__(self, fileName: Optional[str] = None, dictionary: Optional[Dict[str, str]] = None,
                 disableDistributionSampling: bool = False, silent: bool = False, defaultDict=None):
        self.silent = silent
        self.fileName = fileName
        self.disableDistributionSampling = disableDistributionSampling
        self.defaultDict = defaultConfigValues if defaultDict is None else defaultDict
        self.monteCarloLogger = None
        if dictionary is not None:
            self.dict = dictionary
        elif fileName is not None:
            self.dict = self._parseSimDefinitionFile(fileName)
        else:
            raise ValueError("No fileName or dictionary provided to initialize the SimDefinition")

        if not disableDistributionSampling:
            try:
                randomSeed = self.getValue("MonteCarlo.randomSeed")   # a *string* -> CPython str seeding
            except KeyError:
                randomSeed = random.randrange(1000000)
            containsProbabilisticValues = any(k.endswith("_stdDev") for k in self.dict)
            if not silent and containsProbabilisticValues:
                print("Monte Carlo random seed: {}".format(randomSeed))
            self.randomSeed = randomSeed
            self.rng = random.Random(randomSeed)
            self.resampleProbabilisticValues()   # first draw set consumed here (simDefinition.py:216)

    # ---------------------------------------------------------------- parsing
    def _parseSimDefinitionFile(self, fileName: str) -> Dict[str, str]:
        with open(fileName, "r") as f:
            text = f.read()
        text = re.sub(r"(?<!\\)#.*", "", text)      # comments
        text = re.sub(r"\\(?=#)", "", text)         # escaped '#'
        lines = [ln for ln in text.split("\n") if ln.strip() != ""]
        out: Dict[str, str] = {}
        self._parseDictionaryContents(out, lines, 0, "")
        self._resolveFilePaths(out)
        return out

    def _parseDictionaryContents(self, out, lines, start, curr):
        i = start
        while i < len(lines):
            line = lines[i].strip()
            parts = line.split()
            if parts[0] == "!include":
                sub = SimDefinition(self._findFile(line[line.index(" "):].strip()), disableDistributionSampling=True, silent=True)
                for k, v in sub.dict.items():
                    out[k if curr == "" else curr + "." + k] = v
            elif parts[0].startswith("!"):
                raise NotImplementedError("Sim-definition directive '{}' is not supported by the B200 backend".format(parts[0]))
            elif line[-1] == "{":
                name = line[:-1].strip()
                i = self._parseDictionaryContents(out, lines, i + 1, name if curr == "" else curr + "." + name)
            elif line == "}":
                return i
            elif len(parts) > 1:
                key = parts[0] if curr == "" else curr + "." + parts[0]
                if key in out:
                    raise ValueError("Duplicate Key: " + key + " in File: " + str(self.fileName))
                out[key] = " ".join(parts[1:])
            else:
                raise ValueError("Problem parsing line {}: {}".format(i, line))
            i += 1
        return i

    def _findFile(self, val: str) -> Optional[str]:
        """Resolve a path relative to the definition file, the package data dir, a MAPLEAF install or cwd."""
        cands = []
        if os.path.isabs(val):
            cands.append(val)
        if self.fileName is not None:
            cands.append(os.path.join(os.path.dirname(os.path.realpath(self.fileName)), val))
        v = val[2:] if val.startswith("./") else val
        cands.append(os.path.join(DATA_DIR, v))
        if v.startswith("MAPLEAF/Examples/"):
            cands.append(os.path.join(DATA_DIR, v[len("MAPLEAF/Examples/"):]))
        root = os.environ.get("MAPLEAF_ROOT")
        if root:
            cands.append(os.path.join(root, v))
        cands.append(os.path.abspath(v))
        for c in cands:
            if os.path.exists(c):
                return c
        return None

    def _resolveFilePaths(self, d: Dict[str, str]) -> None:
        for k, v in list(d.items()):
            if _looksLikeFileName(v) and ("/" in v or k.endswith("path") or k.endswith("Path")):
                found = self._findFile(v)
                if found is not None:
                    d[k] = found

    # ---------------------------------------------------------------- sampling
    def resampleProbabilisticValues(self, Dict=None) -> None:
        """For each `key` with a `key_stdDev`: draw N(mean, std) (scalar or 3-vector) from self.rng."""
        if Dict is None:
            Dict = self.dict
        if self.disableDistributionSampling:
            return
        for key in list(Dict.keys()):
            stdDevKey = key + "_stdDev"
            if stdDevKey not in Dict:
                continue
            meanKey = key + "_mean"
            try:
                meanString = Dict[meanKey]
            except KeyError:
                meanString = Dict[key]
                Dict[meanKey] = meanString
            try:
                mu = float(meanString)
                sigma = float(Dict[stdDevKey])
                sampledValue = self.rng.gauss(mu, sigma)
                Dict[key] = str(sampledValue)
                logLine = "Sampling scalar parameter: {}, value: {:1.3f}".format(key, sampledValue)
            except ValueError:
                try:
                    muVec = parseVector(meanString)
                    sigmaVec = parseVector(Dict[stdDevKey])
                except ValueError:
                    print("ERROR: Unable to parse probabilistic value: {} for key {} (or {} for key {}). Note that probabilistic values must be either scalars or vectors of length 3.".format(
                        meanString, meanKey, Dict[stdDevKey], stdDevKey))
                    raise
                sampledVec = [self.rng.gauss(m, s) for m, s in zip(muVec, sigmaVec)]
                Dict[key] = formatVector(sampledVec)
                logLine = "Sampling vector parameter: {}, value: ({:1.3f} {:1.3f} {:1.3f})".format(key, *sampledVec)
            if self.monteCarloLogger is not None:
                self.monteCarloLogger.log(logLine)
            elif not self.silent:
                print(logLine)

    def probabilisticKeys(self) -> List[str]:
        """Keys that are re-drawn by every resample, in draw order."""
        return [k for k in self.dict if (k + "_stdDev") in self.dict]

    # ---------------------------------------------------------------- access
    def getValue(self, key: str) -> str:
        key = key.strip()
        if key in self.dict:
            return self.dict[key]
        if key in self.defaultDict:
            return self.defaultDict[key]
        v = self._getClassBasedDefaultValue(key)
        if v is not None:
            return v
        raise KeyError("Key: " + key + " not found in {} or default config values".format(self.fileName))

    def setValue(self, key: str, value) -> None:
        self.dict[key.strip()] = value

    def removeKey(self, key: str):
        return self.dict.pop(key, None)

    def writeToFile(self, fileName: str, writeHeader=True) -> None:
        """Writes the (possibly modified) definition in the .mapleaf format, keys sorted so that dictionaries stay together, tab-indented,
        without comments: the layout of the reference's writer (IO/simDefinition.py:582-645), which its optimisers use for continuation files."""
        from datetime import datetime
        self.fileName = fileName
        lines = []
        if writeHeader:
            lines += ["# MAPLEAF", "# File: {}".format(fileName), "# Autowritten on: " + str(datetime.now())]
        open_ = []                                           # dictionaries currently open, outermost first
        for key in sorted(self.dict):
            *dicts, name = key.split(".")
            if dicts != open_:
                common = 0
                while common < min(len(open_), len(dicts)) and open_[common] == dicts[common]:
                    common += 1
                for depth in range(len(open_), common, -1):
                    lines.append("\t" * (depth - 1) + "}")
                if common == len(dicts):
                    lines.append("")                         # back in an enclosing dictionary: a spacing line before its next keys
                for depth in range(common, len(dicts)):
                    lines.append("")
                    lines.append("\t" * depth + dicts[depth] + "{")
                open_ = dicts
            lines.append("\t" * len(open_) + name + "\t" + str(self.dict[key]))
        for depth in range(len(open_), 0, -1):
            lines.append("\t" * (depth - 1) + "}")
        with open(fileName, "w") as f:
            f.write("\n".join(lines) + "\n")

    def setIfAbsent(self, key: str, value) -> None:
        if key not in self.dict:
            self.setValue(key, value)

    def __contains__(self, key) -> bool:
        try:
            self.getValue(key)
            return True
        except KeyError:
            return False

    def _getClassBasedDefaultValue(self, key: str) -> Optional[str]:
        """'Rocket.S.fins.LeadingEdge.shape' -> class of 'Rocket.S.fins' is FinSet -> default 'FinSet.LeadingEdge.shape'."""
        parts = key.split(".")
        for level in range(len(parts) - 1, 0, -1):
            prefix = ".".join(parts[:level])
            suffix = ".".join(parts[level:])
            className = self.dict.get(prefix + ".class")
            if className is not None:
                return self.defaultDict.get(className + "." + suffix)
        return None

    def getSubKeys(self, key: str) -> List[str]:
        p = key + "."
        return [k for k in self.dict if k.startswith(p)]

    def getImmediateSubKeys(self, key: str) -> List[str]:
        p = key + "."
        n = len(p)
        return [k for k in self.dict if k.startswith(p) and "." not in k[n:]]

    def getImmediateSubDicts(self, key: str) -> List[str]:
        """Immediate sub-dictionaries, in first-appearance order (the reference returns an unordered set;
        every consumer on the path re-sorts, see Rocket/stage.py:47-59)."""
        p = key + "."
        n = len(p)
        seen = []
        for k in self.dict:
            if k.startswith(p):
                rest = k[n:]
                if "." in rest:
                    name = p + rest.split(".")[0]
                    if name not in seen:
                        seen.append(name)
        return seen


class SubDictReader:
    """Typed, path-relative access to a SimDefinition (reference: IO/subDictReader.py:12-121)."""

    def __init__(self, stringPathToThisItemsSubDictionary: str, simDefinition: SimDefinition):
        self.simDefDictPathToReadFrom = stringPathToThisItemsSubDictionary
        self.simDefinition = simDefinition

    def getString(self, key: str) -> str:
        try:
            return self.simDefinition.getValue(self.simDefDictPathToReadFrom + "." + key)
        except KeyError:
            try:
                return self.simDefinition.getValue(key)
            except KeyError:
                raise KeyError("{} and {} not found in {} or in default value dictionary".format(
                    self.simDefDictPathToReadFrom + "." + key, key, self.simDefinition.fileName))

    def getInt(self, key):
        return int(self.getString(key))

    def getFloat(self, key):
        return float(self.getString(key))

    def getVector(self, key):
        return parseVector(self.getString(key))

    def getBool(self, key):
        return strtobool(self.getString(key))

    def _try(self, fn, key, defaultValue):
        try:
            return fn(self.getString(key))
        except KeyError:
            return defaultValue

    def tryGetString(self, key, defaultValue=None):
        return self._try(str, key, defaultValue)

    def tryGetInt(self, key, defaultValue=None):
        return self._try(int, key, defaultValue)

    def tryGetFloat(self, key, defaultValue=None):
        return self._try(float, key, defaultValue)

    def tryGetVector(self, key, defaultValue=None):
        return self._try(parseVector, key, defaultValue)

    def tryGetBool(self, key, defaultValue=None):
        return self._try(strtobool, key, defaultValue)

    def getImmediateSubDicts(self, key=None) -> List[str]:
        if key is None:
            key = self.simDefDictPathToReadFrom
        else:
            result = self.simDefinition.getImmediateSubDicts(self.simDefDictPathToReadFrom + "." + key)
            if len(result) != 0:
                return result
        return self.simDefinition.getImmediateSubDicts(key)

    def _relativeThenAbsolute(self, fn, key):
        if key is None:
            key = self.simDefDictPathToReadFrom
        else:
            result = fn(self.simDefDictPathToReadFrom + "." + key)
            if len(result) != 0:
                return result
        return fn(key)

    def getImmediateSubKeys(self, key=None) -> List[str]:
        return self._relativeThenAbsolute(self.simDefinition.getImmediateSubKeys, key)

    def getSubKeys(self, key=None) -> List[str]:
        return self._relativeThenAbsolute(self.simDefinition.getSubKeys, key)

    def getDictName(self) -> str:
        return self.simDefDictPathToReadFrom[self.simDefDictPathToReadFrom.rfind(".") + 1:]
