"""
Ground-wind sampling from wind roses: the host-side half of `MeanWindModel SampledGroundWindData` and of
`Hellman.groundWindModel SampledGroundWindData`.

The reference draws one mean ground wind per constructed Environment, i.e. once per Monte Carlo run
(ENV/MeanWindModelling.py:41-66,198-329): the wind roses of the sampled locations for the launch month are averaged with the location
weights, a wind speed is drawn from the speed histogram (with a tail of extreme winds appended), then a heading from the histogram of
the headings at that speed (columns interpolated linearly in speed), both with `random.random()` of the GLOBAL random module through
scipy.stats.rv_histogram.ppf; the wind blows from that heading. In a batched run those are two draws per lane, made here in the
reference's order, and the result is the lane's LANE_WIND row: the kernels never see the roses.

Data: `Windrose<Month><Location>.txt`, Iowa Environmental Mesonet layout (36 direction rows x calm + six speed classes, percent
frequency). They are looked up in mapleaf_b200/data/Wind, in `$MAPLEAF_WIND_DATA`, in an installed MAPLEAF's Examples/Wind and in ./.
"""
import math
import os
import random
from typing import List, Sequence

import numpy as np

__all__ = ["WindRoseSampler", "RadioSondeSampler", "windRosePath", "samplerFromDefinition"]

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_MPH = 0.44704
_SPEED_CLASSES_MPH = [1.0, 3.5, 6.0, 8.5, 12.5, 17.5, 23.5]            # representative speed of each column (MeanWindModelling.py:247-255)
_EXTREME_WINDS = [x * _MPH for x in (25, 31, 50)]                      # :208-211
_EXTREME_PROBABILITIES = [0.01, 0.004, 0]


def windRosePath(month: str, location: str) -> str:
    name = "Windrose{}{}.txt".format(month, location)
    cands = [os.path.join(_PKG_DIR, "data", "Wind", name)]
    if os.environ.get("MAPLEAF_WIND_DATA"):
        cands.append(os.path.join(os.environ["MAPLEAF_WIND_DATA"], name))
    try:
        import importlib.util
        spec = importlib.util.find_spec("MAPLEAF")
        if spec is not None and spec.submodule_search_locations:
            cands.append(os.path.join(list(spec.submodule_search_locations)[0], "Examples", "Wind", name))
    except Exception:
        pass
    cands.append(os.path.join("MAPLEAF", "Examples", "Wind", name))
    cands.append(name)
    for c in cands:
        if os.path.exists(c):
            return c
    raise FileNotFoundError("wind rose table {} not found (looked in {})".format(name, ", ".join(cands)))


def _readWindRose(path: str) -> np.ndarray:
    """[36, 8]: percent frequency of calm + the six speed classes, and the mid direction of the row (the first row's set to 0)."""
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.lstrip().startswith("#") and l.strip()]
    for line in lines[1:]:                                   # lines[0] is the column header
        cells = line.rstrip("\n").split(",")
        lo, hi = cells[0].strip().split("-", 1)
        vals = [float(c) if c.strip() else 0.0 for c in cells[1:8]]
        rows.append(vals + [(float(lo) + float(hi)) / 2])
    a = np.array(rows, dtype=np.float64)
    a[0, 7] = 0
    return a


class WindRoseSampler:
    """WindRoseDataSampler.sampleWindRoses (MeanWindModelling.py:213-226) for a fixed set of locations, weights and month."""
    kind = "ground"

    def __init__(self, locations: Sequence[str], weights: Sequence[float], month: str):
        from scipy.stats import rv_histogram
        self._rv = rv_histogram
        rose = np.zeros((36, 8))
        for loc, w in zip(locations, weights):
            rose = rose + _readWindRose(windRosePath(month, loc)) * 1.0 * w          # month weight 1.0, then the location weight (:262-268)
        self.rose = rose
        # speed distribution (:270-288): column sums normalised, extreme-wind tail taken out of the last class
        bySpeed = rose[:, :7].sum(axis=0)
        likelihood = list(bySpeed / bySpeed.sum())
        edges = [x * _MPH for x in (0, 2, 5, 7, 10, 15, 20, 23.5)]
        likelihood[-1] -= sum(_EXTREME_PROBABILITIES)
        self.speedDistribution = rv_histogram((likelihood + _EXTREME_PROBABILITIES, edges + _EXTREME_WINDS), density=True)      # what scipy assumes for unequal bins
        # heading tables (:290-329): row 0 (north) moves to 360, the empty calm column takes the 3.5 mph data and stands for 0 m/s
        table = np.vstack([rose[1:, :7], rose[0:1, :7]])
        table[:, 0] = table[:, 1]
        self._headingTable = table
        self._headingSpeeds = [0.0] + [c * _MPH for c in _SPEED_CLASSES_MPH[1:]]
        self._headings = [0] + [float(x) for x in rose[1:, 7]] + [360]

    def headingDistribution(self, windSpeed: float):
        speeds = self._headingSpeeds
        if windSpeed in speeds:
            likelihood = list(self._headingTable[:, speeds.index(windSpeed)])
        else:
            # linear in speed between the neighbouring classes, the nearest class beyond the ends (pandas interpolate(method="index"))
            likelihood = [float(np.interp(windSpeed, speeds, self._headingTable[i])) for i in range(self._headingTable.shape[0])]
        return self._rv((likelihood, self._headings), density=True)

    def sample(self, rng=random) -> List[float]:
        """One mean ground wind [x, y, z] (m/s): two rng.random() draws, speed then heading."""
        windSpeed = self.speedDistribution.ppf(rng.random())
        heading = self.headingDistribution(windSpeed).ppf(rng.random())
        x, y = math.sin(math.radians(heading)), math.cos(math.radians(heading))
        return [(-1 * x) * windSpeed, (-1 * y) * windSpeed, (-1 * 0) * windSpeed]


def _windDataPath(name: str) -> str:
    cands = [os.path.join(_PKG_DIR, "data", "Wind", name)]
    if os.environ.get("MAPLEAF_WIND_DATA"):
        cands.append(os.path.join(os.environ["MAPLEAF_WIND_DATA"], name))
    try:
        import importlib.util
        spec = importlib.util.find_spec("MAPLEAF")
        if spec is not None and spec.submodule_search_locations:
            cands.append(os.path.join(list(spec.submodule_search_locations)[0], "Examples", "Wind", name))
    except Exception:
        pass
    cands += [os.path.join("MAPLEAF", "Examples", "Wind", name), name]
    for c in cands:
        if os.path.exists(c):
            return c
    raise FileNotFoundError("wind data file {} not found (looked in {})".format(name, ", ".join(cands)))


_MONTHS = ["noZerothMonth", "Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"]


class RadioSondeSampler:
    """RadioSondeDataSampler.getRadioSondeWindProfile (MeanWindModelling.py:331-485): one measured wind profile per constructed Environment.
    Every draw RE-SEEDS the global generator with `SampledRadioSondeData.randomSeed` (None: from the operating system) before choosing a
    location (random.choices) and one of its soundings of the launch month (random.randint) - so with a seed every run flies the same
    profile and the run's later global draws (pink-noise seeds) repeat too, and without one nothing is repeatable; both as in the reference.
    Files: `RadioSonde<Location>_filtered.txt` (reduced IGRA-2: `#station yyyy mm dd hh ...` headers, rows of altitude m ASL, heading deg,
    speed m/s x 10)."""
    kind = "profile"

    def __init__(self, locations: Sequence[str], weights: Sequence[float], aslAltitudes: Sequence[float], month, randomSeed=None):
        self.locations, self.weights, self.aslAltitudes = list(locations), list(weights), list(aslAltitudes)
        self.month = None if month in (None, "Yearly") else month
        self.randomSeed = randomSeed
        self._files = {}
        for loc in self.locations:
            with open(_windDataPath("RadioSonde{}_filtered.txt".format(loc))) as f:
                self._files[loc] = f.readlines()
        # the longest sounding of any candidate file: per-lane wind tables are padded to this many rows
        self.maxRows = 1
        for data in self._files.values():
            n = 0
            for line in data + ["#"]:
                if line[0] == "#":
                    self.maxRows, n = max(self.maxRows, n), 0
                else:
                    n += 1

    def sample(self, rng=random):
        """(altitudes [n] in m AGL ascending, winds [n, 3] in m/s) of one sounding."""
        rng.seed(self.randomSeed)
        loc = rng.choices(self.locations, self.weights, k=1)[0]
        asl = self.aslAltitudes[self.locations.index(loc)]
        data = self._files[loc]
        monthNumber = _MONTHS.index(self.month) if self.month is not None else None
        starts = [i for i, line in enumerate(data) if line[0] == "#" and (monthNumber is None or int(line.split()[2]) == monthNumber)]
        first = starts[rng.randint(0, len(starts) - 1)] + 1
        rows = []
        i = first
        while i < len(data) and data[i][0] != "#":
            rows.append([float(x) for x in data[i].split()])
            i += 1
        a = np.array(rows)
        a[:, 2] = a[:, 2] / 10
        a = a[a[:, 0].argsort()]
        altitudes = a[:, 0] - asl
        winds = np.empty((a.shape[0], 3))
        for k in range(a.shape[0]):
            y, x = math.cos(math.radians(a[k, 1] + 0)), math.sin(math.radians(a[k, 1] + 0))
            winds[k] = ((-1 * x) * a[k, 2], (-1 * y) * a[k, 2], (-1 * 0) * a[k, 2])
        return altitudes, winds


def samplerFromDefinition(envReader):
    """The sampler an Environment of this definition would use (meanWindModelFactory, MeanWindModelling.py:43-53,96-118): a WindRoseSampler
    (kind "ground": sample() is a ground-wind vector), a RadioSondeSampler (kind "profile": sample() is an (altitudes, winds) profile), or
    None when the mean wind is not drawn per Environment."""
    model = envReader.getString("MeanWindModel")
    if model == "SampledRadioSondeData":
        words = envReader.getString("SampledRadioSondeData.locationsToSample").split()
        asl = [float(x) for x in envReader.getString("SampledRadioSondeData.locationASLAltitudes").split()]
        return RadioSondeSampler(words[0::2], [float(w) for w in words[1::2]], asl, envReader.getString("SampledRadioSondeData.launchMonth"),
                                 envReader.tryGetString("SampledRadioSondeData.randomSeed"))
    if not (model == "SampledGroundWindData" or (model == "Hellman" and envReader.getString("Hellman.groundWindModel") == "SampledGroundWindData")):
        return None
    words = envReader.getString("SampledGroundWindData.locationsToSample").split()
    return WindRoseSampler(words[0::2], [float(w) for w in words[1::2]], envReader.getString("SampledGroundWindData.launchMonth"))
