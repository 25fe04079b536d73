"""Writes the environment tables used by the option-coverage fixtures (inputs made for this repository, reference file formats):
   data/Environment/atmosphereTable.txt ... TabulatedAtmosphere (ENV/AtmosphereModelling.py:79-96): h [m ASL], T [K], P [Pa], rho [kg/m3], mu [1e-5 Pa s]
   data/Environment/windProfile.txt ....... CustomWindProfile / InterpolatedWind (ENV/MeanWindModelling.py:165-196): AGL altitude, wind x y z [m/s]
   data/Wind/WindroseApr<TestField>.txt ... two synthetic wind roses for SampledGroundWindData (ENV/MeanWindModelling.py:198-329)
   data/Wind/RadioSonde<TestField>_filtered.txt  two synthetic radio-sonde archives for SampledRadioSondeData (:331-485)"""
import math, os
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(REPO, "mapleaf_b200", "data", "Environment")
os.makedirs(out, exist_ok=True)
# a smooth, slightly non-standard atmosphere (warm day, 1.5 % denser): isothermal above 11.5 km
rows = []
for h in [-500, 0, 250, 500, 900, 1400, 2000, 2750, 3500, 4500, 6000, 7500, 9000, 11500, 14000, 18000, 24000, 32000]:
    T = 293.65 - 0.0064 * min(h, 11500)
    P = 101900.0 * (T / 293.65) ** 5.34 if h <= 11500 else 101900.0 * ((293.65 - 0.0064 * 11500) / 293.65) ** 5.34 * math.exp(-(h - 11500) / 6440.0)
    rho = 1.015 * P / (287.05 * T)
    mu = 1.458e-6 * T ** 1.5 / (T + 110.4) * 1e5
    rows.append((h, T, P, rho, mu))
with open(os.path.join(out, "atmosphereTable.txt"), "w") as f:
    f.write("h\tT\tP\trho\tmu\n")
    for r in rows:
        f.write("%d\t%.3f\t%.1f\t%.5f\t%.4f\n" % r)
# a veering, strengthening wind with a low-level jet and a small updraft band
with open(os.path.join(out, "windProfile.txt"), "w") as f:
    f.write("AGLAltitude(m) WindX(m/s) WindY(m/s) WindZ(m/s)\n")
    for h in [0, 50, 150, 300, 600, 1000, 1600, 2500, 4000, 6000, 9000]:
        speed = 2.5 + 9.0 * (1 - math.exp(-h / 900.0)) + 3.0 * math.exp(-((h - 450) / 200.0) ** 2)
        ang = math.radians(15 + 40 * (1 - math.exp(-h / 1500.0)))
        f.write("%d %.3f %.3f %.3f\n" % (h, speed * math.cos(ang), speed * math.sin(ang), 0.4 * math.exp(-((h - 800) / 400.0) ** 2)))


def rose(name, month, peakDir, spread, calm, scale):
    """A synthetic wind-rose table in the Iowa Environmental Mesonet layout the reference parses (ENV/MeanWindModelling.py:228-260): 36
    direction rows, calm (first row only) + six speed classes in percent frequency, one preferred direction."""
    wdir = os.path.join(REPO, "mapleaf_b200", "data", "Wind")
    os.makedirs(wdir, exist_ok=True)
    lines = ["# Windrose Data Table (Percent Frequency) for %s (synthetic table written for mapleaf_b200's tests, Iowa Environmental Mesonet layout)" % name,
             "# Wind Speed Units: miles per hour", "# First value in table is CALM",
             "Direction,     Calm, 2.0  4.9, 5.0  6.9, 7.0  9.9,10.0 14.9,15.0 19.9,    20.0+"]
    bins = [(0, 10)] + [(10 * (i - 1), 10 * (i + 1)) for i in range(1, 36)]
    for k, (a, b) in enumerate(bins):
        mid = (a + b) / 2
        d = min(abs(mid - peakDir), 360 - abs(mid - peakDir))
        w = math.exp(-(d / spread) ** 2) + 0.15
        vals = [w * scale * f for f in (0.55, 0.45, 0.62, 0.9, 0.5, 0.22)]
        first = "%9.2f" % calm if k == 0 else "         "
        lines.append("%03d-%03d  ,%s,%s" % (a, b, first, ",".join("%9.3f" % v for v in vals)))
    with open(os.path.join(wdir, "Windrose%s%s.txt" % (month, name)), "w") as f:
        f.write("\n".join(lines) + "\n")




def radioSonde(name, elevation, nSoundings=9):
    """A synthetic radio-sonde archive in the reduced IGRA-2 layout the reference reads (ENV/MeanWindModelling.py:386-485): soundings headed
    by `#<station> yyyy mm dd hh ...`, rows of altitude (m ASL), wind heading (deg), wind speed (m/s x 10); different lengths on purpose."""
    wdir = os.path.join(REPO, "mapleaf_b200", "data", "Wind")
    lines = []
    for k in range(nSoundings):
        month = 4 if k % 3 else 7
        lines.append("#SYN%08d 2019 %02d %02d %02d 2315   83 synthetic synthetic  510000 -1120000" % (k, month, 3 + 2 * k, 12 * (k % 2)))
        nLevels = 9 + (k * 5) % 7
        for j in range(nLevels):
            alt = elevation + 30 + 900 * j + 37 * ((k * 7 + j * 3) % 11)
            heading = (200 + 12 * j + 25 * k) % 360
            speed = 40 + 28 * j + 13 * ((k + j) % 5)
            lines.append("%5d   %3d   %3d" % (alt, heading, speed))
    with open(os.path.join(wdir, "RadioSonde%s_filtered.txt" % name), "w") as f:
        f.write("\n".join(lines) + "\n")


radioSonde("TestFieldA", 710)
radioSonde("TestFieldB", 638, nSoundings=7)
rose("TestFieldA", "Apr", 240, 60, 2.1, 0.55)
rose("TestFieldB", "Apr", 300, 35, 0.8, 0.6)
print("wrote", os.listdir(out), os.listdir(os.path.join(REPO, "mapleaf_b200", "data", "Wind")))
