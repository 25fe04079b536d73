"""Writes the two tables of the canard-controlled example (BASELINE.json configs[2]) into mapleaf_b200/data/TabulatedData/:
constant PID gains scheduled by (Mach, Altitude), and canard deflections that are linear in the desired moments
(F1 = 0.1 Mx - 0.1 Mz, F2 = 0.1 My - 0.1 Mz, F3 = -0.1 Mx - 0.1 Mz, F4 = -0.1 My - 0.1 Mz on a +-1e5 N m box), the same
laws as the reference's Examples/TabulatedData/constPIDCoeffs.txt and linearCanardDefls.txt."""
import itertools, os
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mapleaf_b200", "data", "TabulatedData")
with open(os.path.join(out, "constPIDCoeffs.txt"), "w") as f:
    f.write("Mach Altitude Pyz      Iyz  Dyz     Px    Ix   Dx\n")
    for ma, alt in ((0, 0), (0, 100000), (10, 0), (10, 100000)):
        f.write("{}  {}    1000.0   1 100.0   25.0  5 10.0\n".format(ma, alt))
with open(os.path.join(out, "linearCanardDefls.txt"), "w") as f:
    f.write("Ma  Alt Mx  My  Mz  F1  F2  F3  F4\n")
    moments = [(0, 0, 0)]
    for sign in (1, -1):
        for mx, my, mz in ((1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)):
            moments.append((sign * mx * 1e5, sign * my * 1e5, sign * mz * 1e5))
    for mx, my, mz in moments:
        for alt, ma in itertools.product((0, 100000), (0, 5)):
            F = (0.1 * mx - 0.1 * mz, 0.1 * my - 0.1 * mz, -0.1 * mx - 0.1 * mz, -0.1 * my - 0.1 * mz)
            f.write("\t".join(repr(float(v)) if isinstance(v, float) else str(v) for v in (ma, alt, mx, my, mz, *F)) + "\n")
