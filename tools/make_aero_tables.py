"""Writes the multi-key TabulatedAeroForce test tables under mapleaf_b200/data/TabulatedData (scattered key points, so that the Delaunay
triangulation is not a regular grid; some flight conditions of TabulatedComponents2D.mapleaf fall outside the hull and exercise the
nearest-neighbour fallback). Deterministic:  python tools/make_aero_tables.py"""
import os
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(REPO, "mapleaf_b200", "data", "TabulatedData")
rng = np.random.default_rng(20261020)

# 2 keys: Mach, TotalAOA -> CD, CL
mach = np.concatenate([[0, 0, 0.9, 0.9, 0.3], rng.uniform(0, 0.9, 19)])
aoa = np.concatenate([[0, 1.2, 0, 1.2, 0.1], rng.uniform(0, 1.2, 19)])
cd = 0.3 + 0.25 * mach ** 2 + 0.9 * np.sin(aoa) ** 2
cl = 1.8 * np.sin(aoa) * np.cos(aoa) * (1 + 0.2 * mach)
with open(os.path.join(out, "TestAeroForces2D.csv"), "w") as f:
    f.write("Mach,TotalAOA,CD,CL\n")
    for row in zip(mach, aoa, cd, cl):
        f.write(",".join("{:.6f}".format(x) for x in row) + "\n")

# 3 keys: Mach, AOA, AOSS -> CMx, CMy, CMz (AOA and AOSS mutate the cached air velocity in column order)
m3 = np.concatenate([[0, 0, 0, 0, 1, 1, 1, 1], rng.uniform(0, 1, 22)])
a3 = np.concatenate([[-1, -1, 1, 1, -1, -1, 1, 1], rng.uniform(-1, 1, 22)])
s3 = np.concatenate([[-1, 1, -1, 1, -1, 1, -1, 1], rng.uniform(-1, 1, 22)])
cmx = 0.002 * a3 * s3
cmy = -0.6 * np.sin(a3) * (1 + 0.3 * m3)
cmz = 0.6 * np.sin(s3) * (1 + 0.3 * m3)
with open(os.path.join(out, "TestAeroMoments3D.csv"), "w") as f:
    f.write("Mach,AOA,AOSS,CMx,CMy,CMz\n")
    for row in zip(m3, a3, s3, cmx, cmy, cmz):
        f.write(",".join("{:.6f}".format(x) for x in row) + "\n")
print("wrote TestAeroForces2D.csv (24 points), TestAeroMoments3D.csv (30 points)")
