#!/usr/bin/env python
"""
Flattens the reference's own flight regression cases (MAPLEAF/Examples/BatchSims/OpenRocketCases.mapleaf:63-214: B4-4 / C6-3
SimpleRocketCase, HaisunaataCase, RollingRocketCase; NASAVerificationCases.mapleaf:9-262: sphere and brick cases) into model blocks with this repo's builder and stores them, with the final values the
reference records for them, in tests/golden/regression_blobs.npz. The input decks themselves (Jake1.mapleaf, Haisunaata.mapleaf, the Estes
motor tables) stay in /root/reference: what is committed is the derived model block in this repo's own layout (include/mlb_layout.h).
Re-run after a layout change (the tests skip blobs whose H_VERSION / length no longer match):

    python tests/golden/make_regression_blobs.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
REF = "/root/reference"
os.chdir(REF)          # the decks name their motor files relative to the reference's root

from mapleaf_b200 import L, SimDefinition, buildModel  # noqa: E402

J = os.path.join(REF, "MAPLEAF/Examples/Simulations/Jake1.mapleaf")
H = os.path.join(REF, "MAPLEAF/Examples/Simulations/Haisunaata.mapleaf")
C63 = {"Rocket.Sustainer.GeneralMass.mass": "0.00952", "Rocket.Sustainer.Motor.path": "MAPLEAF/Examples/Motors/C6-3.txt",
       "Rocket.Sustainer.Motor.cg": "(4.08e-06 4.08e-06 3.86e-07)"}
CASES = {      # name: (deck, ParameterOverrides, ExpectedFinalValues: position xyz, velocity xyz)
    "B4-4SimpleRocketCase": (J, {"Rocket.Sustainer.GeneralMass.mass": "0.0129", "Rocket.Sustainer.Motor.path": "MAPLEAF/Examples/Motors/B4-4.txt",
                                 "Rocket.Sustainer.Motor.cg": "(5.53E-06 5.53E-06 5.22E-07)", "Rocket.Sustainer.TailFins.finCantAngle": "0",
                                 "SimControl.EndCondition": "Apogee"}, [-1.275, 0.0, 67.818, -0.0714, 0.0, -0.0007425]),
    "C6-3SimpleRocketCase": (J, {**C63, "Rocket.Sustainer.Motor.burnTimeAdjustFactor": "1.16", "Rocket.Sustainer.TailFins.finCantAngle": "0",
                                 "SimControl.EndCondition": "Apogee"}, [1.8459999999999999, 0.0, 156.835, 0.4802, 0.0, -0.0465]),
    "HaisunaataCase": (H, {}, [-138.064, 0.0, 859.154, -9.1115, 0.0, -0.015156]),
    "RollingRocketCase": (J, {**C63, "Rocket.Sustainer.Motor.burnTimeAdjustFactor": "1.15", "Rocket.Sustainer.TailFins.finCantAngle": "5",
                              "Rocket.Sustainer.AltimeterMass.mass": "0.03", "SimControl.EndCondition": "Time", "SimControl.EndConditionValue": "6",
                              "SimControl.timeStep": "0.005"}, [-9.898, 1.51, 123.546, -1.6056, 0.2942, -0.3942]),
}
# NASA verification cases (MAPLEAF/Examples/BatchSims/NASAVerificationCases.mapleaf:9-262, from NASA/TM-2015-218675 check cases): a sphere
# dropped from 9144 m over the round / WGS84 earth without and with drag and wind, a tumbling brick with aerodynamic damping; 30 s
S = os.path.join(REF, "MAPLEAF/Examples/Simulations/NASASphere.mapleaf")
B = os.path.join(REF, "MAPLEAF/Examples/Simulations/NASABrick.mapleaf")
CASES.update({
    "NASADraglessSphere": (S, {"Rocket.Stage.AeroForce.Cd": "0.0"}, [6382876.054, 13969.82281, 0.0, -293.7167128, 465.4464411, 0.0]),
    "NASACdSphere_RoundEarth": (S, {"Environment.EarthModel": "Round"}, [6375952.953, 13954.22714, 0.0, -264.5117166, 464.9271675, 0.0]),
    "NASACdSphere_WGS84": (S, {}, [6383085.212, 13969.82655, 0.0, -264.3698185, 465.4472786, 0.0]),
    "NASACdSphere_WGS84Wind": (S, {"Environment.MeanWindModel": "Constant", "Environment.ConstantMeanWind.velocity": "(6.096 0 0)"},
                               [6383085.4118580, 13978.2005101, 0.0, -264.3589523, 466.3203626, 0.0]),
    "NASADampedBrick": (B, {}, [6382876.054, 13969.82281, 0.0, -293.7167128, 465.4464411, 0.0]),
})
out = {}
for name, (deck, overrides, expected) in CASES.items():
    sd = SimDefinition(deck, silent=True, disableDistributionSampling=True)
    for k, v in overrides.items():
        sd.setValue(k, v)
    m = buildModel(sd)
    out[name + "/blob"] = m.blob
    out[name + "/expected"] = np.array(expected)
    print(name, "blob doubles", m.blob.size)
out["layoutVersion"] = np.array([L.MLB_VERSION, L.H_SIZE, L.FS_SIZE, L.S_SIZE, L.TA_TABLE, L.CS_SIZE])
np.savez_compressed(os.path.join(HERE, "regression_blobs.npz"), **out)
print("wrote regression_blobs.npz", os.path.getsize(os.path.join(HERE, "regression_blobs.npz")), "bytes")
