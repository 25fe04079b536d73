"""Model block vs attributes of the reference's constructed objects (tests/golden/model_montecarlo.json). Bit-exact."""
import os

import numpy as np
import pytest

from conftest import DATA
from mapleaf_b200 import L, SimDefinition, buildModel


@pytest.fixture(scope="module")
def model():
    sd = SimDefinition(os.path.join(DATA, "MonteCarlo.mapleaf"), silent=True, disableDistributionSampling=True)
    sd.setValue("Environment.ConstantMeanWind.velocity", "(2 0 0)")
    return buildModel(sd)


def eq(a, b):
    return np.array_equal(np.atleast_1d(np.asarray(a, dtype=float)), np.atleast_1d(np.asarray(b, dtype=float)))


def test_header_fields(model, golden):
    g = golden("model_montecarlo.json")
    assert model.header("H_MAGIC") == L.MLB_MAGIC and model.header("H_LEN") == model.blob.size
    assert eq(model.header("H_AREF"), g["Aref"])
    assert eq(model.header("H_FINENESS"), g["fineness"])
    assert eq(model.blob[L.H_INIT_STATE:L.H_INIT_STATE + 13], g["init"])
    off = int(model.header("H_USSTD_OFF"))
    assert eq(model.blob[off + 16:off + 24], g["usstdT"]) and eq(model.blob[off + 24:off + 32], g["usstdP"])
    assert model.header("H_INTEGRATOR") == L.INT_RK4 and model.header("H_END_TYPE") == L.END_APOGEE
    mi = model.meta["initialInertia"]
    assert eq([*mi.MOI, *mi.CG, mi.mass], g["inertia0"])


def test_stage_record(model, golden):
    g = golden("model_montecarlo.json")
    S = model.stageRecord(0)
    assert eq(S[L.S_FIXED_MOI:L.S_FIXED_MOI + 3], g["fixedMOI"]) and eq(S[L.S_FIXED_CG:L.S_FIXED_CG + 3], g["fixedCG"])
    assert eq(S[L.S_FIXED_MASS], g["fixedMass"])
    assert int(S[L.S_OVR_FLAGS]) == (L.OVR_CG | L.OVR_MASS | L.OVR_MOI)
    assert eq(S[L.S_OVR_CG:L.S_OVR_CG + 3], [0, 0, -2.65]) and S[L.S_OVR_MASS] == 50 and eq(S[L.S_OVR_MOI:L.S_OVR_MOI + 3], [85, 85, 0.5])
    assert S[L.S_BURN_DURATION] == 5.0 and S[L.S_IGNITION0] == 0.0


def test_component_order_and_records(model, golden):
    """Evaluation order and every precomputed constant."""
    g = golden("model_montecarlo.json")["components"]
    comps = model.components(0)
    assert [c["cls"] for c in g] == ["TabulatedMotor", "NoseCone", "BodyTube", "RecoverySystem", "FixedMass", "FinSet", "BoatTail"]
    assert [int(c[0]) for c in comps] == [L.CT_MOTOR, L.CT_NOSECONE, L.CT_BODYTUBE, L.CT_RECOVERY, L.CT_MASS, L.CT_FINSET, L.CT_BOATTAIL]
    for c, r in zip(comps, g):
        assert eq(c[L.C_POS:L.C_POS + 3], r["position"]), r["cls"]
        if r["cls"] == "NoseCone":
            assert eq(c[L.NC_BASE_AREA], r["baseArea"]) and eq(c[L.NC_LENGTH], r["length"]) and eq(c[L.NC_WETTED], r["wettedArea"])
            assert eq(c[L.NC_CP:L.NC_CP + 3], r["CP"]) and eq(c[L.NC_HALF_ANGLE], r["coneHalfAngle"])
            assert eq(c[L.NC_SUB:L.NC_SUB + 2], r["sub"]) and eq(c[L.NC_TRANS:L.NC_TRANS + 4], r["trans"]) and eq(c[L.NC_ROUGH], r["roughness"])
        elif r["cls"] == "BodyTube":
            assert eq(c[L.BT_LENGTH], r["length"]) and eq(c[L.BT_DIAM], r["outerDiameter"]) and eq(c[L.BT_PLANFORM], r["planformArea"])
            assert eq(c[L.BT_WETTED], r["wettedArea"]) and eq(c[L.BT_CP:L.BT_CP + 3], r["CP"])
        elif r["cls"] == "BoatTail":
            for k, i in (("topArea", L.TR_TOP_AREA), ("bottomArea", L.TR_BOTTOM_AREA), ("frontalArea", L.TR_FRONTAL), ("length", L.TR_LENGTH),
                         ("wettedArea", L.TR_WETTED), ("CdAdjustmentFactor", L.TR_CD_ADJ)):
                assert eq(c[i], r[k]), k
            assert eq(c[L.TR_CP:L.TR_CP + 3], r["CP"])
        elif r["cls"] == "FinSet":
            for k, i in (("numFins", L.FS_NFINS), ("span", L.FS_SPAN), ("rootChord", L.FS_ROOT), ("tipChord", L.FS_TIP), ("thickness", L.FS_THICK),
                         ("surfaceRoughness", L.FS_ROUGH), ("sweepAngle", L.FS_SWEEP), ("midChordSweep", L.FS_MIDSWEEP), ("planformArea", L.FS_PLANFORM),
                         ("wettedArea", L.FS_WETTED), ("aspectRatio", L.FS_AR), ("stallAngle", L.FS_STALL), ("bodyRadius", L.FS_BODY_R),
                         ("bodyOnFinInterferenceFactor", L.FS_BODY_INTERF), ("finNumberInterferenceFactor", L.FS_NUM_INTERF),
                         ("MACLength", L.FS_MAC_LEN), ("MACYPos", L.FS_MAC_Y), ("XMACLeadingEdge", L.FS_XMAC_LE), ("finAngle", L.FS_CANT),
                         ("tipPosition", L.FS_TIP_POS)):
                assert eq(c[i], r[k]), k
            assert eq(c[L.FS_CP_POLY:L.FS_CP_POLY + 6], r["poly"])
            for k, i in (("sliceR", L.FS_SLICE_R), ("sliceA", L.FS_SLICE_A), ("sliceLE", L.FS_SLICE_LE), ("sliceLen", L.FS_SLICE_LEN)):
                assert eq(c[i:i + 10], r[k]), k
            assert eq(c[L.FS_SPANDIR:L.FS_SPANDIR + 8], r["spandir"])
        elif r["cls"] == "TabulatedMotor":
            n = int(c[L.MT_NROWS])
            t = c[L.MT_TABLE:]
            assert eq(t[0:n], r["times"]) and eq(t[n:2 * n], r["thrust"]) and eq(t[2 * n:3 * n], r["oxW"]) and eq(t[3 * n:4 * n], r["fuelW"])
            assert eq(t[4 * n:7 * n].reshape(3, n).T, r["oxMOI"]) and eq(t[7 * n:10 * n].reshape(3, n).T, r["fuelMOI"])
            assert eq(c[L.MT_OX_CG0:L.MT_OX_CG0 + 4], r["cgs"]) and eq(c[L.MT_OX_W0:L.MT_OX_W0 + 2], r["w0"])
        elif r["cls"] == "RecoverySystem":
            assert int(c[L.RC_NSTAGES]) == 2
            s1 = c[L.RC_STAGE0:L.RC_STAGE0 + 5]
            s2 = c[L.RC_STAGE0 + 5:L.RC_STAGE0 + 10]
            assert list(s1) == [L.EV_APOGEE, 30.0, 2.0, 1.5, 2.0] and list(s2) == [L.EV_DESCENDING_ALT, 300.0, 9.0, 1.5, 0.0]


def test_unsupported_component_raises():
    sd = SimDefinition(os.path.join(DATA, "MonteCarlo.mapleaf"), silent=True, disableDistributionSampling=True)
    sd.setValue("Rocket.Sustainer.Odd.class", "StatefulSample")
    with pytest.raises(NotImplementedError, match="StatefulSample"):
        buildModel(sd)
    sd = SimDefinition(os.path.join(DATA, "MonteCarlo.mapleaf"), silent=True, disableDistributionSampling=True)
    sd.setValue("SimControl.timeDiscretization", "RK9")
    with pytest.raises(ValueError):
        buildModel(sd)
