"""
Model-block builder: SimDefinition -> flat FP64 array consumed by the batched trajectory kernels.

This is the host half of the drop-in boundary. It restates what the reference's constructors
precompute once per simulation so that the per-lane device code only has to
evaluate forces and integrate:

  Environment.__init__ ................ MAPLEAF/ENV/environment.py:39-103
  USStandardAtmosphere.__init__ ....... MAPLEAF/ENV/AtmosphereModelling.py:98-145
  integratorFactory / tableaus ........ MAPLEAF/Motion/Integration.py:31-103,144-195,238-329
  Rocket.__init__ and helpers ......... MAPLEAF/Rocket/rocket.py:34-219,251-291,325-337,408-421,489-528
  Stage / CompositeObject ............. MAPLEAF/Rocket/stage.py:13-70, Rocket/CompositeObject.py:23-63
  component constructors .............. Rocket/{noseCone,bodyTube,boatTail,Fins,Propulsion,Recovery,RocketComponents}.py

The layout is defined in include/mlb_layout.h (parsed here, single source of truth).
Component classes the kernels do not implement raise NotImplementedError naming the class
(there is no CPU fallback by design).
"""
import math
import os
import re
from math import cos, sqrt
from typing import Dict, List

import numpy as np

from . import hostmath as hm
from .simdef import SimDefinition, SubDictReader

__all__ = ["L", "buildModel", "Model"]

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))


def _loadLayout() -> Dict[str, int]:
    for cand in (os.path.join(_PKG_DIR, "..", "include", "mlb_layout.h"), os.path.join(_PKG_DIR, "include", "mlb_layout.h")):
        if os.path.exists(cand):
            out = {}
            with open(cand) as f:
                for m in re.finditer(r"^#define\s+(\w+)\s+(-?\d+)\s*(?:/\*.*)?$", f.read(), re.M):
                    out[m.group(1)] = int(m.group(2))
            return out
    raise FileNotFoundError("include/mlb_layout.h not found")


class _Layout:
    def __init__(self, d):
        self.__dict__.update(d)
        self._d = d


L = _Layout(_loadLayout())

_BODY_CLASSES = {"Nosecone", "Bodytube", "BoatTail", "Transition"}
_FIXED_MASS_CLASSES = {"Nosecone", "Bodytube", "BoatTail", "Transition", "FinSet", "Mass", "RecoverySystem"}
_FORCE_ONLY_CLASSES = {"Force", "AeroForce", "AeroDamping", "TabulatedAeroForce", "FractionalJetDamping"}
_SUPPORTED_CLASSES = _FIXED_MASS_CLASSES | {"Motor", "TabulatedInertia"} | _FORCE_ONLY_CLASSES

_INTEGRATORS = {"Euler": L.INT_EULER, "RK2Midpoint": L.INT_RK2_MIDPOINT, "RK2Heun": L.INT_RK2_HEUN, "RK4": L.INT_RK4,
                "RK4_3/8": L.INT_RK4_38, "RK12Adaptive": L.INT_RK12A, "RK23Adaptive": L.INT_RK23A,
                "RK45Adaptive": L.INT_RK45A, "RK78Adaptive": L.INT_RK78A}


def butcherTableau(method: str):
    """Rows of [c_i, a_i1..] followed by the result row(s); zeros are replaced by 1e-16 exactly as the
    reference does before use (Integration.py:51-54, App. A #1)."""
    if method == "Euler":
        return None, 0
    tabs = {
        "RK2Midpoint": ([[0.5, 0.5], [0, 1]], 1),
        "RK2Heun": ([[1, 1], [0.5, 0.5]], 1),
        "RK4": ([[0.5, 0.5], [0.5, 0, 0.5], [1, 0, 0, 1], [1 / 6, 1 / 3, 1 / 3, 1 / 6]], 1),
        "RK4_3/8": ([[1 / 3, 1 / 3], [2 / 3, -1 / 3, 1.0], [1.0, 1.0, -1.0, 1.0], [1 / 8, 3 / 8, 3 / 8, 1 / 8]], 1),
        "RK12Adaptive": ([[0.5, 0.5], [0.0, 1.0], [1.0, 0.0]], 2),
        "RK23Adaptive": ([[0.5, 0.5], [3 / 4, 0.0, 3 / 4], [1.0, 2 / 9, 1 / 3, 4 / 9], [2 / 9, 1 / 3, 4 / 9, 0.0],
                          [7 / 24, 1 / 4, 1 / 3, 1 / 8]], 2),
        "RK45Adaptive": ([[1 / 5, 1 / 5], [3 / 10, 3 / 40, 9 / 40], [4 / 5, 44 / 45, -56 / 15, 32 / 9],
                          [8 / 9, 19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
                          [1.0, 9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
                          [1.0, 35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
                          [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0],
                          [5179 / 57600, 0.0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40]], 2),
        # Dormand-Prince RK8(7)13M, rational approximations of the published coefficients (Prince & Dormand 1981), as the reference
        # enters them (Integration.py:279-297): 12 stage rows after the first evaluation, then the 8th- and the 7th-order result rows
        "RK78Adaptive": ([
            [1 / 18, 1 / 18],
            [1 / 12, 1 / 48, 1 / 16],
            [1 / 8, 1 / 32, 0, 3 / 32],
            [5 / 16, 5 / 16, 0, -75 / 64, 75 / 64],
            [3 / 8, 3 / 80, 0, 0, 3 / 16, 3 / 20],
            [59 / 400, 29443841 / 614563906, 0, 0, 77736538 / 692538347, -28693883 / 1125000000, 23124283 / 1800000000],
            [93 / 200, 16016141 / 946692911, 0, 0, 61564180 / 158732637, 22789713 / 633445777, 545815736 / 2771057229, -180193667 / 1043307555],
            [5490023248 / 9719169821, 39632708 / 573591083, 0, 0, -433636366 / 683701615, -421739975 / 2616292301, 100302831 / 723423059,
             790204164 / 839813087, 800635310 / 3783071287],
            [13 / 20, 246121993 / 1340847787, 0, 0, -37695042795 / 15268766246, -309121744 / 1061227803, -12992083 / 490766935,
             6005943493 / 2108947869, 393006217 / 1396673457, 123872331 / 1001029789],
            [1201146811 / 1299019798, -1028468189 / 846180014, 0, 0, 8478235783 / 508512852, 1311729495 / 1432422823, -10304129995 / 1701304382,
             -48777925059 / 3047939560, 15336726248 / 1032824649, -45442868181 / 3398467696, 3065993473 / 597172653],
            [1, 185892177 / 718116043, 0, 0, -3185094517 / 667107341, -477755414 / 1098053517, -703635378 / 230739211, 5731566787 / 1027545527,
             5232866602 / 850066563, -4093664535 / 808688257, 3962137247 / 1805957418, 65686358 / 487910083],
            [1, 403863854 / 491063109, 0, 0, -5068492393 / 434740067, -411421997 / 543043805, 652783627 / 914296604, 11173962825 / 925320556,
             -13158990841 / 6184727034, 3936647629 / 1978049680, -160528059 / 685178525, 248638103 / 1413531060, 0],
            [14005451 / 335480064, 0, 0, 0, 0, -59238493 / 1068277825, 181606767 / 758867731, 561292985 / 797845732, -1041891430 / 1371343529,
             760417239 / 1151165299, 118820643 / 751138087, -528747749 / 2220607170, 1 / 4],
            [13451932 / 455176623, 0, 0, 0, 0, -808719846 / 976000145, 1757004468 / 5645159321, 656045339 / 265891186, -3867574721 / 1518517206,
             465885868 / 322736535, 53011238 / 667516719, 2 / 45, 0]], 2),
    }
    if method not in tabs:
        raise ValueError("Integration method: {} not implemented. See SimDefinitionTemplate.txt for options.".format(method))
    rows, nResult = tabs[method]
    rows = [[(1e-16 if x == 0 else float(x)) for x in r] for r in rows]
    return rows, nResult


class Inertia:
    """MOI about CG (diagonal), CG, mass. Combination follows Motion/inertia.py:78-115 term by term."""

    def __init__(self, MOI, CG, mass):
        self.MOI, self.CG, self.mass = tuple(MOI), tuple(CG), mass

    @staticmethod
    def combine(items: List["Inertia"]) -> "Inertia":
        totalMass = 0
        totalCG = (0, 0, 0)
        for it in items:
            totalMass += it.mass
            totalCG = hm.v_add(totalCG, hm.v_scale(it.CG, it.mass))
        totalCG = hm.v_div(totalCG, totalMass) if totalMass > 0 else (0, 0, 0)
        mx = my = mz = 0
        for it in items:
            mx += it.MOI[0]
            my += it.MOI[1]
            mz += it.MOI[2]
            d = hm.v_sub(totalCG, it.CG)
            mx += it.mass * (d[1] * d[1] + d[2] * d[2])
            my += it.mass * (d[0] * d[0] + d[2] * d[2])
            mz += it.mass * (d[0] * d[0] + d[1] * d[1])
        return Inertia((mx, my, mz), totalCG, totalMass)


def _linInterp(X, Y, x):
    """Motion/Interpolation.py:15-41 for scalar Y."""
    from bisect import bisect
    i = bisect(X, x)
    if i >= len(X):
        return Y[len(X) - 1]
    if i < 1:
        return Y[0]
    return (Y[i] - Y[i - 1]) * (x - X[i - 1]) / (X[i] - X[i - 1]) + Y[i - 1]


# ======================================================================================
# component builders: each returns (record: list[float], info: dict)
# ======================================================================================
def _fixedMassPrefix(r: SubDictReader, stagePos, ctype, size):
    mass = r.getFloat("mass")
    position = hm.v_add(r.getVector("position"), stagePos)
    cg = hm.v_add(r.getVector("cg"), position)
    try:
        MOI = r.getVector("MOI")
    except Exception:
        MOI = (mass * 0.01, mass * 0.01, mass * 0.01)
    rec = [0.0] * size
    rec[L.C_TYPE] = ctype
    rec[L.C_LEN] = size
    rec[L.C_POS:L.C_POS + 3] = position
    rec[L.C_MOI:L.C_MOI + 3] = MOI
    rec[L.C_CG:L.C_CG + 3] = cg
    rec[L.C_MASS] = mass
    return rec, position, Inertia(MOI, cg, mass)


def _skinFrictionConstants(roughness, length):
    """criticalRe and the roughness-limited coefficient of getSkinFrictionCoefficient (AeroFunctions.py:95-104)."""
    if length == 0:
        return 1e100, 0.0
    if roughness == 0:
        return 1e100, 0.0
    return 51 * (roughness / length) ** (-1.039), 0.032 * (roughness / length) ** 0.2


def _subsonicPolyCoeffs(coneHalfAngle):
    gamma = 1.4
    dCd_dMa_M1 = 4 / (gamma + 1) * (1 - 0.5 * math.sin(coneHalfAngle))
    Cd_M1 = math.sin(coneHalfAngle)
    return [Cd_M1, (dCd_dMa_M1 / Cd_M1)]


def _hoerner(coneHalfAngle, Mach):
    return 2.1 * (math.sin(coneHalfAngle)) ** 2 + 0.5 * math.sin(coneHalfAngle) / math.sqrt(Mach ** 2 - 1)


def cubicInterpInverseMatrix(X1, X2):
    """inverse of the 4x4 Hermite constraint matrix (Motion/Interpolation.py:73-84, scipy.linalg.inv)."""
    import scipy.linalg as linalg
    A = np.array([[1, X1, X1 ** 2, X1 ** 3], [1, X2, X2 ** 2, X2 ** 3], [0, 1, 2 * X1, 3 * X1 ** 2], [0, 1, 2 * X2, 3 * X2 ** 2]])
    return linalg.inv(A)


def _transonicPolyCoeffs(coneHalfAngle):
    gamma = 1.4
    dCd_dMa_M1 = 4 / (gamma + 1) * (1 - 0.5 * math.sin(coneHalfAngle))
    Cd_M1 = math.sin(coneHalfAngle)
    Cd_M13 = _hoerner(coneHalfAngle, 1.3)
    Cd_M1301 = _hoerner(coneHalfAngle, 1.301)
    dCd_dMa_M13 = (Cd_M1301 - Cd_M13) / 0.001
    B = np.array([[Cd_M1], [Cd_M13], [dCd_dMa_M1], [dCd_dMa_M13]])
    return [float(x) for x in cubicInterpInverseMatrix(1.0, 1.3).dot(B)[:, 0]]


def _barrowmanCP(length, areaTop, areaBottom, volume, top=(0, 0, 0)):
    if areaTop == areaBottom:
        d = length / 2
    else:
        d = (length * areaBottom - volume) / (areaBottom - areaTop)
    return (top[0], top[1], top[2] - d)


def _buildNosecone(r, stagePos, rocketRoughness):
    rec, position, inertia = _fixedMassPrefix(r, stagePos, L.CT_NOSECONE, L.NC_SIZE)
    aspectRatio = r.getFloat("aspectRatio")
    baseDiameter = r.getFloat("baseDiameter")
    shape = r.getString("shape")
    roughness = r.tryGetFloat("surfaceRoughness", defaultValue=rocketRoughness)
    length = baseDiameter * aspectRatio
    baseRadius = baseDiameter / 2
    if shape == "tangentOgive":
        rho = (baseRadius ** 2 + length ** 2) / (2 * baseRadius)
        volume = math.pi * (length * rho ** 2 - length ** 3 / 3 - (rho - baseRadius) * rho ** 2 * (math.asin(length / rho)))
        numSteps = 2000
        stepSize = length / numSteps
        wetted = 0
        for x in range(numSteps + 1):
            wetted += 2 * math.pi * ((math.sqrt(rho ** 2 - (length - x * stepSize) ** 2)) + baseRadius - rho) * stepSize
    elif shape == "cone":
        volume = math.pi * (baseDiameter) ** 2 * length / 3
        wetted = math.pi * baseRadius * math.sqrt(baseRadius ** 2 + length ** 2)
    else:
        raise TypeError("The nosecone shape {} is not recognized".format(shape))
    baseArea = baseDiameter ** 2 * math.pi / 4
    CP = _barrowmanCP(length, 0, baseArea, volume)
    half = abs(math.atan((baseDiameter / 2) / length))
    rec[L.NC_BASE_AREA] = baseArea
    rec[L.NC_LENGTH] = length
    rec[L.NC_WETTED] = wetted
    rec[L.NC_CP:L.NC_CP + 3] = CP
    rec[L.NC_HALF_ANGLE] = half
    rec[L.NC_SUB:L.NC_SUB + 2] = _subsonicPolyCoeffs(half)
    rec[L.NC_TRANS:L.NC_TRANS + 4] = _transonicPolyCoeffs(half)
    rec[L.NC_ROUGH] = roughness
    rec[L.NC_CRIT_RE], rec[L.NC_ROUGH_CF] = _skinFrictionConstants(roughness, length)

    def getRadius(d):
        if shape == "cone":
            return (baseDiameter / 2) * (d / length)
        curveRadius = ((baseDiameter / 2) ** 2 + length ** 2) / (baseDiameter)
        y = math.sqrt(curveRadius ** 2 - (length - d) ** 2)
        return y - (curveRadius - (baseDiameter / 2))
    return rec, dict(position=position, inertia=inertia, length=length, maxDiameter=baseDiameter, getRadius=getRadius, body=True,
                     canConnectBelow=True)


def _buildBodytube(r, stagePos, rocketRoughness):
    rec, position, inertia = _fixedMassPrefix(r, stagePos, L.CT_BODYTUBE, L.BT_SIZE)
    length = r.getFloat("length")
    D = r.getFloat("outerDiameter")
    roughness = r.tryGetFloat("surfaceRoughness", defaultValue=rocketRoughness)
    rec[L.BT_LENGTH] = length
    rec[L.BT_DIAM] = D
    rec[L.BT_PLANFORM] = D * length
    rec[L.BT_WETTED] = math.pi * D * length
    rec[L.BT_CP:L.BT_CP + 3] = hm.v_sub(position, (0, 0, length / 2))
    rec[L.BT_ROUGH] = roughness
    rec[L.BT_CRIT_RE], rec[L.BT_ROUGH_CF] = _skinFrictionConstants(roughness, length)
    return rec, dict(position=position, inertia=inertia, length=length, maxDiameter=D, getRadius=lambda _: D / 2, body=True,
                     canConnectBelow=True)


def _transitionRecord(rec, startD, endD, length, position, roughness):
    topArea = startD ** 2 * math.pi / 4
    bottomArea = endD ** 2 * math.pi / 4
    frontal = abs(topArea - bottomArea)
    if length == 0:
        wetted, volume, CP = 0.0, 0.0, position
    else:
        maxD, minD = max(startD, endD), min(startD, endD)
        slope = ((maxD - minD) / length)
        baseConeHeight = maxD / slope
        baseHyp = math.sqrt((maxD / 2) ** 2 + baseConeHeight ** 2)
        fullArea = math.pi * (maxD / 2) * baseHyp
        tipConeHeight = minD / slope
        tipHyp = math.sqrt((minD / 2) ** 2 + tipConeHeight ** 2)
        tipArea = math.pi * (minD / 2) * tipHyp
        wetted = fullArea - tipArea
        volume = topArea * baseConeHeight / 3 - bottomArea * tipConeHeight / 3
        CP = _barrowmanCP(length, topArea, bottomArea, volume, position)
    aspectRatio = 100 if startD == endD else length / abs(startD - endD)
    rec[L.TR_TOP_AREA] = topArea
    rec[L.TR_BOTTOM_AREA] = bottomArea
    rec[L.TR_FRONTAL] = frontal
    rec[L.TR_LENGTH] = length
    rec[L.TR_WETTED] = wetted
    rec[L.TR_CP:L.TR_CP + 3] = CP
    rec[L.TR_ROUGH] = roughness
    rec[L.TR_CRIT_RE], rec[L.TR_ROUGH_CF] = _skinFrictionConstants(roughness, length)
    rec[L.TR_START_D] = startD
    rec[L.TR_END_D] = endD
    if endD <= startD:
        rec[L.TR_GROWING] = 0
        rec[L.TR_CD_ADJ] = 1 if aspectRatio < 1 else ((3 - aspectRatio) / 2 if aspectRatio < 3 else 0)
    else:
        half = math.pi / 2 if length == 0 else math.atan(abs(startD - endD) / 2 / length)
        rec[L.TR_GROWING] = 1
        rec[L.TR_HALF_ANGLE] = half
        rec[L.TR_SUB:L.TR_SUB + 2] = _subsonicPolyCoeffs(half)
        rec[L.TR_TRANS:L.TR_TRANS + 4] = _transonicPolyCoeffs(half)
    return rec


def _buildTransition(r, stagePos, rocketRoughness, ctype):
    rec, position, inertia = _fixedMassPrefix(r, stagePos, ctype, L.TR_SIZE)
    length = r.getFloat("length")
    startD = r.getFloat("startDiameter")
    endD = r.getFloat("endDiameter")
    roughness = r.tryGetFloat("surfaceRoughness", defaultValue=rocketRoughness)
    _transitionRecord(rec, startD, endD, length, position, roughness)
    return rec, dict(position=position, inertia=inertia, length=length, maxDiameter=max(startD, endD), body=True,
                     getRadius=lambda d: (d / length * (endD - startD) + startD) / 2, canConnectBelow=(ctype == L.CT_TRANSITION),
                     isBoatTail=(ctype == L.CT_BOATTAIL))


def _autoBoatTailRecord(diameter, position, roughness):
    rec = [0.0] * L.TR_SIZE
    rec[L.C_TYPE] = L.CT_BOATTAIL
    rec[L.C_LEN] = L.TR_SIZE
    rec[L.C_POS:L.C_POS + 3] = position
    return _transitionRecord(rec, diameter, diameter, 0, position, roughness)


def _buildFinSet(r, stagePos, rocketRoughness):
    rec, position, inertia = _fixedMassPrefix(r, stagePos, L.CT_FINSET, L.FS_SIZE)
    numFins = r.getInt("numFins")
    firstFinAngle = r.getFloat("firstFinAngle")
    finAngle = r.getFloat("finCantAngle")
    sweepAngle = math.radians(r.getFloat("sweepAngle"))
    rootChord = r.getFloat("rootChord")
    tipChord = r.getFloat("tipChord")
    span = r.getFloat("span")
    roughness = r.tryGetFloat("surfaceRoughness", defaultValue=rocketRoughness)
    thickness = r.getFloat("thickness")
    nSlices = r.getInt("numFinSpanSlicesForIntegration")
    if numFins > L.FS_MAX_FINS or nSlices > L.FS_MAX_SLICES:
        raise NotImplementedError("FinSet with more than {} fins or {} span slices".format(L.FS_MAX_FINS, L.FS_MAX_SLICES))
    leShape = r.getString("LeadingEdge.shape")
    if leShape == "Round":
        leThick = r.tryGetFloat("LeadingEdge.radius", defaultValue=thickness / 100) * 2
    elif leShape == "Blunt":
        leThick = r.tryGetFloat("LeadingEdge.thickness", defaultValue=thickness)
    else:
        raise ValueError("ERROR: Leading edge shape: {} not implemented. Use 'Round' or 'Blunt'".format(leShape))
    teShape = r.getString("TrailingEdge.shape")
    if teShape == "Round":
        teThick = r.getFloat("TrailingEdge.radius")
    elif teShape == "Blunt":
        teThick = r.getFloat("TrailingEdge.thickness")
    elif teShape == "Tapered":
        teThick = 0
    else:
        raise ValueError("ERROR: Trailing edge shape: {} not implemented. Use 'Round', 'Blunt', or 'Tapered'".format(teShape))
    rec[L.FS_NFINS] = numFins
    rec[L.FS_SPAN] = span
    rec[L.FS_ROOT] = rootChord
    rec[L.FS_TIP] = tipChord
    rec[L.FS_THICK] = thickness
    rec[L.FS_ROUGH] = roughness
    rec[L.FS_SWEEP] = sweepAngle
    rec[L.FS_LE_SHAPE] = 0 if leShape == "Round" else 1
    rec[L.FS_LE_THICK] = leThick
    rec[L.FS_TE_THICK] = teThick
    rec[L.FS_CANT] = finAngle
    rec[L.FS_NSLICES] = nSlices
    info = dict(position=position, inertia=inertia, body=False, finset=dict(
        numFins=numFins, firstFinAngle=firstFinAngle, sweepAngle=sweepAngle, rootChord=rootChord, tipChord=tipChord, span=span,
        thickness=thickness, nSlices=nSlices))
    return rec, info


def _finishFinSet(rec, fs, bodyRadius):
    """FinSet.precomputeProperties (Fins.py:93-262) once the mounting radius is known."""
    span, rootChord, tipChord, sweepAngle = fs["span"], fs["rootChord"], fs["tipChord"], fs["sweepAngle"]
    thickness, nSlices, numFins = fs["thickness"], fs["nSlices"], fs["numFins"]
    LEtip = span * math.tan(sweepAngle)
    TEtip = tipChord + LEtip - rootChord
    if TEtip == 0:
        teSweep = 0.0
    elif TEtip < 0:
        teSweep = -1 * math.atan2(abs(TEtip), span)
    else:
        teSweep = math.atan2(abs(TEtip), span)
    midChordSweep = (sweepAngle + teSweep) / 2
    dZdS_LE = math.tan(sweepAngle)
    planform = span * (tipChord + rootChord) / 2
    wetted = 2 * planform
    AR = (2 * span) ** 2 / planform
    stall = 15 * (1 + 1.1 / AR)          # degrees, compared against radians downstream (App. A #7)
    bodyInterf = 1 + bodyRadius / (span + bodyRadius)
    if numFins <= 4:
        numInterf = 1.0
    elif numFins <= 8:
        raise AttributeError("'FinSet' object has no attribute 'finNumber'")   # reference bug, Fins.py:154
    else:
        numInterf = 0.75

    def getChord(y):
        return (tipChord - rootChord) * (y / span) + rootChord
    numSpanSteps = round(span / 0.0001)
    step = span / numSpanSteps
    sMAC = sY = sX = 0
    for i in range(numSpanSteps):
        y = i * step + step / 2
        sMAC += ((getChord(y) ** 2) * step)
        sY += y * (getChord(y) * step)
        sX += (dZdS_LE * y) * getChord(y) * step
    MACLength, MACYPos, XMACLeadingEdge = sMAC / planform, sY / planform, sX / planform

    stepSize = span / nSlices
    for i in range(nSlices):
        y = i * stepSize + stepSize * 0.5
        rec[L.FS_SLICE_R + i] = y + bodyRadius
        rec[L.FS_SLICE_A + i] = stepSize * getChord(y)
        rec[L.FS_SLICE_LE + i] = y * dZdS_LE
        rec[L.FS_SLICE_LEN + i] = getChord(y)

    sigma = (AR / 2) * (thickness / rootChord) ** (1 / 3)
    CD_TT_star = 1.15 * (thickness / rootChord) ** (5 / 3) * (1.61 + sigma - sqrt((sigma - 1.43) ** 2 + 0.578))

    f_2 = (AR * (3) ** 0.5 - 0.67) / (2 * (3) ** 0.5 * AR - 1)
    t1 = (2 * (3) ** -0.5 * AR) * (2 * AR * (3) ** 0.5 - 1)
    t2 = ((3) ** 0.5 * AR - 0.67) * (AR * 4 * (3) ** -0.5)
    den = (2 * AR * (3) ** 0.5 - 1) ** 2
    f_prime_2 = (t1 - t2) / den
    a_mat = [[0.5 ** 5, 0.5 ** 4, 0.5 ** 3, 0.5 ** 2, 0.5, 1], [5 * (0.5) ** 4, 4 * (0.5) ** 3, 3 * (0.5) ** 2, 2 * (0.5), 1, 0],
             [2 ** 5, 2 ** 4, 2 ** 3, 2 ** 2, 2, 1], [5 * (2) ** 4, 4 * (2) ** 3, 3 * (2) ** 2, 2 * 2, 1, 0],
             [20 * 2 ** 3, 12 * 2 ** 2, 6 * 2, 2, 0, 0], [60 * 2 ** 2, 24 * 2, 6, 0, 0, 0]]
    poly = np.linalg.solve(a_mat, [0.25, 0, f_2, f_prime_2, 0, 0])

    rec[L.FS_MIDSWEEP] = midChordSweep
    rec[L.FS_PLANFORM] = planform
    rec[L.FS_WETTED] = wetted
    rec[L.FS_AR] = AR
    rec[L.FS_STALL] = stall
    rec[L.FS_BODY_R] = bodyRadius
    rec[L.FS_BODY_INTERF] = bodyInterf
    rec[L.FS_NUM_INTERF] = numInterf
    rec[L.FS_MAC_LEN] = MACLength
    rec[L.FS_MAC_Y] = MACYPos
    rec[L.FS_XMAC_LE] = XMACLeadingEdge
    rec[L.FS_CP_POLY:L.FS_CP_POLY + 6] = [float(x) for x in poly]
    rec[L.FS_TIP_POS] = dZdS_LE * span
    # Transonic blend constants: the four strip sums of Fins.py:508-528 use fixed Mach numbers, so CN_alpha, the Busemann coefficients
    # and each slice's Mach-cone weight (CythonFinFunctions.pyx:92-110) do not depend on the flight state
    def beta(Mach):
        return sqrt(1 - Mach * Mach) if Mach <= 0.9 else sqrt(Mach * Mach - 1)
    for j, Mach in enumerate((0.8, 0.8 + 0.001)):
        t = beta(Mach) * (2 * (span * span) / planform) / math.cos(midChordSweep)
        rec[L.FS_TRANS_CNA + j] = (2 * math.pi * (2 * (span * span) / planform)) / (2 + sqrt(4 + t * t))
    gamma = 1.4
    outerRadius = (nSlices - 1) * stepSize + stepSize * 0.5 + bodyRadius
    for j, Mach in enumerate((1.4, 1.4 + 0.001)):
        b = beta(Mach)
        M2 = Mach * Mach; M4 = M2 * M2; M6 = M4 * M2; M8 = M4 * M4
        b2 = b * b; b4 = b2 * b2; b7 = b4 * b2 * b
        K = [2 / b, ((gamma + 1) * M4 - 4 * b2) / (4 * b4),
             ((gamma + 1) * M8 + (2 * (gamma * gamma) - 7 * gamma - 5) * M6 + 10 * (gamma + 1) * M4 - 12 * M2 + 8) / (6 * b7),
             1 / math.tan(math.asin(1 / Mach))]
        rec[L.FS_TRANS_K + 4 * j:L.FS_TRANS_K + 4 * j + 4] = K
        for i in range(nSlices):
            machCone = (outerRadius - rec[L.FS_SLICE_R + i]) * K[3] + dZdS_LE * span
            fractionOut = min(1.0, abs(machCone - rec[L.FS_SLICE_LE + i]) / rec[L.FS_SLICE_LEN + i])
            rec[L.FS_TRANS_W + 16 * j + i] = fractionOut + 0.5 * (1 - fractionOut)
    rec[L.FS_CDTT_STAR] = CD_TT_star
    rec[L.FS_CRIT_RE], rec[L.FS_ROUGH_CF] = _skinFrictionConstants(rec[L.FS_ROUGH], MACLength)
    sep = 360 / numFins
    for i in range(numFins):
        rec[L.FS_SPANDIR + 2 * i] = math.cos(math.radians(fs["firstFinAngle"] + i * sep))
        rec[L.FS_SPANDIR + 2 * i + 1] = math.sin(math.radians(fs["firstFinAngle"] + i * sep))
        # constant per fin while it is not actuated: deflected normal (Fins.py:479-480), drag direction (:536), |CP span offset| (:489)
        spanDir = (rec[L.FS_SPANDIR + 2 * i], rec[L.FS_SPANDIR + 2 * i + 1], 0.0)
        rot = hm.q_from_axis_angle(spanDir, rec[L.FS_CANT] * (math.pi / 180.0))
        normal, _ = hm.q_rotate(rot, (spanDir[1], -spanDir[0], 0.0))
        rec[L.FS_FIN_NORMAL + 3 * i:L.FS_FIN_NORMAL + 3 * i + 3] = normal
        rec[L.FS_AXIAL_DIR + 3 * i:L.FS_AXIAL_DIR + 3 * i + 3] = hm.v_cross(spanDir, normal)
        rec[L.FS_SPANWISE_CP + i] = hm.v_len(hm.v_scale(spanDir, MACYPos + bodyRadius))


def parseMotorFile(path, stagePosZ, impulseAdjustFactor=1.0, burnTimeAdjustFactor=1.0):
    """Propulsion.py:51-128: returns dict with times, thrust, oxWeights, fuelWeights, MOIs, CG end points."""
    with open(path, "r") as f:
        text = re.sub(re.compile("#.*"), "", f.read())
    lines = [ln for ln in text.split("\n") if ln.strip() != ""]
    cg = [float(lines[i].split()[1]) + stagePosZ for i in range(4)]
    times, thrust, oxFlow, fuelFlow, oxMOIs, fuelMOIs = [], [], [], [], [], []
    from .simdef import parseVector
    for ln in lines[4:]:
        info = ln.split()
        times.append(float(info[0]))
        thrust.append(float(info[1]))
        oxFlow.append(float(info[2]))
        fuelFlow.append(float(info[3]))
        a = ln.index("(")
        b = ln.index(")", a) + 1
        oxMOIs.append(parseVector(ln[a:b]))
        c = ln.index("(", b)
        d = ln.index(")", c) + 1
        fuelMOIs.append(parseVector(ln[c:d]))
    burnDuration0 = times[-1]      # stage.engineShutOffTime is recorded BEFORE the burn-time factor is applied
    rawTimes, rawThrust = list(times), list(thrust)
    thrust = [t * impulseAdjustFactor / burnTimeAdjustFactor for t in thrust]
    times = [t * burnTimeAdjustFactor for t in times]
    initOx = 0
    initFuel = 0
    oxW = [0]
    fuelW = [0]
    for i in range(len(times) - 1, 0, -1):
        dT = times[i] - times[i - 1]
        initOx += dT * (oxFlow[i - 1] + oxFlow[i]) / 2
        oxW.insert(0, initOx)
        initFuel += dT * (fuelFlow[i - 1] + fuelFlow[i]) / 2
        fuelW.insert(0, initFuel)
    return dict(times=times, thrust=thrust, oxW=oxW, fuelW=fuelW, oxMOIs=oxMOIs, fuelMOIs=fuelMOIs, cg=cg,
                initOx=initOx, initFuel=initFuel, engineShutOff=burnDuration0, rawTimes=rawTimes, rawThrust=rawThrust, oxFlow=oxFlow, fuelFlow=fuelFlow)


def motorInertia(m, tSinceIgnition) -> Inertia:
    """TabulatedMotor.getInertia (Propulsion.py:130-135,171-204)."""
    def part(W0, W, cg0, cg1, MOIs):
        if W0 == 0:
            return Inertia((0, 0, 0), (0, 0, 0), 0)
        w = _linInterp(m["times"], W, tSinceIgnition)
        frac = 1 - (w / W0)
        z = cg0 * (1 - frac) + cg1 * frac
        from bisect import bisect
        X = m["times"]
        i = bisect(X, tSinceIgnition)
        if i >= len(X):
            moi = MOIs[len(X) - 1]
        elif i < 1:
            moi = MOIs[0]
        else:
            moi = hm.v_add(hm.v_div(hm.v_scale(hm.v_sub(MOIs[i], MOIs[i - 1]), (tSinceIgnition - X[i - 1])), (X[i] - X[i - 1])), MOIs[i - 1])
        return Inertia(moi, (0, 0, z), w)
    ox = part(m["initOx"], m["oxW"], m["cg"][0], m["cg"][1], m["oxMOIs"])
    fuel = part(m["initFuel"], m["fuelW"], m["cg"][2], m["cg"][3], m["fuelMOIs"])
    return Inertia.combine([fuel, ox])      # ox + fuel == ox.combineInertias([fuel]) -> list [fuel, ox]


def _buildMotor(r, stagePos):
    impulse = r.getFloat("impulseAdjustFactor")
    burn = r.getFloat("burnTimeAdjustFactor")
    path = r.getString("path")
    if not os.path.exists(path):
        found = r.simDefinition._findFile(path)
        if found is None:
            raise FileNotFoundError(path)
        path = found
    m = parseMotorFile(path, stagePos[2], impulse, burn)
    n = len(m["times"])
    size = L.MT_TABLE + L.MT_NCOLS * n
    rec = [0.0] * size
    rec[L.C_TYPE] = L.CT_MOTOR
    rec[L.C_LEN] = size
    rec[L.MT_NROWS] = n
    rec[L.MT_OX_CG0], rec[L.MT_OX_CG1], rec[L.MT_FUEL_CG0], rec[L.MT_FUEL_CG1] = m["cg"]
    rec[L.MT_OX_W0] = m["initOx"]
    rec[L.MT_FUEL_W0] = m["initFuel"]
    t = L.MT_TABLE
    cols = [m["times"], m["thrust"], m["oxW"], m["fuelW"]] + [[v[k] for v in m["oxMOIs"]] for k in range(3)] + \
           [[v[k] for v in m["fuelMOIs"]] for k in range(3)] + [m["rawTimes"], m["rawThrust"], m["oxFlow"], m["fuelFlow"]]
    for c, col in enumerate(cols):
        rec[t + c * n:t + (c + 1) * n] = col
    init = motorInertia(m, 0)
    rec[L.C_POS:L.C_POS + 3] = init.CG      # motor "position" = its initial CG (Propulsion.py:47-49)
    return rec, dict(position=init.CG, motor=m, body=False, motorFactors=(impulse, burn), path=r.simDefDictPathToReadFrom)


_TRIGGERS = {"Apogee": L.EV_APOGEE, "Altitude": L.EV_DESCENDING_ALT, "Time": L.EV_TIME_REACHED}


def _buildRecovery(r, stagePos):
    rec, position, inertia = _fixedMassPrefix(r, stagePos, L.CT_RECOVERY, L.RC_SIZE)
    n = r.getInt("numStages")
    if n > L.RC_MAX_STAGES:
        raise NotImplementedError("RecoverySystem with more than {} stages".format(L.RC_MAX_STAGES))
    rec[L.RC_NSTAGES] = n
    for i in range(1, n + 1):
        trig = r.getString("stage{}Trigger".format(i))
        if trig not in _TRIGGERS:
            raise KeyError(trig)
        try:
            val = r.getFloat("stage{}TriggerValue".format(i))
        except KeyError:
            val = 0.0
        o = L.RC_STAGE0 + (i - 1) * L.RC_STAGE_STRIDE
        rec[o:o + 5] = [_TRIGGERS[trig], val, r.getFloat("stage{}ChuteArea".format(i)), r.getFloat("stage{}Cd".format(i)),
                        r.getFloat("stage{}DelayTime".format(i))]
    return rec, dict(position=position, inertia=inertia, body=False)


def _buildMass(r, stagePos):
    rec, position, inertia = _fixedMassPrefix(r, stagePos, L.CT_MASS, L.C_P)
    return rec, dict(position=position, inertia=inertia, body=False)


_ZERO_INERTIA = None


def _zeroInertiaRecord(ctype, size, position=(0.0, 0.0, 0.0)):
    rec = [0.0] * size
    rec[L.C_TYPE] = ctype
    rec[L.C_LEN] = size
    rec[L.C_POS:L.C_POS + 3] = position
    return rec


def _resolvePath(r, key):
    path = r.getString(key)
    if not os.path.exists(path):
        found = r.simDefinition._findFile(path)
        if found is None:
            raise FileNotFoundError(path)
        path = found
    return path


def _buildFixedForce(r, stagePos):
    """RocketComponents.py:183-204"""
    rec = _zeroInertiaRecord(L.CT_FIXED_FORCE, L.FF_SIZE, r.getVector("position"))
    rec[L.FF_FORCE:L.FF_FORCE + 3] = r.getVector("force")
    rec[L.FF_MOMENT:L.FF_MOMENT + 3] = r.getVector("moment")
    # FixedForce has no `position` attribute: the Z sort falls back to its (zero) inertia's CG (RocketComponents.py:137-141)
    return rec, dict(position=(0.0, 0.0, 0.0), body=False)


def _buildAeroForce(r, stagePos):
    """RocketComponents.py:206-231"""
    position = r.getVector("position")
    rec = _zeroInertiaRecord(L.CT_AERO_FORCE, L.AF_SIZE, position)
    rec[L.AF_AREF] = r.getFloat("Aref")
    rec[L.AF_LREF] = r.getFloat("Lref")
    rec[L.AF_COEFFS:L.AF_COEFFS + 5] = [r.getFloat("Cd"), r.getFloat("Cl"), *r.getVector("momentCoeffs")]
    return rec, dict(position=position, body=False)


def _buildAeroDamping(r, stagePos):
    """RocketComponents.py:233-261 (class-level position = origin)"""
    rec = _zeroInertiaRecord(L.CT_AERO_DAMPING, L.AD_SIZE)
    rec[L.AD_AREF] = r.getFloat("Aref")
    rec[L.AD_LREF] = r.getFloat("Lref")
    rec[L.AD_XCOEFFS:L.AD_XCOEFFS + 3] = r.getVector("xDampingCoeffs")
    rec[L.AD_YCOEFFS:L.AD_YCOEFFS + 3] = r.getVector("yDampingCoeffs")
    rec[L.AD_ZCOEFFS:L.AD_ZCOEFFS + 3] = r.getVector("zDampingCoeffs")
    return rec, dict(position=(0.0, 0.0, 0.0), body=False)


_AERO_KEYS = {"Mach": L.AKEY_MACH, "Altitude": L.AKEY_ALTITUDE, "UnitReynolds": L.AKEY_UNITRE, "TotalAOA": L.AKEY_TOTALAOA,
              "RollAngle": L.AKEY_ROLLANGLE, "AOA": L.AKEY_AOA, "AOSS": L.AKEY_AOSS}
_AERO_COEFFS = ["CD", "CL", "CMx", "CMy", "CMz"]


def _buildTabulatedAeroForce(r, stagePos):
    """RocketComponents.py:263-345; tables with one key column (the reference's 1-D linInterp path)."""
    position = r.getVector("position")
    path = _resolvePath(r, "filePath")
    with open(path) as f:
        columnNames = f.readline().strip().split(",")
    keys = []
    i = 0
    while i < len(columnNames) and columnNames[i] in _AERO_KEYS:
        keys.append(_AERO_KEYS[columnNames[i]])
        i += 1
    coeffIdx = []
    for name in columnNames[i:]:
        if name not in _AERO_COEFFS:
            raise ValueError("ERROR: One of the following columns: {} did not match any of the expected columns names: Keys: {}, values: {}. \
                    Or was in the wrong order. All key columns must come BEFORE value columns.".format(columnNames, list(_AERO_KEYS), _AERO_COEFFS))
        coeffIdx.append(_AERO_COEFFS.index(name))
    if len(keys) < 1 or len(keys) > L.ND_MAX_DIM:
        raise NotImplementedError("TabulatedAeroForce tables with {} key columns are not implemented by the B200 backend".format(len(keys)))
    data = np.loadtxt(path, delimiter=",", skiprows=1)
    if data.ndim == 1:
        data = data.reshape(1, -1)
    nKeys = len(keys)
    n, nVals = data.shape[0], data.shape[1] - nKeys
    if nKeys > 1:
        # scattered linear interpolation over several keys (NoNaNLinearNDInterpolator, RocketComponents.py:316-320): the Delaunay record
        # is appended behind the shared part of the block once that is complete; its offset is patched into TA_ND_OFF
        rec = _zeroInertiaRecord(L.CT_TAB_AERO, L.TA_TABLE, position)
        rec[L.TA_AREF] = r.getFloat("Aref")
        rec[L.TA_LREF] = r.getFloat("Lref")
        rec[L.TA_KEY] = keys[0]
        rec[L.TA_NKEYS] = nKeys
        rec[L.TA_KEYS:L.TA_KEYS + nKeys] = keys
        rec[L.TA_NVALS] = nVals
        rec[L.TA_COEFF_IDX:L.TA_COEFF_IDX + 5] = coeffIdx + [-1] * (5 - len(coeffIdx))
        return rec, dict(position=position, body=False, ndTable=(data[:, 0:nKeys], data[:, nKeys:]))
    size = L.TA_TABLE + n * (1 + nVals)
    rec = _zeroInertiaRecord(L.CT_TAB_AERO, size, position)
    rec[L.TA_AREF] = r.getFloat("Aref")
    rec[L.TA_LREF] = r.getFloat("Lref")
    rec[L.TA_KEY] = keys[0]
    rec[L.TA_NKEYS] = 1
    rec[L.TA_NROWS] = n
    rec[L.TA_NVALS] = nVals
    rec[L.TA_COEFF_IDX:L.TA_COEFF_IDX + 5] = coeffIdx + [-1] * (5 - len(coeffIdx))
    rec[L.TA_TABLE:L.TA_TABLE + n] = [float(x) for x in data[:, 0]]
    for c in range(nVals):
        rec[L.TA_TABLE + n * (1 + c):L.TA_TABLE + n * (2 + c)] = [float(x) for x in data[:, 1 + c]]
    return rec, dict(position=position, body=False)


def _buildTabulatedInertia(r, stagePos):
    """RocketComponents.py:347-376. No `position` attribute: sorts by its CG at t = 0."""
    path = _resolvePath(r, "filePath")
    data = np.loadtxt(path, skiprows=1, delimiter=",")
    if data.shape[1] != 8:
        raise ValueError("Wrong number of columns in inertia table: {}. Expecting 8 columns: \
                Time, Mass, CGx, CGy, CGz, MOIx, MOIy, MOIz")
    n = data.shape[0]
    rec = _zeroInertiaRecord(L.CT_TAB_INERTIA, L.TI_TABLE + 8 * n)
    rec[L.TI_NROWS] = n
    rec[L.TI_TABLE:L.TI_TABLE + 8 * n] = [float(x) for x in data.T.reshape(-1)]
    table = dict(times=[float(x) for x in data[:, 0]], rows=[[float(x) for x in row[1:]] for row in data])
    cg0 = tabulatedInertia(table, 0).CG
    rec[L.C_POS:L.C_POS + 3] = cg0
    return rec, dict(position=cg0, body=False, tabInertia=table)


def tabulatedInertia(table, time) -> Inertia:
    """TabulatedInertia.getInertia: linInterp of the 7 value columns as one numpy row (Interpolation.py:15-41)."""
    from bisect import bisect
    X, Y = table["times"], table["rows"]
    i = bisect(X, time)
    if i >= len(X):
        row = Y[len(X) - 1]
    elif i < 1:
        row = Y[0]
    else:
        row = [(Y[i][k] - Y[i - 1][k]) * (time - X[i - 1]) / (X[i] - X[i - 1]) + Y[i - 1][k] for k in range(7)]
    return Inertia(row[4:7], row[1:4], row[0])


def _buildJetDamping(r, stagePos):
    """RocketComponents.py:378-413"""
    rec = _zeroInertiaRecord(L.CT_JET_DAMPING, L.JD_SIZE)
    rec[L.JD_FRACTION] = r.getFloat("fraction")
    return rec, dict(position=(0.0, 0.0, 0.0), body=False)


def _ndInterpolatorRecord(keys, values):
    """Delaunay triangulation of the key points with scipy (the library the reference calls, Motion/Interpolation.py:109-124)."""
    from scipy.interpolate import LinearNDInterpolator
    keys = np.asarray(keys, dtype=np.float64)
    values = np.asarray(values, dtype=np.float64)
    if keys.shape[1] > L.ND_MAX_DIM or values.shape[1] > L.ND_MAX_VAL:
        raise NotImplementedError("interpolation table with more than {} key or {} value columns".format(L.ND_MAX_DIM, L.ND_MAX_VAL))
    tri = LinearNDInterpolator(keys, values).tri
    rec = [keys.shape[1], values.shape[1], keys.shape[0], tri.simplices.shape[0]]
    rec += [float(x) for x in keys.reshape(-1)] + [float(x) for x in values.reshape(-1)]
    rec += [float(x) for x in tri.simplices.reshape(-1)] + [float(x) for x in tri.transform.reshape(-1)]
    rec += [float(x) for x in tri.neighbors.reshape(-1)]
    return rec


def _buildControlSystem(sd, rk, stages, method, H, append):
    cs = SubDictReader("Rocket.ControlSystem", sd)
    rec = [0.0] * L.CS_SIZE
    # Stabilizer (Navigation.py:15-24)
    direction = hm.v_normalize(cs.getVector("desiredFlightDirection"))
    rec[L.CS_TARGET_Q:L.CS_TARGET_Q + 4] = hm.q_from_axis_angle(hm.v_cross((0, 0, 1), direction), hm.v_angle((0, 0, 1), direction))
    mcType = cs.getString("MomentController.Type")
    if mcType == "ConstantGainPIDRocket":
        rec[L.CS_MC_TYPE] = 0
        rec[L.CS_GAINS:L.CS_GAINS + 6] = [cs.getFloat("MomentController." + k) for k in ("Pxy", "Ixy", "Dxy", "Pz", "Iz", "Dz")]
    elif mcType == "ScheduledGainPIDRocket":
        rec[L.CS_MC_TYPE] = 1
        keyNames = cs.getString("MomentController.scheduledBy").split()
        if len(keyNames) > 4:
            raise NotImplementedError("gain table scheduled by more than 4 keys")
        rec[L.CS_GAIN_NKEYS] = len(keyNames)
        rec[L.CS_GAIN_KEYS:L.CS_GAIN_KEYS + len(keyNames)] = [_AERO_KEYS[k] for k in keyNames]
    elif mcType == "IdealMomentController":
        # MomentControllers.py:85-93: every time step the orientation becomes the target orientation and the angular velocity zero; no
        # controlled system, no actuators, update rate 0 = once per step (ControlSystems.py:79-100,120-122)
        rec[L.CS_MC_TYPE] = 2
        rec[L.CS_STAGE], rec[L.CS_COMP] = -1, -1
        return rec
    else:
        raise ValueError("Moment Controller Type: {} not implemented. Try ScheduledGainPIDRocket or IdealMomentController".format(mcType))
    updateRate = cs.getFloat("updateRate")
    # controlled system: a FinSet (the only ActuatedSystem of the reference, Fins.py:45,87-91)
    path = cs.getString("controlledSystem")
    found = None
    for si, st in enumerate(stages):
        for ci, c in enumerate(st["comps"]):
            if "Rocket.{}.{}".format(st["name"], c["name"]) == path:
                found = (si, ci, c)
    if found is None:
        raise ValueError("Rocket Component: {} not found in {}".format(path, sd.fileName))
    si, ci, comp = found
    if comp["cls"] != "FinSet":
        raise AttributeError("'{}' object has no attribute 'initializeActuators'".format(comp["cls"]))
    rec[L.CS_STAGE], rec[L.CS_COMP] = si, ci
    nAct = int(comp["rec"][L.FS_NFINS])
    rec[L.CS_NACT] = nAct
    act = SubDictReader(path + ".Actuators", sd)
    if act.getString("responseModel") != "FirstOrder":
        raise ValueError("Actuator response model {} not implemented. Try 'FirstOrder'".format(act.getString("responseModel")))
    rec[L.CS_ACT_TAU] = act.getFloat("responseTime")
    mx, mn = act.tryGetFloat("maxDeflection", defaultValue=None), act.tryGetFloat("minDeflection", defaultValue=None)
    if mn is not None:
        rec[L.CS_ACT_HASMIN], rec[L.CS_ACT_MIN] = 1.0, mn
    if mx is not None:
        rec[L.CS_ACT_HASMAX], rec[L.CS_ACT_MAX] = 1.0, mx
    if act.getString("controller") != "TableInterpolating":
        raise ValueError("Actuator controller: {} not implemented. Try TableInterpolating.")
    cols = act.getString("deflectionKeyColumns").split()
    noDesired = []
    for k in cols:
        if "Desired" not in k:
            noDesired.append(k)
        else:
            break
    for k in cols[len(noDesired):]:
        if "Desired" not in k:
            raise ValueError("'Desired' columns such as DesiredMx must come last in deflectionKeyColumns")
    if len(cols) - len(noDesired) != 3 or len(noDesired) > 3:
        raise NotImplementedError("deflection table keys must be up to 3 aero keys followed by DesiredMx DesiredMy DesiredMz")
    rec[L.CS_DEFL_NKEYS] = len(noDesired)
    rec[L.CS_DEFL_KEYS:L.CS_DEFL_KEYS + len(noDesired)] = [_AERO_KEYS[k] for k in noDesired]
    # time stepping (ControlSystems.py:120-153)
    if updateRate > 0:
        controlTimeStep = 1 / updateRate
        rec[L.CS_UPDATE_DT] = controlTimeStep
        if "Adaptive" in method:
            rec[L.CS_FORCE_RK4] = 1.0
            rows, nRes = butcherTableau("RK4")
            rec[L.CS_TABLEAU2_OFF] = append(_flattenTableau(rows, nRes))
        original = float(sd.getValue("SimControl.timeStep"))
        if controlTimeStep / original != round(controlTimeStep / original):
            recommended = controlTimeStep / max(1, round(controlTimeStep / original))
            sd.setValue("SimControl.timeStep", str(recommended))      # the reference rewrites the definition, too (:147)
            H[L.H_DT0] = recommended
    return rec


def _appendInterpolators(sd, rk, H, tail, append):
    """The (large) Delaunay tables go behind H_SHARED_LEN; their offsets are patched into the control record."""
    off = int(H[L.H_CTRL_OFF]) - L.H_SIZE
    cs = SubDictReader("Rocket.ControlSystem", sd)
    if tail[off + L.CS_MC_TYPE] == 1:
        path = _resolvePath(cs, "MomentController.gainTableFilePath")
        data = np.loadtxt(path, skiprows=1)
        n = int(tail[off + L.CS_GAIN_NKEYS])
        tail[off + L.CS_GAIN_INTERP_OFF] = append(_ndInterpolatorRecord(data[:, 0:n], data[:, n:n + 6]))
    act = SubDictReader(cs.getString("controlledSystem") + ".Actuators", sd)
    data = np.loadtxt(_resolvePath(act, "deflectionTablePath"), skiprows=1)
    nKeys = int(tail[off + L.CS_DEFL_NKEYS]) + 3
    nAct = int(tail[off + L.CS_NACT])
    if data.shape[1] - nKeys != nAct:
        raise ValueError("Number of actuators: {}, must match number of actuator deflection columns in deflection table: {}".format(nAct, data.shape[1] - nKeys))
    tail[off + L.CS_DEFL_INTERP_OFF] = append(_ndInterpolatorRecord(data[:, 0:nKeys], data[:, nKeys:]))


# ======================================================================================
class Model:
    """Result of buildModel: `.blob` (np.float64 array) plus host-side metadata."""

    def __init__(self, blob, meta):
        self.blob = blob
        self.meta = meta

    def header(self, name):
        return float(self.blob[getattr(L, name)])

    def stageRecord(self, i):
        off = int(self.blob[L.H_STAGE_OFF + i])
        return self.blob[off:off + L.S_SIZE]

    def components(self, i):
        s = self.stageRecord(i)
        out = []
        for c in range(int(s[L.S_NCOMP])):
            off = int(s[L.S_COMP_OFF + c])
            out.append(self.blob[off:off + int(self.blob[off + L.C_LEN])])
        off = int(s[L.S_AUTO_BT_OFF])
        if off and i == int(self.blob[L.H_NSTAGES]) - 1:      # the bottom stage's automatic zero-length boat tail
            out.append(self.blob[off:off + int(self.blob[off + L.C_LEN])])
        return out


def usStandardAtmosphereBases():
    """USStandardAtmosphere.__init__ (AtmosphereModelling.py:98-145): base temperature/pressure per layer."""
    T0, p0, M, R, G = 288.15, 101325, 28.9644, 8.31432, 9.80665
    baseHeights = [-2000, 11000, 20000, 32000, 47000, 51000, 71000, 84852]
    dt_dh = [-6.5e-3, 0, 1.0e-3, 2.8e-3, 0, -2.8e-3, -2.0e-3, 0]
    temps = [T0 + dt_dh[0] * baseHeights[0]]
    Pb_over_P0 = ((T0 + dt_dh[0] * baseHeights[0]) / T0) ** (-G * M / (R * dt_dh[0] * 1000))
    pressures = [Pb_over_P0 * p0]
    for i in range(1, len(baseHeights)):
        Tb, Pb, lapse = temps[-1], pressures[i - 1], dt_dh[i - 1]
        dh = float(baseHeights[i] - baseHeights[i - 1])
        nextTemp = Tb + lapse * dh
        if lapse == 0.0:
            nextP = math.exp(-G * M * dh / (R * Tb * 1000)) * Pb
        else:
            nextP = ((Tb + lapse * dh) / Tb) ** (-G * M / (R * lapse * 1000)) * Pb
        temps.append(nextTemp)
        pressures.append(nextP)
    return [float(h) for h in baseHeights], dt_dh, temps, pressures


def initialStateLaunchTowerFrame(rocketReader: SubDictReader):
    """rocket.py:251-272 — position, velocity, orientation, angular velocity in the launch-tower frame."""
    initPos = rocketReader.getVector("position")
    initVel = rocketReader.getVector("velocity")
    rotationAxis = rocketReader.tryGetVector("rotationAxis", defaultValue=None)
    if rotationAxis is not None:
        q = hm.q_from_axis_angle(rotationAxis, math.radians(rocketReader.getFloat("rotationAngle")))
    else:
        d = hm.v_normalize(rocketReader.getVector("initialDirection"))
        angle = hm.v_angle((0, 0, 1), d)
        axis = hm.v_cross((0, 0, 1), d)
        q = hm.q_from_axis_angle(axis, angle)
    w = rocketReader.getVector("angularVelocity")
    return initPos, initVel, q, w


def initialStateGlobalFrame(rocketReader: SubDictReader, earth: str, elevation: float, latitude: float, longitude: float, onRail: bool):
    """Environment.convertInitialStateToGlobalFrame (environment.py:105-128): launch-tower frame -> inertial frame of the earth model."""
    initPos, initVel, q0, w0 = initialStateLaunchTowerFrame(rocketReader)
    initPos = hm.v_add(initPos, (0, 0, elevation))
    if onRail:
        w0 = (0.0, 0.0, 0.0)
    if earth in ("Round", "WGS84"):
        # SphericalEarth.convertIntoGlobalFrame (EarthModelling.py:197-227)
        initPos, initVel, q0, w0 = hm.convertIntoGlobalFrame(earth, initPos, initVel, q0, w0, latitude, longitude)
    return initPos, initVel, q0, w0


def _derivedConstants(rec, Aref):
    """Flight-independent quotients of the component force routines (include/mlb_layout.h: *_K_*), evaluated in the reference's order."""
    def quot(a, b):
        return a / b if b != 0 else float("nan")
    ctype = int(rec[L.C_TYPE])
    if ctype == L.CT_NOSECONE:
        rec[L.NC_K_WET] = quot(rec[L.NC_WETTED], Aref)
        rec[L.NC_K_RAD] = quot(quot(rec[L.NC_WETTED], rec[L.NC_LENGTH]), 2 * math.pi)
    elif ctype == L.CT_BODYTUBE:
        rec[L.BT_K_WET] = quot(rec[L.BT_WETTED], Aref)
        rec[L.BT_K_RAD] = quot(quot(rec[L.BT_WETTED], rec[L.BT_LENGTH]), 2 * math.pi)
        rec[L.BT_K_PLAN] = quot(rec[L.BT_PLANFORM], Aref)
        rec[L.BT_DL] = rec[L.BT_LENGTH] / 25
        dL = rec[L.BT_DL]
        for i in range(25):
            rec[L.BT_STRIP_Z + i] = rec[L.C_POS + 2] + (-dL * i - dL / 2)
    elif ctype in (L.CT_BOATTAIL, L.CT_TRANSITION):
        rec[L.TR_K_WET] = quot(rec[L.TR_WETTED], Aref)
        rec[L.TR_K_RAD] = quot(quot(rec[L.TR_WETTED], rec[L.TR_LENGTH]), 2 * math.pi)
        rec[L.TR_K_FRONT] = quot(rec[L.TR_FRONTAL], Aref)
    elif ctype == L.CT_FINSET:
        rec[L.FS_K_SKIN] = quot(rec[L.FS_WETTED], Aref)
        rec[L.FS_K_THICKF] = 1 + 2 * rec[L.FS_THICK] / rec[L.FS_MAC_LEN]
        cs = math.cos(rec[L.FS_SWEEP])
        rec[L.FS_K_LE] = quot(rec[L.FS_LE_THICK] * rec[L.FS_SPAN] * (cs * cs), Aref)
        rec[L.FS_TC] = rec[L.FS_THICK] / rec[L.FS_ROOT]
        rec[L.FS_K_PLAN] = quot(rec[L.FS_PLANFORM], Aref)
        rec[L.FS_COS_MID] = math.cos(rec[L.FS_MIDSWEEP])
        rec[L.FS_RCOS_MID] = 1 / rec[L.FS_COS_MID]
        rec[L.FS_AR2] = 2 * (rec[L.FS_SPAN] * rec[L.FS_SPAN]) / rec[L.FS_PLANFORM]
        r = rec[L.FS_THICK] / rec[L.FS_MAC_LEN]
        rec[L.FS_K_THICK_SUP] = 2.3 * rec[L.FS_AR] * (r * r)
    return rec


def _flattenTableau(rows, nResult):
    """Stage rows, result rows, then (device only) the result rows again as the weights the reference's object algebra ends up applying:
    derivative * c is evaluated as derivative / (1 / c) and the division as * (1 / (1 / c)) (Integration.py:233-236,434-447)."""
    flat = []
    for row in rows:
        flat.extend(row + [0.0] * (L.TABLEAU_STRIDE - len(row)))
    with np.errstate(divide="ignore"):
        for row in rows[len(rows) - nResult:]:
            eff = [float(np.float64(1) / (np.float64(1) / np.float64(c))) for c in row]
            flat.extend(eff + [0.0] * (L.TABLEAU_STRIDE - len(row)))
    return flat


def buildModel(simDefinition: SimDefinition, maxSteps: int = 400000, droppedStage: int = None, droppedFromAltitude: float = None, groundWind=None,
               windProfile=None, windProfileRows: int = None) -> Model:
    """Everything the reference's constructors precompute, flattened into one model block.
    droppedStage = n builds the rocket the reference creates for the n-th dropped stage (`Simulation.createRocket(stage=n)`,
    SingleSimulations.py:161-176,280-299; Rocket.__init__ with stageToInitialize, rocket.py:296-321,339-361): that stage alone, its motor burnt
    out, no staging triggers, no control system, the end condition of SimControl.StageDropPaths; droppedFromAltitude is the altitude of the
    separation (the Altitude end condition looks at it to decide between ascending and descending, SingleSimulations.py:240-247). The lanes
    of such a model start from the separation states (LANE_INIT_STATE, LANE_START)."""
    sd = simDefinition
    env = SubDictReader("Environment", sd)
    rk = SubDictReader("Rocket", sd)
    H = [0.0] * L.H_SIZE
    tail: List[float] = []          # everything after the header; offsets are H_SIZE + index

    def append(rec) -> int:
        off = L.H_SIZE + len(tail)
        tail.extend(float(x) for x in rec)
        return off

    H[L.H_MAGIC] = L.MLB_MAGIC
    H[L.H_VERSION] = L.MLB_VERSION

    # ---- integrator (Integration.py:56-103) ----
    method = rk.getString("SimControl.timeDiscretization")
    if method not in _INTEGRATORS:
        raise ValueError("Integration method: {} not implemented. See SimDefinitionTemplate.txt for options.".format(method))
    H[L.H_INTEGRATOR] = _INTEGRATORS[method]
    H[L.H_DT0] = float(sd.getValue("SimControl.timeStep"))
    rows, nResult = butcherTableau(method)
    if rows is not None:
        H[L.H_TABLEAU_OFF] = append(_flattenTableau(rows, nResult))
        H[L.H_TAB_NSTAGE_ROWS] = len(rows) - nResult
        H[L.H_TAB_NRESULT_ROWS] = nResult
    if "Adapt" in method:
        ad = SubDictReader("SimControl.TimeStepAdaptation", sd)
        controller = ad.getString("controller")
        H[L.H_AD_SAFETY] = 1.0
        if controller == "PID":
            H[L.H_AD_CTRL] = L.CTRL_PID
            H[L.H_AD_P], H[L.H_AD_I], H[L.H_AD_D] = [float(x) for x in ad.getString("PID.coefficients").split()]
        elif controller == "elementary":
            H[L.H_AD_CTRL] = L.CTRL_ELEMENTARY
            H[L.H_AD_SAFETY] = ad.getFloat("Elementary.safetyFactor")
        else:
            # "Constant" included: the reference's integratorFactory leaves safetyFactor unbound for it and the rocket cannot be
            # constructed (Integration.py:74-93); AdaptiveIntegrator itself raises this message for it (:352-357)
            raise ValueError("Adaptive Integrator requires non-constant step size controller such as 'PID' or 'elementary'")
        H[L.H_AD_TARGET] = ad.getFloat("targetError")
        H[L.H_AD_MINF] = ad.getFloat("minFactor")
        H[L.H_AD_MAXF] = ad.getFloat("maxFactor")
        H[L.H_AD_MAXDT] = ad.getFloat("maxTimeStep")
        H[L.H_AD_MINDT] = ad.getFloat("minTimeStep")
    H[L.H_EVENT_DT] = rk.getFloat("SimControl.TimeStepAdaptation.eventTimingAccuracy")
    # the reference has no step limit (a rocket held aloft by an updraft runs forever); a lane that reaches this many
    # steps ends with status ST_STEP_LIMIT. Longest flight of the example cases: 56 k steps (RK4, dt 0.01, to ground).
    H[L.H_MAX_STEPS] = maxSteps
    # inverse Hermite matrix of the transonic fin-force blend between Mach 0.8 and 1.4 (Fins.py:508-528)
    H[L.H_FIN_CUBIC_OFF] = append(cubicInterpInverseMatrix(0.8, 1.4).reshape(-1))

    # ---- earth / atmosphere / wind / turbulence (ENV/*.py factories) ----
    earth = env.getString("EarthModel")
    if earth == "Flat":
        H[L.H_EARTH], H[L.H_EARTH_GM], H[L.H_EARTH_RADIUS] = L.EARTH_FLAT, 398600.5e9, 6371000
    elif earth == "None":
        H[L.H_EARTH] = L.EARTH_NONE
    elif earth == "Round":
        H[L.H_EARTH], H[L.H_EARTH_GM], H[L.H_EARTH_RADIUS] = L.EARTH_ROUND, hm.SPHERE_GM, hm.SPHERE_RADIUS
    elif earth == "WGS84":
        H[L.H_EARTH], H[L.H_EARTH_GM], H[L.H_EARTH_RADIUS] = L.EARTH_WGS84, hm.SPHERE_GM, hm.SPHERE_RADIUS
        H[L.H_WGS_A], H[L.H_WGS_E2] = hm.WGS_A, hm.WGS_E2
    else:
        raise NotImplementedError("Earth model: {} not found. Try 'Flat', 'Round' or 'WGS84'".format(earth))
    H[L.H_EARTH_ROT] = hm.EARTH_ROTATION_RATE if earth in ("Round", "WGS84") else 0.0
    H[L.H_ELEV] = env.tryGetFloat("LaunchSite.elevation")
    H[L.H_LAT] = env.tryGetFloat("LaunchSite.latitude")
    H[L.H_LON] = env.tryGetFloat("LaunchSite.longitude")

    atm = env.getString("AtmosphericPropertiesModel")
    if atm == "Constant":
        H[L.H_ATM] = L.ATM_CONSTANT
        H[L.H_ATM_CONST:L.H_ATM_CONST + 4] = [env.getFloat("ConstantAtmosphere.temp") + 273.15, env.getFloat("ConstantAtmosphere.pressure"),
                                             env.getFloat("ConstantAtmosphere.density"), env.getFloat("ConstantAtmosphere.viscosity")]
    elif atm == "USStandardAtmosphere":
        H[L.H_ATM] = L.ATM_USSTD
        bh, lapse, temps, pres = usStandardAtmosphereBases()
        # rows 4-7 (device only): pressure-ratio exponent -G M / (R lapse 1000), 1 / Tb, the isothermal divisor R Tb 1000 and its reciprocal
        G, M, R = 9.80665, 28.9644, 8.31432
        expo = [(-G * M / (R * lp * 1000)) if lp != 0 else 0.0 for lp in lapse]
        H[L.H_USSTD_OFF] = append(bh + lapse + temps + pres + expo + [1 / t for t in temps] + [R * t * 1000 for t in temps] + [1 / (R * t * 1000) for t in temps])
    elif atm == "TabulatedAtmosphere":
        path = env.getString("TabulatedAtmosphere.filePath")
        table = np.loadtxt(path, skiprows=1)
        table[:, 4] *= 1E-5
        H[L.H_ATM] = L.ATM_TABULATED
        H[L.H_ATM_TAB_N] = table.shape[0]
        H[L.H_ATM_TAB_OFF] = append(table.T.reshape(-1))      # column-major: h, T, P, rho, mu
    else:
        raise ValueError("Atmospheric model: {} not implemented, try using 'USStandardAtmosphere'".format(atm))

    wind = env.getString("MeanWindModel")
    # ground wind drawn from wind roses (MeanWindModelling.py:41-93): a host-side draw per constructed Environment. A Monte Carlo run passes
    # the winds it drew per lane (groundWind / LANE_WIND); building a model on its own draws one here, as constructing the Environment would
    sampledGroundWind = None
    if wind == "SampledGroundWindData" or (wind == "Hellman" and env.getString("Hellman.groundWindModel") == "SampledGroundWindData"):
        if groundWind is not None:
            sampledGroundWind = list(groundWind)
        else:
            from .windsampling import samplerFromDefinition
            sampledGroundWind = samplerFromDefinition(env).sample()
    if wind == "Constant":
        H[L.H_WIND] = L.WIND_CONSTANT
        H[L.H_WIND_V:L.H_WIND_V + 3] = env.getVector("ConstantMeanWind.velocity")
    elif wind == "SampledGroundWindData":
        H[L.H_WIND] = L.WIND_CONSTANT                      # "Wind is not a function of altitude" (:74-78)
        H[L.H_WIND_V:L.H_WIND_V + 3] = sampledGroundWind
    elif wind == "Hellman":
        groundModel = env.getString("Hellman.groundWindModel")
        if groundModel not in ("Constant", "SampledGroundWindData"):
            raise UnboundLocalError("cannot access local variable 'meanGroundWind' where it is not associated with a value")      # the reference's failure (:80-93)
        H[L.H_WIND] = L.WIND_HELLMAN
        H[L.H_WIND_V:L.H_WIND_V + 3] = env.getVector("ConstantMeanWind.velocity") if groundModel == "Constant" else sampledGroundWind
        H[L.H_HELLMAN_ALPHA] = env.getFloat("Hellman.alphaCoeff")
        H[L.H_HELLMAN_LIMIT] = env.getFloat("Hellman.altitudeLimit")
    elif wind == "CustomWindProfile":
        prof = np.loadtxt(env.getString("CustomWindProfile.filePath"), skiprows=1)
        H[L.H_WIND] = L.WIND_INTERPOLATED
        H[L.H_WIND_TAB_N] = prof.shape[0]
        H[L.H_WIND_TAB_OFF] = append(prof.T.reshape(-1))
    elif wind == "SampledRadioSondeData":
        # one measured profile per constructed Environment (MeanWindModelling.py:96-118): given by the Monte Carlo runner, else drawn here.
        # windProfileRows pads the table with copies of its last row (linInterp returns the last row beyond it, Interpolation.py:22-25), so
        # that the per-lane model blocks of a run share one layout whatever the sounding's length
        if windProfile is None:
            from .windsampling import samplerFromDefinition
            windProfile = samplerFromDefinition(env).sample()
        alt, vec = np.asarray(windProfile[0], dtype=np.float64), np.asarray(windProfile[1], dtype=np.float64).reshape(-1, 3)
        prof = np.column_stack([alt, vec])
        if windProfileRows is not None and windProfileRows > prof.shape[0]:
            prof = np.vstack([prof] + [prof[-1:]] * (windProfileRows - prof.shape[0]))
        H[L.H_WIND] = L.WIND_INTERPOLATED
        H[L.H_WIND_TAB_N] = prof.shape[0]
        H[L.H_WIND_TAB_OFF] = append(prof.T.reshape(-1))
    else:
        raise NotImplementedError("MeanWindModel '{}' is not implemented by the B200 backend (SURVEY §8f)".format(wind))

    turb = env.getString("Environment.TurbulenceModel")
    H[L.H_TURB_INTENSITY] = float("nan")
    H[L.H_TURB_STD] = float("nan")
    if turb == "None":
        H[L.H_TURB] = L.TURB_NONE
    elif "PinkNoise" in turb:
        H[L.H_TURB] = {"PinkNoise1D": L.TURB_PINK1D, "PinkNoise2D": L.TURB_PINK2D, "PinkNoise3D": L.TURB_PINK3D}[turb]
        ti = env.tryGetInt("PinkNoiseModel.turbulenceIntensity")
        if ti is not None:
            H[L.H_TURB_INTENSITY] = ti / (2.26 * 100)
        sdv = env.tryGetInt("PinkNoiseModel.velocityStdDeviation")
        if sdv is not None:
            H[L.H_TURB_STD] = sdv / 2.26
        coeffs = [1, ]
        alpha = 5 / 3
        for i in range(2):
            coeffs.append((i - alpha / 2) * coeffs[i] / (i + 1))
        H[L.H_PINK_C0], H[L.H_PINK_C1] = coeffs[1], coeffs[2]
        hasSeed = 0
        for i in range(3):
            s = env.tryGetInt("PinkNoiseModel.randomSeed{}".format(i + 1))
            if s is not None:
                H[L.H_PINK_SEED + i] = s
                hasSeed |= (1 << i)
        H[L.H_PINK_HAS_SEED] = hasSeed
    elif turb == "customSineGust":
        H[L.H_TURB] = L.TURB_SINE_GUST
        H[L.H_GUST_START] = env.getFloat("CustomSineGust.startAltitude")
        H[L.H_GUST_MAG] = env.getFloat("CustomSineGust.magnitude")
        H[L.H_GUST_BLEND] = env.getFloat("CustomSineGust.sineBlendDistance")
        H[L.H_GUST_THICK] = env.getFloat("CustomSineGust.thickness")
        H[L.H_GUST_DIR:L.H_GUST_DIR + 3] = hm.v_normalize(env.getVector("CustomSineGust.direction"))
    else:
        raise ValueError("Turbulence Model {} not found. Please choose from: None, PinkNoise1/2/3D or customSineGust")
    H[L.H_TURB_OFF_CHUTE] = 1.0 if rk.getBool("Environment.turbulenceOffWhenUnderChute") else 0.0

    # ---- initial state, launch rail (rocket.py:251-291, environment.py:57-91,105-128) ----
    railLength = env.getFloat("LaunchSite.railLength")
    H[L.H_RAIL_LEN] = railLength
    initPos, initVel, q0, w0 = initialStateGlobalFrame(rk, earth, H[L.H_ELEV], H[L.H_LAT], H[L.H_LON], railLength > 0)
    if railLength > 0:
        # the rail is aligned with the rocket's initial orientation: rotationAxis / rotationAngle when given, else initialDirection
        # (environment.py:63-80)
        railAxis = env.tryGetVector("Rocket.rotationAxis", defaultValue=None)
        if railAxis is not None:
            qr = hm.q_from_axis_angle(railAxis, math.radians(env.getFloat("Rocket.rotationAngle")))
        else:
            d = hm.v_normalize(env.getVector("Rocket.initialDirection"))
            qr = hm.q_from_axis_angle(hm.v_cross((0, 0, 1), d), hm.v_angle((0, 0, 1), d))
        railPos = hm.v_add(env.getVector("Rocket.position"), (0, 0, H[L.H_ELEV]))
        if earth in ("Round", "WGS84"):
            railPos, _, qr, _ = hm.convertIntoGlobalFrame(earth, railPos, (0, 0, 0), qr, (0, 0, 0), H[L.H_LAT], H[L.H_LON])
        railDir, _ = hm.q_rotate(qr, (0, 0, 1))
        H[L.H_RAIL_POS:L.H_RAIL_POS + 3] = railPos
        H[L.H_RAIL_DIR:L.H_RAIL_DIR + 3] = hm.v_normalize(railDir)

    # ---- rocket-level constants (rocket.py:100-140,221-249) ----
    rocketRoughness = rk.getFloat("Aero.surfaceRoughness")
    H[L.H_ROCKET_ROUGHNESS] = rocketRoughness
    H[L.H_FULLY_TURB] = 1.0 if rk.getBool("Aero.fullyTurbulentBL") else 0.0
    addAutoBT = rk.getBool("Aero.addZeroLengthBoatTailsToAccountForBaseDrag")
    H[L.H_ADD_AUTO_BT] = 1.0 if addAutoBT else 0.0
    stageDicts = [d for d in rk.getImmediateSubDicts() if d not in ("Rocket.ControlSystem", "Rocket.HIL", "Rocket.Aero")]
    # rocket.py:210-217: a control system exists when a controlled system is named OR the moment controller is the ideal one (which has
    # no controlled system: ControlSystems.py:85-100)
    idealControl = rk.tryGetString("ControlSystem.MomentController.Type") == "IdealMomentController"
    hasActuatedSystem = rk.tryGetString("ControlSystem.controlledSystem") is not None and not idealControl
    hasControl = hasActuatedSystem or idealControl
    if droppedStage is not None:
        hasActuatedSystem = idealControl = hasControl = False          # rocket.py:215: only a complete rocket gets the control system
    maxDiameter = 0
    for sdict in stageDicts:
        for cdict in sd.getImmediateSubDicts(sdict):
            if sd.getValue(cdict + ".class") == "Bodytube":
                maxDiameter = max(maxDiameter, float(sd.getValue(cdict + ".outerDiameter")))
    H[L.H_MAXDIAM] = maxDiameter
    H[L.H_AREF] = math.pi * maxDiameter ** 2 / 4
    H[L.H_RAREF] = (1 / H[L.H_AREF]) if H[L.H_AREF] != 0 else float("inf")

    if droppedStage is not None:
        # rocket.py:296-321: the n-th smallest stageNumber; the reference area above stays that of the fully assembled rocket (:100-108)
        numbers = [float(sd.getValue(d + ".stageNumber")) for d in stageDicts]
        if len(set(numbers)) != len(numbers):
            raise ValueError("For multi-stage rockets, each stage must have a unique stage number. Set the Rocket.StageName.stageNumber value for each stage. 0 for first stage, 1 for second, etc...")
        wanted = sorted(numbers)[droppedStage]
        stageDicts = [d for d, n in zip(stageDicts, numbers) if n == wanted]

    # ---- stages and components (stage.py:13-70; sort rules: stage.py:47-59, RocketComponents.py:131-143) ----
    stages = []
    for sdict in stageDicts:
        sr = SubDictReader(sdict, sd)
        stagePos = sr.getVector("position")
        compDicts = sd.getImmediateSubDicts(sdict)
        for c in compDicts:
            cls = sd.getValue(c + ".class")
            if cls not in _SUPPORTED_CLASSES:
                raise NotImplementedError("Rocket component class '{}' ({}) is not implemented by the B200 backend".format(cls, c))
        compDicts.sort(key=lambda c: (0 if sd.getValue(c + ".class") in _BODY_CLASSES else 1, c))
        comps = []
        for c in compDicts:
            cr = SubDictReader(c, sd)
            cls = cr.getString("class")
            if cls == "Nosecone":
                rec, info = _buildNosecone(cr, stagePos, rocketRoughness)
            elif cls == "Bodytube":
                rec, info = _buildBodytube(cr, stagePos, rocketRoughness)
            elif cls == "BoatTail":
                rec, info = _buildTransition(cr, stagePos, rocketRoughness, L.CT_BOATTAIL)
            elif cls == "Transition":
                rec, info = _buildTransition(cr, stagePos, rocketRoughness, L.CT_TRANSITION)
            elif cls == "FinSet":
                rec, info = _buildFinSet(cr, stagePos, rocketRoughness)
            elif cls == "Motor":
                rec, info = _buildMotor(cr, stagePos)
            elif cls == "RecoverySystem":
                rec, info = _buildRecovery(cr, stagePos)
            elif cls == "Force":
                rec, info = _buildFixedForce(cr, stagePos)
            elif cls == "AeroForce":
                rec, info = _buildAeroForce(cr, stagePos)
            elif cls == "AeroDamping":
                rec, info = _buildAeroDamping(cr, stagePos)
            elif cls == "TabulatedAeroForce":
                rec, info = _buildTabulatedAeroForce(cr, stagePos)
            elif cls == "TabulatedInertia":
                rec, info = _buildTabulatedInertia(cr, stagePos)
            elif cls == "FractionalJetDamping":
                rec, info = _buildJetDamping(cr, stagePos)
            else:
                rec, info = _buildMass(cr, stagePos)
            if hasActuatedSystem and cls == "FinSet" and c == rk.getString("ControlSystem.controlledSystem"):
                rec[L.FS_ACTUATED] = 1.0          # fin angles follow the actuators (Fins.py:264-268)
            info["name"] = cr.getDictName()
            info["cls"] = cls
            info["rec"] = rec
            info["order"] = len(comps)          # construction order (body components first, then by dictionary path)
            comps.append(info)
        # fixed-mass inertia pre-sum in construction order (CompositeObject.py:52-63)
        fixed = [c["inertia"] for c in comps if c["cls"] in _FIXED_MASS_CLASSES]
        fixedInertia = Inertia.combine(fixed) if fixed else Inertia((0, 0, 0), (0, 0, 0), 0)
        motors = [c for c in comps if c["cls"] == "Motor"]
        if len(motors) > 1:
            raise NotImplementedError("more than one Motor per stage")
        variable = [c for c in comps if c["cls"] in ("Motor", "TabulatedInertia")]      # construction order
        if len(variable) > L.S_MAX_VAR:
            raise NotImplementedError("more than {} variable-inertia components per stage".format(L.S_MAX_VAR))
        if any(c["cls"] == "FractionalJetDamping" for c in comps) and not motors:
            raise AttributeError("'NoneType' object has no attribute 'ignitionTime'")      # the reference's failure mode (RocketComponents.py:395)
        cgOvr = sr.tryGetVector("constCG", defaultValue=None)
        if cgOvr is not None:
            cgOvr = hm.v_add(cgOvr, stagePos)
        stages.append(dict(name=sr.getDictName(), reader=sr, pos=stagePos, comps=comps, fixed=fixedInertia,
                           motor=(motors[0] if motors else None), variable=variable, cgOvr=cgOvr, massOvr=sr.tryGetFloat("constMass", defaultValue=None),
                           moiOvr=sr.tryGetVector("constMOI", defaultValue=None), stageNumber=sr.getInt("stageNumber"),
                           sepType=sr.getString("separationTriggerType"), sepDelay=sr.getFloat("separationDelay"),
                           sepValue=sr.getFloat("separationTriggerValue")))

    def zKey(c):
        return c["position"][2]
    for st in stages:
        st["comps"].sort(key=zKey, reverse=True)      # stable, top -> bottom
    stages.sort(key=lambda s: s["pos"][2], reverse=True)
    if len(stages) > 4:
        raise NotImplementedError("more than 4 stages")
    if sum(1 for st in stages for c in st["comps"] if c["cls"] == "RecoverySystem") > 1:
        raise NotImplementedError("more than one RecoverySystem in a rocket (the reference keeps only the last one created as Rocket.recoverySystem)")
    if droppedStage is not None:
        for st in stages:
            st["sepType"] = "None"                    # rocket.py:357-361: a dropped stage has no staging events, its motor is burnt out
            if st["motor"] is None:
                raise AttributeError("'NoneType' object has no attribute 'updateIgnitionTime'")
    ignition0 = (lambda st: -1000000000) if droppedStage is not None else (lambda st: 0.0 if st is stages[-1] else 1000000000)
    for st in stages[:-1]:
        if st["motor"] is None:
            raise AttributeError("'NoneType' object has no attribute 'updateIgnitionTime'")      # the reference's failure mode (rocket.py:356)

    # auto-added zero-length boat tail for base drag (rocket.py:489-519)
    def bottomInterfaceZ(st):
        for c in reversed(st["comps"]):
            if c.get("body"):
                return c["position"][2] - c["length"] if c.get("canConnectBelow", True) else None
        return None
    # Every stage that can become the bottom stage gets its record; the kernels evaluate it (after the stage's own components,
    # where the reference appends it) only while the stage IS the bottom one: at launch for the last stage, after each
    # separation for the new last stage (Rocket._stageSeparation -> _ensureBaseDragIsAccountedFor, rocket.py:461-474).
    for st in stages:
        st["autoBT"] = None
        if maxDiameter > 0 and addAutoBT:
            hasBT, sawBody = False, False
            for c in reversed(st["comps"]):
                if c.get("body"):
                    hasBT, sawBody = bool(c.get("isBoatTail")), True
                    break
            if sawBody and not hasBT:
                z = bottomInterfaceZ(st)
                pos = (st["pos"][0], st["pos"][1], z)
                st["autoBT"] = _autoBoatTailRecord(maxDiameter, pos, rocketRoughness)

    # fin sets need the body radius at their mounting station (Fins.py:107-110, stage.py:126-143)
    for st in stages:
        for c in st["comps"]:
            if c["cls"] == "FinSet":
                desiredZ = st["pos"][2] - (st["pos"][2] - c["position"][2])
                radius = None
                for b in st["comps"]:
                    if "length" in b:
                        top = b["position"][2]
                        bot = top - b["length"]
                        if desiredZ <= top and desiredZ >= bot:
                            radius = b["getRadius"](top - desiredZ)
                            break
                if radius is None:
                    raise ValueError("Length {} from the top of stage '{}' does not fall between the top and bottom of any of its components".format(
                        st["pos"][2] - c["position"][2], st["name"]))
                _finishFinSet(c["rec"], c["finset"], radius)

    # fineness ratio (rocket.py:521-528)
    def stageLength(st):
        bodies = [c for c in st["comps"] if c.get("body")]
        if not bodies:
            return None
        return bodies[0]["position"][2] - (bodies[-1]["position"][2] - bodies[-1]["length"])
    for nRemaining in range(1, len(stages) + 1):
        sub = stages[:nRemaining]
        length = 0
        for st in sub:
            ln = stageLength(st)
            if ln is not None:
                length += ln
        md = max(max([c.get("maxDiameter", 0) for c in st["comps"]] + [0]) for st in sub)
        if sub[-1]["autoBT"] is not None:
            md = max(md, maxDiameter)
        H[L.H_FINENESS + nRemaining - 1] = (length / md) if md != 0 else float("nan")
        fr = H[L.H_FINENESS + nRemaining - 1]
        H[L.H_FINENESS_F + nRemaining - 1] = (1 + (0.5 / fr)) if fr != 0 else float("inf")

    # ---- serialise stages ----
    H[L.H_NSTAGES] = len(stages)
    H[L.H_RECOVERY_OFF] = 0
    # mass / CG / MOI cannot change in flight when every stage is motor-less or fully overridden (stage.py:33-41):
    # the CG-migration term of the state derivative (RigidBodies.py:91-97) is then exactly zero
    H[L.H_INERTIA_CONST] = 1.0 if all(not st["variable"] or (st["cgOvr"] is not None and st["massOvr"] is not None and st["moiOvr"] is not None)
                                      for st in stages) else 0.0
    for si, st in enumerate(stages):
        S = [0.0] * L.S_SIZE
        if len(st["comps"]) > L.S_MAX_COMP:
            raise NotImplementedError("more than {} components per stage".format(L.S_MAX_COMP))
        S[L.S_NCOMP] = len(st["comps"])
        S[L.S_POS:L.S_POS + 3] = st["pos"]
        sepTypes = {"None": L.EV_NONE, "apogee": L.EV_APOGEE, "ascendingThroughAltitude": L.EV_ASCENDING_ALT,
                    "descendingThroughAltitude": L.EV_DESCENDING_ALT, "motorBurnout": L.EV_MOTOR_BURNOUT, "timeReached": L.EV_TIME_REACHED}
        S[L.S_SEP_TYPE] = sepTypes[st["sepType"]]
        S[L.S_SEP_VALUE] = st["sepValue"]
        S[L.S_SEP_DELAY] = st["sepDelay"]
        flags = 0
        if st["cgOvr"] is not None:
            flags |= L.OVR_CG
            S[L.S_OVR_CG:L.S_OVR_CG + 3] = st["cgOvr"]
        if st["massOvr"] is not None:
            flags |= L.OVR_MASS
            S[L.S_OVR_MASS] = st["massOvr"]
        if st["moiOvr"] is not None:
            flags |= L.OVR_MOI
            S[L.S_OVR_MOI:L.S_OVR_MOI + 3] = st["moiOvr"]
        S[L.S_OVR_FLAGS] = flags
        S[L.S_FIXED_MOI:L.S_FIXED_MOI + 3] = st["fixed"].MOI
        S[L.S_FIXED_CG:L.S_FIXED_CG + 3] = st["fixed"].CG
        S[L.S_FIXED_MASS] = st["fixed"].mass
        S[L.S_MOTOR] = -1
        S[L.S_BURN_DURATION] = -1
        S[L.S_IGNITION0] = ignition0(st)
        for ci, c in enumerate(st["comps"]):
            off = append(_derivedConstants(c["rec"], H[L.H_AREF]))
            S[L.S_COMP_OFF + ci] = off
            c["blobOffset"] = off
            if c["cls"] == "Motor":
                S[L.S_MOTOR] = ci
                S[L.S_BURN_DURATION] = c["motor"]["engineShutOff"]
            if c["cls"] == "RecoverySystem":
                H[L.H_RECOVERY_OFF] = off
        S[L.S_NVAR] = len(st["variable"])
        for k, c in enumerate(st["variable"]):
            S[L.S_VAR_IDX + k] = st["comps"].index(c)
        for k, c in enumerate(sorted(st["variable"], key=lambda c: st["comps"].index(c))):
            S[L.S_VAR2_IDX + k] = st["comps"].index(c)
        fixedSorted = [c["inertia"] for c in st["comps"] if c["cls"] in _FIXED_MASS_CLASSES]
        fixed2 = Inertia.combine(fixedSorted) if fixedSorted else Inertia((0, 0, 0), (0, 0, 0), 0)
        S[L.S_FIXED2_MOI:L.S_FIXED2_MOI + 3] = fixed2.MOI
        S[L.S_FIXED2_CG:L.S_FIXED2_CG + 3] = fixed2.CG
        S[L.S_FIXED2_MASS] = fixed2.mass
        if st["autoBT"] is not None:
            S[L.S_AUTO_BT_OFF] = append(_derivedConstants(st["autoBT"], H[L.H_AREF]))
        H[L.H_STAGE_OFF + si] = append(S)

    # ---- control system (rocket.py:210-217; GNC/ControlSystems.py:45-153) ----
    H[L.H_SHARED_LEN] = L.H_SIZE + len(tail)          # everything so far is staged in shared memory
    if hasControl:
        H[L.H_CTRL_OFF] = append(_buildControlSystem(sd, rk, stages, method, H, append))
        H[L.H_SHARED_LEN] = L.H_SIZE + len(tail)      # ... including the control record and the RK4 tableau
        if hasActuatedSystem:
            _appendInterpolators(sd, rk, H, tail, append)
    for st in stages:                                  # multi-key aero tables: Delaunay records in the global part of the block
        for c in st["comps"]:
            if c.get("ndTable") is not None:
                tail[c["blobOffset"] - L.H_SIZE + L.TA_ND_OFF] = append(_ndInterpolatorRecord(*c["ndTable"]))

    # ---- initial inertia -> CG shift of the integrated position (rocket.py:416-421) ----
    def stageInertia(st, t):
        items = [st["fixed"]]
        for c in st["variable"]:
            if c["cls"] == "Motor":
                items.append(motorInertia(c["motor"], max(0, t - ignition0(st))))
            else:
                items.append(tabulatedInertia(c["tabInertia"], t))
        comb = Inertia.combine(items)
        if st["cgOvr"] is not None:
            comb.CG = st["cgOvr"]
        if st["moiOvr"] is not None:
            comb.MOI = st["moiOvr"]
        if st["massOvr"] is not None:
            comb.mass = st["massOvr"]
        return comb
    zero = Inertia((0, 0, 0), (0, 0, 0), 0)
    items = [zero] + [stageInertia(st, 0) for st in stages]
    rocketInertia0 = Inertia.combine(items)
    cgGlobal, q0n = hm.q_rotate(q0, rocketInertia0.CG)
    posCG = hm.v_add(initPos, cgGlobal)
    H[L.H_NOSE_POS0:L.H_NOSE_POS0 + 3] = initPos
    H[L.H_INIT_CG0:L.H_INIT_CG0 + 3] = rocketInertia0.CG
    H[L.H_INIT_STATE:L.H_INIT_STATE + 13] = list(posCG) + list(initVel) + list(q0n) + list(w0)

    # ---- end condition (SingleSimulations.py:205-248) ----
    if droppedStage is None:
        endCondition = sd.getValue("SimControl.EndCondition")
        endValue = float(sd.getValue("SimControl.EndConditionValue"))
        startZ = posCG[2]
    else:
        endCondition = sd.getValue("SimControl.StageDropPaths.endCondition")
        endValue = float(sd.getValue("SimControl.StageDropPaths.endConditionValue"))
        startZ = posCG[2] if droppedFromAltitude is None else droppedFromAltitude
    H[L.H_END_VALUE] = endValue
    if endCondition == "Apogee":
        H[L.H_END_TYPE] = L.END_APOGEE
    elif endCondition == "Altitude" and startZ < endValue:
        H[L.H_END_TYPE] = L.END_ALT_ASCENDING
    elif endCondition == "Altitude":
        H[L.H_END_TYPE] = L.END_ALT_DESCENDING
    else:
        H[L.H_END_TYPE] = L.END_TIME

    blob = np.array(H + tail, dtype=np.float64)
    blob[L.H_LEN] = blob.size
    meta = dict(stages=[dict(name=st["name"], components=[(c["name"], c["cls"]) for c in st["comps"]] +
                             ([("Auto-AddedZeroLengthBoatTail", "BoatTail")] if (st is stages[-1] and st["autoBT"] is not None) else []))
                        for st in stages],
                initialInertia=rocketInertia0, method=method, earth=earth,
                # per stage: definition path and (impulseAdjustFactor, burnTimeAdjustFactor) of its motor, for LANE_MOTOR_ADJUST
                motors=[next((dict(path=c["path"], factors=c["motorFactors"]) for c in st["comps"] if c["cls"] == "Motor"), None) for st in stages])
    return Model(blob, meta)
