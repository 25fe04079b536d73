/*
 * mlb_model.cuh — per-lane flight model evaluated inside the trajectory kernels (device side, FP64):
 * environment, component build-up aerodynamics, motor, recovery, gravity, launch rail and the
 * rigid-body state derivative. One lane = one sampled rocket; `B` is the model block of
 * include/mlb_layout.h staged in shared memory (every lane of a CTA reads the same constants, so the
 * reads are shared-memory broadcasts; table rows selected per lane are ordinary LDS).
 *
 * Reference behaviour restated here (paths relative to MAPLEAF/):
 *   ENV/environment.py:146-236, ENV/AtmosphereModelling.py:147-182, ENV/MeanWindModelling.py:133-163,
 *   ENV/TurbulenceModelling.py:84-166, ENV/EarthModelling.py:93-105, ENV/launchRail.py:37-81
 *   Motion/AeroParameters.py:15-163, Motion/forceMomentSystem.py:25-51, Motion/inertia.py:78-115,
 *   Motion/Interpolation.py:15-41,86-107
 *   Rocket/AeroFunctions.py:38-289, noseCone.py:173-212, bodyTube.py:35-97, boatTail.py:127-217,
 *   Fins.py:26-43,263-386,463-551, CythonFinFunctions.pyx:10-116, Propulsion.py:130-204, Recovery.py:50-63
 *   Rocket/rocket.py:531-612, Rocket/CompositeObject.py:82-139, Motion/RigidBodies.py:41-47,88-115
 *
 * Work the reference repeats per evaluation but whose value cannot change is evaluated once here
 * (same arithmetic, same result): fin-slice angles of attack are shared by the four strip sums of the
 * transonic blend and by the mid-span drag angle; the rocket inertia is evaluated once per time instead
 * of three times; roughness-limited skin-friction constants come precomputed in the model block.
 */
#pragma once
#include "../../include/mlb_layout.h"
#include "mlb_math.cuh"
#include "mlb_rng.cuh"

namespace mlb {

/* Component force routines are real (out-of-line) device functions: each gets its own register allocation instead of adding its
   live ranges to the already over-subscribed derivative evaluation (64 registers per lane at 1024 lanes per CTA). Measured +10 %
   against inlining them; pushing more out of line (environment, events, state algebra, the derivative itself) was slower again. */
#ifndef MLB_COMP_NOINLINE
#define MLB_COMP_NOINLINE 15          /* bit mask: 1 nose cone, 2 body tube, 4 boat tail / transition, 8 fin set */
#endif
#if MLB_COMP_NOINLINE & 1
#define MLB_COMP_NOSE static __device__ __noinline__
#else
#define MLB_COMP_NOSE __device__ __forceinline__
#endif
#if MLB_COMP_NOINLINE & 2
#define MLB_COMP_TUBE static __device__ __noinline__
#else
#define MLB_COMP_TUBE __device__ __forceinline__
#endif
#if MLB_COMP_NOINLINE & 4
#define MLB_COMP_TAIL static __device__ __noinline__
#else
#define MLB_COMP_TAIL __device__ __forceinline__
#endif
#if MLB_COMP_NOINLINE & 8
#define MLB_COMP_FINS static __device__ __noinline__
#else
#define MLB_COMP_FINS __device__ __forceinline__
#endif

/* MLB_MAX_STAGES: include/mlb_layout.h */
#define MLB_MAX_FINSETS 4

struct ForceMoment { V3 force, location, moment; };      /* Motion/forceMomentSystem.py */
struct Inertia { V3 MOI, CG; double mass; };             /* Motion/inertia.py */
struct Air { double temp, density, viscosity, asl; V3 wind, meanWind; };
struct Event { int type; int callback; double value; double delay; };

/* Everything one lane carries between time steps. */
struct Lane {
    State s; double time;
    V3 wind;                                    /* sampled Environment.ConstantMeanWind.velocity */
    bool dof3, underChute, onRail, haveLast, haveCache;
    /* control system (GNC/*.py): loop timing, vector PID memory, first-order actuators; RK4 forced during the controlled ascent */
    bool hasControl, forceRK4;
    double ctrlLastRun, gncLastError[3], gncIntegral[3];
    Q4 targetQ; bool haveTargetQ, qAliased;     /* IdealMomentController: the navigator's target quaternion as the reference's object aliasing leaves it */
    int ndHint[2];                              /* simplex that held the previous gain-table / deflection-table query (ndInterpolate) */
    double actLastDefl[CS_MAX_ACT], actLastTime[CS_MAX_ACT], actTarget[CS_MAX_ACT];
    bool resorted;                              /* a stage has been dropped: inertia lists are in top-to-bottom order (rocket.py:476-479) */
    int recoveryStage, nStages, nEvents, turb;
    double dtEntry;                             /* step size handed to Rocket.timeStep for the current step, before the event clamp */
    double* sepOut;                             /* this lane's stage-separation rows (SEP_W doubles each), or nullptr */
    bool eventOverflow;                         /* an event subscription did not fit MLB_MAX_EVENTS: the lane ends with ST_UNSUPPORTED */
    double lastVz, lastTime;                    /* event detector memory (simEventDetector.py:120-132) */
    double pidLastError, pidIntegral;           /* step-size PID (GNC/PID.py:27-51) */
    double ignitionTime[MLB_MAX_STAGES], engineShutOff[MLB_MAX_STAGES];
    /* sampled motor adjust factors (LANE_MOTOR_ADJUST; Propulsion.py:38-40,109-128): the motor tables of this lane are the file's
       columns rescaled on the fly, the initial propellant weights are integrated once at lane start */
    const double* G;                            /* the whole model block in global memory (interpolation records behind H_SHARED_LEN) */
    bool motorAdj;
    /* the rocket's inertia once every motor has burnt out and every inertia table has run past its last row: a constant until the next
       stage separation (rocketInertiaFrozen), kept here instead of being recombined at every derivative evaluation */
    bool inertiaCached; Inertia inertiaC; double massSumC;
    double motorImp[MLB_MAX_STAGES], motorBurn[MLB_MAX_STAGES], motorOxW0[MLB_MAX_STAGES], motorFuelW0[MLB_MAX_STAGES];
    double finThickK[MLB_MAX_FINSETS];          /* Fins.py:198-210, depends on the sampled wind */
    Event events[MLB_MAX_EVENTS];
    PinkNoise png[3];
    long long nSteps, nEvals, nRejects;
    long long nSteps3, nEvals3, nEvalsTransonic;    /* of which: under-chute (3-DoF) steps / evaluations, 6-DoF evaluations at 0.8 < Mach < 1.4 */
};

/* Stage loops run over the model block's stage count and leave at the lane's own (stages drop from the bottom, the remaining ones are a
   prefix): in a model-specialised build (MLB_SPEC) the trip counts are constants, so the stage and component loops unroll into the
   rocket's own straight-line force model with every record offset and constant folded. */
MLB_DEV int stagesOf(const double* B, const Lane& L) { return (B[H_NSTAGES] == 1) ? 1 : L.nStages; }
MLB_DEV bool motorAdjOf(const Lane& L) { return L.motorAdj; }
#ifdef MLB_SPEC
#define MLB_MODEL_LOOP _Pragma("unroll")
#define MLB_COMPONENT_LOOP _Pragma("unroll")
#else
#define MLB_MODEL_LOOP
#define MLB_COMPONENT_LOOP _Pragma("unroll 1")
#endif

MLB_DEV ForceMoment fm(V3 f, V3 loc, V3 m) { ForceMoment r; r.force = f; r.location = loc; r.moment = m; return r; }
MLB_DEV ForceMoment fmZero() { return fm(v3(0, 0, 0), v3(0, 0, 0), v3(0, 0, 0)); }
MLB_DEV ForceMoment fmAt(ForceMoment a, V3 newLoc) {                                        /* forceMomentSystem.py:40-51 */
    V3 newToOld = a.location - newLoc;
    return fm(a.force, newLoc, a.moment + cross(newToOld, a.force));
}
MLB_DEV ForceMoment fmAdd(ForceMoment a, ForceMoment b) {                                   /* :25-32 */
    ForceMoment b2 = fmAt(b, a.location);
    return fm(a.force + b2.force, a.location, a.moment + b2.moment);
}
MLB_DEV V3 ld3(const double* p) { return v3(p[0], p[1], p[2]); }

MLB_DEV int bisectRight(const double* X, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (x < X[mid]) hi = mid; else lo = mid + 1; }
    return lo;
}
MLB_DEV double linInterpAt(const double* X, const double* Y, int n, int i, double x) {      /* Interpolation.py:15-41 */
    if (i >= n) return Y[n - 1];
    if (i < 1) return Y[0];
    return (Y[i] - Y[i - 1]) * (x - X[i - 1]) / (X[i] - X[i - 1]) + Y[i - 1];
}
MLB_DEV double linInterp(const double* X, const double* Y, int n, double x) { return linInterpAt(X, Y, n, bisectRight(X, n, x), x); }
MLB_DEV V3 linInterpV(const double* X, const double* Yx, const double* Yy, const double* Yz, int n, int i, double x) {
    if (i >= n) return v3(Yx[n - 1], Yy[n - 1], Yz[n - 1]);
    if (i < 1) return v3(Yx[0], Yy[0], Yz[0]);
    V3 lg = v3(Yx[i], Yy[i], Yz[i]), sm = v3(Yx[i - 1], Yy[i - 1], Yz[i - 1]);
    return vdiv((lg - sm) * (x - X[i - 1]), (X[i] - X[i - 1])) + sm;
}

/* ---------------------------------------------------------------------------------------------
 * Environment
 * ------------------------------------------------------------------------------------------- */
/* Earth models (ENV/EarthModelling.py): None / Flat use the launch-tower frame as the inertial frame; Round and WGS84 integrate in
   an earth-centred inertial frame and need the geodetic latitude / longitude / height of a position: closed form for the
   sphere (:147-160), Newton-Raphson on the latitude parameter for the ellipsoid (:305-341). Angles in radians here. */
struct Geodetic { double lat, lon, height; };
MLB_DEV bool earthIsRound(const double* B) { int m = (int)B[H_EARTH]; return m == EARTH_ROUND || m == EARTH_WGS84; }
MLB_COLD_GEODETIC Geodetic cartesianToGeodetic(const double* B, V3 pos) {
    Geodetic g;
    const double x = pos.x, y = pos.y, z = pos.z;
    const double p = sqrt(x * x + y * y);
    g.lon = m_atan2(y, x);
    if ((int)B[H_EARTH] == EARTH_ROUND) {
        g.lat = m_atan2(z, p);
        g.height = sqrt(x * x + y * y + z * z) - B[H_EARTH_RADIUS];
        return g;
    }
    const double e2 = B[H_WGS_E2], a = B[H_WGS_A];
    double Kold = 1 / (1 - e2), Knew; const double Korig = Kold;
    for (int it = 0; it < 50; it++) {
        double t = p * p + (1 - e2) * z * z * (Kold * Kold);
        double ci = (t * sqrt(t)) / (a * e2);
        Knew = 1 + (p * p + (1 - e2) * z * z * (Kold * Kold * Kold)) / (ci - p * p);
        if (fabs(Knew - Kold) < 1e-10) break;
        Kold = Knew;
    }
    const double k = Knew;
    g.lat = m_atan2(k * z, p);
    g.height = (1 / e2) * (1 / k - 1 / Korig) * sqrt(p * p + z * z * k * k);
    return g;
}
MLB_DEV Q4 enuRotationOf(const Geodetic& g) {                                               /* getInertialToENUFrameRotation :170-190 */
    /* the reference converts to degrees and back: radians(degrees(lon) + 90), radians(90 - degrees(lat)) */
    double lonDeg = g.lon * (180.0 / MLB_PI), latDeg = g.lat * (180.0 / MLB_PI);
    Q4 rot1 = qAxisAngle(v3(0, 0, 1), (lonDeg + 90) * (MLB_PI / 180.0));
    Q4 rot2 = qAxisAngle(v3(1, 0, 0), (90 - latDeg) * (MLB_PI / 180.0));
    return rot1 * rot2;
}
MLB_DEV double earthAltitude(const double* B, V3 p) {
    if (!earthIsRound(B)) return p.z;                                                       /* No/FlatEarth.getAltitude */
    return cartesianToGeodetic(B, p).height;
}

MLB_DEV void atmosphere(const double* B, double asl, Air& e) {
    int model = (int)B[H_ATM];
    if (model == ATM_CONSTANT) {
        e.temp = B[H_ATM_CONST]; e.density = B[H_ATM_CONST + 2]; e.viscosity = B[H_ATM_CONST + 3];
        return;
    }
    if (model == ATM_TABULATED) {
        int n = (int)B[H_ATM_TAB_N]; const double* T = B + (int)B[H_ATM_TAB_OFF];
        int i = bisectRight(T, n, asl);
        e.temp = linInterpAt(T, T + n, n, i, asl);
        e.density = linInterpAt(T, T + 3 * n, n, i, asl); e.viscosity = linInterpAt(T, T + 4 * n, n, i, asl);
        return;
    }
    /* US Standard Atmosphere 1976 (AtmosphereModelling.py:147-182): baseHeights[8], lapse[8], baseTemps[8], basePressures[8] */
    const double* U = B + (int)B[H_USSTD_OFF];
    const double M = 28.9644, R = 8.31432, G = 9.80665, earthRadius = 6356766;
    double H = (earthRadius * asl) / (earthRadius + asl);
    if (H >= 86000) H = 86000;
    int baseIndex = bisectRight(U, 8, H) - 1;
    double altitudeAboveBase = H - U[baseIndex];
    double lapse = U[8 + baseIndex];
    double Tb = U[16 + baseIndex];
    double temp = Tb + lapse * altitudeAboveBase;
    double Pb = U[24 + baseIndex];
    double pressure;
    /* layer constants from the host: exponent -G M / (R lapse 1000), 1 / Tb, R Tb 1000 and its reciprocal (U rows 4-7) */
    if (lapse == 0) pressure = m_exp(divBy(-G * M * altitudeAboveBase, U[48 + baseIndex], U[56 + baseIndex])) * Pb;
    else pressure = m_pow(divBy(Tb + lapse * altitudeAboveBase, Tb, U[40 + baseIndex]), U[32 + baseIndex]) * Pb;
    double rho = M * pressure / (R * temp * 1000);
    double viscosity = 1.457e-6 * (temp * sqrt(temp)) / (temp + 110.4);                    /* temp**1.5 */
    if (H >= 86000) rho = 0.0;
    e.temp = temp; e.density = rho; e.viscosity = viscosity;
}

MLB_DEV V3 meanWindAt(const double* B, const Lane& L, double agl) {
    int model = (int)B[H_WIND];
    if (model == WIND_CONSTANT) return L.wind;
    if (model == WIND_HELLMAN) {
        double h = fmax(fmin(B[H_HELLMAN_LIMIT], agl), 10.0);
        return L.wind * m_pow(h / 10, B[H_HELLMAN_ALPHA]);
    }
    int n = (int)B[H_WIND_TAB_N]; const double* T = B + (int)B[H_WIND_TAB_OFF];
    int i = bisectRight(T, n, agl);
    return v3(linInterpAt(T, T + n, n, i, agl), linInterpAt(T, T + 2 * n, n, i, agl), linInterpAt(T, T + 3 * n, n, i, agl));
}

MLB_DEV V3 turbulenceAt(const double* B, Lane& L, double altitude, V3 meanWind, double time) {
    int model = (int)B[H_TURB];                             /* L.turb */
    if (model == TURB_NONE) return v3(0, 0, 0);
    if (model == TURB_SINE_GUST) {                                                          /* TurbulenceModelling.py:133-166 */
        double start = B[H_GUST_START], blend = B[H_GUST_BLEND], mag = B[H_GUST_MAG];
        double b1 = start + blend, b2 = start + blend + B[H_GUST_THICK], b3 = start + blend + B[H_GUST_THICK] + blend;
        if (altitude > start) {
            double gustVel;
            if (altitude < b1) gustVel = mag / 2 * (1 - m_cos(MLB_PI * (altitude - start) / blend));
            else if (altitude < b2) gustVel = mag;
            else if (altitude < b3) { double gustThickness = b3 - start; gustVel = mag / 2 * (1 - m_cos(MLB_PI * (altitude - start - gustThickness) / blend)); }
            else return v3(0, 0, 0);
            return ld3(B + H_GUST_DIR) * gustVel;
        }
        return v3(0, 0, 0);
    }
    double c0 = B[H_PINK_C0], c1 = B[H_PINK_C1];
    double stdDev = isnan(B[H_TURB_INTENSITY]) ? B[H_TURB_STD] : len(meanWind) * B[H_TURB_INTENSITY];
    if (model == TURB_PINK1D) return normalize(meanWind) * (pinkValue(L.png[0], time, c0, c1) * stdDev);
    double x = pinkValue(L.png[0], time, c0, c1) * stdDev;
    double y = pinkValue(L.png[1], time, c0, c1) * stdDev;
    if (model == TURB_PINK2D) return v3(x, y, 0);
    double z = pinkValue(L.png[1], time, c0, c1) * stdDev;                                 /* generator 2 again: :128-131 */
    return v3(x, y, z);
}

MLB_DEV Air airProperties(const double* B, Lane& L, V3 position, double time, bool skipTurbulence) {   /* environment.py:146-181 */
    Air e;
    const bool round = earthIsRound(B);
    Geodetic geo;
    if (round) geo = cartesianToGeodetic(B, position);      /* one geodetic solve serves altitude and the ENU rotation */
    double asl = round ? geo.height : position.z;
    e.asl = asl;
    double agl = asl - B[H_ELEV];
    V3 meanWind = meanWindAt(B, L, agl);
    /* under a parachute with turbulenceOffWhenUnderChute the reference still evaluates the turbulence model and
       then discards the value (rocket.py:534-545); nothing reads the generators afterwards, so skip the draw */
    V3 turbWind = skipTurbulence ? v3(0, 0, 0) : turbulenceAt(B, L, agl, meanWind, time);
    double earthRotationSpeed = B[H_EARTH_ROT] * sqrt(position.x * position.x + position.y * position.y);
    meanWind = meanWind + v3(earthRotationSpeed, 0, 0);
    if (round) {                                            /* wind is modelled in the local ENU frame: rotate into the inertial frame */
        Q4 enu = enuRotationOf(geo);
        meanWind = qRotate(enu, meanWind);
        turbWind = qRotate(enu, turbWind);
    }
    e.wind = meanWind + turbWind; e.meanWind = meanWind;
    atmosphere(B, asl, e);
    return e;
}

MLB_DEV V3 gravityForceGlobal(const double* B, double mass, V3 pos) {                       /* EarthModelling.py:93-105,162-168,348-369 */
    const int model = (int)B[H_EARTH];
    if (model == EARTH_NONE) return v3(0, 0, 0);
    if (model == EARTH_FLAT) {
        double r = pos.z + 6371000;
        return v3(0, 0, -1) * (mass * B[H_EARTH_GM] / (r * r));
    }
    const double r = len(pos);
    if (model == EARTH_ROUND) return (-normalize(pos)) * (mass * B[H_EARTH_GM] / (r * r));
    const double J2 = 0.00108262982, re = B[H_WGS_A];                                       /* WGS84: J2 gravity (NESC-RP-12-00770 eq. 29-31) */
    const double r2 = r * r;
    double frac = (3 * J2 * re * re) / (2 * (r2 * r2));
    double multiplier = -B[H_EARTH_GM] * mass / (r2 * r);
    double zz = pos.z * pos.z;
    return v3(pos.x * multiplier * (1 - frac * (5 * zz - r2)),
              pos.y * multiplier * (1 - frac * (5 * zz - r2)),
              pos.z * multiplier * (1 - frac * (5 * zz - 3 * r2)));
}

/* ---------------------------------------------------------------------------------------------
 * Inertia
 * ------------------------------------------------------------------------------------------- */
MLB_DEV Inertia combineInertias(const Inertia* list, int n) {                               /* inertia.py:78-115 */
    double totalMass = 0; V3 totalCG = v3(0, 0, 0);
    for (int i = 0; i < n; i++) { totalMass += list[i].mass; totalCG = totalCG + list[i].CG * list[i].mass; }
    if (totalMass > 0) totalCG = vdiv(totalCG, totalMass); else totalCG = v3(0, 0, 0);
    V3 totalMOI = v3(0, 0, 0);
    for (int i = 0; i < n; i++) {
        totalMOI = totalMOI + list[i].MOI;
        V3 d = totalCG - list[i].CG;
        double xs = d.y * d.y + d.z * d.z, ys = d.x * d.x + d.z * d.z, zs = d.x * d.x + d.y * d.y;
        totalMOI.x += list[i].mass * xs; totalMOI.y += list[i].mass * ys; totalMOI.z += list[i].mass * zs;
    }
    Inertia r; r.MOI = totalMOI; r.CG = totalCG; r.mass = totalMass;
    return r;
}
MLB_DEV Inertia motorInertia(const double* C, double t) {                                   /* Propulsion.py:171-204 */
    int n = (int)C[MT_NROWS]; const double* T = C + MT_TABLE;
    Inertia part[2];                                                                        /* [fuel, ox] (ox + fuel) */
    int i = bisectRight(T, n, t);
    for (int p = 0; p < 2; p++) {
        double W0 = p ? C[MT_OX_W0] : C[MT_FUEL_W0];
        part[p].MOI = v3(0, 0, 0); part[p].CG = v3(0, 0, 0); part[p].mass = 0;
        if (W0 != 0) {
            const double* Wcol = T + (p ? 2 : 3) * n;
            const double* Mcol = T + (p ? 4 : 7) * n;
            double w = linInterpAt(T, Wcol, n, i, t);
            double frac = 1 - (w / W0);
            double cg0 = p ? C[MT_OX_CG0] : C[MT_FUEL_CG0], cg1 = p ? C[MT_OX_CG1] : C[MT_FUEL_CG1];
            part[p].MOI = linInterpV(T, Mcol, Mcol + n, Mcol + 2 * n, n, i, t);
            part[p].CG = v3(0, 0, cg0 * (1 - frac) + cg1 * frac);
            part[p].mass = w;
        }
    }
    return combineInertias(part, 2);
}
/* The same motor with per-lane adjust factors (Propulsion.py:109-128): times are rawTime * burnTimeAdjustFactor, the propellant weight
   at table row k is the trapezoid sum of flowRate over the ADJUSTED times from the last row back to k, accumulated in the reference's
   order (last interval first), so it is rebuilt here from the end of the table down to the interval that brackets t. */
MLB_DEV int bisectRightScaled(const double* X, int n, double x, double scale) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (x < X[mid] * scale) hi = mid; else lo = mid + 1; }
    return lo;
}
MLB_DEV double motorWeightFromEnd(const double* rawT, const double* flow, int n, int k, double burn) {      /* weights[k], :119-128 */
    double sum = 0;
#pragma unroll 1            /* a table-length loop: a specialised build knows n and would otherwise unroll it whole */
    for (int i = n - 1; i > k; i--) { double deltaT = rawT[i] * burn - rawT[i - 1] * burn; sum += deltaT * (flow[i - 1] + flow[i]) / 2; }
    return sum;
}
MLB_DEV Inertia motorInertiaAdjusted(const double* C, double t, double burn, double oxW0, double fuelW0) {
    int n = (int)C[MT_NROWS]; const double* T = C + MT_TABLE; const double* rawT = T + MT_COL_RAW_TIME * n;
    Inertia part[2];                                                                        /* [fuel, ox] (ox + fuel) */
    int i = bisectRightScaled(rawT, n, t, burn);
    for (int p = 0; p < 2; p++) {
        double W0 = p ? oxW0 : fuelW0;
        part[p].MOI = v3(0, 0, 0); part[p].CG = v3(0, 0, 0); part[p].mass = 0;
        if (W0 != 0) {
            const double* flow = T + (p ? MT_COL_OX_FLOW : MT_COL_FUEL_FLOW) * n;
            const double* Mcol = T + (p ? 4 : 7) * n;
            double w; V3 moi;
            if (i >= n) { w = 0; moi = v3(Mcol[n - 1], Mcol[2 * n - 1], Mcol[3 * n - 1]); }            /* weights[-1] = 0 */
            else if (i < 1) { w = W0; moi = v3(Mcol[0], Mcol[n], Mcol[2 * n]); }
            else {
                double wHi = motorWeightFromEnd(rawT, flow, n, i, burn);
                double x0 = rawT[i - 1] * burn, x1 = rawT[i] * burn;
                double wLo = wHi + (x1 - x0) * (flow[i - 1] + flow[i]) / 2;                 /* one more interval: weights[i - 1] */
                w = (wHi - wLo) * (t - x0) / (x1 - x0) + wLo;
                V3 lg = v3(Mcol[i], Mcol[n + i], Mcol[2 * n + i]), sm = v3(Mcol[i - 1], Mcol[n + i - 1], Mcol[2 * n + i - 1]);
                moi = vdiv((lg - sm) * (t - x0), (x1 - x0)) + sm;
            }
            double frac = 1 - (w / W0);
            double cg0 = p ? C[MT_OX_CG0] : C[MT_FUEL_CG0], cg1 = p ? C[MT_OX_CG1] : C[MT_FUEL_CG1];
            part[p].MOI = moi;
            part[p].CG = v3(0, 0, cg0 * (1 - frac) + cg1 * frac);
            part[p].mass = w;
        }
    }
    return combineInertias(part, 2);
}
/* TabulatedInertia.getInertia (RocketComponents.py:369-372): time-tabulated mass, CG, MOI */
MLB_DEV Inertia tabulatedInertia(const double* C, double time) {
    int n = (int)C[TI_NROWS]; const double* T = C + TI_TABLE;
    int i = bisectRight(T, n, time);
    double v[7];
#pragma unroll
    for (int k = 0; k < 7; k++) v[k] = linInterpAt(T, T + (k + 1) * n, n, i, time);
    Inertia r; r.MOI = v3(v[4], v[5], v[6]); r.CG = v3(v[1], v[2], v[3]); r.mass = v[0];
    return r;
}
MLB_DEV Inertia stageInertia(const double* B, const Lane& L, int si, double time) {         /* CompositeObject.py:102-121 */
    const double* S = B + (int)B[H_STAGE_OFF + si];
    int flags = (int)S[S_OVR_FLAGS];
    Inertia r;
    if (flags == (OVR_CG | OVR_MASS | OVR_MOI)) {          /* every term overridden: the combination cannot be observed */
        r.CG = ld3(S + S_OVR_CG); r.MOI = ld3(S + S_OVR_MOI); r.mass = S[S_OVR_MASS];
        return r;
    }
    Inertia list[1 + S_MAX_VAR]; int n = 1;
    list[0].MOI = ld3(S + (L.resorted ? S_FIXED2_MOI : S_FIXED_MOI)); list[0].CG = ld3(S + (L.resorted ? S_FIXED2_CG : S_FIXED_CG));
    list[0].mass = L.resorted ? S[S_FIXED2_MASS] : S[S_FIXED_MASS];
    const int nVar = (int)S[S_NVAR];
    MLB_MODEL_LOOP
    for (int k = 0; k < nVar; k++) {                       /* variable-mass components, in the order the reference combines them */
        const double* C = B + (int)S[S_COMP_OFF + (int)S[(L.resorted ? S_VAR2_IDX : S_VAR_IDX) + k]];
        if ((int)C[C_TYPE] == CT_MOTOR) {
            double tIgn = fmax(0.0, time - L.ignitionTime[si]);
            list[n++] = motorAdjOf(L) ? motorInertiaAdjusted(C, tIgn, L.motorBurn[si], L.motorOxW0[si], L.motorFuelW0[si]) : motorInertia(C, tIgn);
        }
        else list[n++] = tabulatedInertia(C, time);
    }
    r = combineInertias(list, n);
    if (flags & OVR_CG) r.CG = ld3(S + S_OVR_CG);
    if (flags & OVR_MOI) r.MOI = ld3(S + S_OVR_MOI);
    if (flags & OVR_MASS) r.mass = S[S_OVR_MASS];
    return r;
}
MLB_COLD_INERTIA Inertia rocketInertia(const double* B, const Lane& L, double time, double* massSum) {
    Inertia list[MLB_MAX_STAGES + 1]; int n = 0;
    list[n].MOI = v3(0, 0, 0); list[n].CG = v3(0, 0, 0); list[n].mass = 0; n++;            /* rocket-level fixed-mass component (empty) */
    double m = 0;
    const int nStages = stagesOf(B, L), maxStages = (int)B[H_NSTAGES];
    MLB_MODEL_LOOP
    for (int si = 0; si < maxStages; si++) { if (si >= nStages) break; list[n] = stageInertia(B, L, si, time); m += list[n].mass; n++; }
    *massSum = m;                                                                           /* CompositeObject.getMass: plain sum over stages */
    return combineInertias(list, n);
}

/* True when no variable-mass component of the remaining stages can change any more at `time` or later: a motor returns its last table
   row for timeSinceIgnition >= times[-1] (Propulsion.py:171-204 through linInterp's end clamp, Interpolation.py:22-25), a TabulatedInertia
   its last row for time >= its last time. The tests are the clamps' own comparisons, so the frozen value is bit-identical. */
MLB_DEV bool rocketInertiaFrozen(const double* B, const Lane& L, double time) {
    const int nStages = stagesOf(B, L), maxStages = (int)B[H_NSTAGES];
    MLB_MODEL_LOOP
    for (int si = 0; si < maxStages; si++) {
        if (si >= nStages) break;
        const double* S = B + (int)B[H_STAGE_OFF + si];
        if ((int)S[S_OVR_FLAGS] == (OVR_CG | OVR_MASS | OVR_MOI)) continue;
        const int nVar = (int)S[S_NVAR];
        for (int k = 0; k < nVar; k++) {
            const double* C = B + (int)S[S_COMP_OFF + (int)S[S_VAR_IDX + k]];
            if ((int)C[C_TYPE] == CT_MOTOR) {
                const int n = (int)C[MT_NROWS]; const double* T = C + MT_TABLE;
                const double tEnd = motorAdjOf(L) ? T[MT_COL_RAW_TIME * n + n - 1] * L.motorBurn[si] : T[n - 1];
                if (!(fmax(0.0, time - L.ignitionTime[si]) >= tEnd)) return false;
            } else {
                const int n = (int)C[TI_NROWS];
                if (!(time >= C[TI_TABLE + n - 1])) return false;
            }
        }
    }
    return true;
}

/* ---------------------------------------------------------------------------------------------
 * Aerodynamic helper quantities (Motion/AeroParameters.py) and coefficient models (Rocket/AeroFunctions.py)
 * ------------------------------------------------------------------------------------------- */
MLB_DEV double dragToAxialFactor(double AOA);
MLB_DEV double baseDragCoefficient(double Mach);
MLB_DEV double skinFrictionCompressibility(double Mach, bool smooth);

struct Aero {
    double Mach, totalAOA, q, airSpeed, unitRe, density, viscosity, asl, finenessRatio, Aref, time;
    double rAref, finenessF;   /* 1 / Aref and 1 + 0.5 / finenessRatio, host-derived (H_RAREF, H_FINENESS_F) */
    bool qValid;        /* dynamic pressure is computed and cached at its FIRST use in an evaluation (AeroParameters.py:141-150) */
    double sinAOA, axialFactor, baseDrag, sfCorrRough;      /* functions of Mach / total AOA only, shared by every component */
    V3 lfav, normalDir, w;
    bool fullyTurb;
};

MLB_DEV V3 localFrameAirVel(const State& s, V3 wind) { return qConjRotate(s.q, s.vel - wind) * -1.0; }   /* AeroParameters.py:126-129 */
MLB_DEV double betaOf(double Mach) { return (Mach <= 0.9) ? sqrt(1 - Mach * Mach) : sqrt(Mach * Mach - 1); }   /* :152-163 */

MLB_DEV void aeroSetup(const double* B, const Lane& L, Aero& a, const State& s, const Air& e, double time) {
    a.time = time; a.w = s.w; a.density = e.density; a.viscosity = e.viscosity; a.asl = e.asl;
    a.Mach = len(s.vel - e.wind) / sqrt(1.4 * 287 * e.temp);                                /* :15-20 */
    a.lfav = localFrameAirVel(s, e.wind);
    a.airSpeed = len(a.lfav);
    a.totalAOA = (a.airSpeed == 0) ? 0.0 : angleBetween(a.lfav, v3(0, 0, -1));             /* :29-39 */
    if (a.lfav.x == 0.0 && a.lfav.y == 0.0) a.normalDir = v3(1, 0, 0);                      /* :131-139 */
    else a.normalDir = normalize(v3(a.lfav.x, a.lfav.y, 0));
    a.qValid = false; a.q = 0;
    a.unitRe = a.airSpeed * e.density / e.viscosity;                                        /* :22-27 */
    const int nStages = stagesOf(B, L);
    a.finenessRatio = B[H_FINENESS + nStages - 1];
    a.fullyTurb = B[H_FULLY_TURB] != 0;
    a.Aref = B[H_AREF]; a.rAref = B[H_RAREF]; a.finenessF = B[H_FINENESS_F + nStages - 1];
    a.sinAOA = m_sin(a.totalAOA);
    a.axialFactor = dragToAxialFactor(a.totalAOA);
    a.baseDrag = baseDragCoefficient(a.Mach);
    a.sfCorrRough = skinFrictionCompressibility(a.Mach, false);
}

MLB_DEV double dynamicPressure(Aero& a) {
    if (!a.qValid) { a.q = a.airSpeed * a.airSpeed * a.density / 2; a.qValid = true; }
    return a.q;
}
/* getAOA / getAOSS (AeroParameters.py:41-77) zero a component of the CACHED local air-velocity vector: every later use in
   the same evaluation (air speed, Reynolds number, drag direction, a not-yet-cached dynamic pressure) sees the mutated
   vector. Mach, total AOA and the normal-force direction were cached before (rocket.py:563-566). */
MLB_DEV void airVelMutated(Aero& a) { a.airSpeed = len(a.lfav); a.unitRe = a.airSpeed * a.density / a.viscosity; }
MLB_DEV double aeroAOA(Aero& a) {
    if (a.airSpeed == 0) return 0;
    a.lfav.y = 0; airVelMutated(a);
    double ang = angleBetween(a.lfav, v3(0, 0, -1));
    return (a.lfav.x > 0) ? ang : -ang;
}
MLB_DEV double aeroAOSS(Aero& a) {
    if (a.airSpeed == 0) return 0;
    a.lfav.x = 0; airVelMutated(a);
    double ang = angleBetween(a.lfav, v3(0, 0, -1));
    return (a.lfav.y < 0) ? ang : -ang;
}
MLB_DEV double aeroRollAngle(const Aero& a) {                                               /* :79-96 */
    V3 y = a.normalDir * -1.0;
    if (len(y) < 0.01) return 0;
    double rollAngle = m_atan2(-y.y, y.x) * (180.0 / MLB_PI);
    rollAngle = (-1) * rollAngle;
    double r = fmod(rollAngle, 360.0);                                                      /* Python float % */
    if (r != 0 && r < 0) r += 360.0;
    return r;
}
MLB_DEV double sfSub(double Mach, bool smooth) { return smooth ? 1 - 0.10 * (Mach * Mach) : 1 - 0.12 * (Mach * Mach); }   /* AeroFunctions.py:38-51 */
MLB_DEV double sfSup(double Mach, bool smooth) {                                             /* :53-69 */
    double smoothTurbFactor = 1 / m_pow(1 + 0.15 * (Mach * Mach), 0.58);
    if (smooth) return smoothTurbFactor;
    double roughTurbFactor = 1 / (1 + 0.18 * (Mach * Mach));
    return fmax(smoothTurbFactor, roughTurbFactor);
}
MLB_DEV double skinFrictionCompressibility(double Mach, bool smooth) {                     /* :117-134 */
    if (Mach < 0.9) return sfSub(Mach, smooth);
    if (Mach < 1.1) { double sub = sfSub(Mach, smooth), sup = sfSup(Mach, smooth); return sub + ((sup - sub) * (Mach - 0.9) / 0.2); }
    return sfSup(Mach, smooth);
}
/* getSkinFrictionCoefficient (:71-136). criticalRe / roughCf are the component's roughness constants from the model block;
   corrRough is the compressibility correction of a rough surface at this Mach number (computed once per evaluation). */
MLB_DEV double skinFrictionCoefficient(double unitRe, double length, double Mach, double roughness, double criticalRe, double roughCf, bool fullyTurb, double corrRough) {
    double Re = unitRe * length;
    bool smooth = (roughness == 0);
    double cf;
    if (Re > criticalRe) cf = roughCf;
    else if (fullyTurb) {
        if (Re < 1e4) cf = 0.0148;
        else { double t = 1.5 * m_log(Re) - 5.6; cf = 1 / (t * t); }
    } else {
        if (Re < 1e4) cf = 0.01328;
        else if (Re < 5e5) cf = 1.328 / sqrt(Re);
        else { double t = 1.5 * m_log(Re) - 5.6; cf = 1 / (t * t) - 1700 / Re; }
    }
    return cf * (smooth ? skinFrictionCompressibility(Mach, true) : corrRough);
}
/* getCylindricalSkinFrictionDragCoefficientAndRollDampingMoment (:138-168) */
MLB_DEV double cylinderSkinFriction(const Aero& a, double length, double roughness, double criticalRe, double roughCf, double wettedOverAref, double avgRadius, V3& rollDamping) {
    double cd = skinFrictionCoefficient(a.unitRe, length, a.Mach, roughness, criticalRe, roughCf, a.fullyTurb, a.sfCorrRough);
    cd *= a.finenessF;                                      /* (1 + (0.5 / finenessRatio)) */
    cd *= wettedOverAref;                                   /* wettedArea / Aref; avgRadius = (wettedArea / length) / (2 pi): *_K_WET, *_K_RAD */
    double rollingSurfaceVel = avgRadius * a.w.z;
    double totalVelSquared = a.airSpeed * a.airSpeed + rollingSurfaceVel * rollingSurfaceVel;
    if (totalVelSquared == 0) { rollDamping = v3(0, 0, 0); return cd; }
    double totalSurfaceForce = cd * 0.5 * totalVelSquared * a.density * a.Aref;
    double rollDampingForce = totalSurfaceForce * (rollingSurfaceVel / sqrt(totalVelSquared));
    rollDamping = v3(0, 0, -rollDampingForce * avgRadius);
    return cd;
}
MLB_DEV double baseDragCoefficient(double Mach) { return (Mach < 1) ? 0.12 + 0.13 * Mach * Mach : 0.25 / Mach; }   /* :171-175 */
MLB_DEV double cylinderCrossFlowCd(double Mach) {                                           /* :177-189 */
    if (Mach < 0.9) return m_pow(1 - Mach * Mach, -0.417) - 1;
    if (Mach <= 1) return 1 - 1.5 * (Mach - 0.9);
    double m2 = Mach * Mach, m4 = m2 * m2, m6 = m4 * m2;
    return 1.214 - 0.502 / m2 + 0.1095 / m4 + .0231 / m6;
}
MLB_DEV double bluntBodyCd(double Mach) {                                                   /* :191-200 */
    double m2 = Mach * Mach, m4 = m2 * m2, m6 = m4 * m2;
    double qStageOverQ = (Mach < 1) ? 1 + (m2 / 4) + (m4 / 40) : 1.84 - (0.76 / m2) + (0.166 / m4) + (0.035 / m6);
    return 0.85 * qStageOverQ;
}
MLB_DEV double dragToAxialFactor(double AOA) {                                              /* :203-216 */
    AOA = fabs(AOA) * (180.0 / MLB_PI);
    if (AOA <= 17) return -0.000122124 * (AOA * AOA * AOA) + 0.003114164 * (AOA * AOA) + 1;
    return 0.000006684 * (AOA * AOA * AOA) - 0.0010727204 * (AOA * AOA) + 0.030677323 * AOA + 1.055660807;
}
MLB_DEV ForceMoment forceFromCdCN(Aero& a, double Cd, double CN, V3 CP, V3 moment) {   /* :218-236 */
    double CA = a.axialFactor * Cd;
    V3 totalForce = ((v3(0, 0, -CA) + a.normalDir * CN) * a.Aref) * dynamicPressure(a);
    return fm(totalForce, CP, moment);
}
MLB_DEV double barrowmanCN(const Aero& a, double top, double bottom) { return divBy(2 * a.sinAOA * (bottom - top), a.Aref, a.rAref); }   /* :284-289 */
MLB_DEV double conePressureCd(const double* sub, const double* trans, double halfAngle, double Mach) {   /* noseCone.py:199-212 */
    if (Mach < 1) return sub[0] * m_pow(Mach, sub[1]);
    if (Mach > 1 && Mach < 1.3) return trans[0] + trans[1] * Mach + trans[2] * (Mach * Mach) + trans[3] * (Mach * Mach * Mach);
    double sh = m_sin(halfAngle);
    return 2.1 * (sh * sh) + 0.5 * sh / sqrt(Mach * Mach - 1);
}

/* ---------------------------------------------------------------------------------------------
 * Components
 * ------------------------------------------------------------------------------------------- */
MLB_COMP_NOSE ForceMoment noseConeForce(Aero& a, const double* C) {                         /* noseCone.py:173-197 */
    double CN = barrowmanCN(a, 0, C[NC_BASE_AREA]);
    V3 roll;
    double skin = cylinderSkinFriction(a, C[NC_LENGTH], C[NC_ROUGH], C[NC_CRIT_RE], C[NC_ROUGH_CF], C[NC_K_WET], C[NC_K_RAD], roll);
    double pressure = divBy(conePressureCd(C + NC_SUB, C + NC_TRANS, C[NC_HALF_ANGLE], a.Mach) * C[NC_BASE_AREA], a.Aref, a.rAref);
    return forceFromCdCN(a, pressure + skin, CN, ld3(C + NC_CP), roll);
}
MLB_COMP_TUBE ForceMoment bodyTubeForce(Aero& a, const double* C, V3 CG) {                  /* bodyTube.py:35-97 */
    double CN = 1.1 * (a.sinAOA * a.sinAOA);
    CN *= C[BT_K_PLAN];                                     /* BT_PLANFORM / Aref */
    V3 roll;
    double skin = cylinderSkinFriction(a, C[BT_LENGTH], C[BT_ROUGH], C[BT_CRIT_RE], C[BT_ROUGH_CF], C[BT_K_WET], C[BT_K_RAD], roll);
    const int nSegments = 25;
    double dL = C[BT_DL];                                   /* BT_LENGTH / nSegments */
    double dA = dL * C[BT_DIAM];
    V3 position = ld3(C + C_POS);
    V3 total = v3(0, 0, 0);
    const V3 offset = position - CG;           /* only the z component differs between strips */
    if (offset.x == 0 && offset.y == 0) {
        /* axisymmetric layout (tube axis through the CG): psi = (0, 0, z), so w x psi = (w.y z, -w.x z, 0) and the strip
           moment -psi x (v|v|) = (z sy, -z sx, 0): the same products as the general form with its exact-zero terms dropped */
#pragma unroll 5
        for (int i = 0; i < nSegments; i++) {                                               /* longitudinal damping strips */
            double z = C[BT_STRIP_Z + i] - CG.z;            /* (position.z + (-dL i - dL / 2)) - CG.z */
            double vx = a.w.y * z, vy = -(a.w.x * z);
            double sx = vx * fabs(vx), sy = vy * fabs(vy);
            total.x += z * sy; total.y += -(z * sx);
        }
    } else {
#pragma unroll 1
        for (int i = 0; i < nSegments; i++) {
            V3 psi = v3(position.x + 0, position.y + 0, C[BT_STRIP_Z + i]) - CG;
            V3 segVel = cross(a.w, psi);
            V3 sqVel = v3(segVel.x * fabs(segVel.x), segVel.y * fabs(segVel.y), segVel.z * fabs(segVel.z));
            total = total + (-cross(psi, sqVel));
        }
    }
    total = total * ((1.1 / 2) * a.density * dA);
    total = total + roll;
    return forceFromCdCN(a, skin, CN, ld3(C + BT_CP), total);
}
MLB_COMP_TAIL ForceMoment transitionForce(const Lane& L, Aero& a, const double* C, int stageIdx, bool isBoatTail) {   /* boatTail.py:127-217 */
    double CN = barrowmanCN(a, C[TR_TOP_AREA], C[TR_BOTTOM_AREA]);
    double cdp;
    if (isBoatTail) {
        double base = a.baseDrag;
        cdp = base * C[TR_CD_ADJ];
        cdp *= C[TR_K_FRONT];                               /* TR_FRONTAL / Aref */
        bool noEngine = (L.engineShutOff[stageIdx] < 0);
        if (noEngine || a.time > L.engineShutOff[stageIdx]) cdp += divBy(base * C[TR_BOTTOM_AREA], a.Aref, a.rAref);
    } else {
        if (C[TR_GROWING] == 0) cdp = a.baseDrag * C[TR_CD_ADJ];
        else cdp = conePressureCd(C + TR_SUB, C + TR_TRANS, C[TR_HALF_ANGLE], a.Mach);
        cdp *= C[TR_K_FRONT];                               /* TR_FRONTAL / Aref */
    }
    double skin = 0; V3 roll = v3(0, 0, 0);
    if (C[TR_WETTED] != 0) skin = cylinderSkinFriction(a, C[TR_LENGTH], C[TR_ROUGH], C[TR_CRIT_RE], C[TR_ROUGH_CF], C[TR_K_WET], C[TR_K_RAD], roll);
    return forceFromCdCN(a, cdp + skin, CN, ld3(C + TR_CP), roll);
}

MLB_DEV double finCnAlphaSubsonic(const double* C, double Mach) {                           /* Fins.py:26-34 */
    double beta = betaOf(Mach);
    double AR = C[FS_AR2];                                  /* 2 span^2 / planform */
    double t = divBy(beta * AR, C[FS_COS_MID], C[FS_RCOS_MID]);
    return (2 * MLB_PI * AR) / (2 + sqrt(4 + t * t));
}
struct Busemann { double K1, K2, K3, negZPerRadius; };
MLB_DEV Busemann busemann(double Mach) {                                                    /* Fins.py:37-43,494-506 */
    const double gamma = 1.4;
    double beta = betaOf(Mach);
    double M2 = Mach * Mach, M4 = M2 * M2, M6 = M4 * M2, M8 = M4 * M4;
    double b2 = beta * beta, b4 = b2 * b2, b7 = b4 * b2 * beta;
    Busemann k;
    k.K1 = 2 / beta;
    k.K2 = ((gamma + 1) * M4 - 4 * b2) / (4 * b4);
    k.K3 = ((gamma + 1) * M8 + (2 * (gamma * gamma) - 7 * gamma - 5) * M6 + 10 * (gamma + 1) * M4 - 12 * M2 + 8) / (6 * b7);
    k.negZPerRadius = 1 / m_tan(m_asin(1 / Mach));
    return k;
}
/* Interpolation.py:86-107 with the precomputed inverse constraint matrix; the 4x4 by 4x1 product is summed
   (a0 b0 + a2 b2) + (a1 b1 + a3 b3), the order numpy/OpenBLAS uses on the reference's host */
MLB_DEV double cubicBlend(const double* Ainv, double X, double Y1, double Y2, double Y1p, double Y2p, double dx) {
    double b0 = Y1, b1 = Y2, b2 = (Y1p - Y1) / dx, b3 = (Y2p - Y2) / dx, c[4];
    for (int i = 0; i < 4; i++) { const double* r = Ainv + 4 * i; c[i] = (r[0] * b0 + r[2] * b2) + (r[1] * b1 + r[3] * b3); }
    return c[0] + c[1] * X + c[2] * (X * X) + c[3] * (X * X * X);
}
MLB_DEV double finCPX(const double* C, double Mach) {                                       /* Fins.py:360-386 */
    if (Mach < 0.5) return C[FS_XMAC_LE] + 0.25 * C[FS_MAC_LEN];
    if (Mach >= 2.0) {
        double beta = betaOf(Mach), AR = C[FS_AR];
        return C[FS_MAC_LEN] * ((AR * beta - 0.67) / (2 * AR * beta - 1)) + C[FS_XMAC_LE];
    }
    double m2 = Mach * Mach, m3 = m2 * Mach, m4 = m2 * m2, m5 = m4 * Mach;
    double polyval = 0;
    polyval += C[FS_CP_POLY + 0] * m5; polyval += C[FS_CP_POLY + 1] * m4; polyval += C[FS_CP_POLY + 2] * m3;
    polyval += C[FS_CP_POLY + 3] * m2; polyval += C[FS_CP_POLY + 4] * Mach; polyval += C[FS_CP_POLY + 5] * 1.0;
    return polyval * C[FS_MAC_LEN] + C[FS_XMAC_LE];
}

#ifndef MLB_SLICE_UNROLL
#define MLB_SLICE_UNROLL 0
#endif
#if MLB_SLICE_UNROLL
#define MLB_SLICE_LOOP _Pragma("unroll 1")
#else
#define MLB_SLICE_LOOP
#endif
/* Strip sums over the span slices with the slice angles of attack already known (CythonFinFunctions.pyx:52-116) */
MLB_DEV void finStripSubsonic(const double* C, const double* aoa, int n, double spanwiseCP, double CnAlpha, double& force, double& moment) {
    double f = 0, m = 0;
    MLB_SLICE_LOOP
    for (int i = 0; i < n; i++) {
        double sliceForce = CnAlpha * aoa[i] * C[FS_SLICE_A + i];
        f += sliceForce;
        m += sliceForce * (spanwiseCP - C[FS_SLICE_R + i]);
    }
    force = f; moment = m;
}
MLB_DEV void finStripSupersonic(const double* C, const double* aoa, int n, double spanwiseCP, const Busemann& k, double& force, double& moment) {
    double f = 0, m = 0;
    double outerRadius = C[FS_SLICE_R + n - 1];
    MLB_SLICE_LOOP
    for (int i = 0; i < n; i++) {
        double machCone = (outerRadius - C[FS_SLICE_R + i]) * k.negZPerRadius + C[FS_TIP_POS];
        double al = aoa[i];
        double topCp = al * (k.K1 + k.K2 * al + k.K3 * (al * al) - 0 * (al * al));
        double bottomCp = -al * (k.K1 + k.K2 * (-al) + k.K3 * ((-al) * (-al)) - 0 * (al * al));
        double sliceForce = (topCp - bottomCp) * C[FS_SLICE_A + i];
        double fractionOut = fmin(1.0, fabs(machCone - C[FS_SLICE_LE + i]) / C[FS_SLICE_LEN + i]);
        double fractionIn = 1 - fractionOut;
        sliceForce *= fractionOut + 0.5 * (fractionIn);
        f += sliceForce;
        m += sliceForce * (spanwiseCP - C[FS_SLICE_R + i]);
    }
    force = f; moment = m;
}

/* The transonic blend's four strip sums in one pass (Fins.py:508-528): subsonic at CN_alpha(0.8), CN_alpha(0.801), Busemann at Mach 1.4 and
   1.401. Everything Mach-dependent in them is a constant of the fin set (FS_TRANS_*); only the slice angles of attack vary. Each sum adds
   its slices in order, as finStripSubsonic / finStripSupersonic do. */
MLB_DEV void finStripTransonic(const double* C, const double* aoa, int n, double spanwiseCP, double* f, double* m) {
    const double cnA1 = C[FS_TRANS_CNA], cnA2 = C[FS_TRANS_CNA + 1];
    const double* K = C + FS_TRANS_K; const double* W = C + FS_TRANS_W;
    double f0 = 0, m0 = 0, f1 = 0, m1 = 0, f2 = 0, m2 = 0, f3 = 0, m3 = 0;
    MLB_SLICE_LOOP
    for (int i = 0; i < n; i++) {
        const double al = aoa[i], area = C[FS_SLICE_A + i], arm = spanwiseCP - C[FS_SLICE_R + i], al2 = al * al;
        double sf = cnA1 * al * area; f0 += sf; m0 += sf * arm;
        sf = cnA2 * al * area; f1 += sf; m1 += sf * arm;
        double topCp = al * (K[0] + K[1] * al + K[2] * al2 - 0 * al2);
        double bottomCp = -al * (K[0] + K[1] * (-al) + K[2] * ((-al) * (-al)) - 0 * al2);
        sf = (topCp - bottomCp) * area; sf *= W[i]; f2 += sf; m2 += sf * arm;
        topCp = al * (K[4] + K[5] * al + K[6] * al2 - 0 * al2);
        bottomCp = -al * (K[4] + K[5] * (-al) + K[6] * ((-al) * (-al)) - 0 * al2);
        sf = (topCp - bottomCp) * area; sf *= W[16 + i]; f3 += sf; m3 += sf * arm;
    }
    f[0] = f0; f[1] = f1; f[2] = f2; f[3] = f3; m[0] = m0; m[1] = m1; m[2] = m2; m[3] = m3;
}

/* FirstOrderActuator.getDeflection (GNC/Actuators.py:14-29,131-132): analytic first-order lag towards the last target */
MLB_DEV double actuatorDeflection(const double* B, const Lane& L, int i, double time) {
    const double* CS = B + (int)B[H_CTRL_OFF];
    double dt = time - L.actLastTime[i];
    if (dt < 0) dt = 0;
    double lastError = L.actTarget[i] - L.actLastDefl[i];
    return L.actLastDefl[i] + lastError * (1 - m_exp(-dt / CS[CS_ACT_TAU]));
}
#ifndef MLB_FIN_UNIFORM
#define MLB_FIN_UNIFORM 1
#endif
#ifndef MLB_FIN_LINEAR
#define MLB_FIN_LINEAR 1
#endif

MLB_COMP_FINS ForceMoment finSetForce(const double* B, const Lane& L, Aero& a, const double* C, V3 CG, double thickK) {   /* Fins.py:263-358,463-551 */
    const double Mach = a.Mach, Aref = a.Aref, q = dynamicPressure(a);
    /* drag build-up common to all fins of the set (Fins.py:289-358) */
    double cf = skinFrictionCoefficient(a.unitRe, C[FS_MAC_LEN], Mach, C[FS_ROUGH], C[FS_CRIT_RE], C[FS_ROUGH_CF], a.fullyTurb, a.sfCorrRough);
    double skinDrag = cf * C[FS_K_SKIN];                    /* the flight-independent factors are FS_K_* constants from the host */
    skinDrag *= C[FS_K_THICKF];
    double leCd = (C[FS_LE_SHAPE] == 0) ? cylinderCrossFlowCd(Mach) : bluntBodyCd(Mach);
    leCd *= C[FS_K_LE];
    double teCd = divBy(a.baseDrag * C[FS_SPAN] * C[FS_TE_THICK], Aref, a.rAref);
    double tc = C[FS_TC];
    double cm = C[FS_COS_MID];
    double thicknessDrag;
    if (Mach <= 1) {
        double tc2 = tc * tc, d = thickK - (Mach * Mach) * (cm * cm);
        thicknessDrag = 4 * cf * (tc * cm + (30 * (tc2 * tc2) * (cm * cm)) / (d * sqrt(d)));
    } else thicknessDrag = C[FS_K_THICK_SUP];
    thicknessDrag *= C[FS_K_PLAN];
    double totalDrag = (leCd + teCd + thicknessDrag) + skinDrag;

    V3 position = ld3(C + C_POS);
    V3 airVelRelativeToFin = a.lfav - cross(a.w, position - CG);
    double CPXPos = finCPX(C, Mach);
    const int nFins = (int)C[FS_NFINS], nS = (int)C[FS_NSLICES];
    const int mid = (int)rint(nS / 2.0);                                                    /* Python round(): half to even */
    const double stall = C[FS_STALL];
    const double* Ainv = B + (int)B[H_FIN_CUBIC_OFF];
    const int regime = (Mach <= 0.8) ? 0 : (Mach < 1.4 ? 1 : 2);
    const bool actuated = (C[FS_ACTUATED] != 0) && B[H_CTRL_OFF] != 0;      /* L.hasControl */
    double cnA = 0; Busemann kM;
    if (regime == 0) cnA = finCnAlphaSubsonic(C, Mach);
    else if (regime == 2) kM = busemann(Mach);

    ForceMoment total = fm(v3(0, 0, 0), position, v3(0, 0, 0));
#pragma unroll 1
    for (int fi = 0; fi < nFins; fi++) {
        V3 span = v3(C[FS_SPANDIR + 2 * fi], C[FS_SPANDIR + 2 * fi + 1], 0);
        /* deflected fin normal (Fins.py:479-480), drag direction and spanwise CP distance: constants of an un-actuated fin,
           computed by the host with the reference's arithmetic */
        V3 finNormal = ld3(C + FS_FIN_NORMAL + 3 * fi);
        V3 axialDir = ld3(C + FS_AXIAL_DIR + 3 * fi);
        if (actuated) {                                     /* Fins.py:264-268,479-480: the fin angle is the actuator's deflection now */
            Q4 rot = qAxisAngle(span, actuatorDeflection(B, L, fi, a.time) * (MLB_PI / 180.0));
            finNormal = qRotate(rot, v3(span.y, -span.x, 0));
            axialDir = cross(span, finNormal);
        }
        V3 ust = cross(v3(0, 0, a.w.z), span) * -1.0;
        double spanwiseCP = C[FS_SPANWISE_CP + fi];
        /* slice angles of attack (CythonFinFunctions.pyx:10-28): one asin per slice, reused by every strip sum below */
        double aoa[FS_MAX_SLICES];
        /* A rocket that is not rolling has |w.z| at rounding-noise level: the spin term ust * r then vanishes when added to
           the air velocity (checked exactly, for the outermost slice, component by component: round-to-nearest is monotonic,
           so it vanishes for every inner slice too). All slices then see bit-identical velocities: one asin serves them all. */
        const double rOuter = C[FS_SLICE_R + nS - 1];
        const V3 vOuter = airVelRelativeToFin + ust * rOuter;
        const bool uniform = MLB_FIN_UNIFORM && (vOuter.x == airVelRelativeToFin.x) && (vOuter.y == airVelRelativeToFin.y) && (vOuter.z == airVelRelativeToFin.z)
                             && (C[FS_SLICE_R] <= rOuter);
        /* Almost the same holds while the spin term is merely small next to the air speed (roll that grows out of rounding noise on a
           rocket without fin cant: 1e-13 of the air speed in the bench workload; the residual roll of a canard-controlled ascent: up to
           4e-6): with eps = |U| r / |A| <= 1e-5 the slice angle alpha(r) = asin(n.(A + U r) / |A + U r|) is expanded about the axis,
           1 / |A + U r| through the binomial series of (1 + x)^(-1/2) to x^3 (x = (2 A.U r + U.U r^2) / A.A, |x| <= 2.1e-5, next term
           5e-20) and the arcsine to second order in b(r) - b(0) (next term 2e-16 rad). That is at the rounding level of the arcsine /
           square-root / reciprocal chain the reference runs per slice, which this replaces for nine of ten slices per fin; it is the
           GPU path's one departure from the reference's operation sequence in the force model (the oracle keeps the per-slice chain).
           Because warps mix noise-level rollers with exactly non-rolling lanes, the per-slice chain used to run for every warp at a
           third of its lanes (profiles/r2a). A warp takes this path only when all its lanes can (warp vote), so the two paths are never
           both executed; fin cant, real roll and stalled or near-perpendicular flow take the per-slice path below. */
        bool linear = false;
        if (MLB_FIN_LINEAR) {                               /* exactly non-rolling lanes too: one code path for the whole warp */
            const double m2 = dot(airVelRelativeToFin, airVelRelativeToFin), u2 = dot(ust, ust);
            double b0 = 0, c2 = 0, rm = 0;
            bool ok = m2 > 0.04 && u2 * (rOuter * rOuter) < 1e-10 * m2 && C[FS_SLICE_R] >= 0 && C[FS_SLICE_R] <= rOuter;
            if (ok) {
                rm = rcpNormal(sqrt(m2));
                b0 = dot(finNormal, airVelRelativeToFin) * rm;
                c2 = 1 - b0 * b0;
                ok = c2 > 1e-4;
            }
            const unsigned warp = __activemask();
            linear = __all_sync(warp, ok);
            if (linear) {
                const double al0 = m_asin(b0);
                const double rc = 1.0 / sqrt(c2);                          /* d asin / d b */
                const double q = dot(finNormal, ust), au = dot(airVelRelativeToFin, ust);
                if (__all_sync(warp, u2 * (rOuter * rOuter) < 1e-18 * m2)) {
                    /* noise-level roll in every lane of the warp (eps <= 1e-9): the first-order term alone is exact to 1e-18 rad */
                    const double dadr = (q - b0 * (au * rm)) * rm * rc;
                    MLB_SLICE_LOOP
                    for (int i = 0; i < nS; i++) {
                        double al = fma(dadr, C[FS_SLICE_R + i], al0);
                        if (fabs(al) > stall) al *= fabs(stall / al);
                        aoa[i] = al;
                    }
                } else {
                    const double k2 = 0.5 * b0 * (rc * rc * rc);            /* half the second derivative */
                    const double au2 = 2 * au, p0 = dot(finNormal, airVelRelativeToFin), rm2 = rm * rm;
                    MLB_SLICE_LOOP
                    for (int i = 0; i < nS; i++) {
                        const double r = C[FS_SLICE_R + i];
                        const double e = -(r * fma(u2, r, au2)) * rm2;     /* 1 - |A + U r|^2 / |A|^2 */
                        const double poly = fma(e, fma(e, fma(e, 0.3125, 0.375), 0.5), 1.0);
                        const double d = fma(fma(q, r, p0) * rm, poly, -b0);   /* b(r) - b(0) */
                        double al = fma(d, fma(d, k2, rc), al0);
                        if (fabs(al) > stall) al *= fabs(stall / al);
                        aoa[i] = al;
                    }
                }
            }
        }
        const int nAsin = linear ? 0 : (uniform ? 1 : nS);
        for (int i = 0; i < nAsin; i++) {
            V3 finVelocity = airVelRelativeToFin + ust * C[FS_SLICE_R + i];
            double finVelMag = len(finVelocity);
            double al = 0.0;
            if (finVelMag >= 0.1) {
                double body = dot(finNormal, finVelocity) * rcpNormal(finVelMag);      /* 1 / finVelMag, finVelMag >= 0.1 */
                if (fabs(body - 1) < 1e-14) body = 1.0;
                al = m_asin(body);
            }
            if (fabs(al) > stall) al *= fabs(stall / al);
            aoa[i] = al;
        }
        if (uniform && !linear) for (int i = 1; i < nS; i++) aoa[i] = aoa[0];
        double nf, mom;
        if (regime == 0) finStripSubsonic(C, aoa, nS, spanwiseCP, cnA, nf, mom);
        else if (regime == 2) finStripSupersonic(C, aoa, nS, spanwiseCP, kM, nf, mom);
        else {                                                                              /* transonic cubic blend, Fins.py:508-528 */
            double f[4], m[4];                                          /* [0] M 0.8, [1] M 0.801, [2] M 1.4, [3] M 1.401 */
            finStripTransonic(C, aoa, nS, spanwiseCP, f, m);
            nf = cubicBlend(Ainv, Mach, f[0], f[2], f[1], f[3], 0.001);
            mom = cubicBlend(Ainv, Mach, m[0], m[2], m[1], m[3], 0.001);
        }
        V3 normalForce = finNormal * (nf * q);
        mom *= q;
        double axialMag = dragToAxialFactor(aoa[mid]) * totalDrag * Aref * q;
        V3 axialForce = axialDir * axialMag;
        V3 globalCP = span * (C[FS_MAC_Y] + C[FS_BODY_R]) + (position - v3(0, 0, CPXPos));
        total = fmAdd(total, fm(normalForce + axialForce, globalCP, v3(0, 0, mom)));
    }
    double interf = C[FS_BODY_INTERF] * C[FS_NUM_INTERF];
    total.force.x *= interf; total.force.y *= interf; total.moment.x *= interf; total.moment.y *= interf;
    return total;
}

/* AeroFunctions.forceFromCoefficients (:238-269): drag along the air velocity, lift perpendicular to it in the plane of the
   normal-force direction, moments from coefficients */
MLB_DEV ForceMoment forceFromCoefficients(Aero& a, const double* coeffs, V3 cp, double refArea, double refLength) {
    double q = dynamicPressure(a);
    if (q == 0.0) return fmZero();
    double nonDimConstant = q * refArea;
    V3 V = a.lfav;
    V3 dragForce = normalize(V) * coeffs[0];
    V3 nd = a.normalDir;
    Q4 rotator = qAxisAngle(cross(V, nd), MLB_PI / 2 - angleBetween(V, nd));
    V3 liftForce = qRotate(rotator, nd) * coeffs[1];
    V3 totalForce = (dragForce + liftForce) * nonDimConstant;
    V3 moments = (v3(coeffs[2], coeffs[3], coeffs[4]) * nonDimConstant) * refLength;
    return fm(totalForce, cp, moments);
}
MLB_DEV ForceMoment aeroDampingForce(Aero& a, const double* C) {                            /* RocketComponents.py:248-261 */
    double redimConst = C[AD_LREF] / (2 * fmax(a.airSpeed, 0.0000001));
    double coeffs[5];
    coeffs[0] = 0; coeffs[1] = 0;
    coeffs[2] = dot(ld3(C + AD_XCOEFFS), a.w) * redimConst;
    coeffs[3] = dot(ld3(C + AD_YCOEFFS), a.w) * redimConst;
    coeffs[4] = dot(ld3(C + AD_ZCOEFFS), a.w) * redimConst;
    return forceFromCoefficients(a, coeffs, v3(0, 0, 0), C[AD_AREF], C[AD_LREF]);
}
/* N-D scattered linear interpolation with nearest-neighbour fallback (scipy LinearNDInterpolator / NearestNDInterpolator behind
   Motion/Interpolation.py:109-124). The Delaunay triangulation comes from scipy on the host; R points into GLOBAL memory (the
   tables are too large for shared memory, every lane reads the same words: broadcast loads). Point location is a scan over
   the simplices with scipy's inside tolerance; it runs once per control tick, not per derivative evaluation. `hint`: the simplex
   that held the lane's previous query. Between two ticks of a 100 Hz loop the query barely moves, so the hint is tested first, with the
   scan's own inclusion tolerance; when the query has left it, a directed walk through the triangulation's neighbour table (towards the
   neighbour opposite the most negative barycentric coordinate, scipy's own find_simplex strategy) finds the new simplex in a step or two;
   only a walk that leaves the hull or meets a degenerate simplex falls through to the scan (and, outside the hull, to the nearest row).
   A query strictly inside a simplex gets the scan's answer bit for bit. A query ON a shared face (a controller output that is exactly
   zero puts every query of the canard case on one) may be answered from another of the simplices that share the face than the scan's
   lowest-numbered one: the interpolant is continuous across faces, the vertex that differs carries a weight below the tolerance
   (2e-14), so the value agrees to rounding (the parity gates of the canard fixtures, <= 1e-9 against the reference, are measured with
   this rule). Why it matters: with a strict-interior rule nearly every tick of the canard case fell through to the scan, 613 simplices
   of the deflection table = about five derivative evaluations' worth of instructions per step (profiles/r3_nd_hint_ab.txt: 124 -> 227 M steps/s). */
MLB_COLD_NDTABLE void ndInterpolate(const double* __restrict__ R, const double* x, double* out, int& hint) {
    const int d = (int)R[ND_DIM], m = (int)R[ND_NVAL], npts = (int)R[ND_NPTS], ns = (int)R[ND_NSIMP];
    const double* pts = R + ND_DATA; const double* vals = pts + npts * d; const double* simp = vals + npts * m; const double* T = simp + ns * (d + 1);
    const double eps = 100 * 2.220446049250313e-16;
    const bool tryHint = (hint >= 0 && hint < ns);
    int walked = -1;
#ifndef MLB_ND_WALK
#define MLB_ND_WALK 1
#endif
#ifndef MLB_ND_MARGIN
#define MLB_ND_MARGIN (-100 * 2.220446049250313e-16)
#endif
    if (MLB_ND_WALK && tryHint) {
        const double* nb = T + ns * (d + 1) * d;
        int s = hint;
        for (int step = 0; step < 16; step++) {
            const double* Ts = T + s * (d + 1) * d;
            if (Ts[0] != Ts[0]) break;                                                      /* degenerate simplex: leave it to the scan */
            double cLast = 1.0, cMin = 1e300; int kMin = 0;
            for (int i = 0; i < d; i++) {
                double ci = 0;
                for (int j = 0; j < d; j++) ci += Ts[d * i + j] * (x[j] - Ts[d * d + j]);
                cLast -= ci;
                if (ci < cMin) { cMin = ci; kMin = i; }
            }
            if (cLast < cMin) { cMin = cLast; kMin = d; }
            if (cMin >= MLB_ND_MARGIN) { walked = s; break; }                                        /* strictly inside (the scan below re-derives the weights) */
            if (cMin >= -1e-9) break;                                                               /* on or near a face: the scan decides */
            s = (int)nb[s * (d + 1) + kMin];
            if (s < 0) break;                                                               /* left the hull: nearest-neighbour fallback below */
        }
    }
    for (int it = (walked >= 0 || tryHint) ? -1 : 0; it < ns; it++) {
        const int s = (it < 0) ? (walked >= 0 ? walked : hint) : it;
        const double lo = (it < 0) ? MLB_ND_MARGIN : -eps, hi = (it < 0) ? 1 - MLB_ND_MARGIN : 1 + eps;
        const double* Ts = T + s * (d + 1) * d;
        if (Ts[0] != Ts[0]) continue;                                                       /* degenerate simplex */
        double c[ND_MAX_DIM + 1]; double cLast = 1.0; bool inside = true;
        for (int i = 0; i < d && inside; i++) {
            double ci = 0;
            for (int j = 0; j < d; j++) ci += Ts[d * i + j] * (x[j] - Ts[d * d + j]);
            c[i] = ci; cLast -= ci;
            inside = (ci >= lo && ci <= hi);
        }
        if (!inside || !(cLast >= lo && cLast <= hi)) continue;
        c[d] = cLast;
        hint = s;
        for (int k = 0; k < m; k++) {
            double v = 0;
            for (int j = 0; j < d + 1; j++) v += c[j] * vals[(int)simp[s * (d + 1) + j] * m + k];
            out[k] = v;
        }
        return;
    }
    int best = 0; double bestD = 1e300;                                                     /* outside the hull: nearest table row */
    for (int p = 0; p < npts; p++) {
        double dist = 0;
        for (int j = 0; j < d; j++) { double t = pts[p * d + j] - x[j]; dist += t * t; }
        if (dist < bestD) { bestD = dist; best = p; }
    }
    for (int k = 0; k < m; k++) out[k] = vals[best * m + k];
}
MLB_DEV double aeroKeyValue(Aero& a, int id) {                                              /* AeroParameters.stringToAeroFunctionMap */
    switch (id) {
        case AKEY_MACH: return a.Mach;
        case AKEY_ALTITUDE: return a.asl;
        case AKEY_UNITRE: return a.unitRe * 1;
        case AKEY_TOTALAOA: return a.totalAOA;
        case AKEY_ROLLANGLE: return aeroRollAngle(a);
        case AKEY_AOA: return aeroAOA(a);
        default: return aeroAOSS(a);
    }
}
MLB_DEV ForceMoment tabulatedAeroForce(const Lane& L, Aero& a, const double* C) {           /* RocketComponents.py:322-345 */
    if ((int)C[TA_NKEYS] > 1) {
        /* several key columns: scattered linear interpolation with nearest-neighbour fallback (:316-320,331); the keys are evaluated in
           column order (AOA / AOSS mutate the cached air velocity for the keys and components after them, App. A #8) */
        const int nk = (int)C[TA_NKEYS], nv = (int)C[TA_NVALS];
        double keys[ND_MAX_DIM], vals[ND_MAX_VAL];
        for (int k = 0; k < nk; k++) keys[k] = aeroKeyValue(a, (int)C[TA_KEYS + k]);
        int hint = -1;
        ndInterpolate(L.G + (int)C[TA_ND_OFF], keys, vals, hint);
        double coeffs[5] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
        for (int c = 0; c < nv; c++) {
            int k = (int)C[TA_COEFF_IDX + c];
            if (k == 0) coeffs[0] = vals[c]; else if (k == 1) coeffs[1] = vals[c]; else if (k == 2) coeffs[2] = vals[c]; else if (k == 3) coeffs[3] = vals[c]; else coeffs[4] = vals[c];
        }
        return forceFromCoefficients(a, coeffs, ld3(C + C_POS), C[TA_AREF], C[TA_LREF]);
    }
    double key;
    switch ((int)C[TA_KEY]) {
        case AKEY_MACH: key = a.Mach; break;
        case AKEY_ALTITUDE: key = a.asl; break;
        case AKEY_UNITRE: key = a.unitRe * 1; break;
        case AKEY_TOTALAOA: key = a.totalAOA; break;
        case AKEY_ROLLANGLE: key = aeroRollAngle(a); break;
        case AKEY_AOA: key = aeroAOA(a); break;
        default: key = aeroAOSS(a); break;
    }
    const int n = (int)C[TA_NROWS], nv = (int)C[TA_NVALS]; const double* T = C + TA_TABLE;
    const int i = bisectRight(T, n, key);
    double coeffs[5] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
    for (int c = 0; c < nv; c++) {
        double v = linInterpAt(T, T + (c + 1) * n, n, i, key);
        int k = (int)C[TA_COEFF_IDX + c];
        if (k == 0) coeffs[0] = v; else if (k == 1) coeffs[1] = v; else if (k == 2) coeffs[2] = v; else if (k == 3) coeffs[3] = v; else coeffs[4] = v;
    }
    return forceFromCoefficients(a, coeffs, ld3(C + C_POS), C[TA_AREF], C[TA_LREF]);
}
MLB_COLD_INERTIA Inertia rocketInertia(const double* B, const Lane& L, double time, double* massSum);
MLB_DEV ForceMoment jetDampingForce(const double* B, const Lane& L, const Aero& a, const double* C, int si, const Inertia& current) {   /* :392-410 */
    if (a.time > L.ignitionTime[si] && a.time < L.engineShutOff[si]) {
        double m; const double dt = 0.001;
        Inertia next = rocketInertia(B, L, a.time + dt, &m);
        double dampingFactor = ((current.MOI.x - next.MOI.x) / dt) * C[JD_FRACTION];
        return fm(v3(0, 0, 0), v3(0, 0, 0), v3(-a.w.x * dampingFactor, -a.w.y * dampingFactor, 0));
    }
    return fmZero();
}

MLB_DEV double motorThrust(const Lane& L, const double* C, int si, double time) {           /* Propulsion.py:138-151 */
    double t = fmax(0.0, time - L.ignitionTime[si]);
    int n = (int)C[MT_NROWS]; const double* T = C + MT_TABLE;
    if (motorAdjOf(L)) {                                    /* thrust * impulseAdjustFactor / burnTimeAdjustFactor over time * burnTimeAdjustFactor */
        const double* rawT = T + MT_COL_RAW_TIME * n; const double* rawF = T + MT_COL_RAW_THRUST * n;
        const double imp = L.motorImp[si], burn = L.motorBurn[si];
        if (t < 0 || t > rawT[n - 1] * burn) return 0;
        int i = bisectRightScaled(rawT, n, t, burn);
        if (i >= n) return rawF[n - 1] * imp / burn;
        if (i < 1) return rawF[0] * imp / burn;
        double y1 = rawF[i] * imp / burn, y0 = rawF[i - 1] * imp / burn, x0 = rawT[i - 1] * burn, x1 = rawT[i] * burn;
        return (y1 - y0) * (t - x0) / (x1 - x0) + y0;
    }
    if (t < 0 || t > T[n - 1]) return 0;
    return linInterp(T, T + n, n, t);
}

MLB_DEV ForceMoment recoveryForce(const double* B, const Lane& L, const State& s, const Air& e, bool dof3) {   /* Recovery.py:50-63 */
    if (L.recoveryStage == 0) return fmZero();
    const double* st = B + (int)B[H_RECOVERY_OFF] + RC_STAGE0 + (L.recoveryStage - 1) * RC_STAGE_STRIDE;
    V3 airVel = (s.vel - e.wind) * -1.0;
    double velMag = dof3 ? len(airVel) : len(localFrameAirVel(s, e.wind));
    double q = velMag * velMag * e.density / 2;
    return fm(normalize(airVel) * (st[3] * st[2] * q), v3(0, 0, 0), v3(0, 0, 0));
}

/* FinSet._precomputeSubsonicFinThicknessDrag (Fins.py:198-210): evaluated once per lane because the skin-friction
   model sees the sampled wind (and the t = 0 turbulence value) through the Reynolds number */
MLB_DEV double finThicknessK(const double* B, Lane& L, const double* C) {
    State m1; m1.q = q4(B[H_INIT_STATE + 6], B[H_INIT_STATE + 7], B[H_INIT_STATE + 8], B[H_INIT_STATE + 9]);
    m1.vel = v3(0, 0, 340.3); m1.pos = v3(0, 0, 0); m1.w = v3(0, 0, 0);
    Air e = airProperties(B, L, ld3(B + H_NOSE_POS0), 0, false);
    double unitRe = len(localFrameAirVel(m1, e.wind)) * e.density / e.viscosity;
    double CD_f_star = skinFrictionCoefficient(unitRe, C[FS_MAC_LEN], 1.0, C[FS_ROUGH], C[FS_CRIT_RE], C[FS_ROUGH_CF], B[H_FULLY_TURB] != 0,
                                               skinFrictionCompressibility(1.0, false));
    double cm = m_cos(C[FS_MIDSWEEP]);
    double tc = C[FS_THICK] / C[FS_ROOT];
    double t = (C[FS_CDTT_STAR] / CD_f_star - 4 * cm * tc) / (120 * (cm * cm) * ((tc * tc) * (tc * tc)));
    return cm * cm + m_cbrt(t * t);
}

/* ---------------------------------------------------------------------------------------------
 * Rocket._getAppliedForce + RigidBody.getRigidBodyStateDerivative
 * ------------------------------------------------------------------------------------------- */
MLB_DEV ForceMoment launchRailForce(const double* B, const Lane& L, const State& s, double time, ForceMoment unadjusted) {   /* launchRail.py:37-56 */
    if (!(B[H_RAIL_LEN] > 0) || !L.onRail) return unadjusted;
    V3 railDir = ld3(B + H_RAIL_DIR);
    Q4 rot = qAxisAngle(v3(0, 0, 1), B[H_EARTH_ROT] * time);
    V3 currPos = qRotate(rot, ld3(B + H_RAIL_POS));
    if (len(s.pos - currPos) < B[H_RAIL_LEN]) {
        double mag = dot(unadjusted.force, qRotate(rot, railDir));
        return fm(railDir * mag, v3(0, 0, 0), v3(0, 0, 0));
    }
    return unadjusted;
}

struct ForceProbe { ForceMoment total, aero; Inertia inertia; double Mach, density; V3 wind; };

/* PH: the flight phase the caller is compiled for. 0 = free flight (6-DoF body, component build-up aerodynamics), 1 = under a parachute
   (3-DoF body, recovery-system drag only: rocket.py:531-547,700-726). The two are set together by the first chute deployment
   (deployNextRecoveryStage) and never revert, so each trajectory kernel carries only its own half of the force model. */
template <int PH>
MLB_DEV ForceMoment appliedForce(const double* B, Lane& L, double time, State& s, Inertia& inertia, double& massSum, ForceProbe* probe) {
    constexpr bool dof3 = (PH == 1), underChute = (PH == 1);
    L.nEvals++;
    if (dof3) L.nEvals3++;
    Air e = airProperties(B, L, s.pos, time, underChute && B[H_TURB_OFF_CHUTE] != 0);
    if (underChute && B[H_TURB_OFF_CHUTE] != 0) e.wind = e.meanWind;
    if (rocketInertiaFrozen(B, L, time)) {
        if (!L.inertiaCached) { L.inertiaC = rocketInertia(B, L, time, &L.massSumC); L.inertiaCached = true; }
        inertia = L.inertiaC; massSum = L.massSumC;
    } else inertia = rocketInertia(B, L, time, &massSum);
    ForceMoment comp;
    if (!underChute) {
        Aero a;
        aeroSetup(B, L, a, s, e, time);
        if (a.Mach > 0.8 && a.Mach < 1.4) L.nEvalsTransonic++;
        if (probe) { probe->Mach = a.Mach; probe->density = e.density; probe->wind = e.wind; }
        ForceMoment rocketTotal = fm(v3(0, 0, 0), inertia.CG, v3(0, 0, 0));
        int finIdx = 0;
        const int nStages = stagesOf(B, L), maxStages = (int)B[H_NSTAGES];
        MLB_MODEL_LOOP
        for (int si = 0; si < maxStages; si++) {
            if (si >= nStages) break;
            const double* S = B + (int)B[H_STAGE_OFF + si];
            const int nc = (int)S[S_NCOMP];
            ForceMoment stageTotal = fm(v3(0, 0, 0), inertia.CG, v3(0, 0, 0));
            MLB_COMPONENT_LOOP
            for (int ci = 0; ci < nc; ci++) {
                const double* C = B + (int)S[S_COMP_OFF + ci];
                ForceMoment f;
                switch ((int)C[C_TYPE]) {
                    case CT_NOSECONE: f = noseConeForce(a, C); break;
                    case CT_BODYTUBE: f = bodyTubeForce(a, C, inertia.CG); break;
                    case CT_BOATTAIL: f = transitionForce(L, a, C, si, true); break;
                    case CT_TRANSITION: f = transitionForce(L, a, C, si, false); break;
                    case CT_FINSET: f = finSetForce(B, L, a, C, inertia.CG, L.finThickK[finIdx & (MLB_MAX_FINSETS - 1)]); finIdx++; break;
                    case CT_MOTOR: f = fm(v3(0, 0, motorThrust(L, C, si, time)), v3(0, 0, 0), v3(0, 0, 0)); break;
                    case CT_FIXED_FORCE: f = fm(ld3(C + FF_FORCE), ld3(C + C_POS), ld3(C + FF_MOMENT)); break;
                    case CT_AERO_FORCE: f = forceFromCoefficients(a, C + AF_COEFFS, ld3(C + C_POS), C[AF_AREF], C[AF_LREF]); break;
                    case CT_AERO_DAMPING: f = aeroDampingForce(a, C); break;
                    case CT_TAB_AERO: f = tabulatedAeroForce(L, a, C); break;
                    case CT_JET_DAMPING: f = jetDampingForce(B, L, a, C, si, inertia); break;
                    default: f = fmZero(); break;                                          /* FixedMass / RecoverySystem: zero force */
                }
                stageTotal = fmAdd(stageTotal, f);
            }
            /* the automatic zero-length boat tail belongs to whichever stage is currently the bottom one (rocket.py:489-519) */
            if (si == nStages - 1 && S[S_AUTO_BT_OFF] != 0) stageTotal = fmAdd(stageTotal, transitionForce(L, a, B + (int)S[S_AUTO_BT_OFF], si, true));
            rocketTotal = fmAdd(rocketTotal, stageTotal);
        }
        comp = rocketTotal;
    } else {
        comp = recoveryForce(B, L, s, e, dof3);
    }
    comp = fmAt(comp, inertia.CG);
    V3 g = gravityForceGlobal(B, inertia.mass, s.pos);
    if (!dof3) g = qConjRotate(s.q, g);
    ForceMoment total = fmAdd(comp, fm(g, inertia.CG, v3(0, 0, 0)));
    total = launchRailForce(B, L, s, time, total);
    if (probe) { probe->total = total; probe->aero = comp; probe->inertia = inertia; }
    return total;
}

#ifndef MLB_DERIV_NOINLINE
#define MLB_DERIV_NOINLINE 0
#endif
#if MLB_DERIV_NOINLINE
#define MLB_DERIV_FN static __device__ __noinline__
#else
#define MLB_DERIV_FN MLB_DEV
#endif
template <int PH>
MLB_DERIV_FN Deriv stateDerivative(const double* B, Lane& L, double time, State& s) {
    constexpr bool dof3 = (PH == 1);
    Deriv d;
    Inertia inertia; double massSum;
    ForceMoment f = appliedForce<PH>(B, L, time, s, inertia, massSum, nullptr);
    if (dof3) {                                                                             /* RigidBodies.py:41-47 */
        d.vel = s.vel;
        d.acc = vdiv(f.force, massSum);
        d.angvel = v3(0, 0, 0); d.angacc = v3(0, 0, 0);
        return d;
    }
    /* CG migration velocity by forward difference of the inertia model (RigidBodies.py:91-97) */
    V3 CGVel_local = v3(0, 0, 0);
    if (B[H_INERTIA_CONST] == 0 && !rocketInertiaFrozen(B, L, time)) {      /* frozen: both inertias are the same value, the difference is +0 */
        double m2; Inertia inertia2 = rocketInertia(B, L, time + 1e-8, &m2);
        CGVel_local = vdiv(inertia2.CG - inertia.CG, 1e-8);
    }
    V3 CGVel_global = qRotate(s.q, CGVel_local);
    d.acc = vdiv(qRotate(s.q, f.force), inertia.mass);
    d.angvel = qRotate(s.q, s.w);
    V3 moi = inertia.MOI;
    d.angacc = v3((f.moment.x + (moi.y - moi.z) * s.w.y * s.w.z) / moi.x,
                  (f.moment.y + (moi.z - moi.x) * s.w.z * s.w.x) / moi.y,
                  (f.moment.z + (moi.x - moi.y) * s.w.x * s.w.y) / moi.z);
    d.vel = s.vel + CGVel_global;
    return d;
}

}  // namespace mlb
