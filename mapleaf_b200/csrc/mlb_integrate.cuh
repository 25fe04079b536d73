/*
 * mlb_integrate.cuh — per-lane time integration: Butcher-tableau Runge-Kutta steps in the reference's
 * object algebra (fixed-step and adaptive with per-lane PID step control), flight events, the
 * Simulation.run loop with its end conditions, and the running RocketFlight reductions.
 *
 * Reference behaviour restated here (paths relative to MAPLEAF/):
 *   Motion/Integration.py:189-236 (classical), :331-449 (adaptive, retry loop, PID / elementary controllers)
 *   GNC/PID.py:27-51 (step-size PID, integral clamped to +-1)
 *   Rocket/simEventDetector.py:64-162, Rocket/Recovery.py:65-97, Rocket/rocket.py:658-726
 *   ENV/launchRail.py:58-81, ENV/environment.py:195-207
 *   SimulationRunners/SingleSimulations.py:89-139,205-248, IO/rocketFlight.py:27-76
 *
 * The two integrator families take different algebraic routes on purpose: the
 * classical one sums stage derivatives weighted by 1/(1/(a dt)) and converts the sum to a rotation once; the
 * adaptive one converts every stage derivative to a rotation, composes them and rescales the composed angle.
 *
 * Stage derivatives k_i live in per-thread local memory (KStore: [slot][component], indexed dynamically by the
 * tableau loops; the hardware interleaves local memory across the warp, so a slot access is coalesced).
 *
 * Every routine here is compiled once per flight phase (template parameter PH, see appliedForce): PH 0 is the
 * 6-DoF free-flight code with the convoy barriers, PH 1 the 3-DoF under-chute code, which is small enough to run
 * free (no barriers, CONVOY false).
 */
#pragma once
#include "mlb_model.cuh"

namespace mlb {

/* Keep the unit rotation axis and the magnitude of every slot's angular velocity next to the slot, so that a term of the adaptive stage
   composition costs one sincos instead of two normalisations and three square roots, and normalise the composed quaternion once per
   stage. Measured in one run on one box (profiles/r2j_rotation_cache_ab.txt; runs on different boxes of the pool differ by +-3 %, more than
   this effect): 9 % fewer executed instructions, +2.4 % steps/s at 10^6 lanes. */
#ifndef MLB_ROTATION_CACHE
#define MLB_ROTATION_CACHE 1
#endif
/* Slot layout: 12 doubles of the derivative (6 in the 3-DoF kernel, which never touches the rotational half), then, in the free-flight
   kernel, 4 more: the unit rotation axis and the magnitude of the slot's angular velocity (see stageIncrement). */
#define MLB_K_SLOT_6DOF (MLB_ROTATION_CACHE ? 16 : 12)
struct KStore {
    double* base; int slotStride;
    MLB_DEV void put(int slot, const Deriv& d, bool dof3) const {
        double* p = base + slot * slotStride;
        p[0] = d.vel.x; p[1] = d.vel.y; p[2] = d.vel.z;
        p[3] = d.acc.x; p[4] = d.acc.y; p[5] = d.acc.z;
        if (dof3) return;
        p[6] = d.angvel.x; p[7] = d.angvel.y; p[8] = d.angvel.z;
        p[9] = d.angacc.x; p[10] = d.angacc.y; p[11] = d.angacc.z;
#if MLB_ROTATION_CACHE
        const double m = len(d.angvel);
        const V3 u = normalize(d.angvel);
        p[12] = u.x; p[13] = u.y; p[14] = u.z; p[15] = m;
#endif
    }
    MLB_DEV Deriv get(int slot, bool dof3) const {
        const double* p = base + slot * slotStride;
        Deriv d;
        d.vel = v3(p[0], p[1], p[2]);
        d.acc = v3(p[3], p[4], p[5]);
        if (dof3) { d.angvel = v3(0, 0, 0); d.angacc = v3(0, 0, 0); return d; }
        d.angvel = v3(p[6], p[7], p[8]);
        d.angacc = v3(p[9], p[10], p[11]);
        return d;
    }
    MLB_DEV void copy(int to, int from) const {
        double* a = base + to * slotStride; const double* b = base + from * slotStride;
        for (int k = 0; k < slotStride; k++) a[k] = b[k];
    }
#if MLB_ROTATION_CACHE
    /* derivTimes(k_slot, c) of the 6-DoF body (RigidBodyStates.py:173-188): the increment's rotation is the quaternion of the rotation
       vector angvel * c, i.e. angle |angvel| |c| about angvel / |angvel| (CythonAngularVelocity.pyx:78-86, CythonQuaternion.pyx:50-60). The
       reference recomputes axis and magnitude (two normalisations, three square roots) for every (slot, coefficient) pair of the adaptive
       stage composition, 27 times per RK45 attempt; both depend on the slot alone, so they are taken once when the slot is written and a
       term costs one sincos. Same quaternion up to the rounding of |v c| against |v| |c|. */
    MLB_DEV State increment(int slot, double c) const {
        const double* p = base + slot * slotStride;
        State r;
        r.pos = v3(p[0], p[1], p[2]) * c;
        r.vel = v3(p[3], p[4], p[5]) * c;
        r.w = v3(p[9], p[10], p[11]) * c;
        double sn, cs;
        m_sincos((p[15] * c) / 2, &sn, &cs);
        r.q = q4(cs, sn * p[12], sn * p[13], sn * p[14]);
        return r;
    }
#endif
};

MLB_DEV Deriv derivScaled(const Deriv& d, double s, bool dof3) {                            /* RigidBodyStates.py:190-202 (already 1/invScalar) */
    Deriv r;
    r.vel = d.vel * s; r.acc = d.acc * s;
    if (dof3) { r.angvel = v3(0, 0, 0); r.angacc = v3(0, 0, 0); return r; }
    r.angvel = d.angvel * s; r.angacc = d.angacc * s;
    return r;
}
MLB_DEV Deriv derivAdd(const Deriv& a, const Deriv& b, bool dof3) {                         /* :164-171 */
    Deriv r;
    r.vel = a.vel + b.vel; r.acc = a.acc + b.acc;
    if (dof3) { r.angvel = v3(0, 0, 0); r.angacc = v3(0, 0, 0); return r; }
    r.angvel = a.angvel + b.angvel; r.angacc = a.angacc + b.angacc;
    return r;
}

struct StepResult { State newValue; double dt, factor, err; bool rejected; };

MLB_DEV double stepPID(const double* B, Lane& L, double currentError, double dt) {          /* GNC/PID.py:27-51, maxIntegral = 1 */
    double derivative = (currentError - L.pidLastError) / dt;
    L.pidIntegral = L.pidIntegral + (currentError + L.pidLastError) * dt / 2;
    if (fabs(L.pidIntegral) > 1) L.pidIntegral = L.pidIntegral * 1 / fabs(L.pidIntegral);
    L.pidLastError = currentError;
    return B[H_AD_P] * currentError + B[H_AD_I] * L.pidIntegral + B[H_AD_D] * derivative;
}

MLB_DEV bool pyIsClose(double a, double b) {                                                /* math.isclose: rel_tol 1e-9 */
    if (a == b) return true;
    double diff = fabs(b - a);
    return (diff <= fabs(1e-9 * b)) || (diff <= fabs(1e-9 * a));
}

/* Convoy execution. The 6-DoF force model is ~10^4 SASS instructions per evaluation, several times the SM's instruction
   cache, and warps that drift apart each stream it from L2 on their own (measured: 74 % of warp cycles stalled on
   instruction fetch, ICC hit rate 46 %). In the convoy kernels the CTA's warps re-group at a barrier before every
   derivative evaluation, so one warp's fetch serves the others. All control flow around the barriers is CTA-uniform:
   lanes that have nothing to do in a stage / attempt / step stay in the loops with `active` false. The under-chute
   kernel (CONVOY false) fits the instruction cache and runs without barriers: every predicate is then the thread's own. */
#ifndef MLB_EVAL_BARRIER
#define MLB_EVAL_BARRIER 1
#endif
/* MLB_SPLIT_BARRIER: the per-evaluation barrier as a split-phase mbarrier. A warp ARRIVES when it has finished a derivative evaluation and
   WAITS just before the next one; the stage algebra in between (slot write, next stage's increments, error estimate) is slack in which the
   CTA's slower warps catch up, instead of every warp standing at a hard barrier for the slowest one. Warps still enter every evaluation
   together (that is what the instruction cache needs); the arithmetic is untouched. One arrival per warp (lane 0 after __syncwarp). */
#ifndef MLB_SPLIT_BARRIER
#define MLB_SPLIT_BARRIER 0
#endif
struct Convoy {
    unsigned bar;       /* shared-memory address of the CTA's mbarrier */
    unsigned phase;     /* parity of the phase the next wait is for */
};
MLB_DEV void convoyInit(Convoy& c, unsigned long long* barrier, int nWarps) {
    c.bar = (unsigned)__cvta_generic_to_shared(barrier); c.phase = 0;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(c.bar), "r"(nWarps) : "memory");
}
MLB_DEV void convoyArrive(const Convoy& c) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) { unsigned long long st; asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(c.bar) : "memory"); (void)st; }
}
MLB_DEV void convoyWait(Convoy& c) {
    asm volatile("{\n\t.reg .pred p;\n\tMLB_WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra MLB_WAIT_%=;\n\t}" :: "r"(c.bar), "r"(c.phase) : "memory");
    c.phase ^= 1u;
}
template <bool CONVOY> MLB_DEV void convoySync(Convoy& c) {
    if (!CONVOY || !MLB_EVAL_BARRIER) return;
#if MLB_SPLIT_BARRIER
    convoyWait(c);
#else
    __syncthreads();
#endif
}
template <bool CONVOY> MLB_DEV void convoyDone(const Convoy& c) {
#if MLB_SPLIT_BARRIER
    if (CONVOY && MLB_EVAL_BARRIER) convoyArrive(c);
#endif
}

/* One Runge-Kutta attempt of either family. The stage loop has ONE call site of the (large) state-derivative code, shared
   by the first stage and by the classical and the adaptive stage composition, so the force model is instantiated once in the
   kernel. `active` false: the lane only keeps the CTA's barriers company.
   The reference's adaptive integrator retries a step in place with a third of the step size until the error estimate passes
   (Integration.py:331-347). Here a rejected attempt returns (`rejected`, with the step size it used) and the lane re-attempts in
   the time loop's next iteration with `retry` set: same arithmetic, same attempt sequence, but the re-attempt rides along with the
   other lanes' next step instead of holding the whole convoy for a pass in which only the rejected lanes work. */
template <int PH, bool CONVOY>
MLB_DEV StepResult integrateStep(const double* B, Lane& L, const KStore& ks, Convoy& cv, State& y, double t, double dt, bool active, bool retry) {
    /* A fixed-rate control system swaps an adaptive integrator for RK4 while it is in charge (ControlSystems.py:126-135); the
       parachute switch restores the original one (rocket.py:700-726). So the method is a per-lane choice between two tableaus;
       the stage loop below runs to the CTA-uniform maximum row count and lanes with fewer rows sit out the extra stages. */
    const int method0 = (int)B[H_INTEGRATOR];
    const bool rk4Lanes = B[H_CTRL_OFF] != 0 && B[(int)B[H_CTRL_OFF] + CS_FORCE_RK4] != 0;
    const int nRows0 = (method0 == INT_EULER) ? 0 : (int)B[H_TAB_NSTAGE_ROWS];
    const int nRowsLoop = rk4Lanes ? max(nRows0, 3) : nRows0;
    const bool forced = rk4Lanes && L.forceRK4;             /* without such a control system the choice is the model block's alone */
    const int method = forced ? INT_RK4 : method0;
    const bool adaptive = method >= INT_RK12A;
    const double* tab = B + (forced ? (int)B[(int)B[H_CTRL_OFF] + CS_TABLEAU2_OFF] : (int)B[H_TABLEAU_OFF]);
    const int nRows = forced ? 3 : nRows0;
    const int nResultRows = forced ? 1 : (int)B[H_TAB_NRESULT_ROWS];
    const double target = B[H_AD_TARGET], minDt = B[H_AD_MINDT];
    StepResult r; r.factor = 1.0; r.err = 0.0; r.dt = dt; r.rejected = false;
    double err = 10000000;
    constexpr bool dof3 = (PH == 1);
    bool attempting = active;
    if (attempting && adaptive) {
        /* Integration.py:335-340: the first attempt uses max(minDt, (3 dt) / 3), every re-attempt a third of the previous attempt's size */
        if (retry) { L.nRejects++; dt = fmax(minDt, dt / 3); }
        else dt = fmax(minDt, (dt * 3) / 3);
    }
    {
        for (int i = -1; i < nRowsLoop; i++) {
            /* slot 0 may already hold the last stage of the previous accepted step (first-same-as-last) */
            const bool evalNow = attempting && i < nRows && !(i < 0 && adaptive && L.haveCache);
            State evalY; double evalTime = t;
            if (evalNow) {
                if (i < 0) evalY = y;
                else {
                    const double* row = tab + i * TABLEAU_STRIDE;
                    evalTime = t + dt * row[0];
                    if (!adaptive) {                                                        /* :216-231 */
                        Deriv evalK = derivScaled(ks.get(0, dof3), 1 / (1 / (row[1] * dt)), dof3);
                        for (int a = 2; a < i + 2; a++) evalK = derivAdd(evalK, derivScaled(ks.get(a - 1, dof3), 1 / (1 / (row[a] * dt)), dof3), dof3);
                        evalY = stateAdd(y, derivTimes(evalK, 1.0, dof3), dof3);
                    } else {                                                                /* :420-432 */
#if MLB_ROTATION_CACHE
                        if (!dof3) {
                            /* the same sum of increments, composed as unit quaternions and normalised once at the end instead of before
                               and after every product (the products of unit quaternions stay unit to rounding) */
                            State evalK = ks.increment(0, row[1]);
                            for (int a = 2; a < i + 2; a++) {
                                const State term = ks.increment(a - 1, row[a]);
                                evalK.pos = evalK.pos + term.pos; evalK.vel = evalK.vel + term.vel; evalK.w = evalK.w + term.w;
                                evalK.q = term.q * evalK.q;
                            }
                            qNormalize(evalK.q);
                            evalY = stateAdd(y, stateScale(evalK, dt, dof3), dof3);
                        } else
#endif
                        {
                            State evalK = derivTimes(ks.get(0, dof3), row[1], dof3);
                            for (int a = 2; a < i + 2; a++) evalK = stateAdd(evalK, derivTimes(ks.get(a - 1, dof3), row[a], dof3), dof3);
                            evalY = stateAdd(y, stateScale(evalK, dt, dof3), dof3);
                        }
                    }
                }
            }
            convoySync<CONVOY>(cv);
            Deriv kNew;
            if (evalNow) kNew = stateDerivative<PH>(B, L, evalTime, evalY);
            convoyDone<CONVOY>(cv);
            if (evalNow) ks.put(i + 1, kNew, dof3);
        }
        if (attempting) {
            if (method == INT_EULER) {                                                      /* :192-195 */
                r.newValue = stateAdd(y, derivTimes(ks.get(0, dof3), dt, dof3), dof3);
                attempting = false;
            } else {
                /* result-row weights as the reference's algebra applies them, 1 / (1 / c): constants, stored by the host after the rows */
                const double* fr = tab + (nRows + nResultRows) * TABLEAU_STRIDE;
                Deriv k0 = ks.get(0, dof3);
                Deriv fd = derivScaled(k0, fr[0], dof3);
                if (!adaptive) {                                                            /* :233-236 */
                    for (int i = 1; i < nRows + 1; i++) fd = derivAdd(fd, derivScaled(ks.get(i, dof3), fr[i], dof3), dof3);
                    r.newValue = stateAdd(y, derivTimes(fd, dt, dof3), dof3);
                    attempting = false;
                } else {
                    const double* cr = fr + TABLEAU_STRIDE;                                 /* :434-447 */
                    Deriv cd = derivScaled(k0, cr[0], dof3);
                    for (int i = 1; i < nRows + 1; i++) {
                        Deriv ki = ks.get(i, dof3);
                        fd = derivAdd(fd, derivScaled(ki, fr[i], dof3), dof3);
                        cd = derivAdd(cd, derivScaled(ki, cr[i], dof3), dof3);
                    }
                    r.newValue = stateAdd(y, derivTimes(fd, dt, dof3), dof3);
                    State coarsePred = stateAdd(y, derivTimes(cd, dt, dof3), dof3);
                    State neg = stateNeg(coarsePred, dof3);
                    State errEst = stateAdd(neg, r.newValue, dof3);
                    err = stateAbs(errEst, dof3);
                    if (!(err > 20 * target) || pyIsClose(dt, minDt)) attempting = false;   /* :338-347 */
                }
            }
        }
    }
    r.dt = dt;
    if (!active || !adaptive) return r;
    if (attempting) { r.rejected = true; return r; }
    /* accepted adaptive step: FSAL cache, step-size controller (:349-389) */
    if (method == INT_RK45A || method == INT_RK23A) {
        ks.copy(0, nRows);                                   /* becomes k_0 of the next step (attempts never overwrite slot 0) */
        L.haveCache = true;
    }
    const double maxDt = B[H_AD_MAXDT];
    double desired;
    const int ctrl = (int)B[H_AD_CTRL];
    if (ctrl == CTRL_PID) desired = 1 + stepPID(B, L, fmin(1.0, (err - target) / target), dt);
    else if (ctrl == CTRL_ELEMENTARY) desired = sqrt(target / (2 * err));
    else desired = 1;
    double minFactor = fmax(minDt / dt, B[H_AD_MINF]);
    double maxFactor = fmin(maxDt / dt, B[H_AD_MAXF]);
    r.factor = B[H_AD_SAFETY] * fmin(fmax(desired, minFactor), maxFactor);
    r.err = err;
    return r;
}

MLB_DEV bool isAdaptive(const double* B) { return (int)B[H_INTEGRATOR] >= INT_RK12A; }
MLB_DEV bool laneAdaptive(const double* B, const Lane& L) {
    const bool rk4Lanes = B[H_CTRL_OFF] != 0 && B[(int)B[H_CTRL_OFF] + CS_FORCE_RK4] != 0;
    return !(rk4Lanes && L.forceRK4) && isAdaptive(B);
}

/* ---------------------------------------------------------------------------------------------
 * Flight events
 * ------------------------------------------------------------------------------------------- */
MLB_DEV void subscribeEvent(Lane& L, int type, double value, double delay, int cb) {
    if (L.nEvents >= MLB_MAX_EVENTS) { L.eventOverflow = true; return; }      /* never dropped silently: the time loop ends the lane */
    Event& e = L.events[L.nEvents++];
    e.type = type; e.value = value; e.delay = delay; e.callback = cb;
}
MLB_DEV void deployNextRecoveryStage(const double* B, Lane& L) {                            /* Recovery.py:65-97 */
    const double* C = B + (int)B[H_RECOVERY_OFF];
    L.recoveryStage += 1;
    if (L.recoveryStage == 1) {
        /* Rocket._switchTo3DoF (rocket.py:700-726): new 3-DoF body with a fresh integrator (no FSAL cache, fresh PID) */
        L.underChute = true;
        L.dof3 = true; L.s.q = qIdentity(); L.s.w = v3(0, 0, 0);
        L.haveCache = false; L.pidLastError = 0; L.pidIntegral = 0;
        L.forceRK4 = false;
    }
    if (L.recoveryStage < (int)C[RC_NSTAGES]) {
        const double* st = C + RC_STAGE0 + L.recoveryStage * RC_STAGE_STRIDE;
        subscribeEvent(L, (int)st[0], st[1], st[4], CB_RECOVERY_DEPLOY);
    }
}
MLB_DEV void stageSeparation(const double* B, Lane& L) {                                    /* Rocket._stageSeparation, rocket.py:461-479 */
    if (L.nStages <= 1) return;
    if (L.sepOut) {
        /* SingleSimulations.createNewDetachedStage (:280-299): the dropped stage starts from a copy of the whole rocket's state and time,
           with the step size the time loop holds at this moment */
        double* o = L.sepOut + (size_t)((int)B[H_NSTAGES] - L.nStages) * SEP_W;
        o[0] = L.s.pos.x; o[1] = L.s.pos.y; o[2] = L.s.pos.z; o[3] = L.s.vel.x; o[4] = L.s.vel.y; o[5] = L.s.vel.z;
        o[6] = L.s.q.w; o[7] = L.s.q.x; o[8] = L.s.q.y; o[9] = L.s.q.z; o[10] = L.s.w.x; o[11] = L.s.w.y; o[12] = L.s.w.z;
        o[SEP_TIME] = L.time; o[SEP_DT] = L.dtEntry; o[SEP_VALID] = 1.0;
    }
    L.nStages -= 1;                /* the bottom stage leaves; boat-tail ownership and fineness ratio follow the stage count */
    L.resorted = true;
    L.inertiaCached = false;       /* fewer stages, and the next motor is about to ignite */
    const double* S0 = B + (int)B[H_STAGE_OFF];
    L.ignitionTime[0] = L.time;    /* stages[0].motor.updateIgnitionTime(currentTime) — the TOP stage's motor, as in the reference */
    /* ignitionTime + times[-1] (Propulsion.py:153-157): the table's last time AFTER the burn-time factor, unlike the value recorded at
       construction (:102-107, S_BURN_DURATION) */
    double burnTime = S0[S_BURN_DURATION];
    if ((int)S0[S_MOTOR] >= 0) {
        const double* C = B + (int)S0[S_COMP_OFF + (int)S0[S_MOTOR]]; const int n = (int)C[MT_NROWS];
        burnTime = motorAdjOf(L) ? C[MT_TABLE + MT_COL_RAW_TIME * n + n - 1] * L.motorBurn[0] : C[MT_TABLE + n - 1];
    }
    L.engineShutOff[0] = L.time + burnTime;
}
MLB_DEV void triggerEvents(const double* B, Lane& L, double& estTimeToNext, bool& deterministic) {   /* simEventDetector.py:64-118 */
    double est = 1e10; bool det = false;
    /* Environment.convertStateToENUFrame (environment.py:130-144): altitude ASL and the vertical velocity in the local ENU frame */
    double alt = L.s.pos.z, vz = L.s.vel.z;
    if (earthIsRound(B)) {
        Geodetic geo = cartesianToGeodetic(B, L.s.pos);
        alt = geo.height;
        Q4 enuToGlobal = qConj(enuRotationOf(geo));
        vz = qRotate(enuToGlobal, L.s.vel).z;
    }
    const int n0 = L.nEvents;
    unsigned removeMask = 0;
    for (int i = 0; i < n0; i++) {
        Event ev = L.events[i];
        bool occurred = false, timeDet = false; double timeTo = 1e10;
        switch (ev.type) {
            case EV_APOGEE:
                occurred = (vz <= 0 && L.time > 1.0);
                if (L.haveLast && (L.time - L.lastTime) != 0) {
                    double accel = (vz - L.lastVz) / (L.time - L.lastTime);
                    if (accel < 0 && vz > 0) timeTo = -vz / accel;
                }
                break;
            case EV_ASCENDING_ALT:
                occurred = (alt >= ev.value && vz > 0);
                if (vz > 0) timeTo = (ev.value - alt) / vz;
                break;
            case EV_DESCENDING_ALT:
                occurred = (alt <= ev.value && vz < 0);
                if (vz < 0) timeTo = (ev.value - alt) / vz;
                break;
            case EV_MOTOR_BURNOUT:
                occurred = (L.time >= L.engineShutOff[stagesOf(B, L) - 1]);
                timeTo = L.engineShutOff[stagesOf(B, L) - 1] - L.time; timeDet = true;
                break;
            case EV_TIME_REACHED:
                occurred = (L.time >= ev.value);
                timeTo = ev.value - L.time; timeDet = true;
                break;
        }
        if (timeTo < est && !occurred) { est = timeTo; det = timeDet; }
        if (occurred) {
            if (ev.delay == 0) { if (ev.callback == CB_RECOVERY_DEPLOY) deployNextRecoveryStage(B, L); else if (ev.callback == CB_STAGE_SEPARATION) stageSeparation(B, L); }
            else subscribeEvent(L, EV_TIME_REACHED, L.time + ev.delay, 0, ev.callback);
            removeMask |= 1u << i;
        }
    }
    if (removeMask) {
        int w = 0;
        for (int i = 0; i < L.nEvents; i++) if (!((removeMask >> i) & 1u)) { if (w != i) L.events[w] = L.events[i]; w++; }
        L.nEvents = w;
    }
    L.haveLast = true; L.lastVz = vz; L.lastTime = L.time;
    estTimeToNext = est; deterministic = det;
}

MLB_DEV double aeroKeyOf(const State& s, const Air& e, int key) {                           /* AeroParameters.stringToAeroFunctionMap */
    switch (key) {
        case AKEY_MACH: return len(s.vel - e.wind) / sqrt(1.4 * 287 * e.temp);
        case AKEY_ALTITUDE: return e.asl;
        case AKEY_UNITRE: return len(localFrameAirVel(s, e.wind)) * e.density / e.viscosity;
        default: { V3 v = localFrameAirVel(s, e.wind); if (len(v) == 0) return 0; return angleBetween(v, v3(0, 0, -1)); }
    }
}
/* Rocket._runControlSystem (rocket.py:646-656) -> RocketControlSystem.runControlLoopIfRequired (GNC/ControlSystems.py:170-211):
   stabiliser target, orientation error as a rotation vector, (scheduled) PID gains, vector PID, deflection table, actuators.
   G is the model block in global memory (interpolation tables). */
MLB_COLD_CONTROL void runControlSystem(const double* B, const double* G, Lane& L) {
    if (B[H_CTRL_OFF] == 0 || L.underChute) return;        /* L.hasControl */
    const double* CS = B + (int)B[H_CTRL_OFF];
    const double currentTime = L.time;
    Air env = airProperties(B, L, L.s.pos, currentTime, false);
    if (!((currentTime - L.ctrlLastRun) + 0.0000000001 >= CS[CS_UPDATE_DT])) return;
    const double dt = currentTime - L.ctrlLastRun;
    Q4 target = q4(CS[CS_TARGET_Q], CS[CS_TARGET_Q + 1], CS[CS_TARGET_Q + 2], CS[CS_TARGET_Q + 3]);
    if ((int)CS[CS_MC_TYPE] == 2) {
        /* IdealMomentController (MomentControllers.py:85-93): the orientation becomes the target, the angular velocity zero. The
           reference assigns the navigator's quaternion OBJECT, so in-place normalisations of the state's orientation during the step
           reach the target of the following steps: the lane keeps its own copy and takes it back when the step replaces the state. */
        if (!L.haveTargetQ) { L.targetQ = target; L.haveTargetQ = true; }
        L.s.q = L.targetQ; L.s.w = v3(0, 0, 0); L.qAliased = true;
        L.ctrlLastRun = currentTime;
        return;
    }
    /* (target / orientation).toRotationVector() = orientation.inverse() * target (CythonQuaternion.pyx:121-126,160-162,219-226) */
    double n = qNorm(L.s.q); double sc = 1.0 / (n * n);
    Q4 inv = qConj(L.s.q);
    inv = q4(inv.w * sc, inv.x * sc, inv.y * sc, inv.z * sc);
    Q4 errQ = inv * target;
    qNormalize(errQ);
    double nn = sqrt(1 - errQ.w * errQ.w);
    double angle = 2 * m_atan2(nn, errQ.w);
    double err[3] = { 0, 0, 0 };
    if (angle != 0) { err[0] = errQ.x * angle / nn; err[1] = errQ.y * angle / nn; err[2] = errQ.z * angle / nn; }
    double g[ND_MAX_VAL];
    if ((int)CS[CS_MC_TYPE] == 1) {
        double keys[4]; const int nk = (int)CS[CS_GAIN_NKEYS];
        for (int k = 0; k < nk; k++) keys[k] = aeroKeyOf(L.s, env, (int)CS[CS_GAIN_KEYS + k]);
        ndInterpolate(G + (int)CS[CS_GAIN_INTERP_OFF], keys, g, L.ndHint[0]);
    } else for (int k = 0; k < 6; k++) g[k] = CS[CS_GAINS + k];
    double key[ND_MAX_DIM]; const int nk = (int)CS[CS_DEFL_NKEYS];
    for (int k = 0; k < nk; k++) key[k] = aeroKeyOf(L.s, env, (int)CS[CS_DEFL_KEYS + k]);
    for (int k = 0; k < 3; k++) {                           /* PIDController.getNewSetPoint, vector mode, no integral clamp (GNC/PID.py:27-51) */
        const double P = (k < 2) ? g[0] : g[3], I = (k < 2) ? g[1] : g[4], D = (k < 2) ? g[2] : g[5];
        double derivative = (err[k] - L.gncLastError[k]) / dt;
        L.gncIntegral[k] = L.gncIntegral[k] + (err[k] + L.gncLastError[k]) * dt / 2;
        L.gncLastError[k] = err[k];
        key[nk + k] = P * err[k] + I * L.gncIntegral[k] + D * derivative;
    }
    double targets[ND_MAX_VAL];
    ndInterpolate(G + (int)CS[CS_DEFL_INTERP_OFF], key, targets, L.ndHint[1]);
    const int nAct = (int)CS[CS_NACT];
    for (int i = 0; i < nAct; i++) {                        /* FirstOrderActuator.setTargetDeflection (Actuators.py:118-129) */
        L.actLastDefl[i] = actuatorDeflection(B, L, i, currentTime);
        L.actLastTime[i] = currentTime;
        double t = targets[i];
        if (CS[CS_ACT_HASMAX] != 0) t = fmin(t, CS[CS_ACT_MAX]);
        if (CS[CS_ACT_HASMIN] != 0) t = fmax(t, CS[CS_ACT_MIN]);
        L.actTarget[i] = t;
    }
    L.ctrlLastRun = currentTime;
}

MLB_DEV void railMotionConstraint(const double* B, Lane& L) {                                /* launchRail.py:58-81, environment.py:195-207 */
    if (!(B[H_RAIL_LEN] > 0) || !L.onRail) return;
    Q4 rot = qAxisAngle(v3(0, 0, 1), B[H_EARTH_ROT] * L.time);
    V3 currPosition = qRotate(rot, ld3(B + H_RAIL_POS));
    V3 currDirection = qRotate(rot, ld3(B + H_RAIL_DIR));
    double distanceTravelled = dot(L.s.pos - currPosition, currDirection);
    if (distanceTravelled < B[H_RAIL_LEN]) {
        if (distanceTravelled <= 0) {
            L.s.pos = currPosition;
            if (dot(L.s.vel, currDirection) < 0) L.s.vel = cross(v3(0, 0, B[H_EARTH_ROT]), currPosition);
        }
    } else L.onRail = false;
}

/* Rocket.timeStep (rocket.py:658-698) in two halves. The first (rail constraint, event detector, step clamp near events, control
   loop) can deploy the first parachute, which turns the lane into a 3-DoF body: a lane of the free-flight kernel then leaves
   (`left`) with the first half done and its step size already clamped, and the under-chute kernel resumes it with `skipPre`. */
template <int PH, bool CONVOY>
MLB_DEV StepResult rocketTimeStep(const double* B, const double* G, Lane& L, const KStore& ks, Convoy& cv, double& dt, bool active, bool skipPre, bool retry, bool& left) {
    if (active && !skipPre) {
        L.dtEntry = dt;
        railMotionConstraint(B, L);
        double est; bool det;
        triggerEvents(B, L, est, det);
        if (laneAdaptive(B, L) && est < dt) {
            if (det) dt = est + 1e-5;
            else dt = fmax(est / 1.5, B[H_EVENT_DT]);
        }
        if (PH == 0) runControlSystem(B, G, L);             /* `not self.isUnderChute` (rocket.py:646-656) */
        if (L.dof3 != (PH == 1)) { left = true; active = false; }
    }
    StepResult r = integrateStep<PH, CONVOY>(B, L, ks, cv, L.s, L.time, dt, active, retry);
    if (active && !r.rejected) {
        if (PH == 1) L.nSteps3++;
        L.time += r.dt;
        if (L.qAliased) { L.targetQ = L.s.q; L.qAliased = false; }
        L.s = r.newValue;
        L.nSteps++;
    }
    return r;
}

MLB_DEV void endDetector(const double* B, const Lane& L, double dt, bool& end, bool& haveFinalDt, double& finalDt) {   /* SingleSimulations.py:205-248 */
    const double v = B[H_END_VALUE];
    haveFinalDt = false; end = false;
    switch ((int)B[H_END_TYPE]) {
        case END_APOGEE: end = (L.s.vel.z <= 0 && L.time > 1.0); break;
        case END_ALT_ASCENDING: end = (L.s.pos.z >= v); break;
        case END_ALT_DESCENDING: end = (earthAltitude(B, L.s.pos) <= v); break;
        default:
            if (L.time < v && L.time + dt >= v) { haveFinalDt = true; finalDt = v + 1e-14 - L.time; }
            else if (L.time >= v) end = true;
    }
}

/* Per-lane inputs, SoA over lanes ([component][lane], lane fastest) */
struct LaneInputs { const double* wind; const double* initState; const double* pinkSeed; const double* motorAdjust; const double* start; long long nLanes; };

/* fin-set constants that depend on the lane's sampled environment at t = 0 */
MLB_DEV void laneFinConstants(const double* B, Lane& L) {
    int fi = 0;
    for (int k = 0; k < MLB_MAX_FINSETS; k++) L.finThickK[k] = 0;
    for (int si = 0; si < L.nStages; si++) {
        const double* S = B + (int)B[H_STAGE_OFF + si];
        for (int ci = 0; ci < (int)S[S_NCOMP]; ci++) {
            const double* C = B + (int)S[S_COMP_OFF + ci];
            if ((int)C[C_TYPE] == CT_FINSET && fi < MLB_MAX_FINSETS) L.finThickK[fi++] = finThicknessK(B, L, C);
        }
    }
}

MLB_DEV void laneInit(const double* B, Lane& L, const LaneInputs& in, long long lane, uint32_t* mtBase, bool live) {
    L.wind = ld3(B + H_WIND_V);
    if (in.wind) L.wind = v3(in.wind[lane], in.wind[in.nLanes + lane], in.wind[2 * in.nLanes + lane]);
    double st[13];
    for (int i = 0; i < 13; i++) st[i] = in.initState ? in.initState[i * in.nLanes + lane] : B[H_INIT_STATE + i];
    L.s.pos = v3(st[0], st[1], st[2]); L.s.vel = v3(st[3], st[4], st[5]);
    L.s.q = q4(st[6], st[7], st[8], st[9]); L.s.w = v3(st[10], st[11], st[12]);
    L.time = in.start ? in.start[lane] : B[H_START_TIME];
    L.dtEntry = 0; L.sepOut = nullptr;
    L.dof3 = false; L.underChute = false; L.haveLast = false; L.haveCache = false; L.resorted = false;
    L.lastVz = 0; L.lastTime = 0; L.pidLastError = 0; L.pidIntegral = 0;
    L.nSteps = 0; L.nEvals = 0; L.nRejects = 0; L.nEvents = 0; L.eventOverflow = false; L.nSteps3 = 0; L.nEvals3 = 0; L.nEvalsTransonic = 0;
    L.turb = (int)B[H_TURB];
    for (int g = 0; g < 3; g++) { L.png[g].rng.mt = nullptr; L.png[g].rng.idx = 624; L.png[g].rng.hasNext = 0; L.png[g].rng.next = 0; L.png[g].last0 = 0; L.png[g].last1 = 0; L.png[g].nextSlot = 0; L.png[g].lastValTime = 0; }
    if (live && L.turb >= TURB_PINK1D && L.turb <= TURB_PINK3D) {
        const int ng = (L.turb == TURB_PINK1D) ? 1 : (L.turb == TURB_PINK2D ? 2 : 3);
        for (int g = 0; g < ng; g++) {
            double seed = in.pinkSeed ? in.pinkSeed[g * in.nLanes + lane] : B[H_PINK_SEED + g];
            L.png[g].rng.mt = mtBase + ((size_t)lane * 3 + g) * 624;
            pinkInit(L.png[g], (uint64_t)seed, B[H_PINK_C0], B[H_PINK_C1]);
        }
    }
    L.nStages = (int)B[H_NSTAGES];
    L.onRail = B[H_RAIL_LEN] > 0;
    L.hasControl = B[H_CTRL_OFF] != 0;
    L.forceRK4 = L.hasControl && B[(int)B[H_CTRL_OFF] + CS_FORCE_RK4] != 0;
    L.ctrlLastRun = 0;                                      /* RocketControlSystem(initTime=0) */
    for (int k = 0; k < 3; k++) { L.gncLastError[k] = 0; L.gncIntegral[k] = 0; }
    L.ndHint[0] = -1; L.ndHint[1] = -1;
    L.haveTargetQ = false; L.qAliased = false; L.targetQ = qIdentity();
    for (int k = 0; k < CS_MAX_ACT; k++) { L.actLastDefl[k] = 0; L.actLastTime[k] = 0; L.actTarget[k] = 0; }
    for (int si = 0; si < MLB_MAX_STAGES; si++) { L.ignitionTime[si] = 0; L.engineShutOff[si] = -1; }
    for (int si = 0; si < L.nStages; si++) {
        const double* S = B + (int)B[H_STAGE_OFF + si];
        L.ignitionTime[si] = S[S_IGNITION0];
        L.engineShutOff[si] = S[S_BURN_DURATION];
    }
    L.inertiaCached = false;
    L.motorAdj = in.motorAdjust != nullptr;
    for (int si = 0; si < MLB_MAX_STAGES; si++) { L.motorImp[si] = 1; L.motorBurn[si] = 1; L.motorOxW0[si] = 0; L.motorFuelW0[si] = 0; }
    if (L.motorAdj) for (int si = 0; si < L.nStages; si++) {
        const double* S = B + (int)B[H_STAGE_OFF + si];
        if ((int)S[S_MOTOR] < 0) continue;
        const double* C = B + (int)S[S_COMP_OFF + (int)S[S_MOTOR]]; const int n = (int)C[MT_NROWS]; const double* T = C + MT_TABLE;
        L.motorImp[si] = in.motorAdjust[(2 * si) * in.nLanes + lane];
        L.motorBurn[si] = in.motorAdjust[(2 * si + 1) * in.nLanes + lane];
        L.motorOxW0[si] = motorWeightFromEnd(T + MT_COL_RAW_TIME * n, T + MT_COL_OX_FLOW * n, n, 0, L.motorBurn[si]);
        L.motorFuelW0[si] = motorWeightFromEnd(T + MT_COL_RAW_TIME * n, T + MT_COL_FUEL_FLOW * n, n, 0, L.motorBurn[si]);
    }
    if (L.motorAdj) {
        /* the state's position is the CG's (rocket.py:251-291,416-421) and the propellant weights move the CG: launch position of the
           nose + rotate(q0, this lane's CG at t = 0). With a sampled initial state as well, the nose position is recovered from it. */
        L.resorted = false;
        double massSum; const Inertia I0 = rocketInertia(B, L, 0.0, &massSum);
        const V3 nose = in.initState ? (L.s.pos - qRotateUnit(L.s.q, ld3(B + H_INIT_CG0))) : ld3(B + H_NOSE_POS0);
        L.s.pos = nose + qRotateUnit(L.s.q, I0.CG);
    }
    /* RecoverySystem.__init__ -> _deployNextStage(): chute stage 0, subscribes the first trigger */
    L.recoveryStage = -1;
    if (B[H_RECOVERY_OFF] != 0) deployNextRecoveryStage(B, L);
    /* Rocket._initializeStagingTriggers (rocket.py:339-361): one separation event per stage that has a trigger */
    for (int si = 0; si < L.nStages; si++) {
        const double* S = B + (int)B[H_STAGE_OFF + si];
        if ((int)S[S_SEP_TYPE] != EV_NONE) subscribeEvent(L, (int)S[S_SEP_TYPE], S[S_SEP_VALUE], S[S_SEP_DELAY], CB_STAGE_SEPARATION);
    }
    laneFinConstants(B, L);
}

}  // namespace mlb
