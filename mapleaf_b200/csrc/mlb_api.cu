/*
 * mlb_api.cu — trajectory kernels and the C ABI of libmapleaf_b200.so (include/mapleaf_b200.h).
 *
 * Kernels (sm_100a, FP64, one lane = one sampled rocket = one thread). A run is a fixed, data-independent sequence of launches;
 * between launches a lane lives in its LaneRec in HBM (written once per launch, ~1.7 KB):
 *   initLanes ......... lane prologue: sampled inputs -> initial state, CPython-exact MT19937 seeding of the pink-noise generators,
 *                       per-lane fin constants, event subscriptions.
 *   flightLanes<B> .... FREE FLIGHT (6-DoF): events, RK stages with the whole component build-up force model fused in, convoy
 *                       execution (one barrier per derivative evaluation, see mlb_integrate.cuh). Launched MLB_GENERATIONS times: a
 *                       CTA whose live lanes drop below a fraction of what it started with hands the stragglers to the next
 *                       generation's queue and retires, so lock-step tails are re-packed into full CTAs of lanes that are at the
 *                       same point of their flights. A lane leaves for good when its end condition fires (results written) or when
 *                       its first parachute opens (queued for chuteLanes).
 *   chuteLanes ........ UNDER PARACHUTE (3-DoF): the same integrator compiled without the component force model; small enough for the
 *                       instruction cache, so it runs free of barriers with more registers per lane.
 *   forceEvalKernel ... one Rocket._getAppliedForce per thread (wind-tunnel known-answer path).
 *   momentsPartial/Final  per-column count / mean / M2 of the result block (Chan's pairwise combination), the per-GPU half of the
 *                       dispersion statistics.
 * The model block is staged in shared memory once per CTA; RK stage derivatives live in per-thread local memory.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <map>
#include <mutex>

#include "../../include/mapleaf_b200.h"
#include "mlb_integrate.cuh"

using namespace mlb;

#define MLB_BLOCK 128               /* threads per CTA of the small helper kernels */
#define MLB_K_SLOTS_MAX 14          /* RK78: 13 stages + first-same-as-last */
/* The free-flight kernel is built for three CTA shapes. Larger CTAs make larger convoys (see mlb_integrate.cuh) and more lanes per SM;
   registers per lane shrink accordingly (255 / 128 / 64). mlb_run picks the largest shape that still gives every SM a CTA. */
/* (CTA size, CTAs per SM) of every instantiation; MLB_BLOCK_SIZE selects one by its CTA size. Tuning builds replace the list:
   a header with `#define MLB_FLIGHT_SHAPES(X) X(768,1) X(640,1)` passed with -include (tools/tune.py) */
#ifndef MLB_FLIGHT_SHAPES
#define MLB_FLIGHT_SHAPES(X) X(128, 2) X(512, 1) X(1024, 1)
#endif
#define MLB_PERLANE_BLOCK 512        /* CTA shape of the free-flight kernel that reads per-lane model blocks */
/* under-chute kernel: CTA size and CTAs per SM (registers per lane = 65536 / (block x CTAs)). Same-box sweep (profiles/r2o): 256 x 2 at 128
   registers is 3 % faster than 128 x 4; 96 or 80 registers per lane (128 x 5 / 6) cost 25-30 %. */
#ifndef MLB_CHUTE_BLOCK
#define MLB_CHUTE_BLOCK 256
#endif
#ifndef MLB_CHUTE_CTAS
#define MLB_CHUTE_CTAS 2
#endif
/* launches of the free-flight kernel per run; the last one runs its lanes to the end. A CTA hands its stragglers on when fewer than
   bailPermille of its lanes are still flying: 750 / 12 generations is 1.3 % faster than 500 / 6 at 10^6 lanes (profiles/r2o) */
#ifndef MLB_GENERATIONS
#define MLB_GENERATIONS 12
#endif
#define TRACE_W 16
#define MLB_NCOUNTERS 6

/* A lane between two launches. Everything the time loop carries from step to step. */
struct LaneRec {
    Lane L;
    double k0[12];                  /* first-same-as-last derivative (slot 0 of the stage store) */
    double dt, finalDt;
    double apogee, maxSpeed, maxH;  /* RocketFlight reductions (rocketFlight.py:27-50), kept as running values */
    V3 prevPos, lastPos;
    long long rows, traceStride;    /* flight log */
    int st;
    int flags;                      /* REC_* */
};
#define REC_END 1
#define REC_HAVE_FINAL 2
#define REC_PRE_DONE 4              /* left the free-flight kernel after the first half of Rocket.timeStep (see rocketTimeStep) */
#define REC_RETRY 8                 /* the last attempt was rejected: `dt` is the size it used (see integrateStep) */

struct RunArgs {
    const double* blob; int blobLen; int blobPad; int nSlots;
    LaneInputs in;
    double* results; double* finals; int* status; long long* counters;
    uint32_t* mt;
    double* separations;                  /* [lane][MLB_MAX_STAGES - 1][SEP_W] */
    double* trace; long long traceLane, traceCount, traceMaxRows, traceEvery; long long* traceRows;     /* flight log of lanes [traceLane, traceLane + traceCount) */
    const double* dtReplay; long long nReplay;
    LaneRec* recs;
    const double* laneModels; long long laneModelStride;      /* per-lane model blocks (mlb_set_lane_models): lane i reads its constants from laneModels + i * stride */
    /* lane queues: entries are lane indices. qIn == nullptr: the identity over [0, nLanes) */
    const int* qIn; const int* qInCount;
    int* qNext; int* qNextCount;          /* stragglers of this generation (free flight) */
    int* qChute; int* qChuteCount;        /* lanes whose first parachute has opened */
    int bailPermille;                     /* a free-flight CTA retires when fewer than this share of its initial lanes are still flying; 0: never */
    int ctaSlots;                         /* free-flight CTAs resident at a time (SMs x CTAs per SM): the queue is spread evenly over whole waves of them */
};

/* One row of a lane's flight log (RocketFlight.times / rigidBodyStates, IO/rocketFlight.py:12-25). The log keeps every `stride`-th
   accepted step; when the lane's buffer is full it drops every other stored row and doubles the stride, so a flight of any length
   ends up as maxRows/2 .. maxRows evenly strided states. `last`: the final state is always stored. */
__device__ __noinline__ void traceRow(const RunArgs& a, double* base, const Lane& L, long long& rows, long long& stride, double dt, double err, bool last) {
    if (a.traceEvery == 0) { if (rows >= a.traceMaxRows) return; }                          /* mlb_set_trace: the first maxRows steps, then stop */
    else if (!last && (L.nSteps % stride) != 0) return;
    if (rows >= a.traceMaxRows) {
        if (a.traceMaxRows < 4) { if (!last) return; rows = a.traceMaxRows - 1; }
        else {
            const long long keep = (a.traceMaxRows + 1) / 2;
            for (long long i = 1; i < keep; i++) for (int k = 0; k < TRACE_W; k++) base[i * TRACE_W + k] = base[2 * i * TRACE_W + k];
            rows = keep; stride *= 2;
            if (!last && (L.nSteps % stride) != 0) return;
        }
    }
    double* T = base + rows * TRACE_W;
    T[0] = L.time; T[1] = L.s.pos.x; T[2] = L.s.pos.y; T[3] = L.s.pos.z; T[4] = L.s.vel.x; T[5] = L.s.vel.y; T[6] = L.s.vel.z;
    T[7] = L.s.q.w; T[8] = L.s.q.x; T[9] = L.s.q.y; T[10] = L.s.q.z; T[11] = L.s.w.x; T[12] = L.s.w.y; T[13] = L.s.w.z;
    T[14] = dt; T[15] = err; rows++;
}

MLB_DEV bool laneTraced(const RunArgs& a, long long lane) { return a.trace != nullptr && lane >= a.traceLane && lane < a.traceLane + a.traceCount; }

/* Lane prologue: Simulation.__init__ .. the first endDetector call of Simulation.run (SingleSimulations.py:89-101). */
/* The model block every thread of a CTA reads: staged in shared memory once per CTA. PERLANE kernels (Monte Carlo keys that change the
   model itself: masses, geometry, launch site ...) skip this and read each lane's own block from global memory. */
template <bool PERLANE>
MLB_DEV const double* stageModel(const RunArgs& a) {
    extern __shared__ double smem[];
    if constexpr (PERLANE) return nullptr;
#ifdef MLB_SPEC
    else return kSpecBlob;          /* model-specialised build (tools/spec_header.py): constant-index reads fold into the code */
#else
    else {
        for (int i = threadIdx.x; i < a.blobLen; i += blockDim.x) smem[i] = a.blob[i];
        __syncthreads();
        return smem;
    }
#endif
}

template <bool PERLANE>
__global__ void __launch_bounds__(MLB_BLOCK) initLanes(const RunArgs a) {
    const double* B = stageModel<PERLANE>(a);
    const long long lane = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= a.in.nLanes) return;
    if (PERLANE) B = a.laneModels + lane * a.laneModelStride;
    LaneRec& R = a.recs[lane];
    Lane L;
    laneInit(B, L, a.in, lane, a.mt, true);
    L.G = a.blob;
    R.dt = a.in.start ? a.in.start[a.in.nLanes + lane] : B[H_DT0]; R.finalDt = 0;
    for (int k = 0; k < (MLB_MAX_STAGES - 1) * SEP_W; k++) a.separations[(size_t)lane * (MLB_MAX_STAGES - 1) * SEP_W + k] = 0;
    R.apogee = -10000000; R.maxSpeed = 0; R.maxH = 0;
    R.prevPos = L.s.pos; R.lastPos = L.s.pos;
    R.st = ST_OK;
    if (L.s.pos.z > R.apogee) R.apogee = L.s.pos.z;
    R.maxSpeed = fmax(R.maxSpeed, len(L.s.vel));
    R.maxH = fmax(R.maxH, sqrt(L.s.vel.x * L.s.vel.x + L.s.vel.y * L.s.vel.y));
    long long rows = 0, stride = a.traceEvery > 0 ? a.traceEvery : 1;
    if (laneTraced(a, lane)) traceRow(a, a.trace + (size_t)(lane - a.traceLane) * a.traceMaxRows * TRACE_W, L, rows, stride, 0.0, 0.0, false);
    R.rows = rows; R.traceStride = stride;
    bool end, haveFinal; double finalDt;
    endDetector(B, L, R.dt, end, haveFinal, finalDt);
    R.finalDt = finalDt;
    if (L.eventOverflow) { R.st = ST_UNSUPPORTED; end = true; }
    R.flags = (end ? REC_END : 0) | (haveFinal ? REC_HAVE_FINAL : 0);
    for (int k = 0; k < 12; k++) R.k0[k] = 0;
    R.L = L;
}

/* RocketFlight.getLandingLocation (rocketFlight.py:52-76) and the other per-flight outputs */
MLB_DEV void finishLane(const RunArgs& a, long long lane, const Lane& L, V3 prevPos, V3 lastPos, double apogee, double maxSpeed, double maxH, int st, long long rows) {
    double* R = a.results + lane * RES_SIZE;
    double dz = lastPos.z - prevPos.z;
    if (L.nSteps >= 1 && lastPos.z > dz * 2) {                /* linear extrapolation of the last two positions to z = 0 */
        double dxdz = (lastPos.x - prevPos.x) / dz, dydz = (lastPos.y - prevPos.y) / dz;
        double dzLand = -lastPos.z;
        R[RES_LAND_X] = lastPos.x + dxdz * dzLand; R[RES_LAND_Y] = lastPos.y + dydz * dzLand; R[RES_LAND_Z] = 0.0;
    } else { R[RES_LAND_X] = lastPos.x; R[RES_LAND_Y] = lastPos.y; R[RES_LAND_Z] = lastPos.z; }
    R[RES_APOGEE] = apogee; R[RES_MAX_SPEED] = maxSpeed; R[RES_FLIGHT_TIME] = L.time; R[RES_MAX_HVEL] = maxH;
    double* F = a.finals + lane * FIN_SIZE;
    F[0] = L.s.pos.x; F[1] = L.s.pos.y; F[2] = L.s.pos.z; F[3] = L.s.vel.x; F[4] = L.s.vel.y; F[5] = L.s.vel.z;
    F[6] = L.s.q.w; F[7] = L.s.q.x; F[8] = L.s.q.y; F[9] = L.s.q.z; F[10] = L.s.w.x; F[11] = L.s.w.y; F[12] = L.s.w.z; F[13] = L.time;
    a.status[lane] = st;
    long long* cn = a.counters + lane * MLB_NCOUNTERS;
    cn[0] = L.nSteps; cn[1] = L.nEvals; cn[2] = L.nRejects; cn[3] = L.nSteps3; cn[4] = L.nEvals3; cn[5] = L.nEvalsTransonic;
    if (laneTraced(a, lane)) a.traceRows[lane - a.traceLane] = rows;
}

/* The Simulation.run loop (SingleSimulations.py:102-139) of one flight phase. PH 0: free flight, convoy execution, CTA-uniform control
   flow (out-of-range and finished threads stay and keep the barriers company). PH 1: under parachute, every thread on its own. */
template <int PH, bool CONVOY, bool PERLANE>
MLB_DEV void flightLoop(const RunArgs& a, const double* B) {
    const long long count = a.qIn ? (long long)*a.qInCount : a.in.nLanes;
    /* Lanes per CTA. The convoy kernel spreads the queue evenly over whole waves of CTAs (a.ctaSlots CTAs run at a time): 10^5 lanes are
       143 CTAs of 704 lanes, one on every SM, rather than 98 full ones; 10^6 lanes 7 equal waves rather than 6.6. A wave's duration follows
       its warps per CTA closely (see mlb_run), so this is worth what the empty SMs were. Warps past `per` exit before the first barrier. */
    int per = blockDim.x;
    if (CONVOY && a.ctaSlots > 0) {
        const long long full = (long long)a.ctaSlots * blockDim.x, waves = (count + full - 1) / full;
        if (waves > 0) per = (int)min((long long)blockDim.x, (((count + a.ctaSlots * waves - 1) / (a.ctaSlots * waves)) + 31) & ~31LL);
    }
    if ((long long)blockIdx.x * per >= count || (int)threadIdx.x >= per) return;          /* nothing queued for this CTA / warp */
    const long long slot = (long long)blockIdx.x * per + threadIdx.x;
    const bool inRange = slot < count;
    const long long lane = inRange ? (a.qIn ? (long long)a.qIn[slot] : slot) : 0;
    if (PERLANE) B = a.laneModels + lane * a.laneModelStride;
    /* RK stage derivatives: per-thread local memory, indexed dynamically by the tableau loops */
    double kLocal[MLB_K_SLOTS_MAX * (PH == 1 ? 6 : MLB_K_SLOT_6DOF)];
    KStore ks; ks.base = kLocal; ks.slotStride = (PH == 1 ? 6 : MLB_K_SLOT_6DOF);

    LaneRec* const rec = a.recs + lane;
    Lane L;
    double dt = 0, finalDt = 0, apogee = 0, maxSpeed = 0, maxH = 0;
    V3 prevPos = v3(0, 0, 0), lastPos = v3(0, 0, 0);
    long long rows = 0, traceStride = 1;
    int st = ST_OK;
    bool end = true, haveFinal = false, skipPre = false, left = false, retry = false;
    if (inRange) {
        L = rec->L;
        dt = rec->dt; finalDt = rec->finalDt; apogee = rec->apogee; maxSpeed = rec->maxSpeed; maxH = rec->maxH;
        prevPos = rec->prevPos; lastPos = rec->lastPos; rows = rec->rows; traceStride = rec->traceStride; st = rec->st;
        end = (rec->flags & REC_END) != 0; haveFinal = (rec->flags & REC_HAVE_FINAL) != 0; skipPre = (rec->flags & REC_PRE_DONE) != 0; retry = (rec->flags & REC_RETRY) != 0;
        if (L.haveCache) {
            Deriv k0;
            k0.vel = v3(rec->k0[0], rec->k0[1], rec->k0[2]); k0.acc = v3(rec->k0[3], rec->k0[4], rec->k0[5]);
            k0.angvel = v3(rec->k0[6], rec->k0[7], rec->k0[8]); k0.angacc = v3(rec->k0[9], rec->k0[10], rec->k0[11]);
            ks.put(0, k0, PH == 1);
        }
        L.G = a.blob; L.sepOut = a.separations + (size_t)lane * (MLB_MAX_STAGES - 1) * SEP_W;
    } else L.forceRK4 = false;          /* threads past the end of the queue only keep the barriers company: the one lane field that path reads */
    const bool tracing = inRange && laneTraced(a, lane);
    double* const traceBase = tracing ? a.trace + (size_t)(lane - a.traceLane) * a.traceMaxRows * TRACE_W : nullptr;
    const long long maxSteps = (long long)B[H_MAX_STEPS];

    Convoy cv; cv.bar = 0; cv.phase = 0;
#if MLB_SPLIT_BARRIER
    if (CONVOY) {
        /* one arrival per warp of the CTA's `per` threads; the first wait of the time loop answers the arrival made here */
        __shared__ unsigned long long convoyBarrier;
        convoyInit(cv, &convoyBarrier, (per + 31) >> 5);
        __syncthreads();
        convoyDone<CONVOY>(cv);
    }
#endif
    int bailBelow = 0;
    if (CONVOY && a.bailPermille > 0) {
        const long long inCta = min((long long)per, count - (long long)blockIdx.x * per);
        bailBelow = (int)((inCta * a.bailPermille + 999) / 1000);
    }
    for (;;) {
        const bool run = !end && !left;
        if (CONVOY) { const int flying = __syncthreads_count(run); if (flying == 0 || flying < bailBelow) break; }
        else if (!run) break;
        if (run && !skipPre && !retry) {
            if (haveFinal) dt = finalDt;
            if (a.dtReplay && L.nSteps < a.nReplay) dt = a.dtReplay[L.nSteps];
        }
        StepResult r = rocketTimeStep<PH, CONVOY>(B, a.blob, L, ks, cv, dt, run, skipPre || retry, retry, left);
        skipPre = false;
        if (run && !left && r.rejected) { retry = true; dt = r.dt; }
        else if (run && !left) {
            retry = false;
            dt = r.dt * r.factor;
            prevPos = lastPos; lastPos = L.s.pos;
            if (L.s.pos.z > apogee) apogee = L.s.pos.z;
            double sp = len(L.s.vel); if (sp > maxSpeed) maxSpeed = sp;
            double hv = sqrt(L.s.vel.x * L.s.vel.x + L.s.vel.y * L.s.vel.y); if (hv > maxH) maxH = hv;
            if (!(L.s.pos.z == L.s.pos.z) || !(L.s.vel.x == L.s.vel.x)) { st = ST_NAN; end = true; }
            else if (L.nSteps >= maxSteps) { st = ST_STEP_LIMIT; end = true; }
            else if (L.eventOverflow) { st = ST_UNSUPPORTED; end = true; }
            else endDetector(B, L, dt, end, haveFinal, finalDt);
            if (tracing) traceRow(a, traceBase, L, rows, traceStride, r.dt, r.err, end);
        }
    }

    /* hand-over: finished lanes write their results; the others go back to HBM and into the queue of the kernel that continues them */
    const bool toChute = inRange && !end && left;
    const bool toNext = inRange && !end && !left;
    if (inRange && end) finishLane(a, lane, L, prevPos, lastPos, apogee, maxSpeed, maxH, st, rows);
    if (toChute || toNext) {
        rec->L = L;
        rec->dt = dt; rec->finalDt = finalDt; rec->apogee = apogee; rec->maxSpeed = maxSpeed; rec->maxH = maxH;
        rec->prevPos = prevPos; rec->lastPos = lastPos; rec->rows = rows; rec->traceStride = traceStride; rec->st = st;
        rec->flags = (haveFinal ? REC_HAVE_FINAL : 0) | (left ? REC_PRE_DONE : 0) | (retry ? REC_RETRY : 0);
        if (L.haveCache) for (int k = 0; k < (PH == 1 ? 6 : 12); k++) rec->k0[k] = kLocal[k];
    }
    if (CONVOY) {
        /* one atomic per CTA and queue: warp ballots, then the warp totals through shared memory */
        __shared__ int warpBase[2][32];
        __shared__ int ctaBase[2];
        const int w = threadIdx.x >> 5, ln = threadIdx.x & 31, nWarps = (per + 31) >> 5;
        const unsigned mNext = __ballot_sync(0xffffffffu, toNext), mChute = __ballot_sync(0xffffffffu, toChute);
        if (ln == 0) { warpBase[0][w] = __popc(mNext); warpBase[1][w] = __popc(mChute); }
        __syncthreads();
        if (threadIdx.x < 2) {
            int sum = 0;
            for (int i = 0; i < nWarps; i++) { int c = warpBase[threadIdx.x][i]; warpBase[threadIdx.x][i] = sum; sum += c; }
            int* counter = threadIdx.x == 0 ? a.qNextCount : a.qChuteCount;
            ctaBase[threadIdx.x] = (sum > 0 && counter) ? atomicAdd(counter, sum) : 0;
        }
        __syncthreads();
        const unsigned below = (1u << ln) - 1u;
        if (toNext && a.qNext) a.qNext[ctaBase[0] + warpBase[0][w] + __popc(mNext & below)] = (int)lane;
        if (toChute && a.qChute) a.qChute[ctaBase[1] + warpBase[1][w] + __popc(mChute & below)] = (int)lane;
    }
}

template <int BLOCK, int MIN_BLOCKS, bool PERLANE>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) flightLanes(const RunArgs a) {
    flightLoop<0, true, PERLANE>(a, stageModel<PERLANE>(a));
}

template <bool PERLANE>
__global__ void __launch_bounds__(MLB_CHUTE_BLOCK, MLB_CHUTE_CTAS) chuteLanes(const RunArgs a) {
    flightLoop<1, false, PERLANE>(a, stageModel<PERLANE>(a));
}


__global__ void __launch_bounds__(MLB_BLOCK) forceEvalKernel(const double* blob, int blobLen, long long n, const double* states, const double* times,
                                                             const double* winds, uint32_t* mt, double* out) {
    extern __shared__ double smem[];
    double* B = smem;
    for (int i = threadIdx.x; i < blobLen; i += blockDim.x) B[i] = blob[i];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Lane L;
    LaneInputs in; in.wind = nullptr; in.initState = nullptr; in.pinkSeed = nullptr; in.motorAdjust = nullptr; in.start = nullptr; in.nLanes = 1;
    laneInit(B, L, in, 0, mt ? mt + (size_t)i * 3 * 624 : nullptr, mt != nullptr);      /* each state gets its own generator block */
    L.G = blob;
    if (winds) { L.wind = v3(winds[3 * i], winds[3 * i + 1], winds[3 * i + 2]); laneFinConstants(B, L); }
    const double* s13 = states + 13 * i;
    State s; s.pos = v3(s13[0], s13[1], s13[2]); s.vel = v3(s13[3], s13[4], s13[5]);
    s.q = q4(s13[6], s13[7], s13[8], s13[9]); s.w = v3(s13[10], s13[11], s13[12]);
    ForceProbe p; Inertia inertia; double massSum;
    p.Mach = 0; p.density = 0; p.wind = v3(0, 0, 0);
    ForceMoment f = appliedForce<0>(B, L, times[i], s, inertia, massSum, &p);
    double* o = out + 18 * i;
    o[0] = f.force.x; o[1] = f.force.y; o[2] = f.force.z; o[3] = f.moment.x; o[4] = f.moment.y; o[5] = f.moment.z;
    o[6] = p.aero.force.x; o[7] = p.aero.force.y; o[8] = p.aero.force.z; o[9] = p.Mach; o[10] = p.density;
    o[11] = p.wind.x; o[12] = p.wind.y; o[13] = p.wind.z;
    o[14] = inertia.CG.x; o[15] = inertia.CG.y; o[16] = inertia.CG.z; o[17] = inertia.mass;
}

/* Per-column (count, mean, M2) over lanes with status OK. One CTA per 7 columns x chunk; partials are combined by
   one thread per column with Chan's update, written to out[15]. Two passes of a tiny kernel: the result block is
   56 B/lane, this is bandwidth-trivial next to the integration. */
__global__ void momentsPartial(const double* results, const int* status, long long nLanes, double* partials /* [grid][15] */) {
    __shared__ double sh[MLB_BLOCK * 3];
    for (int col = 0; col < RES_SIZE; col++) {
        double n = 0, mean = 0, M2 = 0;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nLanes; i += (long long)gridDim.x * blockDim.x) {
            if (status[i] != ST_OK) continue;
            double x = results[i * RES_SIZE + col];
            n += 1; double d = x - mean; mean += d / n; M2 += d * (x - mean);
        }
        sh[threadIdx.x * 3] = n; sh[threadIdx.x * 3 + 1] = mean; sh[threadIdx.x * 3 + 2] = M2;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                double na = sh[threadIdx.x * 3], ma = sh[threadIdx.x * 3 + 1], Sa = sh[threadIdx.x * 3 + 2];
                double nb = sh[(threadIdx.x + s) * 3], mb = sh[(threadIdx.x + s) * 3 + 1], Sb = sh[(threadIdx.x + s) * 3 + 2];
                double nt = na + nb;
                if (nt > 0) { double d = mb - ma; sh[threadIdx.x * 3] = nt; sh[threadIdx.x * 3 + 1] = ma + d * (nb / nt); sh[threadIdx.x * 3 + 2] = Sa + Sb + d * d * (na * nb / nt); }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) { double* P = partials + (size_t)blockIdx.x * 15; if (col == 0) P[0] = sh[0]; P[1 + 2 * col] = sh[1]; P[2 + 2 * col] = sh[2]; }
        __syncthreads();
    }
}
__global__ void momentsFinal(const double* partials, int nPartials, double* out15) {
    int col = threadIdx.x;
    if (col >= RES_SIZE) return;
    double n = 0, mean = 0, M2 = 0;
    for (int p = 0; p < nPartials; p++) {
        const double* P = partials + (size_t)p * 15;
        double nb = P[0], mb = P[1 + 2 * col], Sb = P[2 + 2 * col];
        double nt = n + nb;
        if (nt > 0) { double d = mb - mean; mean = mean + d * (nb / nt); M2 = M2 + Sb + d * d * (n * nb / nt); n = nt; }
    }
    if (col == 0) out15[0] = n;
    out15[1 + 2 * col] = mean; out15[2 + 2 * col] = M2;
}

/* ---------------------------------------------------------------------------------------------
 * Host side
 * ------------------------------------------------------------------------------------------- */
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(MLB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

struct mlb_handle {
    int device = 0;
    std::vector<double> blob;
    double* d_blob = nullptr;
    long long nLanes = 0;
    long long capLanes = 0;             /* lanes the per-lane device buffers below are sized for (kept across mlb_set_lanes calls) */
    double* d_lane[LANE_NUM_ARRAYS] = {};
    size_t laneBytes[LANE_NUM_ARRAYS] = {};        /* capacity of the library-owned copies */
    double* d_laneOwned[LANE_NUM_ARRAYS] = {};
    bool ownResults = true;
    double* d_results = nullptr; double* d_finals = nullptr; int* d_status = nullptr; long long* d_counters = nullptr;
    double* d_resultsOwned = nullptr; int* d_statusOwned = nullptr;
    uint32_t* d_mt = nullptr;
    double* d_separations = nullptr;
    LaneRec* d_recs = nullptr;
    double* d_laneModels = nullptr; size_t laneModelBytes = 0; long long laneModelStride = 0; bool haveLaneModels = false;
    int* d_queue[3] = { nullptr, nullptr, nullptr };      /* two straggler queues (ping-pong) and the under-chute queue */
    int* d_queueCount = nullptr;                          /* MLB_GENERATIONS + 1 counters */
    double* d_trace = nullptr; long long traceLane = -1, traceCount = 0, traceMaxRows = 0, traceEvery = 1; long long* d_traceRows = nullptr;
    double* d_replay = nullptr; long long nReplay = 0;
    double* d_partials = nullptr; double* d_moments = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evFlight0 = nullptr, evFlight1 = nullptr;
    cudaStream_t stream = nullptr;
    bool ran = false;
    long long launches = 0;
    int nSlots = 0; size_t smemBytes = 0;       /* the staged model block (forceEvalKernel; the trajectory kernels of the generic build) */
    size_t trajSmemBytes = 0;                   /* dynamic shared memory of the trajectory kernels: smemBytes, or 0 in a model-specialised build (no staging) */
    int nSM = 148; int forceBlock = 0; int lastBlock = 0; int lastGenerations = 0; int bailPermille = 750; int generations = MLB_GENERATIONS; bool evenSplit = true;
    size_t sharedLen = 0;
    long long tuneLanes = 0; double tuneMs[2] = { -1, -1 };    /* free-flight time of this batch size in the 1024- and the 512-lane shape (shape tuning, mlb_run) */
    int tuneRunBlock = 0;                                      /* shape of the last run if its time has not been read yet */
};

#ifdef MLB_SPEC
/* Shape tuning of a model-specialised library. Once the model block is compiled in, 512 lanes x 128 registers and 1024 x 64 are within a
   few per cent of each other and which one wins depends on the rocket (same-box sweep of the five bench workloads, profiles/r3l: 512 is
   +4 % / +13 % / +5 % on the single-stage adaptive / fixed-step / NASA cases, 1024 is +15 % on the canard case, whose force model and
   control loop are twice the code). Results do not depend on the shape (same arithmetic per lane), so a handle simply measures: the
   first multi-wave run of a batch size flies in the 1024-lane shape, the second in the 512-lane one, later runs in the faster of the
   two; the choice is kept per library (= per model block) and batch size for every handle of the process. */
static std::mutex g_tuneMutex;
static std::map<long long, int> g_tunedBlock;
static int tunedShape(mlb_handle* h) {
    std::lock_guard<std::mutex> lock(g_tuneMutex);
    auto it = g_tunedBlock.find(h->nLanes);
    if (it != g_tunedBlock.end()) return it->second;
    if (h->tuneLanes != h->nLanes) { h->tuneLanes = h->nLanes; h->tuneMs[0] = h->tuneMs[1] = -1; h->tuneRunBlock = 0; }
    else if (h->tuneRunBlock && cudaEventQuery(h->evFlight1) == cudaSuccess) {      /* the previous run of this batch size has finished: take its time */
        float ms = 0;
        if (cudaEventElapsedTime(&ms, h->evFlight0, h->evFlight1) == cudaSuccess) h->tuneMs[h->tuneRunBlock == 1024 ? 0 : 1] = ms;
    }
    h->tuneRunBlock = 0;
    (void)cudaGetLastError();                               /* cudaEventQuery of a pending event reports cudaErrorNotReady: not an error */
    if (h->tuneMs[0] < 0) return 1024;
    if (h->tuneMs[1] < 0) return 512;
    const int best = h->tuneMs[1] < h->tuneMs[0] ? 512 : 1024;
    g_tunedBlock[h->nLanes] = best;
    return best;
}
#endif

static int laneArrayComponents(int id) { return id == LANE_WIND ? 3 : (id == LANE_INIT_STATE ? 13 : (id == LANE_PINK_SEED ? 3 : (id == LANE_MOTOR_ADJUST ? 2 * MLB_MAX_STAGES : (id == LANE_START ? 2 : -1)))); }

static void freeLanes(mlb_handle* h) {
    for (int i = 0; i < LANE_NUM_ARRAYS; i++) { cudaFree(h->d_laneOwned[i]); h->d_laneOwned[i] = nullptr; h->laneBytes[i] = 0; h->d_lane[i] = nullptr; }
    cudaFree(h->d_resultsOwned); cudaFree(h->d_statusOwned);
    h->d_resultsOwned = nullptr; h->d_statusOwned = nullptr; h->ownResults = true;
    cudaFree(h->d_finals); cudaFree(h->d_counters); cudaFree(h->d_mt); cudaFree(h->d_recs); cudaFree(h->d_laneModels); cudaFree(h->d_separations); h->d_separations = nullptr;
    h->d_laneModels = nullptr; h->laneModelBytes = 0; h->haveLaneModels = false;
    for (int i = 0; i < 3; i++) { cudaFree(h->d_queue[i]); h->d_queue[i] = nullptr; }
    h->d_results = h->d_finals = nullptr; h->d_status = nullptr; h->d_counters = nullptr; h->d_mt = nullptr; h->d_recs = nullptr;
    h->nLanes = 0; h->capLanes = 0; h->ran = false;
}

extern "C" const char* mlb_last_error(void) { return g_err.c_str(); }
#ifdef MLB_SPEC
extern "C" const char* mlb_version(void) { return "mapleaf_b200 0.2 (sm_100a, model-specialised)"; }
extern "C" int mlb_is_specialized(void) { return 1; }
#else
extern "C" const char* mlb_version(void) { return "mapleaf_b200 0.2 (sm_100a)"; }
extern "C" int mlb_is_specialized(void) { return 0; }
#endif

static int createOnDevice(mlb_handle* h, const double* b, size_t len) {
    CU(cudaMalloc(&h->d_blob, len * sizeof(double)));
    CU(cudaMemcpy(h->d_blob, b, len * sizeof(double), cudaMemcpyHostToDevice));
    /* the kernels' working set is per-thread local memory: leave to L1 everything the staged model block does not need */
    const int carveout = getenv("MLB_SMEM_CARVEOUT") ? atoi(getenv("MLB_SMEM_CARVEOUT")) : (int)((h->trajSmemBytes + 1024) * 100 / (228 * 1024)) + 1;
#define MLB_SET_SMEM(N, M) CU((cudaFuncSetAttribute(flightLanes<N, M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->trajSmemBytes))); \
    if (carveout >= 0) CU((cudaFuncSetAttribute(flightLanes<N, M, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout)));
    MLB_FLIGHT_SHAPES(MLB_SET_SMEM)
    CU(cudaFuncSetAttribute(chuteLanes<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->trajSmemBytes));
    CU(cudaFuncSetAttribute(initLanes<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->trajSmemBytes));
    CU(cudaFuncSetAttribute(forceEvalKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smemBytes));
    { cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, h->device)); h->nSM = prop.multiProcessorCount; }
    CU(cudaEventCreate(&h->ev0)); CU(cudaEventCreate(&h->ev1)); CU(cudaEventCreate(&h->evFlight0)); CU(cudaEventCreate(&h->evFlight1));
    CU(cudaMalloc(&h->d_partials, 148 * 15 * sizeof(double)));
    CU(cudaMalloc(&h->d_moments, 15 * sizeof(double)));
    CU(cudaMalloc(&h->d_queueCount, (MLB_GENERATIONS + 2) * sizeof(int)));
    return MLB_OK;
}

extern "C" int mlb_create(const void* model_blob, size_t n, int device, mlb_handle** out) {
    if (!model_blob || !out || n < H_SIZE * sizeof(double) || n % sizeof(double)) return fail(MLB_ERR_ARG, "mlb_create: model block missing or too short");
    const double* b = (const double*)model_blob;
    size_t len = n / sizeof(double);
    if ((int)b[H_MAGIC] != MLB_MAGIC || (int)b[H_VERSION] != MLB_VERSION) return fail(MLB_ERR_ARG, "mlb_create: not a model block (magic / version mismatch)");
    if ((size_t)b[H_LEN] != len) return fail(MLB_ERR_ARG, "mlb_create: model block length field does not match the buffer");
#ifdef MLB_SPEC
    /* model-specialised build: the kernels carry ONE model block as constants and serve no other */
    {
        const size_t shared = (size_t)b[H_SHARED_LEN] > 0 && (size_t)b[H_SHARED_LEN] <= len ? (size_t)b[H_SHARED_LEN] : len;
        if (shared != (size_t)MLB_SPEC_LEN) return fail(MLB_ERR_ARG, "mlb_create: this library is specialised for another model block (length)");
        for (size_t i = 0; i < shared; i++)
            if (memcmp(&b[i], &kSpecBlobHost[i], sizeof(double)) != 0 && !(b[i] != b[i] && kSpecBlobHost[i] != kSpecBlobHost[i]))
                return fail(MLB_ERR_ARG, "mlb_create: this library is specialised for another model block (word " + std::to_string(i) + " differs)");
    }
#endif
    int nStages = (int)b[H_NSTAGES];
    if (nStages < 1 || nStages > MLB_MAX_STAGES) return fail(MLB_ERR_UNSUPPORTED, "mlb_create: stage count out of range");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(MLB_ERR_CUDA, std::string("mlb_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
    if (device < 0 || device >= count) return fail(MLB_ERR_ARG, "mlb_create: device index out of range");
    CU(cudaSetDevice(device));
    mlb_handle* h = new mlb_handle();
    h->device = device;
    h->blob.assign(b, b + len);
    int method = (int)b[H_INTEGRATOR];
    int nRows = (int)b[H_TAB_NSTAGE_ROWS];
    h->nSlots = (method == INT_EULER) ? 1 : nRows + 1;
    h->sharedLen = (size_t)b[H_SHARED_LEN] > 0 && (size_t)b[H_SHARED_LEN] <= len ? (size_t)b[H_SHARED_LEN] : len;      /* the rest stays in global memory */
    size_t blobPad = (h->sharedLen + 1) & ~(size_t)1;
    h->smemBytes = blobPad * sizeof(double);
    if (h->smemBytes > 226 * 1024) { delete h; return fail(MLB_ERR_UNSUPPORTED, "mlb_create: the shared part of the model block exceeds shared memory"); }
#ifdef MLB_SPEC
    h->trajSmemBytes = 0;
#else
    h->trajSmemBytes = h->smemBytes;
#endif
    { const char* e = getenv("MLB_BLOCK_SIZE"); h->forceBlock = e ? atoi(e) : 0; }
    { const char* e = getenv("MLB_BAIL_PERMILLE"); if (e) h->bailPermille = atoi(e); }
    { const char* e = getenv("MLB_EVEN_SPLIT"); if (e) h->evenSplit = atoi(e) != 0; }
    { const char* e = getenv("MLB_NUM_GENERATIONS"); if (e) { int g = atoi(e); if (g >= 1 && g <= MLB_GENERATIONS) h->generations = g; } }
    int rc = createOnDevice(h, b, len);
    if (rc != MLB_OK) { mlb_destroy(h); return rc; }       /* g_err keeps the message of the failed call */
    *out = h;
    return MLB_OK;
}

extern "C" void mlb_destroy(mlb_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    freeLanes(h);
    cudaFree(h->d_blob); cudaFree(h->d_trace); cudaFree(h->d_traceRows); cudaFree(h->d_replay); cudaFree(h->d_partials); cudaFree(h->d_moments); cudaFree(h->d_queueCount);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->evFlight0) cudaEventDestroy(h->evFlight0);
    if (h->evFlight1) cudaEventDestroy(h->evFlight1);
    delete h;
}

/* Per-lane device buffers are sized once for the largest batch seen and kept: a Monte Carlo driver that calls mlb_set_lanes with a
   fresh sample block of the same size every pass re-uses them (at 10^6 lanes the generator states alone are 7.5 GB). */
static int reserveLanes(mlb_handle* h, long long nLanes) {
    if (nLanes <= h->capLanes) return MLB_OK;
    cudaFree(h->d_resultsOwned); cudaFree(h->d_statusOwned); cudaFree(h->d_finals); cudaFree(h->d_counters); cudaFree(h->d_mt); cudaFree(h->d_recs); cudaFree(h->d_separations);
    h->d_separations = nullptr;
    for (int i = 0; i < 3; i++) { cudaFree(h->d_queue[i]); h->d_queue[i] = nullptr; }
    h->d_resultsOwned = nullptr; h->d_statusOwned = nullptr; h->d_finals = nullptr; h->d_counters = nullptr; h->d_mt = nullptr; h->d_recs = nullptr;
    h->capLanes = 0;
    if (nLanes > 0x7fffffffLL) return fail(MLB_ERR_UNSUPPORTED, "mlb_set_lanes: more than 2^31 - 1 lanes per handle");
    CU(cudaMalloc(&h->d_resultsOwned, (size_t)nLanes * RES_SIZE * sizeof(double)));
    CU(cudaMalloc(&h->d_finals, (size_t)nLanes * FIN_SIZE * sizeof(double)));
    CU(cudaMalloc(&h->d_statusOwned, (size_t)nLanes * sizeof(int)));
    CU(cudaMalloc(&h->d_counters, (size_t)nLanes * MLB_NCOUNTERS * sizeof(long long)));
    CU(cudaMalloc(&h->d_recs, (size_t)nLanes * sizeof(LaneRec)));
    CU(cudaMalloc(&h->d_separations, (size_t)nLanes * (MLB_MAX_STAGES - 1) * SEP_W * sizeof(double)));
    for (int i = 0; i < 3; i++) CU(cudaMalloc(&h->d_queue[i], (size_t)nLanes * sizeof(int)));
    int turb = (int)h->blob[H_TURB];
    if (turb >= TURB_PINK1D && turb <= TURB_PINK3D) CU(cudaMalloc(&h->d_mt, (size_t)nLanes * 3 * 624 * sizeof(uint32_t)));
    h->capLanes = nLanes;
    return MLB_OK;
}

static int setLanes(mlb_handle* h, int64_t nLanes, const double* const* arrays, const int* ids, int nArrays, bool onDevice) {
    if (!h) return fail(MLB_ERR_ARG, "null handle");
    if (nLanes <= 0) return fail(MLB_ERR_ARG, "mlb_set_lanes: nLanes must be positive");
    if (nArrays < 0 || (nArrays > 0 && (!arrays || !ids))) return fail(MLB_ERR_ARG, "mlb_set_lanes: array list missing");
    for (int k = 0; k < nArrays; k++) {
        if (laneArrayComponents(ids[k]) < 0) return fail(MLB_ERR_ARG, "mlb_set_lanes: unknown array id");
        if (!arrays[k]) return fail(MLB_ERR_ARG, "mlb_set_lanes: null array");
    }
    CU(cudaSetDevice(h->device));
    h->nLanes = 0; h->ran = false;                         /* the handle has no runnable batch until everything below has succeeded */
    int rc = reserveLanes(h, nLanes);
    if (rc != MLB_OK) { freeLanes(h); return rc; }
    for (int i = 0; i < LANE_NUM_ARRAYS; i++) h->d_lane[i] = nullptr;
    for (int k = 0; k < nArrays; k++) {
        const int id = ids[k];
        if (onDevice) { h->d_lane[id] = (double*)arrays[k]; continue; }
        size_t bytes = (size_t)laneArrayComponents(id) * nLanes * sizeof(double);
        if (bytes > h->laneBytes[id]) {
            cudaFree(h->d_laneOwned[id]); h->d_laneOwned[id] = nullptr; h->laneBytes[id] = 0;
            cudaError_t e = cudaMalloc(&h->d_laneOwned[id], bytes);
            if (e != cudaSuccess) { freeLanes(h); return fail(MLB_ERR_CUDA, std::string("mlb_set_lanes: cudaMalloc: ") + cudaGetErrorString(e)); }
            h->laneBytes[id] = bytes;
        }
        cudaError_t e = cudaMemcpy(h->d_laneOwned[id], arrays[k], bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { freeLanes(h); return fail(MLB_ERR_CUDA, std::string("mlb_set_lanes: cudaMemcpy: ") + cudaGetErrorString(e)); }
        h->d_lane[id] = h->d_laneOwned[id];
    }
    h->d_results = h->d_resultsOwned; h->d_status = h->d_statusOwned; h->ownResults = true;
    h->haveLaneModels = false;
    h->nLanes = nLanes;
    return MLB_OK;
}
extern "C" int mlb_set_lanes(mlb_handle* h, int64_t nLanes, const double* const* arrays, const int* ids, int nArrays) { return setLanes(h, nLanes, arrays, ids, nArrays, false); }
extern "C" int mlb_set_lanes_device(mlb_handle* h, int64_t nLanes, const double* const* arrays, const int* ids, int nArrays) { return setLanes(h, nLanes, arrays, ids, nArrays, true); }

extern "C" int mlb_set_lane_models(mlb_handle* h, const double* models, int64_t strideDoubles) {
    if (!h) return fail(MLB_ERR_ARG, "null handle");
    if (h->nLanes <= 0) return fail(MLB_ERR_STATE, "mlb_set_lane_models: call mlb_set_lanes first");
    if (!models) { h->haveLaneModels = false; return MLB_OK; }
#ifdef MLB_SPEC
    return fail(MLB_ERR_UNSUPPORTED, "mlb_set_lane_models: a model-specialised library flies one model block; per-lane model blocks need the generic library");
#endif
    if (strideDoubles < (int64_t)h->sharedLen) return fail(MLB_ERR_ARG, "mlb_set_lane_models: stride shorter than the shared part of the model block (H_SHARED_LEN)");
    /* every lane's block must be a variant of the handle's: same layout, same length, same integrator and structure offsets */
    static const int same[] = { H_MAGIC, H_VERSION, H_LEN, H_SHARED_LEN, H_NSTAGES, H_INTEGRATOR, H_TAB_NSTAGE_ROWS, H_TAB_NRESULT_ROWS, H_TABLEAU_OFF, H_CTRL_OFF, H_RECOVERY_OFF, H_STAGE_OFF, H_TURB };
    for (int64_t i = 0; i < h->nLanes; i++)
        for (int k : same)
            if (models[i * strideDoubles + k] != h->blob[k]) return fail(MLB_ERR_ARG, "mlb_set_lane_models: lane " + std::to_string(i) + " has a different model structure than the handle's model block (header word " + std::to_string(k) + ")");
    CU(cudaSetDevice(h->device));
    const size_t bytes = (size_t)h->nLanes * strideDoubles * sizeof(double);
    if (bytes > h->laneModelBytes) {
        cudaFree(h->d_laneModels); h->d_laneModels = nullptr; h->laneModelBytes = 0;
        CU(cudaMalloc(&h->d_laneModels, bytes));
        h->laneModelBytes = bytes;
    }
    CU(cudaMemcpy(h->d_laneModels, models, bytes, cudaMemcpyHostToDevice));
    h->laneModelStride = strideDoubles; h->haveLaneModels = true; h->ran = false;
    return MLB_OK;
}

extern "C" int mlb_run(mlb_handle* h, void* stream) {
    if (!h) return fail(MLB_ERR_ARG, "null handle");
    if (h->nLanes <= 0) return fail(MLB_ERR_STATE, "mlb_run: call mlb_set_lanes first");
    CU(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    h->stream = s;
    RunArgs a;
    a.blob = h->d_blob; a.blobLen = (int)h->sharedLen; a.blobPad = (int)((h->sharedLen + 1) & ~(size_t)1); a.nSlots = h->nSlots;
    a.in.wind = h->d_lane[LANE_WIND]; a.in.initState = h->d_lane[LANE_INIT_STATE]; a.in.pinkSeed = h->d_lane[LANE_PINK_SEED]; a.in.motorAdjust = h->d_lane[LANE_MOTOR_ADJUST]; a.in.start = h->d_lane[LANE_START]; a.in.nLanes = h->nLanes;
    a.results = h->d_results; a.finals = h->d_finals; a.status = h->d_status; a.counters = h->d_counters; a.mt = h->d_mt; a.separations = h->d_separations;
    a.trace = (h->traceLane >= 0 && h->traceLane < h->nLanes) ? h->d_trace : nullptr;
    a.traceLane = h->traceLane; a.traceCount = h->traceCount; a.traceMaxRows = h->traceMaxRows; a.traceEvery = h->traceEvery; a.traceRows = h->d_traceRows;
    if (a.trace) CU(cudaMemsetAsync(h->d_traceRows, 0, (size_t)h->traceCount * sizeof(long long), s));
    a.dtReplay = h->d_replay; a.nReplay = h->nReplay;
    a.recs = h->d_recs;
    a.laneModels = h->haveLaneModels ? h->d_laneModels : nullptr; a.laneModelStride = h->laneModelStride;
    const bool perLane = h->haveLaneModels;
    a.qIn = nullptr; a.qInCount = nullptr; a.qNext = nullptr; a.qNextCount = nullptr; a.qChute = h->d_queue[2]; a.qChuteCount = h->d_queueCount + MLB_GENERATIONS; a.bailPermille = 0;
    /* CTA shape of the free-flight kernel: the one with the fewest (waves of CTAs) x (duration of a wave of that shape relative to a
       1024-lane wave). A wave's duration grows a little slower than its lane count (0.45 / 0.55 / 1 for 2 x 128 / 512 / 1024 lanes per SM,
       profiles/r2c_shape_sweep.txt; the ratios are a property of the convoy, the same within a few per cent for every workload
       measured), so large batches want the largest convoy, a batch that leaves SMs empty in one shape may fill them in a smaller
       one, and a batch just over one wave is better off in a larger shape than in two waves of a smaller one. MLB_BLOCK_SIZE overrides. */
    int block = 128; double bestCost = 1e300;
    {
        static const int shapeBlock[3] = { 128, 512, 1024 }, shapeCtas[3] = { 2, 1, 1 };
        static const double shapeWave[3] = { 0.45, 0.55, 1.0 };
        for (int k = 0; k < 3; k++) {
            const long long ctas = (h->nLanes + shapeBlock[k] - 1) / shapeBlock[k], slots = (long long)h->nSM * shapeCtas[k];
            const double cost = (double)((ctas + slots - 1) / slots) * shapeWave[k];
            if (cost < bestCost) { bestCost = cost; block = shapeBlock[k]; }
        }
    }
#ifdef MLB_SPEC
    { const char* e = getenv("MLB_SHAPE_TUNING");
      if (block == 1024 && !h->forceBlock && !perLane && h->nLanes >= 2LL * h->nSM * 1024 && !(e && atoi(e) == 0)) { block = tunedShape(h); h->tuneRunBlock = block; } }
#endif
#define MLB_ALLOW_SHAPE(N, M) if (h->forceBlock == N) block = N;
    MLB_FLIGHT_SHAPES(MLB_ALLOW_SHAPE)
    if (perLane) block = MLB_PERLANE_BLOCK;                    /* one shape: every constant is a global load there, fewer lanes per SM keep them cached */
    h->lastBlock = block;
    int ctasPerSM = 1;
#define MLB_SHAPE_CTAS(N, M) if (block == N) ctasPerSM = M;
    MLB_FLIGHT_SHAPES(MLB_SHAPE_CTAS)
    const int ctaSlots = h->nSM * ctasPerSM;
    const unsigned grid = (unsigned)((h->nLanes + block - 1) / block) + (h->evenSplit ? ctaSlots + 1 : 0);      /* even split: at most this many CTAs */
    a.ctaSlots = h->evenSplit ? ctaSlots : 0;
    h->launches = 0;
    CU(cudaEventRecord(h->ev0, s));
    CU(cudaMemsetAsync(h->d_queueCount, 0, (MLB_GENERATIONS + 2) * sizeof(int), s));
#ifndef MLB_SPEC
    if (perLane) initLanes<true><<<(unsigned)((h->nLanes + MLB_BLOCK - 1) / MLB_BLOCK), MLB_BLOCK, 0, s>>>(a);
    else
#endif
    initLanes<false><<<(unsigned)((h->nLanes + MLB_BLOCK - 1) / MLB_BLOCK), MLB_BLOCK, h->trajSmemBytes, s>>>(a);
    h->launches++;
    CU(cudaGetLastError());
    /* Free flight, MLB_GENERATIONS launches over ping-pong straggler queues. Every launch is sized for the whole batch (CTAs past the end
       of their queue exit at once), so the sequence needs no host round trip. */
    CU(cudaEventRecord(h->evFlight0, s));
    /* a batch that fits one wave gains nothing from re-packing (no CTA is waiting for the SM): one generation runs it to the end */
    const bool singleWave = (h->nLanes + block - 1) / block <= (long long)ctaSlots;
    const int G = singleWave ? 1 : h->generations;
    h->lastGenerations = G;
    for (int g = 0; g < G; g++) {
        a.qIn = g == 0 ? nullptr : h->d_queue[(g - 1) & 1];
        a.qInCount = g == 0 ? nullptr : h->d_queueCount + (g - 1);
        a.qNext = h->d_queue[g & 1]; a.qNextCount = h->d_queueCount + g;
        a.bailPermille = (g == G - 1 || singleWave) ? 0 : h->bailPermille;
        bool launched = perLane;
#define MLB_LAUNCH_SHAPE(N, M) if (block == N && !perLane) { flightLanes<N, M, false><<<grid, N, h->trajSmemBytes, s>>>(a); launched = true; }
        MLB_FLIGHT_SHAPES(MLB_LAUNCH_SHAPE)
        if (!launched) return fail(MLB_ERR_STATE, "mlb_run: this build has no free-flight kernel of " + std::to_string(block) + " threads (tuning build with a reduced MLB_FLIGHT_SHAPES list)");
#ifndef MLB_SPEC
        if (perLane) flightLanes<MLB_PERLANE_BLOCK, 1, true><<<grid, MLB_PERLANE_BLOCK, 0, s>>>(a);
#endif
        h->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(h->evFlight1, s));
    /* Under parachute: everything the free-flight launches queued */
    a.qIn = h->d_queue[2]; a.qInCount = h->d_queueCount + MLB_GENERATIONS; a.qNext = nullptr; a.qNextCount = nullptr; a.qChute = nullptr; a.qChuteCount = nullptr; a.bailPermille = 0; a.ctaSlots = 0;
#ifndef MLB_SPEC
    if (perLane) chuteLanes<true><<<(unsigned)((h->nLanes + MLB_CHUTE_BLOCK - 1) / MLB_CHUTE_BLOCK), MLB_CHUTE_BLOCK, 0, s>>>(a);
    else
#endif
    chuteLanes<false><<<(unsigned)((h->nLanes + MLB_CHUTE_BLOCK - 1) / MLB_CHUTE_BLOCK), MLB_CHUTE_BLOCK, h->trajSmemBytes, s>>>(a);
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev1, s));
    h->ran = true;
    return MLB_OK;
}

extern "C" int mlb_sync(mlb_handle* h) {
    if (!h) return fail(MLB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return MLB_OK;
}

static int needRun(mlb_handle* h, const char* who) {
    if (!h) return fail(MLB_ERR_ARG, "null handle");
    if (!h->ran) return fail(MLB_ERR_STATE, std::string(who) + ": no completed mlb_run");
    return mlb_sync(h);
}

extern "C" int mlb_results(mlb_handle* h, double* results, int32_t* status) {
    int rc = needRun(h, "mlb_results"); if (rc) return rc;
    if (results) CU(cudaMemcpy(results, h->d_results, (size_t)h->nLanes * RES_SIZE * sizeof(double), cudaMemcpyDeviceToHost));
    if (status) CU(cudaMemcpy(status, h->d_status, (size_t)h->nLanes * sizeof(int), cudaMemcpyDeviceToHost));
    return MLB_OK;
}
extern "C" int mlb_finals(mlb_handle* h, double* finals) {
    int rc = needRun(h, "mlb_finals"); if (rc) return rc;
    if (!finals) return fail(MLB_ERR_ARG, "mlb_finals: null buffer");
    CU(cudaMemcpy(finals, h->d_finals, (size_t)h->nLanes * FIN_SIZE * sizeof(double), cudaMemcpyDeviceToHost));
    return MLB_OK;
}
extern "C" int mlb_stage_separations(mlb_handle* h, double* out) {
    int rc = needRun(h, "mlb_stage_separations"); if (rc) return rc;
    if (!out) return fail(MLB_ERR_ARG, "mlb_stage_separations: null buffer");
    CU(cudaMemcpy(out, h->d_separations, (size_t)h->nLanes * (MLB_MAX_STAGES - 1) * SEP_W * sizeof(double), cudaMemcpyDeviceToHost));
    return MLB_OK;
}
extern "C" int mlb_lane_counters(mlb_handle* h, int64_t* out) {
    int rc = needRun(h, "mlb_lane_counters"); if (rc) return rc;
    if (!out) return fail(MLB_ERR_ARG, "mlb_lane_counters: null buffer");
    CU(cudaMemcpy(out, h->d_counters, (size_t)h->nLanes * MLB_NCOUNTERS * sizeof(long long), cudaMemcpyDeviceToHost));
    return MLB_OK;
}
extern "C" int mlb_counters(mlb_handle* h, mlb_counters_t* out) {
    int rc = needRun(h, "mlb_counters"); if (rc) return rc;
    if (!out) return fail(MLB_ERR_ARG, "mlb_counters: null buffer");
    std::vector<long long> c((size_t)h->nLanes * MLB_NCOUNTERS); std::vector<int> st((size_t)h->nLanes);
    CU(cudaMemcpy(c.data(), h->d_counters, c.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(st.data(), h->d_status, st.size() * sizeof(int), cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof(*out));
    out->lanes = h->nLanes;
    for (long long i = 0; i < h->nLanes; i++) {
        const long long* ci = c.data() + MLB_NCOUNTERS * i;
        out->steps += ci[0]; out->derivative_evals += ci[1]; out->rejected_steps += ci[2];
        out->steps_3dof += ci[3]; out->derivative_evals_3dof += ci[4]; out->derivative_evals_transonic += ci[5];
        if (st[i] == ST_OK) out->lanes_ok++; else if (st[i] == ST_STEP_LIMIT) out->lanes_step_limit++; else if (st[i] == ST_NAN) out->lanes_nan++;
    }
    float ms = 0, msFlight = 0, msChute = 0;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    CU(cudaEventElapsedTime(&msFlight, h->evFlight0, h->evFlight1));
    CU(cudaEventElapsedTime(&msChute, h->evFlight1, h->ev1));
    out->last_run_ms = ms; out->dominant_kernel_ms = msFlight; out->kernel_launches = h->launches;
    out->chute_kernel_ms = msChute; out->flight_block = h->lastBlock; out->flight_generations = h->lastGenerations;
    return MLB_OK;
}
extern "C" int mlb_device_results(mlb_handle* h, void** results, void** status) {
    if (!h || h->nLanes <= 0) return fail(MLB_ERR_STATE, "mlb_device_results: no lanes");
    if (results) *results = h->d_results;
    if (status) *status = h->d_status;
    return MLB_OK;
}
extern "C" int mlb_set_result_buffers(mlb_handle* h, void* results, void* status) {
    if (!h || h->nLanes <= 0) return fail(MLB_ERR_STATE, "mlb_set_result_buffers: call mlb_set_lanes first");
    if (!results || !status) return fail(MLB_ERR_ARG, "mlb_set_result_buffers: null buffer");
    CU(cudaSetDevice(h->device));
    h->d_results = (double*)results; h->d_status = (int*)status; h->ownResults = false; h->ran = false;
    return MLB_OK;
}
extern "C" int mlb_moments(mlb_handle* h, double* out15) {
    int rc = needRun(h, "mlb_moments"); if (rc) return rc;
    if (!out15) return fail(MLB_ERR_ARG, "mlb_moments: null buffer");
    int grid = 148;
    momentsPartial<<<grid, MLB_BLOCK, 0, h->stream>>>(h->d_results, h->d_status, h->nLanes, h->d_partials);
    momentsFinal<<<1, 32, 0, h->stream>>>(h->d_partials, grid, h->d_moments);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(out15, h->d_moments, 15 * sizeof(double), cudaMemcpyDeviceToHost));
    return MLB_OK;
}

extern "C" int mlb_set_flight_log(mlb_handle* h, int64_t firstLane, int64_t nLanes, int64_t every, int64_t maxRowsPerLane) {
    if (!h) return fail(MLB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    cudaFree(h->d_trace); h->d_trace = nullptr; cudaFree(h->d_traceRows); h->d_traceRows = nullptr;
    h->traceLane = -1; h->traceCount = 0; h->traceMaxRows = 0; h->traceEvery = 1;
    if (firstLane >= 0 && nLanes > 0 && maxRowsPerLane > 0) {
        if (every < 0) return fail(MLB_ERR_ARG, "mlb_set_flight_log: every must be >= 1 (0: every step, stop when the buffer is full)");
        CU(cudaMalloc(&h->d_trace, (size_t)nLanes * maxRowsPerLane * TRACE_W * sizeof(double)));
        CU(cudaMalloc(&h->d_traceRows, (size_t)nLanes * sizeof(long long)));
        CU(cudaMemset(h->d_traceRows, 0, (size_t)nLanes * sizeof(long long)));
        h->traceLane = firstLane; h->traceCount = nLanes; h->traceMaxRows = maxRowsPerLane; h->traceEvery = every;
    }
    return MLB_OK;
}
extern "C" int mlb_get_flight_log(mlb_handle* h, double* rows, int64_t* rowsPerLane) {
    int rc = needRun(h, "mlb_get_flight_log"); if (rc) return rc;
    if (!h->d_trace || h->traceCount <= 0) return fail(MLB_ERR_STATE, "mlb_get_flight_log: no flight log was requested before the run");
    static_assert(sizeof(long long) == sizeof(int64_t), "row counters are copied as int64");
    if (rowsPerLane) CU(cudaMemcpy(rowsPerLane, h->d_traceRows, (size_t)h->traceCount * sizeof(long long), cudaMemcpyDeviceToHost));
    if (rows) CU(cudaMemcpy(rows, h->d_trace, (size_t)h->traceCount * h->traceMaxRows * TRACE_W * sizeof(double), cudaMemcpyDeviceToHost));
    return MLB_OK;
}
extern "C" int mlb_set_trace(mlb_handle* h, int64_t lane, int64_t maxRows) { return mlb_set_flight_log(h, lane, 1, 0, maxRows); }
extern "C" int mlb_get_trace(mlb_handle* h, double* rows, int64_t* nRows) {
    int rc = needRun(h, "mlb_get_trace"); if (rc) return rc;
    long long n = 0;
    if (h->d_traceRows) CU(cudaMemcpy(&n, h->d_traceRows, sizeof(long long), cudaMemcpyDeviceToHost));
    if (nRows) *nRows = n;
    if (rows && n > 0 && h->d_trace) CU(cudaMemcpy(rows, h->d_trace, (size_t)n * TRACE_W * sizeof(double), cudaMemcpyDeviceToHost));
    return MLB_OK;
}
extern "C" int mlb_set_dt_replay(mlb_handle* h, const double* dts, int64_t n) {
    if (!h) return fail(MLB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    cudaFree(h->d_replay); h->d_replay = nullptr; h->nReplay = 0;
    if (dts && n > 0) {
        CU(cudaMalloc(&h->d_replay, (size_t)n * sizeof(double)));
        CU(cudaMemcpy(h->d_replay, dts, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
        h->nReplay = n;
    }
    return MLB_OK;
}

extern "C" int mlb_force_eval(mlb_handle* h, int64_t n, const double* states, const double* times, const double* winds, double* out) {
    if (!h || n <= 0 || !states || !times || !out) return fail(MLB_ERR_ARG, "mlb_force_eval: bad arguments");
    CU(cudaSetDevice(h->device));
    double *d_s = nullptr, *d_t = nullptr, *d_w = nullptr, *d_o = nullptr; uint32_t* d_mt = nullptr;
    int turb = (int)h->blob[H_TURB];
    if (turb >= TURB_PINK1D && turb <= TURB_PINK3D) CU(cudaMalloc(&d_mt, (size_t)n * 3 * 624 * sizeof(uint32_t)));
    CU(cudaMalloc(&d_s, (size_t)n * 13 * sizeof(double))); CU(cudaMalloc(&d_t, (size_t)n * sizeof(double))); CU(cudaMalloc(&d_o, (size_t)n * 18 * sizeof(double)));
    CU(cudaMemcpy(d_s, states, (size_t)n * 13 * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_t, times, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    if (winds) { CU(cudaMalloc(&d_w, (size_t)n * 3 * sizeof(double))); CU(cudaMemcpy(d_w, winds, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice)); }
    forceEvalKernel<<<(unsigned)((n + MLB_BLOCK - 1) / MLB_BLOCK), MLB_BLOCK, h->smemBytes>>>(h->d_blob, (int)h->sharedLen, n, d_s, d_t, d_w, d_mt, d_o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out, d_o, (size_t)n * 18 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_s); cudaFree(d_t); cudaFree(d_w); cudaFree(d_o); cudaFree(d_mt);
    if (e != cudaSuccess) return fail(MLB_ERR_CUDA, std::string("mlb_force_eval: ") + cudaGetErrorString(e));
    return MLB_OK;
}

extern "C" int mlb_host_gauss_stream(const uint32_t* key, int keyLen, int64_t skip, const double* mu, const double* sigma, int nPer, int64_t nDraws, double* out) {
    if (!key || keyLen <= 0 || !mu || !sigma || nPer <= 0 || nDraws < 0 || !out) return fail(MLB_ERR_ARG, "mlb_host_gauss_stream: bad arguments");
    std::vector<uint32_t> mt(624);
    PyRandom r; r.mt = mt.data();
    mtInitByArray(r.mt, key, keyLen);
    r.idx = 624; r.hasNext = 0; r.next = 0;
    for (int64_t i = 0; i < skip; i++) pyRandomGauss(r, 0, 1);
    for (int64_t i = 0; i < nDraws; i++) out[i] = pyRandomGauss(r, mu[i % nPer], sigma[i % nPer]);
    return MLB_OK;
}

extern "C" int mlb_host_randrange_stream(const uint32_t* key, int keyLen, int64_t skip, uint32_t n, int64_t nDraws, double* out) {
    /* random.randrange(n) for 0 < n < 2^32: CPython's _randbelow_with_getrandbits (Lib/random.py): k = n.bit_length();
       r = getrandbits(k) until r < n; getrandbits(k <= 32) = genrand_uint32() >> (32 - k) */
    if (!key || keyLen <= 0 || n == 0 || nDraws < 0 || !out) return fail(MLB_ERR_ARG, "mlb_host_randrange_stream: bad arguments");
    std::vector<uint32_t> mt(624);
    PyRandom r; r.mt = mt.data();
    mtInitByArray(r.mt, key, keyLen);
    r.idx = 624; r.hasNext = 0; r.next = 0;
    int k = 0; for (uint32_t v = n; v; v >>= 1) k++;
    for (int64_t i = 0; i < skip + nDraws; i++) {
        uint32_t v = mtGenrand(r) >> (32 - k);
        while (v >= n) v = mtGenrand(r) >> (32 - k);
        if (i >= skip) out[i - skip] = (double)v;
    }
    return MLB_OK;
}

/* The same two streams continued from a live CPython generator: `state` = the 624 words of random.Random.getstate()[1][:624], *pos its
   last element, hasNext / next the cached second Gaussian (getstate()[2]). All four are updated in place, so the caller hands the
   advanced state back with setstate() and the Python object carries on as if it had made the draws itself. */
extern "C" int mlb_host_gauss_from_state(uint32_t* state, int* pos, int* hasNext, double* next, const double* mu, const double* sigma, int nPer, int64_t nDraws, double* out) {
    if (!state || !pos || !hasNext || !next || !mu || !sigma || nPer <= 0 || nDraws < 0 || !out) return fail(MLB_ERR_ARG, "mlb_host_gauss_from_state: bad arguments");
    PyRandom r; r.mt = state; r.idx = *pos; r.hasNext = *hasNext; r.next = *next;
    for (int64_t i = 0; i < nDraws; i++) out[i] = pyRandomGauss(r, mu[i % nPer], sigma[i % nPer]);
    *pos = r.idx; *hasNext = r.hasNext; *next = r.next;
    return MLB_OK;
}
extern "C" int mlb_host_randrange_from_state(uint32_t* state, int* pos, uint32_t n, int64_t nDraws, double* out) {
    if (!state || !pos || n == 0 || nDraws < 0 || !out) return fail(MLB_ERR_ARG, "mlb_host_randrange_from_state: bad arguments");
    PyRandom r; r.mt = state; r.idx = *pos; r.hasNext = 0; r.next = 0;
    int k = 0; for (uint32_t v = n; v; v >>= 1) k++;
    for (int64_t i = 0; i < nDraws; i++) {
        uint32_t v = mtGenrand(r) >> (32 - k);
        while (v >= n) v = mtGenrand(r) >> (32 - k);
        out[i] = (double)v;
    }
    *pos = r.idx;
    return MLB_OK;
}

/* FP64 pipe ceiling of this GPU, measured: 16 independent DFMA chains per thread, every SM full. The roofline
   denominator of the trajectory kernels (MEASURED_PEAKS.json carries HBM and bf16 tensor figures only). */
__global__ void __launch_bounds__(256) fp64PeakKernel(double* out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = threadIdx.x * 1e-3 + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = fma(x[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s += x[k];
    if (s == 12345.678) out[0] = s;
}
extern "C" int mlb_fp64_peak(int device, double* tflops) {
    if (!tflops) return fail(MLB_ERR_ARG, "mlb_fp64_peak: null output");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, device));
    double* d_out = nullptr; CU(cudaMalloc(&d_out, sizeof(double)));
    cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, iters = 20000;
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CU(cudaEventRecord(e0, 0));
        fp64PeakKernel<<<blocks, 256>>>(d_out, iters, 0.999999, 1e-7);
        CU(cudaEventRecord(e1, 0));
        CU(cudaEventSynchronize(e1));
        float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
        double tf = 2.0 * 16 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    *tflops = best;
    return MLB_OK;
}
