/*
 * mlb_math.cuh — FP64 value types of the batched trajectory kernels (device side).
 *
 * 3-vector, quaternion and rigid-body-state algebra with the operation order of the reference's
 * native value types, so that fixed-step trajectories agree with the reference to ~1e-13:
 *   MAPLEAF/Motion/CythonVector.pyx:13-139, CythonQuaternion.pyx:13-226, CythonAngularVelocity.pyx:78-86
 *   MAPLEAF/Motion/RigidBodyStates.py:12-88 (3-DoF), :110-211 (6-DoF state / derivative algebra)
 * Quirks that are observable in results are kept: in-place normalize that
 * skips the divide when the norm is exactly 1, rotationAxis() dividing by sqrt(norm - q0^2),
 * rotationAngle() = 2 atan2(sqrt(1 - q0^2), q0).
 */
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define MLB_DEV __device__ __forceinline__
/* routines off the steady-state path (variable inertia while a motor burns, geodetic solve, control tick, N-D table look-up): one
   out-of-line copy each instead of several inlined ones, so the per-evaluation instruction stream the convoy shares stays compact.
   MLB_COLD_MASK selects them: 1 rocket inertia, 2 geodetic solve, 4 N-D look-up, 8 control system. Measured on the bench workload:
   mask 15 shrinks the kernel from 48.9 k to 36.0 k SASS instructions at unchanged speed (264-267 M steps/s for masks 0, 1, 15); the
   default keeps out of line only what the bench workload proved neutral AND is cold in every configuration (the inertia). */
#ifndef MLB_COLD_MASK
#define MLB_COLD_MASK 1
#endif
#if MLB_COLD_MASK & 1
#define MLB_COLD_INERTIA static __device__ __noinline__
#else
#define MLB_COLD_INERTIA MLB_DEV
#endif
#if MLB_COLD_MASK & 2
#define MLB_COLD_GEODETIC static __device__ __noinline__
#else
#define MLB_COLD_GEODETIC MLB_DEV
#endif
#if MLB_COLD_MASK & 4
#define MLB_COLD_NDTABLE static __device__ __noinline__
#else
#define MLB_COLD_NDTABLE MLB_DEV
#endif
#if MLB_COLD_MASK & 8
#define MLB_COLD_CONTROL static __device__ __noinline__
#else
#define MLB_COLD_CONTROL MLB_DEV
#endif
#define MLB_PI 3.141592653589793

#ifndef MLB_MATH_NOINLINE
#define MLB_MATH_NOINLINE 0
#endif
#if MLB_MATH_NOINLINE
#define MLB_MATHFN static __device__ __noinline__      /* one copy of each libm routine: smaller instruction footprint */
#else
#define MLB_MATHFN static __device__ __forceinline__
#endif

namespace mlb {

MLB_MATHFN double m_sin(double x) { return sin(x); }
MLB_MATHFN double m_cos(double x) { return cos(x); }
MLB_MATHFN void m_sincos(double x, double* s, double* c) { sincos(x, s, c); }
MLB_MATHFN double m_tan(double x) { return tan(x); }
MLB_MATHFN double m_asin(double x) { return asin(x); }
MLB_MATHFN double m_acos(double x) { return acos(x); }
MLB_MATHFN double m_atan2(double y, double x) { return atan2(y, x); }
MLB_MATHFN double m_exp(double x) { return exp(x); }
MLB_MATHFN double m_log(double x) { return log(x); }
/* x^y of the power laws (atmosphere layers, Hellman wind profile, skin-friction and cross-flow correlations): positive bases, |y ln x| of
   order 1-10. MLB_POW_EXPLOG evaluates exp(y log x) instead of the library's pow, whose extended-precision logarithm is a third of the
   under-chute kernel's instruction stream (profiles/r2_chute_by_enclosing_function.txt); the result differs from pow's by |y ln x| ulp.
   Measured +1.7 % (configs[1]) / +2.6 % (configs[4]) in a same-box A/B (profiles/r2n_pow_explog_ab.txt); fixed-step and lock-step gates stay
   green, but two free-running adaptive gates (the NASA case's 0.2 % regression gate, the elementary-controller flight) sit on the edge of
   their chaotic bands and tip over with the changed rounding: off. */
#ifndef MLB_POW_EXPLOG
#define MLB_POW_EXPLOG 0
#endif
MLB_MATHFN double m_pow(double x, double y) {
#if MLB_POW_EXPLOG
    if (x > 0 && x < 1.7e308) return exp(y * log(x));
#endif
    return pow(x, y);
}
MLB_MATHFN double m_cbrt(double x) { return cbrt(x); }

struct V3 { double x, y, z; };
struct Q4 { double w, x, y, z; };

MLB_DEV V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
MLB_DEV V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
MLB_DEV V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
MLB_DEV V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }          /* a + (-b): identical in IEEE */
MLB_DEV V3 operator*(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
MLB_DEV V3 vdiv(V3 a, double s) { return a * (1 / s); }                                    /* CythonVector.pyx:105 */
MLB_DEV double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
MLB_DEV double len(V3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
/* x / n for several x sharing one divisor: one IEEE reciprocal, then Markstein's correction step per quotient
   (q = x r; q' = q + r (x - n q), both fused). With r the correctly rounded 1/n this returns the correctly rounded x / n
   (Markstein 1990), i.e. the same bits as the division it replaces, at a third of the instructions. */
MLB_DEV double divBy(double x, double n, double r) {
    double q = x * r;
    return fma(fma(-n, q, x), r, q);
}
/* 1 / x for a finite x well inside the normal range (callers guarantee it): the hardware seed and the two refinement steps of the
   IEEE reciprocal sequence, without its range tests and slow-path call. The last step rounds an 80-bit-accurate value once. */
MLB_DEV double rcpNormal(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
MLB_DEV V3 normalize(V3 a) {                                                               /* :118 — zero stays zero */
    double l = len(a);
    if (l > 0) { double r = 1.0 / l; return v3(divBy(a.x, l, r), divBy(a.y, l, r), divBy(a.z, l, r)); }
    return v3(0, 0, 0);
}
MLB_DEV V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
MLB_DEV double angleBetween(V3 a, V3 b) {                                                  /* :132 */
    double interior = dot(a, b) / (len(a) * len(b));
    if (fabs(interior - 1) < 0.00000001) return 0.0;                                       /* acos(1) */
    return m_acos(interior);
}

MLB_DEV Q4 q4(double w, double x, double y, double z) { Q4 r; r.w = w; r.x = x; r.y = y; r.z = z; return r; }
MLB_DEV Q4 qIdentity() { return q4(1, 0, 0, 0); }
MLB_DEV Q4 qAxisAngle(V3 axis, double angle) {                                             /* CythonQuaternion.pyx:50 */
    V3 ax = normalize(axis);
    double s, c;
    m_sincos(angle / 2, &s, &c);
    return q4(c, s * ax.x, s * ax.y, s * ax.z);
}
MLB_DEV Q4 operator*(Q4 a, Q4 b) {                                                         /* :102 */
    return q4(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
              a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
              a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w);
}
MLB_DEV double qNorm(Q4 q) { return sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z); }
MLB_DEV void qNormalize(Q4& q) {                                                           /* :144 */
    double n = qNorm(q);
    if (n == 1.0) return;
    double r = 1.0 / n;
    q.w = divBy(q.w, n, r); q.x = divBy(q.x, n, r); q.y = divBy(q.y, n, r); q.z = divBy(q.z, n, r);
}
MLB_DEV Q4 qConj(Q4 q) { return q4(q.w, -q.x, -q.y, -q.z); }
/* q (0,v) q* for a unit q, written out: the (0,v) factor makes a quarter of the products of two full
   quaternion multiplications vanish, and the scalar part of the result is never used. Term order follows
   the two Hamilton products of CythonQuaternion.pyx:164-169 with the zero terms dropped. */
MLB_DEV V3 qRotateUnit(Q4 q, V3 v) {
    double tw = -q.x * v.x - q.y * v.y - q.z * v.z;
    double tx = q.w * v.x + q.y * v.z - q.z * v.y;
    double ty = q.w * v.y - q.x * v.z + q.z * v.x;
    double tz = q.w * v.z + q.x * v.y - q.y * v.x;
    return v3(-tw * q.x + tx * q.w - ty * q.z + tz * q.y,
              -tw * q.y + tx * q.z + ty * q.w - tz * q.x,
              -tw * q.z - tx * q.y + ty * q.x + tz * q.w);
}
MLB_DEV V3 qRotate(Q4& q, V3 v) { qNormalize(q); return qRotateUnit(q, v); }               /* normalizes the receiver */
MLB_DEV V3 qConjRotate(Q4 q, V3 v) { Q4 c = qConj(q); return qRotate(c, v); }
MLB_DEV double qRotationAngle(Q4& q) {                                                     /* :205 */
    qNormalize(q);
    return 2 * m_atan2(sqrt(1 - q.w * q.w), q.w);
}
MLB_DEV V3 qRotationAxis(Q4& q) {                                                          /* :197 */
    /* `if abs(self.rotationAngle()) > 0`: the angle 2 atan2(sqrt(1 - w^2), w) is only tested against zero here. With t = 1 - w^2 it is
       non-zero exactly when t > 0 (atan2 of a positive sine) or t == 0 and w < 0 (atan2(+0, negative) = pi); t < 0 makes it NaN and the
       test false. Same branch, same in-place normalisation, without the second arctangent and square root of every angle rescaling. */
    qNormalize(q);
    const double t = 1 - q.w * q.w;
    if (t > 0 || (t == 0 && q.w < 0)) {
        double n = sqrt(qNorm(q) - q.w * q.w);
        double r = 1.0 / n;
        return v3(divBy(q.x, n, r), divBy(q.y, n, r), divBy(q.z, n, r));
    }
    return v3(1, 0, 0);
}
MLB_DEV Q4 qScaleRotation(Q4& q, double s) {                                               /* :128 */
    double a = qRotationAngle(q);
    if (a == 0) return q;
    return qAxisAngle(qRotationAxis(q), s * a);
}
MLB_DEV Q4 angVelToQuat(V3 w) { return qAxisAngle(normalize(w), len(w)); }                 /* CythonAngularVelocity.pyx:78 */

/* Rigid-body state and its time derivative. Under a parachute the body is 3-DoF (rocket.py:700-726): only
   pos / vel are integrated; the caller passes `dof3` and orientation / angular velocity stay identity / zero. */
struct State { V3 pos, vel; Q4 q; V3 w; };
struct Deriv { V3 vel, acc, angvel, angacc; };

MLB_DEV State stateAdd(State& self, const State& o, bool dof3) {                          /* RigidBodyStates.py:110-117 / :24-29 */
    State r;
    r.pos = self.pos + o.pos;
    r.vel = self.vel + o.vel;
    if (dof3) { r.q = qIdentity(); r.w = v3(0, 0, 0); return r; }
    r.w = self.w + o.w;
    qNormalize(self.q);
    r.q = o.q * self.q;
    qNormalize(r.q);
    return r;
}
MLB_DEV State stateScale(State& self, double s, bool dof3) {                               /* :119-131 / :31-37 */
    State r;
    r.pos = self.pos * s;
    r.vel = self.vel * s;
    if (dof3) { r.q = qIdentity(); r.w = v3(0, 0, 0); return r; }
    r.w = self.w * s;
    r.q = qScaleRotation(self.q, s);
    return r;
}
MLB_DEV State stateNeg(const State& self, bool dof3) {                                     /* :154-155 / :49-50 */
    State r;
    r.pos = self.pos * -1.0; r.vel = self.vel * -1.0;
    if (dof3) { r.q = qIdentity(); r.w = v3(0, 0, 0); return r; }
    r.q = qConj(self.q); r.w = self.w * -1.0;
    return r;
}
MLB_DEV double stateAbs(State& self, bool dof3) {                                          /* :133-139 / :40-44 */
    double positionMag = len(self.pos) + len(self.vel);
    if (dof3) return positionMag;
    double orientationMag = fabs(qRotationAngle(self.q)) + len(self.w);
    return orientationMag * 100 + positionMag;
}
MLB_DEV State derivTimes(const Deriv& d, double dt, bool dof3) {                           /* :173-188: derivative * dt -> state increment */
    State r;
    r.pos = d.vel * dt;
    r.vel = d.acc * dt;
    if (dof3) { r.q = qIdentity(); r.w = v3(0, 0, 0); return r; }
    r.q = angVelToQuat(d.angvel * dt);
    qNormalize(r.q);
    r.w = d.angacc * dt;
    return r;
}

}  // namespace mlb
