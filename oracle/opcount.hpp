/* TEST / MEASUREMENT INFRASTRUCTURE ONLY (like everything under oracle/).
 *
 * Op-counting scalar for the CPU oracle: oracle/opcount_build.cpp compiles mapleaf_oracle.c as C++ with `double` replaced by
 * this type, so every FP64 operation the restatement of the reference's arithmetic performs is counted exactly. The counts
 * are the ALGORITHMIC work per unit of SURVEY.md §8(d) (add/sub/mul/div/sqrt/compare-select = 1 FLOP each, a transcendental
 * call = 1, negation / fabs / copies = 0), bucketed by what the kernel's own counters can see: 6-DoF derivative evaluation
 * (outside / inside the transonic fin blend), 3-DoF evaluation, and the rest of a time step (integrator algebra, events).
 * tools/count_flops.py runs it and prints the table bench.py's FLOP_* constants come from. */
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

struct OpCounters {
    uint64_t add, mul, div, sqrt_, cmp, trans, other;        /* other: fabs, negation, rounding (not counted as FLOPs) */
    uint64_t flops() const { return add + mul + div + sqrt_ + cmp + trans; }
};
extern OpCounters g_ops;

struct CD {
    double v;
    CD() = default;
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> CD(T x) : v((double)x) {}
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type> explicit operator T() const { return (T)v; }
    CD& operator+=(CD o) { g_ops.add++; v += o.v; return *this; }
    CD& operator-=(CD o) { g_ops.add++; v -= o.v; return *this; }
    CD& operator*=(CD o) { g_ops.mul++; v *= o.v; return *this; }
    CD& operator/=(CD o) { g_ops.div++; v /= o.v; return *this; }
};
static_assert(sizeof(CD) == sizeof(double), "CD must alias double arrays at the C ABI");
inline CD operator+(CD a, CD b) { g_ops.add++; return CD(a.v + b.v); }
inline CD operator-(CD a, CD b) { g_ops.add++; return CD(a.v - b.v); }
inline CD operator*(CD a, CD b) { g_ops.mul++; return CD(a.v * b.v); }
inline CD operator/(CD a, CD b) { g_ops.div++; return CD(a.v / b.v); }
inline CD operator-(CD a) { g_ops.other++; return CD(-a.v); }
inline CD operator+(CD a) { return a; }
inline bool operator<(CD a, CD b) { g_ops.cmp++; return a.v < b.v; }
inline bool operator>(CD a, CD b) { g_ops.cmp++; return a.v > b.v; }
inline bool operator<=(CD a, CD b) { g_ops.cmp++; return a.v <= b.v; }
inline bool operator>=(CD a, CD b) { g_ops.cmp++; return a.v >= b.v; }
inline bool operator==(CD a, CD b) { g_ops.cmp++; return a.v == b.v; }
inline bool operator!=(CD a, CD b) { g_ops.cmp++; return a.v != b.v; }
inline bool operator!(CD a) { return a.v == 0; }
#define OPC_TRANS1(f) inline CD f(CD a) { g_ops.trans++; return CD(::f(a.v)); }
OPC_TRANS1(sin) OPC_TRANS1(cos) OPC_TRANS1(tan) OPC_TRANS1(asin) OPC_TRANS1(acos) OPC_TRANS1(atan) OPC_TRANS1(exp) OPC_TRANS1(log)
inline CD atan2(CD a, CD b) { g_ops.trans++; return CD(::atan2(a.v, b.v)); }
inline CD pow(CD a, CD b) { g_ops.trans++; return CD(::pow(a.v, b.v)); }
inline CD sqrt(CD a) { g_ops.sqrt_++; return CD(::sqrt(a.v)); }
inline CD fabs(CD a) { g_ops.other++; return CD(::fabs(a.v)); }
inline CD fmax(CD a, CD b) { g_ops.cmp++; return CD(::fmax(a.v, b.v)); }
inline CD fmin(CD a, CD b) { g_ops.cmp++; return CD(::fmin(a.v, b.v)); }
inline CD fmod(CD a, CD b) { g_ops.div++; return CD(::fmod(a.v, b.v)); }
inline CD round(CD a) { g_ops.other++; return CD(::round(a.v)); }
inline CD nearbyint(CD a) { g_ops.other++; return CD(::nearbyint(a.v)); }
inline CD floor(CD a) { g_ops.other++; return CD(::floor(a.v)); }
#undef isnan
inline bool isnan(CD a) { return a.v != a.v; }

/* buckets: 0 = 6-DoF evaluation, 1 = 6-DoF evaluation inside 0.8 < Mach < 1.4, 2 = 3-DoF evaluation,
   3 = rest of a 6-DoF time step, 4 = rest of a 3-DoF time step; [k][0] = FLOPs, [k][1] = how many */
extern uint64_t g_bucket[5][2];
extern uint64_t g_evalStart, g_stepStart, g_stepEvalFlops;
#define OPC_EVAL_BEGIN() (g_evalStart = g_ops.flops())
#define OPC_EVAL_END(kind) do { uint64_t d_ = g_ops.flops() - g_evalStart; g_bucket[kind][0] += d_; g_bucket[kind][1]++; g_stepEvalFlops += d_; } while (0)
#define OPC_STEP_BEGIN() do { g_stepStart = g_ops.flops(); g_stepEvalFlops = 0; } while (0)
#define OPC_STEP_END(kind, attempts) do { g_bucket[3 + (kind)][0] += g_ops.flops() - g_stepStart - g_stepEvalFlops; g_bucket[3 + (kind)][1] += (attempts); } while (0)
