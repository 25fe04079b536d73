/*
 * mlb_rng.cuh — bit-exact CPython `random.Random` stream (MT19937 + random() + gauss()) and the
 * reference's pink-noise generator, usable from host and device code.
 *
 * Why bit-exact: the reference draws its Monte Carlo samples and its turbulence from CPython's
 * generator (MAPLEAF/IO/simDefinition.py:205-216,471-532; MAPLEAF/ENV/TurbulenceModelling.py:168-247),
 * and results are compared sample for sample.
 * Algorithms: Matsumoto & Nishimura MT19937 (init_genrand / init_by_array / genrand_int32),
 * CPython's random() = (a>>5, b>>6) 53-bit construction and gauss() = Box-Muller with a cached second
 * variate; Kasdin's 2-pole AR pink-noise filter on a 20 Hz grid with linear inter/extrapolation.
 *
 * The 624-word state lives wherever the caller puts it (one contiguous 2496-byte block per generator):
 * on the device that is HBM, one block per lane and generator, so that a lane's word-at-a-time twist
 * walks its own cache lines.
 */
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define MLB_HD __host__ __device__ __forceinline__
/* cold, loop-heavy routines: one out-of-line copy keeps the trajectory kernel's instruction footprint small */
#define MLB_HD_COLD __host__ __device__ __noinline__
#else
#define MLB_HD inline
#define MLB_HD_COLD inline
#endif

namespace mlb {

struct PyRandom {
    uint32_t* mt;      /* 624 words */
    int idx;           /* next word to temper; 624 = state exhausted */
    int hasNext;       /* gauss(): cached second variate */
    double next;
};

MLB_HD void mtInitGenrand(uint32_t* mt, uint32_t s) {
    mt[0] = s;
#pragma unroll 1
    for (int i = 1; i < 624; i++) mt[i] = 1812433253U * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
}
MLB_HD_COLD static void mtInitByArray(uint32_t* mt, const uint32_t* key, int keyLen) {
    mtInitGenrand(mt, 19650218U);
    int i = 1, j = 0;
    int k = (624 > keyLen) ? 624 : keyLen;
#pragma unroll 1
    for (; k; k--) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525U)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
        if (j >= keyLen) j = 0;
    }
#pragma unroll 1
    for (k = 623; k; k--) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941U)) - (uint32_t)i;
        i++;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000U;
}
/* random.Random(int seed): key = 32-bit little-endian chunks of |seed| */
MLB_HD void pyRandomSeedInt(PyRandom& r, uint64_t seed) {
    uint32_t key[2];
    key[0] = (uint32_t)(seed & 0xffffffffU);
    key[1] = (uint32_t)(seed >> 32);
    mtInitByArray(r.mt, key, key[1] ? 2 : 1);
    r.idx = 624; r.hasNext = 0; r.next = 0;
}
MLB_HD_COLD static void mtTwist(uint32_t* mt) {
    int kk; uint32_t y;
#pragma unroll 4
    for (kk = 0; kk < 624 - 397; kk++) { y = (mt[kk] & 0x80000000U) | (mt[kk + 1] & 0x7fffffffU); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U); }
#pragma unroll 4
    for (; kk < 623; kk++) { y = (mt[kk] & 0x80000000U) | (mt[kk + 1] & 0x7fffffffU); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U); }
    y = (mt[623] & 0x80000000U) | (mt[0] & 0x7fffffffU); mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
}
/* Device code regenerates the state ONE WORD PER DRAW, in place, instead of all 624 words when the block runs out: word i of the next
   block is a function of mt[i], mt[i + 1] and mt[i + 397] (indices mod 624), and walking i upwards in place reads exactly the old / new
   words the block twist reads (genrand_int32 is that same loop), so the stream is bit-identical. Why: a lane's block twist is ~9 k
   instructions executed by ONE thread of its warp while the CTA's other warps wait at the convoy barrier (ncu: the twist ran at 1.03
   threads per instruction, profiles/r3b). In this representation words at and above idx are still the previous block's; idx 624 and
   idx 0 are the same position. Host code (the Monte Carlo samplers, which take over CPython generator states) keeps the block twist. */
#ifndef MLB_MT_INCREMENTAL
#define MLB_MT_INCREMENTAL 1
#endif
MLB_HD uint32_t mtGenrand(PyRandom& r) {
#if defined(__CUDA_ARCH__) && MLB_MT_INCREMENTAL
    uint32_t* mt = r.mt;
    const int i = r.idx >= 624 ? 0 : r.idx;
    const int i1 = (i == 623) ? 0 : i + 1, im = (i < 624 - 397) ? i + 397 : i + (397 - 624);
    uint32_t y = (mt[i] & 0x80000000U) | (mt[i1] & 0x7fffffffU);
    y = mt[im] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
    mt[i] = y;
    r.idx = i + 1;
#else
    if (r.idx >= 624) { mtTwist(r.mt); r.idx = 0; }
    uint32_t y = r.mt[r.idx++];
#endif
    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680U; y ^= (y << 15) & 0xefc60000U; y ^= (y >> 18);
    return y;
}
MLB_HD double pyRandomRandom(PyRandom& r) {
    uint32_t a = mtGenrand(r) >> 5, b = mtGenrand(r) >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}
MLB_HD double pyRandomGauss(PyRandom& r, double mu, double sigma) {
    double z;
    if (r.hasNext) { z = r.next; r.hasNext = 0; }
    else {
        double x2pi = pyRandomRandom(r) * (2.0 * 3.141592653589793);
        double g2rad = sqrt(-2.0 * log(1.0 - pyRandomRandom(r)));
        z = cos(x2pi) * g2rad;
        r.next = sin(x2pi) * g2rad; r.hasNext = 1;
    }
    return mu + z * sigma;
}

/* PinkNoiseGenerator (TurbulenceModelling.py:168-247), nPoles = 2, 20 Hz */
struct PinkNoise {
    PyRandom rng;
    double last0, last1;      /* ring of the two most recent filtered samples */
    int nextSlot;             /* slot the next sample is written to */
    double lastValTime;
};

MLB_HD double pinkNewSample(PinkNoise& g, double c0, double c1) {
    /* coefficient i multiplies the sample written (i+1) calls ago */
    double newest = g.nextSlot ? g.last0 : g.last1;
    double older = g.nextSlot ? g.last1 : g.last0;
    double white = pyRandomGauss(g.rng, 0, 1);
    white -= c0 * newest;
    white -= c1 * older;
    if (g.nextSlot) g.last1 = white; else g.last0 = white;
    g.nextSlot ^= 1;
    return white;
}
MLB_HD_COLD static void pinkInit(PinkNoise& g, uint64_t seed, double c0, double c1) {
    pyRandomSeedInt(g.rng, seed);
    g.nextSlot = 0; g.lastValTime = 0;
    g.last0 = pyRandomGauss(g.rng, 0, 1);
    g.last1 = pyRandomGauss(g.rng, 0, 1);
    for (int i = 0; i < 20; i++) pinkNewSample(g, c0, c1);
}
/* inline: two calls per force evaluation on the turbulence path; out of line it cost 1-2 % in call overhead (267 -> 271 M steps/s) */
MLB_HD double pinkValue(PinkNoise& g, double time, double c0, double c1) {
    const double timeStep = 1.0 / 20;
    while (time > g.lastValTime) {
        double s = pinkNewSample(g, c0, c1);
        g.lastValTime += timeStep;
        if (g.lastValTime == time) return s;
    }
    double lastVal = g.nextSlot ? g.last0 : g.last1;
    double secondLastVal = g.nextSlot ? g.last1 : g.last0;
    double secondLastValTime = g.lastValTime - timeStep;
    /* (lastVal - secondLastVal) / timeStep with the exactly rounded reciprocal 1 / 0.05 = 20 and one correction step (Markstein): the
       correctly rounded quotient, without the full division sequence */
    const double diff = lastVal - secondLastVal, q = diff * 20.0;
    double dValdt = fma(fma(-timeStep, q, diff), 20.0, q);
    return dValdt * (time - secondLastValTime) + secondLastVal;
}

}  // namespace mlb
