/* TEST / MEASUREMENT INFRASTRUCTURE ONLY. The oracle, compiled as C++ with an op-counting scalar in place of `double`
 * (oracle/opcount.hpp). Same C ABI as libmapleaf_oracle.so plus oracle_opcount_reset / oracle_opcount_get. */
#include "opcount.hpp"
OpCounters g_ops;
uint64_t g_bucket[5][2];
uint64_t g_evalStart, g_stepStart, g_stepEvalFlops;
#define double CD
extern "C" {
#include "mapleaf_oracle.c"
}
#undef double
extern "C" void oracle_opcount_reset(void) { memset(&g_ops, 0, sizeof(g_ops)); memset(g_bucket, 0, sizeof(g_bucket)); }
/* out[0..6] = add, mul, div, sqrt, compare, transcendental, uncounted; out[7 + 2k], out[8 + 2k] = FLOPs and count of bucket k */
extern "C" void oracle_opcount_get(uint64_t* out) {
    out[0] = g_ops.add; out[1] = g_ops.mul; out[2] = g_ops.div; out[3] = g_ops.sqrt_; out[4] = g_ops.cmp; out[5] = g_ops.trans; out[6] = g_ops.other;
    for (int k = 0; k < 5; k++) { out[7 + 2 * k] = g_bucket[k][0]; out[8 + 2 * k] = g_bucket[k][1]; }
}
