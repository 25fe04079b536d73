"""Exact algorithmic FLOP counts of the hot path, from the op-counting build of the CPU oracle (oracle/opcount.hpp):
every FP64 add/sub/mul/div/sqrt/compare = 1, a transcendental call = 1, over N sampled lanes of a definition. Prints the per-unit table
bench.py's FLOP_* constants are taken from. CPU only:  python tools/count_flops.py [lanes] [definition]"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
import numpy as np
import bench
from mapleaf_b200 import L, SimDefinition, buildModel
from oracle import oracle

lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 32
definition = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(bench.DEFINITION)
sd = SimDefinition(os.path.join(REPO, "mapleaf_b200", "data", "Simulations", definition), silent=True, disableDistributionSampling=True)
m = buildModel(sd)
arrays = {}
if int(m.header("H_WIND")) in (L.WIND_CONSTANT, L.WIND_HELLMAN):
    wind, seeds = bench.sampleInputs(0, lanes)
    arrays[L.LANE_WIND] = wind
    if int(m.header("H_TURB")) in (L.TURB_PINK1D, L.TURB_PINK2D, L.TURB_PINK3D):
        arrays[L.LANE_PINK_SEED] = seeds
out, ops = oracle.runCounted(m.blob, lanes, arrays)
steps = int(out["counters"][:, 0].sum()); evals = int(out["counters"][:, 1].sum()); rejects = int(out["counters"][:, 2].sum())
b = ops["buckets"]
table = {k: (v["flops"] / v["count"] if v["count"] else 0.0) for k, v in b.items()}
print(json.dumps(dict(definition=definition, lanes=lanes, steps=steps, derivative_evals=evals, rejected=rejects,
                      ops={k: ops[k] for k in ("add", "mul", "div", "sqrt", "compare", "transcendental", "uncounted")},
                      flops=ops["flops"], flops_per_step=ops["flops"] / steps,
                      counts={k: v["count"] for k, v in b.items()}, flops_per_unit=table,
                      flops_outside_steps=ops["flops"] - sum(v["flops"] for v in b.values())), indent=1))
