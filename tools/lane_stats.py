"""How much of the CTA lock-step time is idle lanes? Uses per-lane counters of one pass (run on the GPU box)."""
import os, sys, json
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
import bench
from mapleaf_b200 import L, SimDefinition, backend, buildModel
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 151552
m = bench.buildConfigModel(bench.CONFIGS[1])
wind, seeds = bench.sampleInputs(0, lanes)
with backend.TrajectoryBatch(m.blob) as b:
    b.setLanes(lanes, {L.LANE_WIND: wind, L.LANE_PINK_SEED: seeds}); b.run(); c = b.laneCounters(); tot = b.counters()
n = c[:, 0]; n3 = c[:, 3]; n6 = n - n3
print("lanes", lanes, "ms", tot["last_run_ms"])
print("steps/lane mean %.0f min %d max %d p99 %.0f | 6dof mean %.0f max %d | 3dof mean %.0f max %d" % (n.mean(), n.min(), n.max(), np.percentile(n, 99), n6.mean(), n6.max(), n3.mean(), n3.max()))
c6, c3 = 1.0, 0.1          # relative cost of a 6-DoF and a 3-DoF step
ideal = (n6 * c6 + n3 * c3).sum()
for blk in (32, 128, 512, 1024):
    k = lanes // blk * blk
    a6 = n6[:k].reshape(-1, blk); a3 = n3[:k].reshape(-1, blk); at = n[:k].reshape(-1, blk)
    # lock-step cost of a group: while any lane is 6-DoF an iteration costs c6 (+c3 if any lane is 3-DoF); afterwards c3
    m6 = a6.max(axis=1); mt = at.max(axis=1); first3 = a6.min(axis=1)
    cost = (m6 * c6 + (m6 - first3) * c3 + (mt - m6) * c3) * blk
    print("group %5d: lock-step efficiency vs phase-sorted ideal %.3f   (max total steps / mean %.3f)" % (blk, (n6[:k] * c6 + n3[:k] * c3).sum() / cost.sum(), mt.mean() / n.mean()))
print("6dof steps percentiles 1/10/50/90/99/max:", [int(np.percentile(n6, p)) for p in (1, 10, 50, 90, 99, 100)])
print("3dof steps percentiles 1/10/50/90/99/max:", [int(np.percentile(n3, p)) for p in (1, 10, 50, 90, 99, 100)])
w6 = n6[:lanes // 32 * 32].reshape(-1, 32)
print("per-warp: mean(max n6)/mean(n6) = %.3f, mean(min n6)/mean(n6) = %.3f" % (w6.max(axis=1).mean() / n6.mean(), w6.min(axis=1).mean() / n6.mean()))
np.save(os.path.join(REPO, "gpurun_out", "lane_counters.npy"), c)
