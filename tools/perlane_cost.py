"""Cost of the per-lane model-block path (mlb_set_lane_models) against the shared-memory model block, same lanes, same results:
    python tools/perlane_cost.py [lanes] [config]"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
import numpy as np
import bench
from mapleaf_b200 import backend
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 151552
cfg = bench.CONFIGS[int(sys.argv[2]) if len(sys.argv) > 2 else 1]
m = bench.buildConfigModel(cfg)
arrays = bench.laneInputs(cfg, m, 0, lanes)
shared = int(m.header("H_SHARED_LEN")) or m.blob.size
with backend.TrajectoryBatch(m.blob) as b:
    b.setLanes(lanes, arrays)
    b.run(); b.sync(); b.run(); b.sync()
    c0 = b.counters(); r0, _ = b.results()
    b.setLaneModels(np.tile(m.blob[:shared], (lanes, 1)))
    b.run(); b.sync(); b.run(); b.sync()
    c1 = b.counters(); r1, _ = b.results()
print(json.dumps({"lanes": lanes, "shared_ms": c0["last_run_ms"], "perlane_ms": c1["last_run_ms"], "ratio": c1["last_run_ms"] / c0["last_run_ms"],
                  "shared_flight_block": c0["flight_block"], "perlane_flight_block": c1["flight_block"], "model_bytes_per_lane": shared * 8,
                  "max_rel_diff_apogee": float(np.max(np.abs(r1[:, 3] - r0[:, 3]) / np.abs(r0[:, 3])))}))
