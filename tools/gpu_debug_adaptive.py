"""Debug aid: compare the CUDA adaptive path with the CPU oracle step by step (run on the GPU box)."""
import json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mapleaf_b200 import L, SimDefinition, buildModel, backend
from oracle import oracle

DATA = os.path.join(REPO, "mapleaf_b200", "data", "Simulations")
g = json.load(open(os.path.join(REPO, "tests", "golden", sys.argv[1] if len(sys.argv) > 1 else "rk45_pink.json")))
sd = SimDefinition(os.path.join(DATA, g["definition"]), silent=True, disableDistributionSampling=True)
for k, v in (g.get("overrides") or {}).items():
    sd.setValue(k, v)
m = buildModel(sd)
runs = g["runs"]
w = np.array([r["wind"] for r in runs], dtype=float).T.copy()
seeds = np.array([r["pinkSeeds"] + [0] for r in runs], dtype=float).T.copy()
arr = {L.LANE_WIND: w, L.LANE_PINK_SEED: seeds}
for lane in range(len(runs)):
    o = oracle.run(m.blob, len(runs), arr, traceLane=lane, traceMaxRows=4000)
    with backend.TrajectoryBatch(m.blob) as b:
        b.setLanes(len(runs), arr); b.setTrace(lane, 4000); b.run()
        res, st = b.results(); tr = b.trace(); cnt = b.laneCounters()
    ot = o["trace"]
    n = min(len(tr), len(ot))
    dtime = np.abs(tr[:n, 0] - ot[:n, 0])
    first = int(np.argmax(dtime > 1e-9)) if (dtime > 1e-9).any() else -1
    print("lane", lane, "steps gpu/oracle/ref", cnt[lane, 0], o["counters"][lane, 0], runs[lane]["steps"], "rejects", cnt[lane, 2], o["counters"][lane, 2],
          "first time divergence at step", first)
    print("  results gpu   ", res[lane])
    print("  results oracle", o["results"][lane])
    print("  results ref   ", np.array(runs[lane]["results"]))
    if first > 0:
        for i in range(max(0, first - 3), min(n, first + 3)):
            print("   step", i, "t", tr[i, 0], ot[i, 0], "dt", tr[i, 14], ot[i, 14], "err", tr[i, 15], ot[i, 15], "z", tr[i, 3], ot[i, 3])
    # relative error-estimate differences before divergence
    k = first if first > 0 else n
    rel = np.abs(tr[1:k, 15] - ot[1:k, 15]) / np.maximum(np.abs(ot[1:k, 15]), 1e-300)
    print("  err-estimate rel diff before divergence: max %.3e median %.3e" % (rel.max() if rel.size else 0, np.median(rel) if rel.size else 0))
    pos = np.abs(tr[1:k, 1:4] - ot[1:k, 1:4]).max() if k > 1 else 0
    print("  max abs position diff before divergence: %.3e" % pos)
