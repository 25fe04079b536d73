"""Exact algorithmic FLOP counts of the hot path, from the op-counting build of the CPU oracle (oracle/opcount.hpp):
every FP64 add/sub/mul/div/sqrt/compare = 1, a transcendental call = 1, over N sampled lanes of a bench configuration. Prints the per-unit
table bench.py's CONFIGS[...]["flops"] entries are taken from. CPU only:  python tools/count_flops.py [lanes] [--config C]"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO)
import bench
from oracle import oracle

args = [a for a in sys.argv[1:] if not a.startswith("--")]
lanes = int(args[0]) if args else 32
config = int(sys.argv[sys.argv.index("--config") + 1]) if "--config" in sys.argv else 1
cfg = bench.CONFIGS[config]
m = bench.buildConfigModel(cfg)
arrays = bench.laneInputs(cfg, m, 0, lanes)
out, ops = oracle.runCounted(m.blob, lanes, arrays)
steps = int(out["counters"][:, 0].sum()); evals = int(out["counters"][:, 1].sum()); rejects = int(out["counters"][:, 2].sum())
b = ops["buckets"]
table = {k: (v["flops"] / v["count"] if v["count"] else 0.0) for k, v in b.items()}
print(json.dumps(dict(config=config, definition=cfg["definition"], lanes=lanes, steps=steps, derivative_evals=evals, rejected=rejects,
                      ops={k: ops[k] for k in ("add", "mul", "div", "sqrt", "compare", "transcendental", "uncounted")},
                      flops=ops["flops"], flops_per_step=ops["flops"] / steps,
                      counts={k: v["count"] for k, v in b.items()}, flops_per_unit=table,
                      flops_outside_steps=ops["flops"] - sum(v["flops"] for v in b.values())), indent=1))
