"""Run-time tuning sweep (on the GPU box): times one pass of a workload for combinations of the library's tuning switches.

    python tools/sweep.py [--lanes N] [--def FILE.mapleaf] [--libs a.so,b.so] [--blocks 1024,512] [--bails 500,750] [--gens 6] [--set key=value ...]

Each line is one JSON record: total / free-flight / under-chute device time of the second of two passes and the counters."""
import argparse, itertools, json, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def child(args):
    os.environ.setdefault("MLB_SHAPE_TUNING", "0")         # sweeps time the shape they name, not the library's measured choice
    import bench
    from mapleaf_b200 import L, SimDefinition, backend, buildModel
    if args.config >= 0:
        cfg = dict(bench.CONFIGS[args.config])
        if args.no_motor:
            cfg["motor"] = False
        m = bench.buildConfigModel(cfg)
        lanes = bench.laneInputs(cfg, m, 0, args.lanes)
        args.definition = "config%d%s" % (args.config, "-nomotor" if args.no_motor else "")
    else:
        sd = SimDefinition(os.path.join(REPO, "mapleaf_b200", "data", "Simulations", args.definition), silent=True, disableDistributionSampling=True)
        for kv in args.set or []:
            k, v = kv.split("=", 1); sd.setValue(k, v)
        m = buildModel(sd)
        wind, seeds = bench.sampleInputs(0, args.lanes)
        lanes = {L.LANE_WIND: wind}
        if int(m.blob[L.H_TURB]) in (L.TURB_PINK1D, L.TURB_PINK2D, L.TURB_PINK3D):
            lanes[L.LANE_PINK_SEED] = seeds
    for block, bail, gens in itertools.product(args.blocks.split(","), args.bails.split(","), args.gens.split(",")):
        for key, val in (("MLB_BLOCK_SIZE", block), ("MLB_BAIL_PERMILLE", bail), ("MLB_NUM_GENERATIONS", gens)):
            if val in ("", "default"):
                os.environ.pop(key, None)          # the library's own default
            else:
                os.environ[key] = val
        with backend.TrajectoryBatch(m.blob) as b:
            b.setLanes(args.lanes, lanes)
            b.run(); b.sync()
            b.run(); b.sync()
            c = b.counters()
        print(json.dumps({"lib": os.path.basename(os.environ.get("MLB_LIB", "default")), "def": args.definition, "lanes": args.lanes, "block": c["flight_block"], "bail": bail or "default",
                          "gens": c["flight_generations"], "ms": round(c["last_run_ms"], 2), "flight_ms": round(c["dominant_kernel_ms"], 2), "chute_ms": round(c["chute_kernel_ms"], 2),
                          "Msteps_per_s": round(c["steps"] / c["last_run_ms"] / 1e3, 2), "steps": c["steps"], "ok": c["lanes_ok"], "steps_3dof": c["steps_3dof"],
                          "evals": c["derivative_evals"], "rejects": c["rejected_steps"]}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--lanes", type=int, default=151552)
    ap.add_argument("--def", dest="definition", default="MonteCarloAdaptive.mapleaf")
    ap.add_argument("--libs", default="")
    ap.add_argument("--blocks", default="", help="comma list of CTA shapes (MLB_BLOCK_SIZE); empty: the library's choice")
    ap.add_argument("--bails", default="", help="comma list of hand-over thresholds in permille (MLB_BAIL_PERMILLE); empty: the library's default")
    ap.add_argument("--gens", default="", help="comma list of generation counts (MLB_NUM_GENERATIONS); empty: the library's default")
    ap.add_argument("--set", action="append")
    ap.add_argument("--config", type=int, default=-1, help="a bench.py configuration (definition, overrides and sampled inputs) instead of --def")
    ap.add_argument("--no-motor", action="store_true", help="with --config 3: leave the motor factors unsampled")
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child or not args.libs:
        child(args)
    else:
        for lib in args.libs.split(","):
            env = dict(os.environ)
            if lib != "default":
                env["MLB_LIB"] = lib if os.path.isabs(lib) else os.path.join(REPO, lib)
            subprocess.call([sys.executable, os.path.abspath(__file__), "--child"] + [a for a in sys.argv[1:]], env=env)
