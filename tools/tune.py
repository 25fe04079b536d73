"""Kernel tuning aid. `tune.py build` compiles variants of the CUDA library into gpurun_out/variants/ (here, no GPU needed);
`tune.py run [lanes]` (on the GPU box) times one pass of the bench workload with each variant."""
import json, os, subprocess, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
VDIR = os.path.join(REPO, "mapleaf_b200", "lib", "variants")
# name -> extra nvcc flags; edit for the experiment at hand (see DESIGN.md for the switches: MLB_COMP_NOINLINE, MLB_COLD_MASK,
# MLB_SYNC_STRIDE, MLB_TUNE_SHAPES, MLB_TUNE_512X2, MLB_TUNE_640X2, MLB_MATH_NOINLINE, MLB_LOCKSTEP, MLB_FIN_UNIFORM)
VARIANTS = {
    "default": [],
    "cold15": ["-DMLB_COLD_MASK=15"],
}

def build(names=None):
    from mapleaf_b200 import backend
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for name, flags in VARIANTS.items():
        if names and name not in names: continue
        out = os.path.join(VDIR, name + ".so")
        cmd = ["nvcc"] + backend.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-o", out, os.path.join(backend.CSRC_DIR, "mlb_api.cu")]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        txt = p.communicate()[0]
        lines = txt.splitlines()
        for i, l in enumerate(lines):
            if "integrateLanes" in l and "Compiling" in l:
                print(name, "|", lines[i + 2].strip(), "|", lines[i + 3].strip())
        if p.returncode: print(name, "FAILED\n", txt[-2000:])

def runOne(lib, lanes, definition, overrides=()):
    os.environ["MLB_LIB"] = lib
    import numpy as np
    import bench
    from mapleaf_b200 import L, SimDefinition, backend, buildModel
    sd = SimDefinition(os.path.join(REPO, "mapleaf_b200", "data", "Simulations", definition), silent=True, disableDistributionSampling=True)
    for kv in overrides:
        k, v = kv.split("=", 1); sd.setValue(k, v)
    m = buildModel(sd)
    wind, seeds = bench.sampleInputs(0, lanes)
    with backend.TrajectoryBatch(m.blob) as b:
        b.setLanes(lanes, {L.LANE_WIND: wind, L.LANE_PINK_SEED: seeds})
        b.run(); b.sync()
        b.run(); b.sync()
        c = b.counters()
    print(json.dumps({"lib": os.path.basename(lib), "def": definition, "lanes": lanes, "ms": c["last_run_ms"], "steps_per_s": c["steps"] / c["last_run_ms"] * 1e3,
                      "steps": c["steps"], "ok": c["lanes_ok"], "steps_3dof": c["steps_3dof"], "evals": c["derivative_evals"], "overrides": list(overrides)}), flush=True)

if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "one":
        runOne(sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5:])
    else:
        lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 151552
        defs = sys.argv[3:] or ["MonteCarloAdaptive.mapleaf"]
        for f in sorted(os.listdir(VDIR)):
            for d in defs:
                subprocess.call([sys.executable, __file__, "one", os.path.join(VDIR, f), str(lanes), d])
