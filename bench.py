#!/usr/bin/env python
"""
bench.py — Monte Carlo trajectory-steps/s of the B200 trajectory kernels (and of the CPU reference beside them).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--lanes LANES_PER_GPU] [--scaling weak|strong]
                    [--total-lanes T] [--impl b200|reference] [--skip-cpu]

Workloads are the configurations BASELINE.json names (`--config`, default 1 = the one its metric is quoted on):
  0  single-stage rocket, fixed-step RK4 0.01 s, sampled wind, to apogee                       (configs[0])
  1  same rocket, RK45Adaptive + PID, Hellman wind + PinkNoise2D, to the ground, 10^6 lanes     (configs[1])
  2  canard-controlled rocket: 100 Hz GNC PID + 5-D deflection table, RK4 forced in the ascent  (configs[2])
  3  NASA two-stage orbital rocket, WGS84 + J2, tabulated aero, sampled motor factors, 10^5     (configs[3])
  4  landing-dispersion sweep: the single-stage rocket flown to the ground with RK4, 10^7 lanes
     in total with `--scaling strong` (split over the GPUs) or 1.25 x 10^6 per GPU (weak)      (configs[4])
A "step" of this bench is one pass of the hot path over one batch: every lane integrated from launch to its end condition.
Inputs are synthetic in the sense of the contract (no data files): winds are drawn exactly as the reference draws them
(random.Random("2394781").gauss per component), pink-noise seeds like random.Random(2394781).randrange(10^6).

What the JSON line reports:
By default the batch flies on the library compiled for the configuration's model block (backend.specializedLibrary: compiled ahead of
time by __graft_entry__.build(), or here before any timed region; `config.library` says which); `--generic` uses the generic library.
  value ......... accepted integrator steps of all lanes of all GPUs / device time, inputs resident in HBM
  e2e ........... same metric through the public call sequence with HOST buffers: mlb_set_lanes (H2D of the sampled
                  inputs from pinned memory) -> mlb_run -> mlb_results (D2H of the 7 result columns + status) each step
  roofline ...... FP64, for the dominant kernel (the free-flight kernel, all its generations): algorithmic FLOPs of its launches
                  (counted per-evaluation figures x the kernel's own evaluation counters) / its CUDA-event duration, against the FP64
                  FMA peak measured in this run; `traffic` and the executed DFMA/DADD/DMUL counts come from the committed ncu
                  metrics pass of the same code and workload (profiles/, named in `traffic_source`), scaled by lanes
  cpu_baseline .. the reference (or its C restatement) flying a bounded sample of the same definition on host cores
N > 1 (torchrun): every rank integrates its own contiguous block of samples, then results are gathered with one NCCL
all-gather written straight from the kernel's output buffer, and the per-GPU moments are combined.
"""
import argparse
import csv
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

DATA = os.path.join(REPO, "mapleaf_b200", "data", "Simulations")
METRIC = "MC trajectory-steps/sec (FP64)"
UNIT = "trajectory-steps/s"
SEED = "2394781"

# Algorithmic FP64 work per unit (add/sub/mul/div/sqrt/compare = 1, a transcendental call = 1): EXACT counts of the
# reference's arithmetic from the CPU oracle compiled with an op-counting scalar (oracle/opcount.hpp; `python tools/count_flops.py N
# --config C` prints this table; tests/test_host_runner.py re-counts config 1). eval6 / eval6t: one Rocket._getAppliedForce + rigid-body
# derivative outside / inside 0.8 < Mach < 1.4 (the transonic fin blend evaluates four strip sums, Fins.py:508-528); eval3: under a
# parachute; step6 / step3: the rest of an integrator attempt (stage composition, solution rows, error norm, events, controllers).
CONFIGS = {
    0: dict(definition="MonteCarlo.mapleaf", overrides={}, lanes=1000000,
            workload="single-stage rocket, fixed-step RK4 0.01 s, sampled wind, to apogee (BASELINE configs[0])",
            flops=dict(eval6=4486.5, eval6t=11112.2, eval3=0.0, step6=706.0, step3=0.0)),
    1: dict(definition="MonteCarloAdaptive.mapleaf", overrides={}, lanes=1000000,
            workload="single-stage rocket, RK45Adaptive+PID, Hellman wind + PinkNoise2D, to ground (BASELINE configs[1])",
            flops=dict(eval6=4533.0, eval6t=11147.0, eval3=521.6, step6=3617.0, step3=576.2)),
    2: dict(definition="Canards.mapleaf", overrides={}, lanes=1000000,
            workload="canard-controlled rocket, 100 Hz GNC PID + 5-D deflection table, RK4 forced in the ascent, to apogee (BASELINE configs[2])",
            flops=dict(eval6=6683.6, eval6t=6683.6, eval3=0.0, step6=11823.1, step3=0.0)),       # step6 includes the 100 Hz control loop's table scans
    3: dict(definition="NASATwoStageOrbitalRocket.mapleaf", overrides={}, lanes=100000, motor=True,
            workload="NASA two-stage orbital rocket, RK45Adaptive, WGS84 + J2, tabulated aero, sampled motor factors, 200 s (BASELINE configs[3])",
            flops=dict(eval6=2758.6, eval6t=3723.4, eval3=0.0, step6=3903.0, step3=0.0)),
    4: dict(definition="MonteCarlo.mapleaf", overrides={"SimControl.EndCondition": "Altitude", "SimControl.EndConditionValue": "0"}, lanes=1250000,
            workload="landing-dispersion sweep: single-stage rocket, RK4 0.01 s, sampled wind, to ground (BASELINE configs[4])",
            flops=dict(eval6=4485.6, eval6t=11112.2, eval3=466.0, step6=705.4, step3=181.4)),
}


def buildConfigModel(cfg):
    from mapleaf_b200 import SimDefinition, buildModel
    sd = SimDefinition(os.path.join(DATA, cfg["definition"]), silent=True, disableDistributionSampling=True)
    for k, v in cfg["overrides"].items():
        sd.setValue(k, v)
    return buildModel(sd)


def sampleInputs(first, n):
    """Winds and pink-noise seeds of samples [first, first + n), in the reference's draw order (C++ sampler)."""
    from mapleaf_b200 import backend
    # draw set 0 is consumed when the SimDefinition is constructed (IO/simDefinition.py:216)
    w = backend.hostGaussStream(SEED, [2.0, 0.0, 0.0], [5.0, 5.0, 0.5], n, skipDraws=3 * (first + 1))
    s = backend.hostRandrangeStream(int(SEED), 1000000, 2 * n, skipDraws=2 * first).reshape(n, 2)
    wind = np.ascontiguousarray(w.T)
    seeds = np.zeros((3, n))
    seeds[:2] = s.T
    return wind, seeds


def sampleInputsPython(first, n):
    """The same two streams drawn with CPython's own `random` (what the reference does): the reference arm's inputs."""
    r = random.Random(SEED)
    for _ in range(3 * (first + 1)):
        r.gauss(0, 1)
    wind = np.array([[r.gauss(2.0, 5.0), r.gauss(0.0, 5.0), r.gauss(0.0, 0.5)] for _ in range(n)]).T.copy()
    g = random.Random(int(SEED))
    for _ in range(2 * first):
        g.randrange(1000000)
    seeds = np.zeros((3, n))
    seeds[:2] = np.array([[g.randrange(1000000), g.randrange(1000000)] for _ in range(n)], dtype=float).T
    return wind, seeds


def motorFactors(first, n, python=False):
    """impulseAdjustFactor ~ N(1, 0.02), burnTimeAdjustFactor ~ N(1, 0.05) for the motors of stages 0 and 1 (config 3): per-lane
    LANE_MOTOR_ADJUST rows. Separation and upper-stage ignition then happen at a different time in every lane."""
    from mapleaf_b200 import L
    mu, sg = [1.0, 1.0, 1.0, 1.0], [0.02, 0.05, 0.02, 0.05]
    if python:
        r = random.Random(SEED + "m")
        for _ in range(4 * first):
            r.gauss(0, 1)
        f = np.array([[r.gauss(m, s) for m, s in zip(mu, sg)] for _ in range(n)])
    else:
        from mapleaf_b200 import backend
        f = backend.hostGaussStream(SEED + "m", mu, sg, n, skipDraws=4 * first)
    a = np.ones((2 * L.MLB_MAX_STAGES, n))
    a[:4] = f.T
    return a


def laneInputs(cfg, model, first, n, python=False):
    """{LANE_* id: [components, n] array} of samples [first, first + n) for a configuration."""
    from mapleaf_b200 import L
    wind, seeds = (sampleInputsPython if python else sampleInputs)(first, n)
    arrays = {L.LANE_WIND: wind}
    if int(model.header("H_TURB")) in (L.TURB_PINK1D, L.TURB_PINK2D, L.TURB_PINK3D):
        arrays[L.LANE_PINK_SEED] = seeds
    if cfg.get("motor"):
        arrays[L.LANE_MOTOR_ADJUST] = motorFactors(first, n, python)
    return arrays


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.stop = gpu, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.5)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=15)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def algorithmicFlops(c, f):
    """(free-flight kernel FLOPs, under-chute kernel FLOPs) of one pass from the kernels' counters and a per-unit table."""
    e3 = c["derivative_evals_3dof"]
    e6 = c["derivative_evals"] - e3
    et = c["derivative_evals_transonic"]
    attempts3 = c["steps_3dof"]
    attempts6 = c["steps"] - attempts3 + c["rejected_steps"]
    return (f["eval6"] * (e6 - et) + f["eval6t"] * et + f["step6"] * attempts6, f["eval3"] * e3 + f["step3"] * attempts3)


def measuredKernelMetrics(configId, specialised=False, block=1024):
    """Per-lane DRAM traffic and executed FP64 instructions of the free-flight kernel from the committed ncu metrics pass of this
    code and workload (profiles/*_full_occupancy_metrics_lanes<N>.csv: one launch over N lanes, newest first; files of another
    configuration carry `config<C>` in their name): (dict per lane, source file) or (None, None)."""
    import re
    pdir = os.path.join(REPO, "profiles")
    want = "config%d" % configId
    names = [f for f in os.listdir(pdir) if re.search(r"_full_occupancy_metrics_lanes\d+\.csv$", f) and (want in f or (configId == 1 and "config" not in f))
             and (("_spec_" in f) == bool(specialised))          # passes of the model-specialised build carry `_spec_` in their name
             and (("_block%d_" % block in f) if block != 1024 else ("_block" not in f))]      # ... and `_block512_` when they captured that CTA shape
    for name in sorted(names, reverse=True):
        lanes = int(re.search(r"_lanes(\d+)\.csv$", name).group(1))
        vals = {}
        with open(os.path.join(pdir, name)) as fh:
            for row in csv.reader(l for l in fh if l.startswith('"')):
                if len(row) >= 15 and row[0] != "ID" and "flightLanes" in row[4]:
                    # a pass over several launches (the generations of one flight: `-c 12`) adds up its `.sum` metrics
                    v = float(row[14].replace(",", ""))
                    vals[row[12]] = vals.get(row[12], 0.0) + v if row[12].endswith(".sum") else v
        if "dram__bytes_read.sum" in vals:
            out = {"dram_bytes": (vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]) / lanes}
            for k, short in (("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "dfma"), ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dadd"),
                             ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "dmul")):
                if k in vals:
                    out[short] = vals[k] / lanes
            return out, "profiles/" + name
    return None, None


def runB200(args):
    import torch
    import torch.distributed as dist
    from mapleaf_b200 import L, backend, distributed

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: mapleaf_b200 has no CPU path")
    backend.lib()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = CONFIGS[args.config]
    if args.scaling == "strong":
        total = args.total_lanes or 10000000
        n = total // world
    else:
        n = args.lanes or cfg["lanes"]
    model = buildConfigModel(cfg)
    tSample = time.perf_counter()
    arrays = laneInputs(cfg, model, rank * n, n)
    tSample = time.perf_counter() - tSample

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # model-specialised library (default): compiled for this model block, normally ahead of time by __graft_entry__.build(); a missing one is
    # compiled here by rank 0, before any timed region
    specInfo = None
    if not args.generic:
        cached = os.path.exists(backend.specializedLibraryPath(model.blob))
        tBuild = time.perf_counter()
        if rank == 0:
            backend.buildSpecializedLibrary(model.blob)
        if world > 1:
            dist.barrier()
        specInfo = {"library": "mapleaf_b200/lib/spec/" + os.path.basename(backend.specializedLibraryPath(model.blob)),
                    "compiled": "ahead of time (cached by content hash)" if cached else "in this run, %.1f s of nvcc before the timed region" % (time.perf_counter() - tBuild)}
    batch = backend.TrajectoryBatch(model.blob, device=local, specialize=not args.generic)
    devArrays = {k: torch.from_numpy(a).to(dev) for k, a in arrays.items()}
    # the kernels write their results straight into the all-gather send buffer
    gathered = torch.empty((world, n, L.RES_SIZE), dtype=torch.float64, device=dev)
    gatheredStatus = torch.empty((world, n), dtype=torch.int32, device=dev)
    mine, mineStatus = gathered[rank], gatheredStatus[rank]
    batch.setLanesDevice(n, {k: t.data_ptr() for k, t in devArrays.items()}, keepAlive=devArrays)
    batch.setResultBuffers(mine.data_ptr(), mineStatus.data_ptr(), keepAlive=(gathered, gatheredStatus))
    if args.log_every > 0:
        # north_star's log companion: every lane keeps every log_every-th accepted state (16 doubles = 128 B per kept step) in HBM
        batch.setFlightLog(0, n, args.log_every, args.log_rows)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def onePass():
        flush.zero_()
        batch.run()
        if world > 1:
            # the only exchange of the path: landing locations / apogees / ... of every sample, and the moments
            dist.all_gather_into_tensor(gathered.view(-1), mine.reshape(-1))
            dist.all_gather_into_tensor(gatheredStatus.view(-1), mineStatus)
            cnt, mean, M2 = batch.moments()               # per-GPU count / mean / M2 (device kernels), combined over ranks
            distributed.gatherMoments(cnt, mean, M2, device=dev)

    for _ in range(args.warmup):
        onePass()
        batch.sync()          # untimed: lets a model-specialised library settle its CTA shape on measured passes (mlb_run, shape tuning)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.steps):
            onePass()
        e1.record()
        barrier()
    totalMs = e0.elapsed_time(e1)
    c = batch.counters()                      # counters of the last pass (every pass integrates the same lanes)
    logRows = int(batch.flightLogRowCounts().sum()) if args.log_every > 0 else 0
    launches = args.steps * (c["kernel_launches"] + (2 if world > 1 else 0))     # this library's kernels: prologue, free-flight generations, under-chute (+ the two moment kernels)
    if world > 1:
        t = torch.tensor([totalMs], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        totalMs = float(t.item())
        tot = torch.tensor([c["steps"], c["lanes_ok"], c["derivative_evals"], c["rejected_steps"]], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        stepsAll, okAll, evalsAll, rejAll = [float(x) for x in tot.tolist()]
    else:
        stepsAll, okAll, evalsAll, rejAll = float(c["steps"]), float(c["lanes_ok"]), float(c["derivative_evals"]), float(c["rejected_steps"])
    msPerStep = totalMs / args.steps
    value = stepsAll / (msPerStep * 1e-3)

    # ---- end to end through the public call sequence with host buffers ----
    batch.close()
    del gathered, gatheredStatus, flush
    torch.cuda.empty_cache()
    h2d = sum(a.nbytes for a in arrays.values())
    d2h = n * (L.RES_SIZE * 8 + 4)
    e2eValue = e2eS = None
    if not args.skip_e2e:
        pinned = {k: torch.from_numpy(a).pin_memory() for k, a in arrays.items()}
        hostArrays = {k: t.numpy() for k, t in pinned.items()}
        e2eBatch = backend.TrajectoryBatch(model.blob, device=local, specialize=not args.generic)

        def e2ePass():
            e2eBatch.setLanes(n, hostArrays)                     # H2D of the step's inputs
            e2eBatch.run()
            return e2eBatch.results()                             # D2H of the step's results
        e2ePass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res, st = e2ePass()
        torch.cuda.synchronize()
        e2eS = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2eS], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2eS = float(t.item())
        e2eValue = stepsAll / (e2eS / args.steps)
        e2eBatch.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops6, flops3 = algorithmicFlops(c, cfg["flops"])
    peak = backend.fp64Peak(local)
    flightMs, chuteMs = c["dominant_kernel_ms"], c["chute_kernel_ms"]
    achieved = flops6 / (flightMs * 1e-3) / 1e12
    hbmBytes = n * (sum(a.shape[0] for a in arrays.values()) * 8 + L.RES_SIZE * 8 + L.FIN_SIZE * 8 + 4 + 6 * 8)
    perLane, source = measuredKernelMetrics(args.config, specialised=not args.generic, block=int(c["flight_block"]))
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": perLane["dram_bytes"] * n if perLane else None,
                "traffic_source": (source + " (ncu dram__bytes_read.sum + dram__bytes_write.sum of one full-occupancy launch of this code and workload, per lane x lanes)") if perLane else None,
                "hbm_algorithmic_bytes": hbmBytes,
                "kernel": "flightLanes<%d> x %d generations" % (c["flight_block"], c["flight_generations"]), "kernel_ms": flightMs, "algorithmic_flops_per_launch": flops6,
                "share_of_step": flightMs / c["last_run_ms"],
                "peak_source": "FP64 DFMA microbenchmark run on this GPU by mlb_fp64_peak (MEASURED_PEAKS.json has no FP64 entry); nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2",
                "other_kernels": {"chuteLanes": {"kernel_ms": chuteMs, "algorithmic_flops": flops3, "achieved": flops3 / (chuteMs * 1e-3) / 1e12 if chuteMs > 0 else None},
                                  "initLanes_ms": c["last_run_ms"] - flightMs - chuteMs},
                "note": "FP64-pipe bound by design (~1e5 FLOP per algorithmic byte); the counted FLOPs are the reference's arithmetic (it recomputes the inertia three times per evaluation, 41 arcsines per fin in the transonic blend), the kernel executes less"}
    if perLane and "dfma" in perLane:
        ex = {k: perLane[k] * n for k in ("dfma", "dadd", "dmul") if k in perLane}
        roofline["executed_fp64_thread_instructions"] = ex
        roofline["executed_flops"] = 2 * ex.get("dfma", 0) + ex.get("dadd", 0) + ex.get("dmul", 0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": msPerStep, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "config_id": args.config, "definition": "mapleaf_b200/data/Simulations/" + cfg["definition"],
                   "overrides": cfg["overrides"], "lanes_per_gpu": n,
                   "library": ("model-specialised build: the model block compiled into the kernels as constants (backend.specializedLibrary); " + json.dumps(specInfo))
                              if specInfo else "generic build (interprets the model block from shared memory)",
                   "lanes_total": n * world, "accepted_steps_per_pass": stepsAll, "derivative_evals_per_pass": evalsAll,
                   "rejected_steps_per_pass": rejAll, "lanes_ok": okAll,
                   "rank0_steps_3dof": c["steps_3dof"], "rank0_evals_3dof": c["derivative_evals_3dof"], "rank0_evals_transonic": c["derivative_evals_transonic"],
                   "host_sampling_s": tSample,
                   "sampling": "random.Random('%s') winds, Random(%s).randrange seeds%s (C++ bit-exact sampler, %.2f s host)" % (SEED, SEED, ", motor factors" if cfg.get("motor") else "", tSample),
                   "l2": "256 MiB device buffer zeroed between passes (L2 flush)", "parallelism": "lanes sharded contiguously over %d GPU(s); NCCL all-gather of the 7 result columns + moments" % world},
        "sims_per_sec": n * world / (msPerStep * 1e-3),
        "e2e": ({"value": e2eValue, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "sims_per_sec": n * world / (e2eS / args.steps)}
                if e2eValue is not None else None),
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": roofline,
    }
    if args.log_every > 0:
        hbmPeak, hbmSource = 6541.1, "MEASURED_PEAKS.json hbm_gbs of this pool's B200 (driver-written) not found: the pool's recorded 6541.1 GB/s"
        try:
            with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
                hbmPeak, hbmSource = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)"
        except Exception:
            pass
        logBytes = logRows * 16 * 8
        gbs = logBytes / (c["last_run_ms"] * 1e-3) / 1e9
        line["log_roofline"] = {"bound": "hbm", "achieved": gbs, "peak": hbmPeak, "unit": "GB/s", "frac": gbs / hbmPeak, "bytes_written": logBytes,
                                "rows_kept": logRows, "every": args.log_every, "rows_per_lane_max": args.log_rows, "peak_source": hbmSource,
                                "note": "flight log of EVERY lane: 128 B per kept step written by the lane itself (one 128-byte row = four full 32-byte sectors); a lane whose "
                                        "buffer fills drops every other row in place and doubles its stride. The log rides on FP64-bound kernels: its write rate is what "
                                        "the trajectory rate allows, far below what HBM could take"}
    if world == 1 and not args.skip_cpu:
        line["cpu_baseline"] = cpuBaseline(cfg, min(os.cpu_count() or 1, 16), 3, warmup=1)      # one untimed pass: worker start-up, first-call parsing
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpuJobs(cfg, kind, cores, perWorker):
    """Bounded sample of the workload for the CPU arm. The reference arm's inputs are drawn with CPython's `random` alone."""
    n = cores * perWorker
    if kind == "reference":
        wind, seeds = sampleInputsPython(0, n)
        motor = motorFactors(0, n, python=True) if cfg.get("motor") else None
        motorPaths = [m["path"] for m in buildConfigModel(cfg).meta["motors"] if m is not None] if motor is not None else []
        jobs = []
        for i in range(n):
            ov = {"Environment.ConstantMeanWind.velocity": "({} {} {})".format(*wind[:, i]),
                  "Environment.PinkNoiseModel.randomSeed1": str(int(seeds[0, i])), "Environment.PinkNoiseModel.randomSeed2": str(int(seeds[1, i]))}
            for si, path in enumerate(motorPaths):          # LANE_MOTOR_ADJUST rows 2 si / 2 si + 1 = the factors of stage si's motor
                ov[path + ".impulseAdjustFactor"] = repr(float(motor[2 * si, i]))
                ov[path + ".burnTimeAdjustFactor"] = repr(float(motor[2 * si + 1, i]))
            jobs.append(ov)
        return jobs
    model = buildConfigModel(cfg)
    arrays = laneInputs(cfg, model, 0, n)
    return [{k: a[:, i * perWorker:(i + 1) * perWorker].copy() for k, a in arrays.items()} for i in range(cores)]


def cpuBaseline(cfg, cores, steps, warmup=0, kindWanted=None):
    """Times the CPU arm on a bounded sample; returns the cpu_baseline object (value in trajectory-steps/s)."""
    from oracle import ref_bench
    kind = kindWanted or ("reference" if ref_bench.referenceAvailable() else "port")
    perWorker = 4 if kind == "reference" else 32
    arm = ref_bench.CpuArm(kind, os.path.join(DATA, cfg["definition"]), cfg["overrides"], cores)
    try:
        jobs = cpuJobs(cfg, kind, cores, perWorker)
        for _ in range(warmup):
            arm.step(jobs[:cores])
        tot, sims, wall, per = 0, 0, 0.0, []
        for _ in range(steps):
            s, m, w = arm.step(jobs)
            tot, sims, wall = tot + s, sims + m, wall + w
            per.append(w)
    finally:
        arm.close()
    what = "unmodified MAPLEAF (baseline/_ref) Simulation.run()" if kind == "reference" else "C restatement (oracle/mapleaf_oracle.c)"
    return {"value": tot / wall, "unit": UNIT, "cores": cores, "kind": kind, "sims_per_sec": sims / wall, "ms_per_step": 1e3 * wall / steps,
            "sample": "%d trajectories per step (%d per worker process x %d processes, handed out one at a time) of the same definition and sampled inputs, %s; %d steps, %d accepted integrator steps in %.1f s"
                      % (cores * perWorker, perWorker, cores, what, steps, tot, wall)}


def runReference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    b = cpuBaseline(cfg, cores, args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": b["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": b["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["workload"], "config_id": args.config, "definition": "mapleaf_b200/data/Simulations/" + cfg["definition"], "overrides": cfg["overrides"]},
            "cpu_baseline": b, "sims_per_sec": b["sims_per_sec"],
            "e2e": {"value": b["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS), help="BASELINE.json configs index (default 1: the headline workload)")
    ap.add_argument("--lanes", type=int, default=0, help="sampled rockets per GPU (default: the configuration's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: --total-lanes split over the GPUs")
    ap.add_argument("--total-lanes", type=int, default=0, help="with --scaling strong (default 10^7, BASELINE configs[4])")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--generic", action="store_true", help="fly on the generic library (interprets the model block) instead of the one compiled for this model block")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs: leave out the cpu_baseline leg")
    ap.add_argument("--skip-e2e", action="store_true", help="record runs of the large configurations: leave out the end-to-end leg (e2e: null)")
    ap.add_argument("--log-every", type=int, default=0, help="keep every n-th accepted state of EVERY lane in the device flight log (0: no log) and report its HBM write rate")
    ap.add_argument("--log-rows", type=int, default=256, help="log rows per lane (128 B each)")
    args = ap.parse_args()
    if args.impl == "reference":
        runReference(args)
    else:
        runB200(args)


if __name__ == "__main__":
    main()
