#!/usr/bin/env python
"""
bench.py — Monte Carlo trajectory-steps/s of the B200 trajectory kernels (and of the CPU reference beside them).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--lanes LANES_PER_GPU] [--impl b200|reference]

Workload (BASELINE.json configs[1], named in `config.workload`): the single-stage example rocket, RK45Adaptive with
PID step-size control, Hellman wind + PinkNoise2D turbulence, flown to the ground; 10^6 sampled rockets per GPU.
A "step" of this bench is one pass of the hot path over one batch: every lane integrated from launch to touchdown.
Inputs are synthetic in the sense of the contract (no data files): winds are drawn exactly as the reference draws them
(random.Random("2394781").gauss per component) and pink-noise seeds like random.Random(2394781).randrange(10^6).

What the JSON line reports:
  value ......... accepted integrator steps of all lanes of all GPUs / device time, inputs resident in HBM
  e2e ........... same metric through the public call sequence with HOST buffers: mlb_set_lanes (H2D of the sampled
                  inputs from pinned memory) -> mlb_run -> mlb_results (D2H of the 7 result columns + status) each step
  roofline ...... FP64: algorithmic FLOPs of the launch (SURVEY.md §8d per-evaluation figures x the kernel's own
                  evaluation counters) / integrate-kernel duration, against the FP64 FMA peak measured in this run
  cpu_baseline .. the reference (or its C restatement) flying a bounded sample of the same definition on host cores
N > 1 (torchrun): every rank integrates its own contiguous block of samples (weak scaling), then results are gathered
with one NCCL all-gather written straight from the kernel's output buffer, and the per-GPU moments are combined.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

DEFINITION = os.path.join(REPO, "mapleaf_b200", "data", "Simulations", "MonteCarloAdaptive.mapleaf")
WORKLOAD = "single-stage rocket, RK45Adaptive+PID, Hellman wind + PinkNoise2D, to ground (BASELINE configs[1])"
METRIC = "MC trajectory-steps/sec (FP64)"
UNIT = "trajectory-steps/s"
SEED = "2394781"

# algorithmic FP64 work per unit (add/sub/mul/div/sqrt/compare = 1, a transcendental call = 1). EXACT counts of the
# reference's arithmetic, from the CPU oracle compiled with an op-counting scalar (oracle/opcount.hpp, tools/count_flops.py, 64 sampled lanes
# of this workload: 8.8e8 add, 9.7e8 mul, 1.0e8 div, 4.6e7 sqrt, 1.3e8 compare, 6.4e7 transcendental); tests/test_host_runner.py re-counts them.
FLOP_EVAL_6DOF = 4533.0             # one Rocket._getAppliedForce + rigid-body derivative, Mach <= 0.8 or >= 1.4
FLOP_EVAL_6DOF_TRANSONIC = 11147.0  # the transonic fin blend evaluates four strip sums (Fins.py:508-528)
FLOP_EVAL_3DOF = 521.6              # under parachute
FLOP_STEP_ALGEBRA_6DOF = 3617.0     # rest of an RK45 attempt: stage composition, two solution rows, error norm, events, step-size PID
FLOP_STEP_ALGEBRA_3DOF = 576.2


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


def algorithmicFlops(c):
    e3 = c["derivative_evals_3dof"]
    e6 = c["derivative_evals"] - e3
    et = c["derivative_evals_transonic"]
    attempts3 = c["steps_3dof"]
    attempts6 = c["steps"] - attempts3 + c["rejected_steps"]
    return (FLOP_EVAL_6DOF * (e6 - et) + FLOP_EVAL_6DOF_TRANSONIC * et + FLOP_EVAL_3DOF * e3
            + FLOP_STEP_ALGEBRA_6DOF * attempts6 + FLOP_STEP_ALGEBRA_3DOF * attempts3)


def runB200(args):
    import torch
    import torch.distributed as dist
    from mapleaf_b200 import L, SimDefinition, backend, buildModel, distributed

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

    n = args.lanes
    sd = SimDefinition(DEFINITION, silent=True, disableDistributionSampling=True)
    model = buildModel(sd)
    tSample = time.perf_counter()
    wind, seeds = sampleInputs(rank * n, n)
    tSample = time.perf_counter() - tSample

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch = backend.TrajectoryBatch(model.blob, device=local)
    dWind = torch.from_numpy(wind).to(dev)
    dSeeds = torch.from_numpy(seeds).to(dev)
    # the kernel writes its results straight into the all-gather send buffer
    gathered = torch.empty((world, n, L.RES_SIZE), dtype=torch.float64, device=dev)
    gatheredStatus = torch.empty((world, n), dtype=torch.int32, device=dev)
    mine, mineStatus = gathered[rank], gatheredStatus[rank]
    batch.setLanesDevice(n, {L.LANE_WIND: dWind.data_ptr(), L.LANE_PINK_SEED: dSeeds.data_ptr()}, keepAlive=(dWind, dSeeds))
    batch.setResultBuffers(mine.data_ptr(), mineStatus.data_ptr(), keepAlive=(gathered, gatheredStatus))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def onePass():
        flush.zero_()
        batch.run()
        launches = 1
        if world > 1:
            # the only exchange of the path: landing locations / apogees / ... of every sample, and the moments
            dist.all_gather_into_tensor(gathered.view(-1), mine.reshape(-1))
            dist.all_gather_into_tensor(gatheredStatus.view(-1), mineStatus)
            cnt, mean, M2 = batch.moments()               # per-GPU count / mean / M2 (device kernels), combined over ranks
            launches += 2
            distributed.gatherMoments(cnt, mean, M2, device=dev)
        return launches

    for _ in range(args.warmup):
        onePass()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.steps):
            launches += onePass()
        e1.record()
        barrier()
    totalMs = e0.elapsed_time(e1)
    c = batch.counters()                      # counters of the last pass (every pass integrates the same lanes)
    kernelMsLast = c["dominant_kernel_ms"]
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
    pinWind = torch.from_numpy(wind).pin_memory()
    pinSeeds = torch.from_numpy(seeds).pin_memory()
    e2eBatch = backend.TrajectoryBatch(model.blob, device=local)
    hostArrays = {L.LANE_WIND: pinWind.numpy(), L.LANE_PINK_SEED: pinSeeds.numpy()}

    def e2ePass():
        e2eBatch.setLanes(n, hostArrays)                     # H2D of the step's inputs
        e2eBatch.run()
        res, st = e2eBatch.results()                          # D2H of the step's results
        return res, st
    batch.close()
    del gathered, gatheredStatus, flush
    torch.cuda.empty_cache()
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
    h2d = n * (3 + 3) * 8
    d2h = n * (L.RES_SIZE * 8 + 4)
    e2eBatch.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops = algorithmicFlops(c)
    peak = backend.fp64Peak(local)
    achieved = flops / (kernelMsLast * 1e-3) / 1e12
    hbmBytes = n * ((3 + 3) * 8 + L.RES_SIZE * 8 + L.FIN_SIZE * 8 + 4 + 6 * 8)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": msPerStep, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "definition": "mapleaf_b200/data/Simulations/MonteCarloAdaptive.mapleaf", "lanes_per_gpu": n,
                   "lanes_total": n * world, "accepted_steps_per_pass": stepsAll, "derivative_evals_per_pass": evalsAll,
                   "rejected_steps_per_pass": rejAll, "lanes_ok": okAll,
                   "rank0_steps_3dof": c["steps_3dof"], "rank0_evals_3dof": c["derivative_evals_3dof"], "rank0_evals_transonic": c["derivative_evals_transonic"], "sampling": "random.Random('%s') winds, Random(%s).randrange seeds (C++ bit-exact sampler, %.2f s host)" % (SEED, SEED, tSample),
                   "l2": "256 MiB device buffer zeroed between passes (L2 flush)", "parallelism": "lanes sharded contiguously over %d GPU(s); NCCL all-gather of the 7 result columns + moments" % world},
        "sims_per_sec": n * world / (msPerStep * 1e-3),
        "e2e": {"value": e2eValue, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "sims_per_sec": n * world / (e2eS / args.steps)},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     # DRAM bytes per launch: 18.4 MB per lane measured by ncu at full occupancy (profiles/r1d_full_occupancy_metrics.csv:
                     # 1.36 TB read + 1.43 TB written for 151 552 lanes; local-memory spill traffic, the algorithmic figure is
                     # hbm_algorithmic_bytes), scaled to this launch
                     "traffic": 18.4e6 * n, "hbm_algorithmic_bytes": hbmBytes,
                     "kernel": "integrateLanes", "kernel_ms": kernelMsLast, "algorithmic_flops_per_launch": flops,
                     "peak_source": "FP64 DFMA microbenchmark run on this GPU by mlb_fp64_peak (MEASURED_PEAKS.json has no FP64 entry); nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2",
                     "hbm_algorithmic_gbs": hbmBytes / (kernelMsLast * 1e-3) / 1e9, "note": "not HBM- or tensor-bound: ~1e5 FLOP per byte of state moved"},
    }
    if world == 1 and not args.skip_cpu:
        line["cpu_baseline"] = cpuBaseline(min(os.cpu_count() or 1, 16), 1, warmup=1)      # one untimed pass: worker start-up, first-call parsing
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpuJobs(kind, cores, perWorker):
    wind, seeds = sampleInputs(0, cores * perWorker)
    if kind == "reference":
        return [(tuple(wind[:, i]), (seeds[0, i], seeds[1, i])) for i in range(cores * perWorker)]
    return [(wind[:, i * perWorker:(i + 1) * perWorker].copy(), seeds[:, i * perWorker:(i + 1) * perWorker].copy()) for i in range(cores)]


def cpuBaseline(cores, steps, warmup=0, kindWanted=None):
    """Times the CPU arm on a bounded sample; returns the cpu_baseline object (value in trajectory-steps/s)."""
    from oracle import ref_bench
    kind = kindWanted or ("reference" if ref_bench.referenceAvailable() else "port")
    perWorker = 1 if kind == "reference" else 32
    arm = ref_bench.CpuArm(kind, DEFINITION, {}, cores)
    try:
        jobs = cpuJobs(kind, cores, perWorker)
        for _ in range(warmup):
            arm.step(jobs)
        tot, sims, wall, per = 0, 0, 0.0, []
        for _ in range(steps):
            s, m, w = arm.step(jobs)
            tot, sims, wall = tot + s, sims + m, wall + w
            per.append(w)
    finally:
        arm.close()
    what = "unmodified MAPLEAF (baseline/_ref) Simulation.run()" if kind == "reference" else "C restatement (oracle/mapleaf_oracle.c)"
    return {"value": tot / wall, "unit": UNIT, "cores": cores, "kind": kind, "sims_per_sec": sims / wall, "ms_per_step": 1e3 * wall / steps,
            "sample": "%d trajectories per step (%d per worker process x %d processes) of the same definition and sampled inputs, %s; %d accepted steps in %.1f s"
                      % (cores * perWorker, perWorker, cores, what, tot, wall)}


def runReference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    b = cpuBaseline(cores, args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": b["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": b["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "definition": "mapleaf_b200/data/Simulations/MonteCarloAdaptive.mapleaf"},
            "cpu_baseline": b, "sims_per_sec": b["sims_per_sec"],
            "e2e": {"value": b["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--lanes", type=int, default=1000000, help="sampled rockets per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs: leave out the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        runReference(args)
    else:
        runB200(args)


if __name__ == "__main__":
    main()
