/*
 * mapleaf_b200.h — C ABI of libmapleaf_b200.so: batched Monte Carlo 6-DOF trajectory integration on one B200.
 *
 * The reference (henrystoldt/MAPLEAF) has no native boundary for this path: a Monte Carlo run is a Python loop
 * that builds and flies one rocket per sample. Each entry point below therefore names the reference code whose
 * work it takes over (paths relative to MAPLEAF/):
 *
 *   mlb_create ........ Simulation.__init__ / Environment.__init__ / Rocket.__init__ once per run instead of once per
 *                       sample (SimulationRunners/SingleSimulations.py:30-51, ENV/environment.py:39-103,
 *                       Rocket/rocket.py:34-219); the flattened result is the model block of include/mlb_layout.h
 *   mlb_set_lanes ..... the sampled values of SimDefinition.resampleProbabilisticValues (IO/simDefinition.py:471-532),
 *                       one column per Monte Carlo sample ("lane")
 *   mlb_run ........... the body of _runSimulations_SingleThreaded's loop: Simulation.run -> Rocket.timeStep ->
 *                       integrator -> Rocket._getAppliedForce (SimulationRunners/MonteCarlo.py:64-76,
 *                       SingleSimulations.py:53-139, Rocket/rocket.py:549-698, Motion/Integration.py:189-449)
 *   mlb_results ....... the RocketFlight reductions used by postProcess (IO/rocketFlight.py:27-76,
 *                       SimulationRunners/MonteCarlo.py:52-62)
 *   mlb_finals ........ the last RigidBodyState of each flight (what the reference's batch regression compares,
 *                       SimulationRunners/Batch.py:488-521)
 *   mlb_moments ....... the mean / standard deviation summaries of _showResults (MonteCarlo.py:148-166, IO/Plotting.py:690-735)
 *   mlb_force_eval .... one Rocket._getAppliedForce per state: the wind-tunnel sweep of WindTunnelSimulation.runSweep
 *                       (SingleSimulations.py:444-475)
 *
 * Conventions. All functions return 0 on success and a non-zero code on failure; mlb_last_error() then returns a
 * message (thread-local). The caller owns every host buffer; the library copies inputs during the call and owns its
 * device memory until mlb_destroy. A handle is bound to one CUDA device and is not thread-safe. Per-lane failures
 * never abort the batch: they are reported in `status` (ST_* of mlb_layout.h). There is no CPU fallback: every entry
 * point fails with MLB_ERR_CUDA when no usable device is present.
 */
#ifndef MAPLEAF_B200_H
#define MAPLEAF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mlb_handle mlb_handle;

#define MLB_OK 0
#define MLB_ERR_ARG 1       /* bad argument / malformed model block */
#define MLB_ERR_CUDA 2      /* CUDA runtime error (message has the CUDA error string) */
#define MLB_ERR_STATE 3     /* call order (e.g. run before set_lanes) */
#define MLB_ERR_UNSUPPORTED 4

typedef struct {
    int64_t lanes;
    int64_t steps;              /* accepted integrator steps (calls of Rocket.timeStep), summed over lanes */
    int64_t derivative_evals;   /* Rocket._getAppliedForce evaluations */
    int64_t rejected_steps;     /* adaptive attempts discarded (Integration.py:338-347) */
    int64_t steps_3dof;                 /* of `steps`: taken under a parachute (3-DoF body, rocket.py:700-726) */
    int64_t derivative_evals_3dof;      /* of `derivative_evals`: 3-DoF */
    int64_t derivative_evals_transonic; /* of the 6-DoF evaluations: 0.8 < Mach < 1.4 (cubic fin-force blend, Fins.py:508-528) */
    int64_t lanes_ok, lanes_step_limit, lanes_nan;
    int64_t kernel_launches;    /* CUDA kernels launched by the last mlb_run */
    double last_run_ms;         /* device time of the last mlb_run (CUDA events on the launch stream) */
    double dominant_kernel_ms;  /* of which: the free-flight (6-DoF) kernel, all its generations */
    double chute_kernel_ms;     /* of which: the under-chute (3-DoF) kernel */
    int64_t flight_block;       /* CTA shape of the free-flight kernel in the last run */
    int64_t flight_generations; /* launches of the free-flight kernel in the last run */
} mlb_counters_t;

/* model_blob: the FP64 model block (include/mlb_layout.h), n = its size in BYTES */
int mlb_create(const void* model_blob, size_t n, int device, mlb_handle** out);
void mlb_destroy(mlb_handle* h);
const char* mlb_last_error(void);
const char* mlb_version(void);
/* 1 when this build of the library is MODEL-SPECIALISED: compiled with one model block baked into the kernels as constants
   (mapleaf_b200.backend.specializedLibrary: the host writes the block as a header and compiles csrc/mlb_api.cu with -DMLB_SPEC). The
   reference rebuilds its Python object graph per sample (Rocket/rocket.py:34-219); the generic library interprets one flattened block for
   all samples; a specialised one has the block's structure (stages, component list, tableau, environment models) resolved at compile time.
   Same ABI, same results up to the rounding of folded constants; mlb_create refuses any other model block and mlb_set_lane_models is
   unsupported (MLB_ERR_UNSUPPORTED). 0 for the generic library. */
int mlb_is_specialized(void);

/* Per-lane inputs. perLaneArrays[k] is a host array of [components(arrayIds[k])][nLanes] doubles, lane fastest
   (LANE_WIND: 3, LANE_INIT_STATE: 13, LANE_PINK_SEED: 3, LANE_MOTOR_ADJUST: 8 = impulseAdjustFactor and burnTimeAdjustFactor of the
   motor of stages 0-3, Propulsion.py:38-40,109-128; LANE_START: 2 = start time and first step size). Arrays not given use the model block's value for every lane. */
int mlb_set_lanes(mlb_handle* h, int64_t nLanes, const double* const* perLaneArrays, const int* arrayIds, int nArrays);
/* Same, with the arrays already resident in this device's memory (no copy is made; the caller keeps them alive). */
int mlb_set_lanes_device(mlb_handle* h, int64_t nLanes, const double* const* devLaneArrays, const int* arrayIds, int nArrays);

/* Per-lane MODEL blocks, for Monte Carlo keys that change what the constructors precompute (masses, geometry, launch site, rail ...;
   SimDefinition.resampleProbabilisticValues re-draws ANY key that has a `_stdDev` companion, IO/simDefinition.py:471-532, and the reference
   then rebuilds the whole rocket): models = nLanes x strideDoubles doubles, row i = the first H_SHARED_LEN doubles (or more) of lane i's own
   model block, built by the host from that sample's values. Every row must have the structure of the handle's block (same layout, stages,
   components, integrator); wind, initial state, rail and end condition are then read from the lane's own header. Call after mlb_set_lanes
   (which clears it); models = NULL switches back to the handle's block for every lane. Slower than the shared block (each constant is a
   global load): it is the general path, the common keys (wind, initial state, motor factors) stay per-lane arrays. */
int mlb_set_lane_models(mlb_handle* h, const double* models, int64_t strideDoubles);

/* Integrates every lane to its end condition. `stream` is a cudaStream_t (NULL = default stream). Asynchronous with
   respect to the host; mlb_results / mlb_finals / mlb_counters / mlb_sync wait for completion. */
int mlb_run(mlb_handle* h, void* stream);
int mlb_sync(mlb_handle* h);

/* results: nLanes x 7 row-major (RES_* of mlb_layout.h); status: nLanes (ST_*). Either may be NULL. */
int mlb_results(mlb_handle* h, double* results, int32_t* status);
/* finals: nLanes x 14 row-major: position, velocity, orientation quaternion, body angular velocity, time */
int mlb_finals(mlb_handle* h, double* finals);
int mlb_counters(mlb_handle* h, mlb_counters_t* out);
/* per lane (nLanes x 6): steps, derivative evaluations, rejected attempts, 3-DoF steps, 3-DoF evaluations,
   transonic 6-DoF evaluations */
int mlb_lane_counters(mlb_handle* h, int64_t* out);
/* Stage separations of every lane: nLanes x (MLB_MAX_STAGES - 1) x SEP_W doubles (mlb_layout.h SEP_*): the whole rocket's state, the time
   and the step size at each separation, i.e. what a dropped stage's own flight starts from (Rocket._stageSeparation ->
   SingleSimulations.createNewDetachedStage, rocket.py:461-474, SingleSimulations.py:280-299). The caller flies the dropped stages as a second
   batch: a handle built from the dropped stage's model block, LANE_INIT_STATE = the separation state, LANE_START = (time, step size). */
int mlb_stage_separations(mlb_handle* h, double* out);
/* Device addresses of the result block (nLanes x 7 doubles) and status (nLanes int32): the send buffers of the
   multi-GPU gather. Valid until the next mlb_set_lanes* or mlb_destroy. */
int mlb_device_results(mlb_handle* h, void** results, void** status);
/* Makes the next runs write results / status into caller-owned device buffers of the same shapes (e.g. the send
   buffer of a collective). Call after mlb_set_lanes*; the caller keeps them alive until the next mlb_set_lanes*. */
int mlb_set_result_buffers(mlb_handle* h, void* results, void* status);

/* Column statistics of the result block, computed on the device over lanes with status ST_OK:
   out[0] = count, then for each of the 7 columns: mean, M2 (sum of squared deviations). 15 doubles. */
int mlb_moments(mlb_handle* h, double* out15);

/* Flight paths of lanes [firstLane, firstLane + nLanes): what RocketFlight.times / rigidBodyStates hold for those samples
   (IO/rocketFlight.py:12-25), the input of `MonteCarlo.output flightPaths` (MonteCarlo.py:60-62, Plotting._keepNTimeSteps). Rows of 16
   doubles: time, position 3, velocity 3, quaternion 4, body-frame angular velocity 3, step size taken, error estimate. Row 0 is the initial
   state; then the state after every `every`-th accepted step; the final state is always the last row. A lane whose buffer fills up drops
   every other stored row and doubles its stride, so any flight ends as maxRowsPerLane/2 .. maxRowsPerLane evenly strided states
   (112 B of state per kept step: the HBM-write-bound companion of the trajectory kernel). Call before mlb_run; firstLane < 0 switches
   the log off. mlb_get_flight_log: rows = nLanes x maxRowsPerLane x 16 doubles, rowsPerLane = nLanes counts. */
int mlb_set_flight_log(mlb_handle* h, int64_t firstLane, int64_t nLanes, int64_t every, int64_t maxRowsPerLane);
int mlb_get_flight_log(mlb_handle* h, double* rows, int64_t* rowsPerLane);

/* Test / analysis hooks */
/* the first maxRows steps of one lane, every step (later steps are not recorded): mlb_set_flight_log(h, lane, 1, 0, maxRows) */
int mlb_set_trace(mlb_handle* h, int64_t lane, int64_t maxRows);
int mlb_get_trace(mlb_handle* h, double* rows, int64_t* nRows);
/* force the first n step sizes of every lane (lock-step replay of a recorded adaptive step sequence) */
int mlb_set_dt_replay(mlb_handle* h, const double* dts, int64_t n);
/* One force evaluation per state (states: n x 13 row-major, times: n, winds: n x 3 or NULL). out: n x 18:
   total force xyz (body frame), total moment about CG xyz, aerodynamic force xyz, Mach, density, wind xyz, CG xyz, mass */
int mlb_force_eval(mlb_handle* h, int64_t n, const double* states, const double* times, const double* winds, double* out);

/* Host-side CPython-compatible sampler (random.Random: MT19937 init_by_array + gauss). key: the 32-bit words CPython
   derives from the seed (little-endian chunks of the integer seed; for a str seed the integer is
   int.from_bytes(s + sha512(s))). Fills out[nDraws] with gauss(mu[i % nPer], sigma[i % nPer]) in draw order after
   skipping `skip` draws. Mirrors simDefinition.rng (IO/simDefinition.py:205-216,505-517). */
int mlb_host_gauss_stream(const uint32_t* key, int keyLen, int64_t skip, const double* mu, const double* sigma, int nPer, int64_t nDraws, double* out);

/* random.randrange(n) stream of the same generator (pink-noise seeds: ENV/TurbulenceModelling.py:191-195) */
int mlb_host_randrange_stream(const uint32_t* key, int keyLen, int64_t skip, uint32_t n, int64_t nDraws, double* out);
/* Both streams continued from a live CPython generator (random.Random.getstate()): the 624 state words, the position and the cached
   second Gaussian are read and updated in place, so the Python object can be set to the advanced state afterwards. This is what lets
   the Monte Carlo runner draw 10^6 runs' inputs at C speed while the SimDefinition's own generator stays the single source of the
   stream (IO/simDefinition.py:216,471-532; ENV/TurbulenceModelling.py:191-195 for the seeds drawn from the global `random` module). */
int mlb_host_gauss_from_state(uint32_t* state624, int* pos, int* hasNext, double* next, const double* mu, const double* sigma, int nPer, int64_t nDraws, double* out);
int mlb_host_randrange_from_state(uint32_t* state624, int* pos, uint32_t n, int64_t nDraws, double* out);

/* Measured FP64 FMA throughput of `device` in TFLOP/s (independent DFMA chains on every SM): the denominator of the
   FP64 roofline this path is reported against. */
int mlb_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
