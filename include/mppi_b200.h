/*
 * mppi_b200.h - C ABI of the B200-native MPPI rollout engine (libmppi_b200.so).
 *
 * Drop-in boundary for ONE hot path of kohonda/mppi_playground: a
 * `pi_mpc.MPPI.forward` solve (reference src/pi_mpc/mppi.py:223-460) plus the
 * pieces of solver state around it (constructor :24-210, reset :212-221,
 * get_top_samples :462-487). The reference is pure Python and has no FFI of
 * its own; these are the entry points a Python shim binds with ctypes (see
 * INTEGRATION.md and mppi_playground_b200/_capi.py). Plain pointers and
 * sizes only - no torch types cross this boundary.
 *
 * Conventions
 *   - every function returns MPPI_OK (0) or a negative MppiStatus; the text of
 *     the last error of the calling thread is mppi_last_error();
 *   - "d_" pointers are device pointers on the handle's device, "h_" pointers
 *     are host pointers; all arrays are fp32, row-major, dense;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *     stream); calls are asynchronous on that stream unless stated;
 *   - one handle is driven by one host thread at a time (the reference is not
 *     re-entrant either: global torch RNG, src/pi_mpc/mppi.py:93);
 *   - there is no CPU fallback: without a usable sm_100 device mppi_create
 *     fails with MPPI_ERR_CUDA.
 */
#ifndef MPPI_B200_H_
#define MPPI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPPI_ABI_VERSION 1
#define MPPI_MAX_DIM_CONTROL 4
#define MPPI_MAX_DIM_STATE 8
#define MPPI_MAX_SG_WINDOW 33
#define MPPI_MAX_MODEL_PARAMS 32

typedef enum MppiStatus {
  MPPI_OK = 0,
  MPPI_ERR_INVALID = -1, /* bad argument / config (the reference's AssertionError / ValueError cases) */
  MPPI_ERR_CUDA = -2,    /* CUDA runtime error, text in mppi_last_error() */
  MPPI_ERR_STATE = -3,   /* call sequence error (e.g. top samples before any solve) */
  MPPI_ERR_UNSUPPORTED = -4
} MppiStatus;

/* Built-in env models compiled as __device__ functions. */
typedef enum MppiModel {
  MPPI_MODEL_PENDULUM = 0,     /* example/pendulum.py:17-47 */
  MPPI_MODEL_CARTPOLE = 1,     /* example/cartpole.py:17-81 */
  MPPI_MODEL_MOUNTAINCAR = 2,  /* example/mountaincar.py:17-55 */
  MPPI_MODEL_NAVIGATION2D = 3, /* src/envs/navigation_2d.py:218-279 */
  MPPI_MODEL_RACING = 4,       /* src/envs/racing_env.py:327-372 + example/racing.py:110-159 */
  MPPI_MODEL_CARTPOLE_CONTINUOUS = 5, /* example/mujoco_cartpole.py:20-80 */
  MPPI_MODEL_GOAL_IN_DANGER_ZONE = 6  /* src/envs/goal_in_danger_zone.py:113-156 */
} MppiModel;

/* lambda_ argument of the constructor (src/pi_mpc/mppi.py:183-210). */
typedef enum MppiLambdaMode {
  MPPI_LAMBDA_FIXED = 0,
  MPPI_LAMBDA_MPO = 1,
  MPPI_LAMBDA_LBPS = 2,
  MPPI_LAMBDA_ESSPS = 3
} MppiLambdaMode;

/* Layout of MppiConfig.model_params (floats), per model.
 *   PENDULUM / CARTPOLE / MOUNTAINCAR: none (the reference hard-codes them).
 *   NAVIGATION2D (12): v_min v_max w_min w_max  goal_x goal_y  x_lo x_hi y_lo y_hi  dt  obstacle_weight
 *   RACING (17):       a_min a_max s_min s_max  wheelbase v_max  x_lo x_hi y_lo y_hi  dt  Qc Ql Qv Qo Qin Qdin
 *   CARTPOLE_CONTINUOUS: none.
 *   GOAL_IN_DANGER_ZONE (11): v_min v_max w_min w_max  dt  goal_x goal_y  centre_x centre_y  radius  collision_cost
 */
#define MPPI_NAV2D_NUM_PARAMS 12
#define MPPI_RACING_NUM_PARAMS 17
#define MPPI_GOAL_ZONE_NUM_PARAMS 11

/* Mirrors the keyword arguments of MPPI.__init__ (src/pi_mpc/mppi.py:24-47). */
/* MppiConfig.flags: force the loop form of racing / navigation2d rollouts (default: picked from the sample
 * count, see paired_loop_pays in csrc/mppi_engine.cu) */
#define MPPI_CFG_FORCE_PAIRED 1 /* two samples per thread, packed fp32 */
#define MPPI_CFG_FORCE_SINGLE 2 /* one sample per thread */

typedef struct MppiConfig {
  int32_t abi_version; /* MPPI_ABI_VERSION */
  int32_t model;       /* MppiModel */
  int32_t horizon;     /* T */
  int32_t num_samples; /* K rolled by THIS handle (the shard size when sharded) */
  int32_t dim_state;
  int32_t dim_control;
  float u_min[MPPI_MAX_DIM_CONTROL];
  float u_max[MPPI_MAX_DIM_CONTROL];
  float sigmas[MPPI_MAX_DIM_CONTROL];
  int32_t lambda_mode; /* MppiLambdaMode */
  double lambda_;      /* used when lambda_mode == FIXED */
  double lbps_delta;
  double essps_target_ess; /* <= 0: default total_samples / 10 (mppi.py:185-187) */
  double lambda_min;
  double lambda_max;
  double exploration; /* fraction of zero-mean samples (mppi.py:266) */
  int32_t use_sg_filter;
  int32_t sg_window_size;
  int32_t sg_poly_order;
  int32_t sg_coeffs_given; /* 1: use sg_coeffs as given (e.g. torch's fp32 pinv), 0: engine solves them */
  float sg_coeffs[MPPI_MAX_SG_WINDOW];
  uint64_t seed;
  int32_t device; /* CUDA device ordinal */
  /* sample sharding: this handle owns global sample ids
   * [sample_offset, sample_offset + num_samples) out of total_samples.
   * Single GPU: sample_offset = 0, total_samples = num_samples (or 0). */
  int64_t sample_offset;
  int64_t total_samples;
  int32_t num_model_params;
  float model_params[MPPI_MAX_MODEL_PARAMS];
  int32_t block_size; /* 0: engine picks */
  int32_t flags;      /* MPPI_CFG_*: launch-geometry overrides (tests, tuning); 0: engine picks */
} MppiConfig;

typedef struct MppiHandle MppiHandle;

/* ---- lifetime ---------------------------------------------------------------- */
/* MPPI.__init__ (mppi.py:24-210): validates the config, allocates device state
 * (warm start [T,du], SG history [T-1,du], costs [K], partial buffers). */
int mppi_create(const MppiConfig* cfg, MppiHandle** out);
void mppi_destroy(MppiHandle* h);
/* MPPI.reset (mppi.py:212-221): zero the warm start and the SG history. */
int mppi_reset(MppiHandle* h, void* stream);
const char* mppi_last_error(void);
int mppi_abi_version(void);

/* ---- model data ---------------------------------------------------------------- */
/* Replace the model parameter block (cost weights can change between solves:
 * example/racing.py:41-46 are plain attributes). n must match the model. */
int mppi_set_model_params(MppiHandle* h, const float* params, int32_t n);
/* Occupancy grid for slot 0 (obstacle map; NAVIGATION2D and RACING) or slot 1
 * (lane map; RACING). `grid` is the reference's [W,H] fp32 0/1 map, x on the
 * slow axis (src/envs/obstacle_map_2d.py:195), on the device if on_device != 0.
 * It is bit-packed once into the layout the kernels stage into shared memory;
 * grids too large for shared memory (beyond ~200 kB packed per solve, e.g. two
 * 2000 x 2000 maps) stay in global memory and are read through L2 by a separate
 * kernel instantiation (the reference's lookup has no size limit,
 * obstacle_map_2d.py:168-200). Synchronous. */
int mppi_set_map(MppiHandle* h, int32_t slot, const float* grid, int32_t on_device, int32_t width, int32_t height,
                 float cell_size, float origin_x, float origin_y);

/* Device rasteriser ("next" row 4): builds the occupancy grid of `slot` on the device instead of the
 * reference's Python loops and emits the packed layout the kernels read directly (no fp32 grid round trip).
 *   MPPI_RASTER_OBSTACLE  ObstacleMap.add_circle_obstacle / add_rectangle_obstacle
 *                         (src/envs/obstacle_map_2d.py:103-123, :125-160): free background, shapes paint 1.
 *                         h_discs [n_discs,3] = (centre cell x, centre cell y, radius_occ^2) - indices that fall
 *                         outside the map are clipped onto its border like the reference's np.clip (:121-122);
 *                         h_rects [n_rects,4] = (x_init, x_end, y_init, y_end), the already clipped half-open
 *                         slice of :150-157.
 *   MPPI_RASTER_LANE      LaneMap.populate_map (src/envs/lane_map_2d.py:68-88): occupied background, every
 *                         disc (centre-line cell inside the map, r2 = largest integer squared cell distance with
 *                         sqrt(r2) <= (lane_width / 2) / cell_size in fp64) paints 0 - the thresholded Euclidean
 *                         distance transform of the centre line, exactly. No rectangles.
 * The world -> cell conversions are the reference's fp64 numpy expressions and stay with the caller
 * (mppi_playground_b200/maps.py). d_grid_out: optional [W,H] fp32 0/1 grid (the reference's _map_torch), or NULL.
 * Re-rasterising a slot with unchanged geometry reuses its buffers and division proofs. Synchronous. */
#define MPPI_RASTER_OBSTACLE 0
#define MPPI_RASTER_LANE 1
int mppi_raster_map(MppiHandle* h, int32_t slot, int32_t mode, int32_t width, int32_t height, float cell_size,
                    float origin_x, float origin_y, const int32_t* h_discs, int32_t n_discs, const int32_t* h_rects,
                    int32_t n_rects, float* d_grid_out);

/* ---- one solve (MPPI.forward, mppi.py:223-460) ----------------------------------- */
/* d_state [ds]; d_refpath [T+1,4] (RACING only, else NULL: example/racing.py:73-81);
 * d_noise NULL for the in-kernel Philox sampler, or [K,T,du] = sigma*eps as
 * MultivariateNormal.rsample returns it (parity mode, mppi.py:261-263).
 * Outputs: d_action_seq [T,du], d_state_seq [T+1,ds] (the reference returns
 * the latter as [1,T+1,ds]). Asynchronous on `stream`. */
int mppi_solve(MppiHandle* h, const float* d_state, const float* d_refpath, const float* d_noise,
               float* d_action_seq, float* d_state_seq, void* stream);
/* Same solve with HOST buffers: pinned staging, H2D of state/refpath, the
 * solve, D2H of both outputs, and a stream synchronize before returning. */
int mppi_solve_host(MppiHandle* h, const float* h_state, const float* h_refpath, float* h_action_seq,
                    float* h_state_seq);

/* ---- sample-sharded solve (one handle per GPU, K split across ranks) ------------- */
/* Stage 1: roll this shard. Fixed-lambda / MPO modes also reduce the shard to
 * its partial (see mppi_partial_floats). LBPS / ESSPS stop after the costs. */
int mppi_shard_rollout(MppiHandle* h, const float* d_state, const float* d_refpath, const float* d_noise,
                       void* stream);
/* Stage 2 (LBPS / ESSPS only): d_costs_all [total_samples] = all shards' costs
 * gathered by the caller; runs the lambda search (identical on every rank)
 * and reduces this shard to its partial. */
int mppi_shard_lambda(MppiHandle* h, const float* d_costs_all, void* stream);
/* Stage 3: d_partials [n_shards, mppi_partial_floats()] gathered by the
 * caller; combines them and finishes the solve (SG filter, optimal-trajectory
 * rollout, state carry). Every rank computes identical outputs. */
int mppi_shard_finish(MppiHandle* h, const float* d_partials, int32_t n_shards, const float* d_state,
                      float* d_action_seq, float* d_state_seq, void* stream);
int32_t mppi_partial_floats(const MppiHandle* h);
/* Fused peer exchange (one process per GPU on one NVLink/NVSwitch node): every rank exports a CUDA IPC
 * handle of its mailbox (64 bytes), the caller all-gathers the handles ([world][64]) and connects. From
 * then on mppi_solve / mppi_solve_host on these shard handles (fixed lambda / MPO) is ONE kernel launch
 * per GPU: the finishing block stores the shard partial into every peer's mailbox over NVLink as 8-byte
 * (payload, sequence number) words, polls its own mailbox until every rank's words carry this solve's number
 * and finishes the solve - no collective call, no second launch.
 * LBPS / ESSPS handles keep using the staged mppi_shard_* path. mppi_p2p_status reports a timed-out
 * exchange (a peer that never launched its solve). */
int mppi_p2p_export(MppiHandle* h, uint8_t handle_out[64]);
int mppi_p2p_connect(MppiHandle* h, const uint8_t* handles, int32_t world, int32_t rank);
int mppi_p2p_status(MppiHandle* h, int32_t* timed_out);
/* Device-side barrier across the connected ranks on `stream` (one tiny kernel: every rank flags every peer's
 * mailbox over NVLink and waits for all flags in its own). All ranks leave within one NVLink latency of each
 * other; bench.py uses it to start every timed step on all GPUs together. Every rank must call it the same
 * number of times; a missing rank times out like the exchange (mppi_p2p_status reports a negative number). */
int mppi_p2p_barrier(MppiHandle* h, void* stream);
/* Same wiring for shard handles that live in ONE process (one process driving several GPUs with peer
 * access enabled, or several shards on one GPU): the mailboxes are plain device pointers. The shard
 * solves must then run concurrently (one stream per handle). */
int mppi_p2p_mailbox_ptr(MppiHandle* h, uint64_t* ptr);
int mppi_p2p_connect_local(MppiHandle* h, const uint64_t* mailbox_ptrs, int32_t world, int32_t rank);
/* Device buffers owned by the handle, valid until mppi_destroy. */
int mppi_costs_ptr(MppiHandle* h, const float** d_costs);     /* [num_samples] of the last solve */
int mppi_partial_ptr(MppiHandle* h, const float** d_partial); /* [mppi_partial_floats()] */

/* ---- inspection ------------------------------------------------------------------ */
/* softmax(-costs/lambda) of the last solve (mppi.py:376), d_weights [num_samples]. */
int mppi_weights(MppiHandle* h, float* d_weights, void* stream);
/* get_top_samples (mppi.py:462-487): the n highest-weight samples of the last
 * solve, weight-descending. Trajectories are not stored during the solve;
 * they are re-rolled from the sampler key (or from d_noise, which must still
 * be valid, in parity mode). d_traj [n,T+1,ds], d_w [n]. n <= 1024 runs the
 * radix select + multi-block re-roll of mppi_step_epilogue; larger n falls
 * back to a full radix sort of the costs. */
int mppi_top_samples(MppiHandle* h, int32_t n, float* d_traj, float* d_w, void* stream);
/* ---- control-step epilogue ("next" row 2) ---------------------------------------------------------------
 * What the reference's control loops run between two solves (example/racing.py:233-237,
 * example/navigation2d.py:39-44) in 1-3 small launches (step + flags + final select | winners' re-roll on one
 * block each | one select level per 8192 candidates above 8192):
 *   env.step(action_seq[0])          src/envs/racing_env.py:142-163 / navigation_2d.py:97-117: clamp, one
 *                                    dynamics step, goal test norm(next[:2] - goal) < threshold;
 *   env.collision_check(state_seq)   racing_env.py:374-384 / navigation_2d.py:281-291: obstacle-map value of
 *                                    every predicted position;
 *   solver.get_top_samples(top_n)    mppi.py:462-487: radix SELECT of the top_n lowest costs (= highest weights;
 *                                    ties by the lower sample id) instead of a sort of all K, winners re-rolled
 *                                    (racing / navigation2d: block-parallel rollout, one block per winner).
 * Every group is optional (NULL outputs / top_n = 0). */
typedef struct MppiStepEpilogue {
  const float* d_state;      /* [ds] env state before the step; NULL: the state of the last solve */
  const float* d_action_seq; /* [T,du] of the last solve (row 0 is executed) */
  const float* d_state_seq;  /* [T+1,ds] of the last solve */
  float goal_x, goal_y, goal_threshold; /* goal_threshold <= 0: no goal test (flag 0) */
  int32_t top_n;             /* 0: no top samples; at most 1024 on this path */
  /* sample-sharded solvers: the ranks' mppi_top_candidates outputs gathered by the caller ([n_cand] each);
   * NULL / 0: select from this handle's own costs */
  const float* d_cand_cost;
  const int32_t* d_cand_id;
  int32_t n_cand;
  const float* d_noise_global; /* parity mode on a sharded solver: injected noise indexed by GLOBAL sample id, else NULL */
  float* d_next_state;       /* out [ds], or NULL */
  float* d_flags;            /* out [1 + T+1]: is_goal_reached, then is_collisions[0..T] (0/1 like compute_cost); needs d_next_state */
  float* d_top_traj;         /* out [top_n, T+1, ds] */
  float* d_top_w;            /* out [top_n], descending */
  float* d_top_cost;         /* out [top_n] optional: the winners' costs (ascending) */
  int32_t* d_top_id;         /* out [top_n] optional: the winners' global sample ids */
} MppiStepEpilogue;
int mppi_step_epilogue(MppiHandle* h, const MppiStepEpilogue* args, void* stream);
/* This handle's n best (cost ascending, then id) samples of the last solve as (cost, GLOBAL sample id) pairs:
 * the per-rank half of get_top_samples on a sharded solver. n <= 1024. */
int mppi_top_candidates(MppiHandle* h, int32_t n, float* d_cand_cost, int32_t* d_cand_id, void* stream);
/* Kernels the last mppi_step_epilogue / mppi_top_candidates / mppi_top_samples (select path) launched. */
int32_t mppi_last_epilogue_launches(const MppiHandle* h);
/* _states_prediction (mppi.py:508-524): roll n action sequences d_actions [n,T,du]
 * from d_state; d_traj [n,T+1,ds]. Used by get_samples_from_posterior (mppi.py:489-506). */
int mppi_rollout_actions(MppiHandle* h, const float* d_state, const float* d_actions, int32_t n, float* d_traj,
                         void* stream);
/* lambda the last solve's weights used, and the one the next solve will use
 * (they differ in MPO mode, mppi.py:376 vs :398). Synchronises `stream`. */
int mppi_get_lambda(MppiHandle* h, double* lambda_used, double* lambda_next, void* stream);
/* Warm start [T,du] and SG history [T-1,du] (the solver's whole carried state
 * besides MPO's rho/Adam moments; mppi.py:157,163,452-458). */
int mppi_get_carry(MppiHandle* h, float* d_prev_action_seq, float* d_history, void* stream);
int mppi_set_carry(MppiHandle* h, const float* d_prev_action_seq, const float* d_history, void* stream);
/* Device pointer of the live warm start buffer [T,du] (read-only for callers). */
int mppi_prev_action_ptr(MppiHandle* h, const float** d_prev_action_seq);
/* Number of kernels the last mppi_solve / mppi_solve_host launched. */
int32_t mppi_last_launch_count(const MppiHandle* h);
/* Launch geometry picked for the rollout kernel. */
int mppi_launch_info(const MppiHandle* h, int32_t* grid, int32_t* block, int32_t* smem_bytes);
/* How the cell index x / cell_size of map `slot` is evaluated: fast_division = 1 when the
 * 3-instruction exact sequence was proven bit-identical to the IEEE division for this cell size
 * by an exhaustive device check over all 2^32 inputs (mismatches = 0), else the true division is
 * used. model_flags: bit0 both racing grids share one geometry, bit1 unit wheelbase. */
int mppi_map_info(const MppiHandle* h, int32_t slot, int32_t* fast_division, uint64_t* mismatches,
                  int32_t* model_flags);
/* Profiling aid: with enable != 0 every block of the solve kernel stamps %globaltimer (ns) into a row of 24
 * slots; h_out (may be NULL) receives [min(max_blocks, grid), 24] stamps of the last solve. Row 0 is the
 * finisher block (0 start, 1 warm-up pass over, 2 all workers' tickets seen, 3 partials in shared memory,
 * 8-9 inside the combine, 5 combined, 7 SG / carry done, 10-13 inside the optimal-trajectory rollout,
 * 6 finished; sharded fused solves: 16 exchange entered, 17 partial sent to the peers, 18 all peers' words seen and
 * gathered (19 = 18)); rows 1.. are the worker blocks (0 start, 1 inputs staged, 2 costs done, 3 weights done,
 * 4 partial written). Unused slots keep 0. */
int mppi_block_trace(MppiHandle* h, int32_t enable, uint64_t* h_out, int32_t max_blocks);
/* Exhaustive device self-test of the bounded arithmetic helpers against the general ones, over every
 * fp32 input of their claimed range: mismatches[0] tan_quarter vs tanf (|x| <= 0.78), [1]
 * wrap_angle_bounded vs wrap_angle (|x| < 9), [2] floored_remainder vs the fmodf form, [3] sincos_bounded
 * vs sincosf (|x| <= 4). All must be 0. */
int mppi_selftest(int32_t device, uint64_t mismatches[4]);
/* ---- racing reference path on the device ("next" row: example/racing.py:161-218) ------------------
 * racing_controller.calc_ref_trajectory runs on the host before every solve in the reference (a Python
 * loop over the N centre-line points). MppiRefPath keeps the centre line [n,3] (x, y, yaw), the per-row
 * index offsets int(round(travel / DL)) (formed on the host in fp64 like the reference, [rows] = T+1) and
 * the carried path index on the device; mppi_refpath_update writes the [T+1,4] reference path for d_state
 * into d_refpath_out with one small kernel - no host round trip between env step and solve.
 * mppi_refpath_index reads (current != NULL) and/or sets (set_value >= 0) the carried index. */
typedef struct MppiRefPath MppiRefPath;
int mppi_refpath_create(int32_t device, const float* h_path, int32_t n, const int32_t* h_index_offsets, int32_t rows,
                        float v_max, MppiRefPath** out);
void mppi_refpath_destroy(MppiRefPath* r);
int mppi_refpath_update(MppiRefPath* r, const float* d_state, float* d_refpath_out, void* stream);
int mppi_refpath_index(MppiRefPath* r, int32_t set_value, int32_t* current, void* stream);
/* Time the dominant (rollout) kernel of subsequent solves with CUDA events on
 * the launching stream: enable with 1, read back the mean/launch count with
 * mppi_kernel_time_ms (which synchronises the events). */
int mppi_kernel_timing(MppiHandle* h, int32_t enable);
int mppi_kernel_time_ms(MppiHandle* h, double* mean_ms, int32_t* launches);

/* Measured fp32 denominators of the roofline (csrc/mppi_microbench.cu): full-chip throughput of FFMA
 * (3-register form, 8 independent chains per thread, 2 x 1024 threads per SM), of the packed FFMA2, of the
 * never-contracted FMUL+FADD pair the reference's op order leaves (-fmad=false) and of its packed form, in
 * TFLOP/s (an FMA counts 2); dependent-issue latencies in SM cycles; the SM clock during the run
 * (clock64 / globaltimer). Synchronous, a few ms. MEASURED_PEAKS.json has no fp32 figure, hence this. */
typedef struct MppiFp32Report {
  double ffma_tflops, ffma2_tflops, fmul_fadd_tflops, fmul2_fadd2_tflops;
  double lat_ffma, lat_ffma2, lat_fadd2, lat_fmnmx, lat_fadd, lat_fmul2;
  double sm_clock_mhz;
  int32_t sms;
  int32_t reserved;
} MppiFp32Report;
int mppi_fp32_microbench(int32_t device, MppiFp32Report* out);

/* Philox4x32-10 known-answer hook for tests: out[4] = block(counter, key). Host code. */
void mppi_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);
/* Index the NEXT solve will key its sampler with (== number of solves so far). */
uint64_t mppi_solve_index(const MppiHandle* h);
/* sigma*eps the in-kernel sampler draws for solve `solve_index`, in the
 * reference's [K,T,du] layout (what MultivariateNormal.rsample returns,
 * mppi.py:261-263). Lets a test feed the engine's own noise to the oracle. */
int mppi_sample_noise(MppiHandle* h, uint64_t solve_index, float* d_noise_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPPI_B200_H_ */
