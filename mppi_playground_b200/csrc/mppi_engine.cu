// mppi_engine.cu - host side of libmppi_b200.so: handle management, launch
// geometry, and the C ABI declared in include/mppi_b200.h.
//
// Built for sm_100a only:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false ...
// (-fmad=false: see mppi_models.cuh - the reference never contracts a*b+c).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <cub/device/device_radix_sort.cuh>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mppi_b200.h"
#include "mppi_kernels.cuh"
#include "mppi_epilogue.cuh"

using namespace mppi;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return fail(MPPI_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));    \
  } while (0)

constexpr int kNumSMs = 148;
constexpr unsigned kMaxSmem = 227 * 1024;

// Every entry point runs on the handle's device and leaves the caller's current device as it found it
// (torch reads cudaGetDevice: a solve or a destructor on cuda:1 must not move later `device="cuda"` work).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) err = cudaSetDevice(device);
  }
  ~DeviceGuard() {
    int now = -1;
    if (prev >= 0 && cudaGetDevice(&now) == cudaSuccess && now != prev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define ON_DEVICE(dev)                                                                              \
  DeviceGuard _guard(dev);                                                                          \
  if (_guard.err != cudaSuccess) return fail(MPPI_ERR_CUDA, "cudaSetDevice(%d): %s", (dev), cudaGetErrorString(_guard.err))

struct ModelInfo {
  int ds, du, maps, n_params;
  bool refpath;
  int tail_per_step;
};
ModelInfo model_info(int model) {
  switch (model) {
    case MPPI_MODEL_PENDULUM: return {Pendulum::DS, Pendulum::DU, 0, 0, false, 0};
    case MPPI_MODEL_CARTPOLE: return {Cartpole::DS, Cartpole::DU, 0, 0, false, 0};
    case MPPI_MODEL_MOUNTAINCAR: return {MountainCar::DS, MountainCar::DU, 0, 0, false, 0};
    case MPPI_MODEL_NAVIGATION2D: return {Navigation2D::DS, Navigation2D::DU, 1, MPPI_NAV2D_NUM_PARAMS, false, tail_per_step<Navigation2D>()};
    case MPPI_MODEL_RACING: return {Racing::DS, Racing::DU, 2, MPPI_RACING_NUM_PARAMS, true, tail_per_step<Racing>()};
    case MPPI_MODEL_CARTPOLE_CONTINUOUS: return {CartpoleContinuous::DS, CartpoleContinuous::DU, 0, 0, false, 0};
    case MPPI_MODEL_GOAL_IN_DANGER_ZONE:
      return {GoalInDangerZone::DS, GoalInDangerZone::DU, 0, MPPI_GOAL_ZONE_NUM_PARAMS, false, 0};
    default: return {0, 0, 0, 0, false, 0};
  }
}

// Savitzky-Golay coefficients = first row of pinv(vander(-h..h, order+1))
// (mppi.py:568-596) via the normal equations in fp64, rounded to fp32.
bool savgol_coeffs(int window, int order, float* out) {
  const int h = (window - 1) / 2, n = order + 1;
  std::vector<double> G(n * n, 0.0), inv(n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int x = -h; x <= h; ++x) s += pow((double)x, i) * pow((double)x, j);
      G[i * n + j] = s;
    }
  // invert G by Gauss-Jordan with partial pivoting
  for (int i = 0; i < n; ++i) inv[i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (fabs(G[r * n + c]) > fabs(G[piv * n + c])) piv = r;
    if (fabs(G[piv * n + c]) < 1e-300) return false;
    for (int j = 0; j < n; ++j) {
      std::swap(G[c * n + j], G[piv * n + j]);
      std::swap(inv[c * n + j], inv[piv * n + j]);
    }
    double d = G[c * n + c];
    for (int j = 0; j < n; ++j) {
      G[c * n + j] /= d;
      inv[c * n + j] /= d;
    }
    for (int r = 0; r < n; ++r)
      if (r != c) {
        double f = G[r * n + c];
        for (int j = 0; j < n; ++j) {
          G[r * n + j] -= f * G[c * n + j];
          inv[r * n + j] -= f * inv[c * n + j];
        }
      }
  }
  // pinv(A)[0, m] = sum_j inv[0][j] * x_m^j
  for (int m = 0; m < window; ++m) {
    double s = 0;
    for (int j = 0; j < n; ++j) s += inv[j] * pow((double)(m - h), j);
    out[m] = (float)s;
  }
  return true;
}

}  // namespace

struct MppiHandle {
  MppiConfig cfg{};
  ModelInfo mi{};
  int device = 0;
  int E = 0, E_pad = 0, P = 0;
  // launch geometry of the rollout kernel: [0] the in-kernel sampler (two samples per thread when the model's
  // bounded loop applies), [1] injected noise (always one sample per thread)
  struct Geometry {
    int spt = 1, block = 0, grid = 0;
    unsigned smem = 0, stage_bytes = 0;
  } geo[2];
  SolveParams base{};  // everything that does not change between solves
  // device buffers
  float* d_prev_action = nullptr;
  float* d_history = nullptr;
  float* d_nominal_snapshot = nullptr;
  float* d_state_snapshot = nullptr;
  DeviceScalars* d_sc = nullptr;
  float* d_costs = nullptr;
  float* d_block_partials = nullptr;
  float* d_rank_partial = nullptr;
  float* d_dry = nullptr;  // dummy targets of the finisher block's warm-up pass
  unsigned int* d_counter = nullptr;
  uint32_t* d_map[2] = {nullptr, nullptr};
  bool map_set[2] = {false, false};
  bool maps_global = false;  // grids exceed shared memory: they stay in global memory (solve_kernel<.., kGlobalMaps>)
  float proved_cell[2] = {0.0f, 0.0f}, proved_ox[2] = {0.0f, 0.0f}, proved_oy[2] = {0.0f, 0.0f};  // division proofs cached
  bool proved[2] = {false, false};
  unsigned long long fastdiv_mismatches[2] = {0, 0};
  bool tiny_quotient_ok[2] = {false, false};  // check_tiny_quotient_kernel found no mismatch for this slot
  float proved_wheelbase = -1.0f, wheelbase_rcp = 1.0f;
  bool wheelbase_exact = false;
  // host-call staging (mppi_solve_host)
  float* h_pinned = nullptr;  // state | refpath | action_seq | state_seq
  float* d_stage = nullptr;
  cudaStream_t own_stream = nullptr;
  // top samples: select tree candidates (ping-pong), then the full-sort fallback for n > kTopMax
  float* d_cand_cost[2] = {nullptr, nullptr};
  int* d_cand_id[2] = {nullptr, nullptr};
  size_t cand_cap = 0;
  float* d_win_cost = nullptr;
  int* d_win_id = nullptr;
  RasterShape* d_shapes = nullptr;
  size_t shapes_cap = 0;
  int* d_idx_in = nullptr;
  int* d_idx_out = nullptr;
  float* d_keys_out = nullptr;
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  // bookkeeping
  uint64_t solve_count = 0;  // index of the NEXT solve
  bool solved = false;
  const float* last_noise = nullptr;
  int last_launches = 0;
  int last_epilogue_launches = 0;
  // fused peer exchange (mppi_p2p_*)
  float* d_mailbox = nullptr;
  float* d_gather_scratch = nullptr;
  int* d_error_flag = nullptr;
  int p2p_world = 0, p2p_rank = 0;
  unsigned barrier_seq = 0;
  float* peer_mailbox[kMaxPeers] = {};
  bool peer_opened[kMaxPeers] = {};
  const float* inline_state = nullptr;  // set for the duration of a host-call solve
  const float* inline_ref = nullptr;
  unsigned* host_done = nullptr;  // set for the duration of a host-call solve: completion word in h_pinned
  unsigned host_done_seq = 0;
  bool timing = false;
  unsigned long long* d_trace = nullptr;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
};

namespace {

size_t pad16(size_t bytes) { return (bytes + 15) / 16 * 16; }

// Launch geometry for the rollout kernel. The path is latency bound at the
// sizes MPPI runs (K/32 warps spread over 592 schedulers), so the model is:
// time ~ waves * w / ipc(w) with w = max warps on one scheduler.
int pick_block(const MppiHandle* h, int n_maps, const unsigned* map_bytes, unsigned pa_bytes, int spt) {
  const long long K = (h->cfg.num_samples + spt - 1) / spt;  // threads needed
  int best = 0;
  double best_t = 1e300;
  const int cands1[] = {512, 256, 128, 64}, cands2[] = {256, 128, 64, 0};  // paired kernel: <= 256 threads
  const int* cands = spt == 2 ? cands2 : cands1;
  for (int ci = 0; ci < 4; ++ci) {
    const int bs = cands[ci];
    if (bs == 0) continue;
    SmemLayout L = make_layout(n_maps, map_bytes, h->cfg.horizon, h->E_pad, pa_bytes, h->mi.refpath, bs / 32,
                               h->mi.tail_per_step, spt, 0);
    if (L.total > kMaxSmem) continue;
    long long blocks = (K + bs - 1) / bs;
    const long long regs = spt == 2 ? 200 : 128;
    int per_sm = (int)std::min<long long>({2048 / bs, (long long)(kMaxSmem / L.total), 65536 / (regs * bs)});
    if (per_sm < 1) per_sm = 1;
    long long conc = (long long)kNumSMs * per_sm;
    long long waves = (blocks + conc - 1) / conc;
    long long res_blocks = std::min<long long>(per_sm, (blocks + kNumSMs - 1) / kNumSMs);
    double w = std::max(1.0, ceil(res_blocks * (bs / 32) / 4.0));
    double ipc = std::min(0.9, 0.28 * pow(w, 0.7));
    double t = waves * w / ipc;
    if (t < best_t * 0.999 || best == 0) {
      best_t = t;
      best = bs;
    }
  }
  return best;
}

// Two samples per thread (packed fp32) or one? Pass 1 is bound by the slower of (a) one warp's dependent chain
// and (b) the issue slots of the warps that share a scheduler. Measured on the racing model (B200, round 2):
// 277 vs 184 SASS instructions per warp and time step; a lone warp sustains ~0.34 instructions per cycle, a
// scheduler ~0.65 (paired: every packed instruction holds the FMA pipe two cycles) / ~0.75 (single).
// K = 65536: 2 paired warps per scheduler (852 cycles per step) beat 4 single ones (981); at K <= ~49k - a shard
// of a multi-GPU solve, the K = 4000 of the reference's examples - the single loop's shorter chain wins.
bool paired_loop_pays(long long K) {
  auto cycles_per_step = [&](int spt) {
    const double warps = ceil((double)K / spt / 32.0);
    const double w = std::max(1.0, ceil(warps / ((kNumSMs - 1) * 4.0)));  // warps on the fullest scheduler
    const double instr = spt == 2 ? 277.0 : 184.0, ipc_chain = 0.34, ipc_sched = spt == 2 ? 0.65 : 0.75;
    return std::max(instr / ipc_chain, w * instr / ipc_sched);
  };
  return cycles_per_step(2) <= cycles_per_step(1);
}

template <class M>
int launch_solve(MppiHandle* h, const SolveParams& p, int mode, bool inject, cudaStream_t st) {
  const MppiHandle::Geometry& g = h->geo[inject ? 1 : 0];
  void (*k)(SolveParams) = nullptr;
#define PICK(MODE)                                                                   \
  do {                                                                               \
    if (h->maps_global) {                                                            \
      if constexpr (M::kMaps > 0) {                                                  \
        k = inject ? (void (*)(SolveParams))solve_kernel<M, true, MODE, 1, true>     \
                   : (void (*)(SolveParams))solve_kernel<M, false, MODE, 1, true>;   \
      }                                                                              \
    } else if (inject) {                                                             \
      k = (void (*)(SolveParams))solve_kernel<M, true, MODE, 1>;                     \
    } else if (g.spt == 2) {                                                         \
      if constexpr (M::kHasBounded) k = (void (*)(SolveParams))solve_kernel<M, false, MODE, 2>; \
    } else {                                                                         \
      k = (void (*)(SolveParams))solve_kernel<M, false, MODE, 1>;                    \
    }                                                                                \
  } while (0)
  if (mode == kFused)
    PICK(kFused);
  else if (mode == kCosts)
    PICK(kCosts);
  else
    PICK(kReduce);
#undef PICK
  if (!k) return fail(MPPI_ERR_STATE, "no kernel for this launch geometry");
  {  // opt in to the large dynamic shared memory once per kernel instantiation and device, and again only if
     // it GROWS: the attribute belongs to the function (per device), not to a handle - a second handle with a
     // smaller layout must not lower the limit under a first one (its launches would fail: invalid argument)
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, unsigned> opted;
    std::lock_guard<std::mutex> lock(mu);
    unsigned& have = opted[{h->device, (const void*)k}];
    if (have < g.smem && g.smem > 48u * 1024u) {
      CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
      have = g.smem;
    }
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->timing && mode != kReduce) {
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, st));
  }
  // kFused / kReduce: block 0 is the finisher, blocks 1 .. g.grid are the workers
  k<<<g.grid + (mode != kCosts ? 1 : 0), g.block, g.smem, st>>>(p);
  if (e0) {
    CUDA_TRY(cudaEventRecord(e1, st));
    h->events.emplace_back(e0, e1);
  }
  CUDA_TRY(cudaGetLastError());
  h->last_launches++;
  return MPPI_OK;
}

int dispatch_solve(MppiHandle* h, const SolveParams& p, int mode, bool inject, cudaStream_t st) {
  switch (h->cfg.model) {
    case MPPI_MODEL_PENDULUM: return launch_solve<Pendulum>(h, p, mode, inject, st);
    case MPPI_MODEL_CARTPOLE: return launch_solve<Cartpole>(h, p, mode, inject, st);
    case MPPI_MODEL_MOUNTAINCAR: return launch_solve<MountainCar>(h, p, mode, inject, st);
    case MPPI_MODEL_NAVIGATION2D: return launch_solve<Navigation2D>(h, p, mode, inject, st);
    case MPPI_MODEL_RACING: return launch_solve<Racing>(h, p, mode, inject, st);
    case MPPI_MODEL_CARTPOLE_CONTINUOUS: return launch_solve<CartpoleContinuous>(h, p, mode, inject, st);
    case MPPI_MODEL_GOAL_IN_DANGER_ZONE: return launch_solve<GoalInDangerZone>(h, p, mode, inject, st);
  }
  return fail(MPPI_ERR_INVALID, "unknown model %d", h->cfg.model);
}

template <class M>
int launch_finish(MppiHandle* h, const SolveParams& p, const float* parts, int n, cudaStream_t st) {
  unsigned sm = finish_scratch_bytes(h->E_pad, h->cfg.horizon, h->mi.tail_per_step);
  finish_kernel<M><<<1, 256, sm, st>>>(p, parts, n);
  CUDA_TRY(cudaGetLastError());
  h->last_launches++;
  return MPPI_OK;
}

int dispatch_finish(MppiHandle* h, const SolveParams& p, const float* parts, int n, cudaStream_t st) {
  switch (h->cfg.model) {
    case MPPI_MODEL_PENDULUM: return launch_finish<Pendulum>(h, p, parts, n, st);
    case MPPI_MODEL_CARTPOLE: return launch_finish<Cartpole>(h, p, parts, n, st);
    case MPPI_MODEL_MOUNTAINCAR: return launch_finish<MountainCar>(h, p, parts, n, st);
    case MPPI_MODEL_NAVIGATION2D: return launch_finish<Navigation2D>(h, p, parts, n, st);
    case MPPI_MODEL_RACING: return launch_finish<Racing>(h, p, parts, n, st);
    case MPPI_MODEL_CARTPOLE_CONTINUOUS: return launch_finish<CartpoleContinuous>(h, p, parts, n, st);
    case MPPI_MODEL_GOAL_IN_DANGER_ZONE: return launch_finish<GoalInDangerZone>(h, p, parts, n, st);
  }
  return fail(MPPI_ERR_INVALID, "unknown model %d", h->cfg.model);
}

int launch_search(MppiHandle* h, const float* costs, long long n, cudaStream_t st) {
  SearchParams q{};
  q.costs = costs;
  q.n = n;
  q.mode = h->cfg.lambda_mode;
  q.lambda_min = h->cfg.lambda_min;
  q.lambda_max = h->cfg.lambda_max;
  q.lbps_delta = h->cfg.lbps_delta;
  q.essps_target = h->cfg.essps_target_ess;
  q.sc = h->d_sc;
  cudaLaunchConfig_t cfg{};
  // costs in registers for the whole search when they fit (kSearchPerThread per thread), else re-read from L2
  const bool in_regs = n <= (long long)kSearchCluster * kSearchThreadsReg * kSearchPerThread;
  cfg.gridDim = dim3(kSearchCluster);
  cfg.blockDim = dim3(in_regs ? kSearchThreadsReg : kSearchThreadsGlobal);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kSearchCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (in_regs)
    CUDA_TRY(cudaLaunchKernelEx(&cfg, lambda_search_kernel<kSearchPerThread>, q));
  else
    CUDA_TRY(cudaLaunchKernelEx(&cfg, lambda_search_kernel<0>, q));
  CUDA_TRY(cudaGetLastError());
  h->last_launches++;
  return MPPI_OK;
}

int check_ready(MppiHandle* h) {
  if (!h) return fail(MPPI_ERR_INVALID, "null handle");
  for (int i = 0; i < h->mi.maps; ++i)
    if (!h->map_set[i]) return fail(MPPI_ERR_STATE, "occupancy map slot %d has not been set (mppi_set_map)", i);
  return MPPI_OK;
}

// Per-solve parameter block: base + this call's pointers + sampler counter.
int make_params(MppiHandle* h, const float* d_state, const float* d_refpath, const float* d_noise, float* d_action,
                float* d_state_seq, int n_shards, SolveParams* out, bool use_p2p = false) {
  if (!d_state) return fail(MPPI_ERR_INVALID, "state is null");
  if (h->mi.refpath && !d_refpath) return fail(MPPI_ERR_INVALID, "this model needs a reference path [T+1,4]");
  SolveParams p = h->base;
  p.state = d_state;
  p.refpath = d_refpath;
  p.ref_bulk_ok = d_refpath && (((uintptr_t)d_refpath & 15) == 0);
  p.noise = d_noise;
  p.stage_bytes = h->geo[d_noise ? 1 : 0].stage_bytes;
  p.action_out = d_action;
  p.state_seq_out = d_state_seq;
  p.key.solve_lo = (uint32_t)h->solve_count;
  p.key.solve_hi = (uint32_t)(h->solve_count >> 32);
  p.n_shards = n_shards;
  p.trace = h->d_trace;
  p.p2p_world = 0;
  if (n_shards > 1 && h->p2p_world > 1 && use_p2p) {
    p.p2p_world = h->p2p_world;
    p.p2p_rank = h->p2p_rank;
    p.p2p_seq = (unsigned)(h->solve_count + 1);
    for (int r = 0; r < h->p2p_world; ++r) p.peer_mailbox[r] = h->peer_mailbox[r];
    p.gather_scratch = h->d_gather_scratch;
    p.error_flag = h->d_error_flag;
  }
  p.host_done = h->host_done;
  p.host_done_seq = h->host_done_seq;
  p.inline_inputs = 0;
  if (h->inline_state) {
    p.inline_inputs = 1;
    for (int i = 0; i < h->mi.ds; ++i) p.state_inline[i] = h->inline_state[i];
    if (h->inline_ref) memcpy(p.ref_inline, h->inline_ref, (size_t)(h->cfg.horizon + 1) * 16);
  }
  *out = p;
  return MPPI_OK;
}

template <class M>
int launch_reroll(MppiHandle* h, const SolveParams& p, int n, float* traj, float* w, cudaStream_t st) {
  if (p.noise)
    reroll_kernel<M, true><<<(n + 127) / 128, 128, 0, st>>>(p, h->d_idx_out, n, traj, w);
  else
    reroll_kernel<M, false><<<(n + 127) / 128, 128, 0, st>>>(p, h->d_idx_out, n, traj, w);
  CUDA_TRY(cudaGetLastError());
  return MPPI_OK;
}

template <class M>
int launch_rollout_actions(MppiHandle* h, const SolveParams& p, const float* actions, int n, float* traj,
                           cudaStream_t st) {
  rollout_actions_kernel<M><<<(n + 127) / 128, 128, 0, st>>>(p, actions, n, traj);
  CUDA_TRY(cudaGetLastError());
  return MPPI_OK;
}

// Exhaustively prove (cell-size style) that x / c == the 3-instruction exact form for this divisor.
bool prove_exact_division(float c, float* rcp_out) {
  const float rcp = (float)(1.0 / (double)c);
  *rcp_out = rcp;
  unsigned long long* d_bad = nullptr;
  unsigned long long bad = 1;
  if (cudaMalloc((void**)&d_bad, 8) == cudaSuccess) {
    cudaMemset(d_bad, 0, 8);
    check_fastdiv_kernel<<<148 * 8, 256>>>(c, rcp, d_bad);
    if (cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost) != cudaSuccess) bad = 1;
    cudaFree(d_bad);
  }
  return bad == 0;
}

// every position inside the dynamics' clamp box maps to a cell index in [0, W] x [0, H] (the bordered
// grid then answers without bounds logic); same fp32 arithmetic as the device (IEEE divide, add, rint)
bool box_maps_inside(const SolveParams& b, int slot, float x_lo, float x_hi, float y_lo, float y_hi) {
  auto cell = [&](float x, float o) { return (long long)nearbyintf(x / b.map_cell[slot] + o); };
  return cell(x_lo, b.map_ox[slot]) >= 0 && cell(x_hi, b.map_ox[slot]) <= b.map_W[slot] &&
         cell(y_lo, b.map_oy[slot]) >= 0 && cell(y_hi, b.map_oy[slot]) <= b.map_H[slot] && x_lo <= x_hi && y_lo <= y_hi;
}

void refresh_model_flags(MppiHandle* h) {
  SolveParams& b = h->base;
  int flags = 0;
  const float* v = b.mp.v;
  // solver control bounds inside the env's own clamp (both models: params 0..3 = lo0 hi0 lo1 hi1)
  const bool clamp_redundant = b.u_min[0] >= v[0] && b.u_max[0] <= v[1] && b.u_min[1] >= v[2] && b.u_max[1] <= v[3];
  if (h->cfg.model == MPPI_MODEL_RACING) {
    if (h->map_set[0] && h->map_set[1] && b.map_W[0] == b.map_W[1] && b.map_H[0] == b.map_H[1] &&
        b.map_cell[0] == b.map_cell[1] && b.map_ox[0] == b.map_ox[1] && b.map_oy[0] == b.map_oy[1] &&
        b.map_fastdiv[0] == b.map_fastdiv[1])
      flags |= kFlagSameMapGeometry;
    // wheelbase division: v[17] holds RN(1/L); proven per distinct L (cached), trivially exact for L == 1
    if (v[4] != h->proved_wheelbase) {
      float rcp = 1.0f;
      h->wheelbase_exact = (v[4] == 1.0f) ? true : prove_exact_division(v[4], &rcp);
      if (v[4] == 1.0f) rcp = 1.0f;
      h->wheelbase_rcp = rcp;
      h->proved_wheelbase = v[4];
    }
    b.mp.v[17] = h->wheelbase_rcp;
    if (h->wheelbase_exact) flags |= kFlagUnitWheelbase;  // "exact division available"
    // bounded-variant preconditions (fp64, with margin): |steer| <= 0.78 and |v_max tan(s) / L dt| < 6
    const double smax = std::max(fabs((double)v[2]), fabs((double)v[3]));
    if (smax <= 0.78 && v[4] > 0.0f) {
      const double yaw = fabs((double)v[5]) * tan(smax) / (double)v[4] * fabs((double)v[10]);
      if (yaw < 6.0 && clamp_redundant && (flags & kFlagSameMapGeometry) &&
          box_maps_inside(b, 0, v[6], v[7], v[8], v[9]) && b.map_fastdiv[0] && h->wheelbase_exact &&
          h->tiny_quotient_ok[0] && fabsf(b.map_ox[0]) >= 1e-20f && fabsf(b.map_oy[0]) >= 1e-20f)
        flags |= kFlagBounded;
    }
    if (v[4] == 1.0f) flags |= kFlagUnitL;
  } else if (h->cfg.model == MPPI_MODEL_NAVIGATION2D) {
    const double wmax = std::max(fabs((double)v[2]), fabs((double)v[3]));
    if (wmax * fabs((double)v[10]) < 6.0 && clamp_redundant && h->map_set[0] &&
        box_maps_inside(b, 0, v[6], v[7], v[8], v[9]) && b.map_fastdiv[0] && h->tiny_quotient_ok[0] &&
        fabsf(b.map_ox[0]) >= 1e-20f && fabsf(b.map_oy[0]) >= 1e-20f)
      flags |= kFlagBounded;
  }
  b.mp.flags = flags;
}

void refresh_launch_geometry(MppiHandle* h) {
  unsigned mb[2] = {h->base.map_bytes[0], h->base.map_bytes[1]};
  {  // grids that do not fit beside the smallest block's buffers stay in global memory (general loop only)
    SmemLayout Lmin = make_layout(h->mi.maps, mb, h->cfg.horizon, h->E_pad, h->base.prev_action_bytes, h->mi.refpath, 2,
                                  h->mi.tail_per_step, 1, 0);
    h->maps_global = h->mi.maps > 0 && Lmin.total > kMaxSmem;
    if (h->maps_global) {
      mb[0] = mb[1] = 0;
      h->base.mp.flags &= ~kFlagBounded;  // the bounded loops read the staged grids
    }
  }
  const bool pair_model = h->cfg.model == MPPI_MODEL_RACING || h->cfg.model == MPPI_MODEL_NAVIGATION2D;
  for (int v = 0; v < 2; ++v) {
    MppiHandle::Geometry& g = h->geo[v];
    // two samples per thread: the model's bounded loop is available (host-verified flags) and the noise
    // comes from the in-kernel sampler; the kernel re-checks the solve's initial state
    const bool want_pair = (h->cfg.flags & MPPI_CFG_FORCE_PAIRED) ||
                           (!(h->cfg.flags & MPPI_CFG_FORCE_SINGLE) && paired_loop_pays(h->cfg.num_samples));
    g.spt = (v == 0 && pair_model && (h->base.mp.flags & kFlagBounded) && want_pair) ? 2 : 1;
    int bs = h->cfg.block_size > 0 ? h->cfg.block_size : pick_block(h, h->mi.maps, mb, h->base.prev_action_bytes, g.spt);
    if (bs <= 0) bs = 64;  // nothing fits (oversized grids): the shared-memory check below reports it
    if (g.spt == 2 && bs > 256) bs = 256;
    g.block = bs;
    const long long per_block = (long long)bs * g.spt;
    g.grid = (int)((h->cfg.num_samples + per_block - 1) / per_block);
    // landing zone for the block partials in the last block: all of them if that fits beside the rest
    const unsigned want = (unsigned)std::min<long long>((long long)g.grid * h->P * 4, (long long)kMaxSmem);
    SmemLayout L = make_layout(h->mi.maps, mb, h->cfg.horizon, h->E_pad, h->base.prev_action_bytes, h->mi.refpath,
                               bs / 32, h->mi.tail_per_step, g.spt, want);
    g.stage_bytes = want;
    if (L.total > kMaxSmem) {  // does not fit: combine straight from global memory
      g.stage_bytes = 0;
      L = make_layout(h->mi.maps, mb, h->cfg.horizon, h->E_pad, h->base.prev_action_bytes, h->mi.refpath, bs / 32,
                      h->mi.tail_per_step, g.spt, 0);
    }
    g.smem = L.total;
  }
}

}  // namespace

// ================================ C ABI ========================================
extern "C" {

const char* mppi_last_error(void) { return g_last_error.c_str(); }
int mppi_abi_version(void) { return MPPI_ABI_VERSION; }

void mppi_philox4x32_10(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
  Philox::block(counter[0], counter[1], counter[2], counter[3], key[0], key[1], out);
}

int mppi_create(const MppiConfig* cfg, MppiHandle** out) {
  if (!cfg || !out) return fail(MPPI_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->abi_version != MPPI_ABI_VERSION)
    return fail(MPPI_ERR_INVALID, "abi_version %d != %d", cfg->abi_version, MPPI_ABI_VERSION);
  ModelInfo mi = model_info(cfg->model);
  if (mi.ds == 0) return fail(MPPI_ERR_INVALID, "unknown model %d", cfg->model);
  if (cfg->dim_state != mi.ds || cfg->dim_control != mi.du)
    return fail(MPPI_ERR_INVALID, "model %d has dim_state=%d dim_control=%d, got %d/%d", cfg->model, mi.ds, mi.du,
                cfg->dim_state, cfg->dim_control);
  if (cfg->horizon < 1 || cfg->num_samples < 1) return fail(MPPI_ERR_INVALID, "horizon and num_samples must be >= 1");
  if ((long long)cfg->horizon * mi.du > 4096) return fail(MPPI_ERR_UNSUPPORTED, "horizon * dim_control > 4096");
  if (cfg->lambda_mode < MPPI_LAMBDA_FIXED || cfg->lambda_mode > MPPI_LAMBDA_ESSPS)
    return fail(MPPI_ERR_INVALID, "lambda_ must be 'MPO', 'LBPS', 'ESSPS', or a float value.");
  if (cfg->lambda_mode == MPPI_LAMBDA_FIXED && !(cfg->lambda_ > 0.0))
    return fail(MPPI_ERR_INVALID, "lambda_ must be positive");
  if (cfg->num_model_params != mi.n_params)
    return fail(MPPI_ERR_INVALID, "model %d takes %d parameters, got %d", cfg->model, mi.n_params,
                cfg->num_model_params);
  if (cfg->use_sg_filter) {
    if (cfg->sg_window_size % 2 == 0 || cfg->sg_window_size <= cfg->sg_poly_order)
      return fail(MPPI_ERR_INVALID, "window_size must be odd and greater than poly_order.");  // mppi.py:580-581
    if (cfg->sg_window_size > MPPI_MAX_SG_WINDOW) return fail(MPPI_ERR_UNSUPPORTED, "sg_window_size > %d", MPPI_MAX_SG_WINDOW);
    if (2 * cfg->horizon - 1 < cfg->sg_window_size / 2)
      return fail(MPPI_ERR_INVALID, "horizon too short for the Savitzky-Golay window");
  }
  if (cfg->block_size != 0 && (cfg->block_size < 64 || cfg->block_size > 512 || cfg->block_size % 32))
    return fail(MPPI_ERR_INVALID, "block_size must be 0 or a multiple of 32 in [64, 512]");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(MPPI_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(MPPI_ERR_INVALID, "device %d out of range", cfg->device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(MPPI_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device,
                prop.major, prop.minor);
  ON_DEVICE(cfg->device);

  MppiHandle* h = new (std::nothrow) MppiHandle();
  if (!h) return fail(MPPI_ERR_INVALID, "out of host memory");
  h->cfg = *cfg;
  h->mi = mi;
  h->device = cfg->device;
  if (h->cfg.total_samples <= 0) {
    h->cfg.total_samples = cfg->num_samples;
    h->cfg.sample_offset = 0;
  }
  if (h->cfg.essps_target_ess <= 0.0) h->cfg.essps_target_ess = (double)h->cfg.total_samples / 10.0;  // mppi.py:185-187
  const int T = cfg->horizon, DU = mi.du, DS = mi.ds, K = cfg->num_samples;
  h->E = T * DU;
  h->E_pad = (h->E + 31) / 32 * 32;
  h->P = kPartialHeader + h->E_pad;

  auto cleanup = [&](int rc) {
    mppi_destroy(h);
    return rc;
  };
#define ALLOC(ptr, bytes)                                                                                   \
  do {                                                                                                      \
    cudaError_t _e = cudaMalloc((void**)&(ptr), pad16(bytes) ? pad16(bytes) : 16);                         \
    if (_e != cudaSuccess) return cleanup(fail(MPPI_ERR_CUDA, "cudaMalloc(%zu): %s", (size_t)(bytes),      \
                                               cudaGetErrorString(_e)));                                    \
    cudaMemset((ptr), 0, pad16(bytes) ? pad16(bytes) : 16);                                                 \
  } while (0)
  ALLOC(h->d_prev_action, (size_t)h->E_pad * 4);
  ALLOC(h->d_history, (size_t)std::max(1, (T - 1) * DU) * 4);
  ALLOC(h->d_nominal_snapshot, (size_t)h->E_pad * 4);
  ALLOC(h->d_state_snapshot, 64);
  ALLOC(h->d_sc, sizeof(DeviceScalars));
  ALLOC(h->d_costs, (size_t)K * 4);
  ALLOC(h->d_rank_partial, (size_t)h->P * 4);
  ALLOC(h->d_counter, 16);
  ALLOC(h->d_dry, dry_scratch_floats(h->E_pad, T, DS, DU) * 4);
  // staging for mppi_solve_host: state | refpath | action | state_seq
  size_t stage_floats = 8 + (size_t)(T + 1) * 4 + (size_t)h->E_pad + (size_t)(T + 1) * DS + 8 + 16;  // + done word
  ALLOC(h->d_stage, stage_floats * 4);
  if (cudaHostAlloc((void**)&h->h_pinned, stage_floats * 4, cudaHostAllocMapped) != cudaSuccess)
    return cleanup(fail(MPPI_ERR_CUDA, "cudaMallocHost failed"));
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess)
    return cleanup(fail(MPPI_ERR_CUDA, "cudaStreamCreate failed"));

  DeviceScalars sc{};
  sc.lambda = (cfg->lambda_mode == MPPI_LAMBDA_FIXED) ? cfg->lambda_ : 1.0;  // MPO starts at 1.0 (mppi.py:193)
  sc.lambda_used = sc.lambda;
  sc.rho = 0.0f;  // log(1.0)
  if (cudaMemcpy(h->d_sc, &sc, sizeof sc, cudaMemcpyHostToDevice) != cudaSuccess)
    return cleanup(fail(MPPI_ERR_CUDA, "cudaMemcpy(scalars) failed"));

  SolveParams& b = h->base;
  b.K = K;
  b.T = T;
  b.k_offset = h->cfg.sample_offset;
  // threshold = int(num_samples * (1 - exploration)) on the GLOBAL sample count (mppi.py:266)
  b.explore_threshold = (long long)((double)h->cfg.total_samples * (1.0 - cfg->exploration));
  for (int d = 0; d < DU; ++d) {
    b.u_min[d] = cfg->u_min[d];
    b.u_max[d] = cfg->u_max[d];
    b.sigma[d] = cfg->sigmas[d];
  }
  for (int i = 0; i < mi.n_params; ++i) b.mp.v[i] = cfg->model_params[i];
  refresh_model_flags(h);
  b.prev_action = h->d_prev_action;
  b.prev_action_bytes = (unsigned)pad16((size_t)h->E_pad * 4);
  b.history = h->d_history;
  b.nominal_snapshot = h->d_nominal_snapshot;
  b.state_snapshot = h->d_state_snapshot;
  b.sc = h->d_sc;
  b.costs = h->d_costs;
  b.rank_partial = h->d_rank_partial;
  b.counter = h->d_counter;
  b.dry_scratch = h->d_dry;
  b.key.seed_lo = (uint32_t)cfg->seed;
  b.key.seed_hi = (uint32_t)(cfg->seed >> 32);
  b.lambda_mode = cfg->lambda_mode;
  b.mpo_epsilon = 0.1f;  // mppi.py:194
  b.use_sg = cfg->use_sg_filter ? 1 : 0;
  b.sg_window = cfg->sg_window_size;
  if (cfg->use_sg_filter) {
    if (cfg->sg_coeffs_given) {
      for (int i = 0; i < cfg->sg_window_size; ++i) b.sg_coeffs[i] = cfg->sg_coeffs[i];
    } else if (!savgol_coeffs(cfg->sg_window_size, cfg->sg_poly_order, b.sg_coeffs)) {
      return cleanup(fail(MPPI_ERR_INVALID, "singular Savitzky-Golay system"));
    }
  }
  b.E = h->E;
  b.E_pad = h->E_pad;
  b.P = h->P;
  refresh_launch_geometry(h);
  if (std::max(h->geo[0].smem, h->geo[1].smem) > kMaxSmem)
    return cleanup(fail(MPPI_ERR_UNSUPPORTED, "shared memory budget exceeded (%u B)",
                        std::max(h->geo[0].smem, h->geo[1].smem)));
  // block partials: sized for the smallest block the engine may pick later (maps change smem)
  ALLOC(h->d_block_partials, (size_t)((K + 63) / 64) * h->P * 4);
  b.block_partials = h->d_block_partials;
#undef ALLOC
  *out = h;
  return MPPI_OK;
}

void mppi_destroy(MppiHandle* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  for (auto& ev : h->events) {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  cudaFree(h->d_prev_action);
  cudaFree(h->d_history);
  cudaFree(h->d_nominal_snapshot);
  cudaFree(h->d_state_snapshot);
  cudaFree(h->d_sc);
  cudaFree(h->d_costs);
  cudaFree(h->d_block_partials);
  cudaFree(h->d_rank_partial);
  cudaFree(h->d_counter);
  cudaFree(h->d_dry);
  cudaFree(h->d_map[0]);
  cudaFree(h->d_map[1]);
  cudaFree(h->d_stage);
  for (int i = 0; i < 2; ++i) {
    cudaFree(h->d_cand_cost[i]);
    cudaFree(h->d_cand_id[i]);
  }
  cudaFree(h->d_win_cost);
  cudaFree(h->d_win_id);
  cudaFree(h->d_shapes);
  cudaFree(h->d_idx_in);
  cudaFree(h->d_idx_out);
  cudaFree(h->d_keys_out);
  cudaFree(h->d_sort_tmp);
  cudaFree(h->d_trace);
  for (int r = 0; r < kMaxPeers; ++r)
    if (h->peer_opened[r]) cudaIpcCloseMemHandle(h->peer_mailbox[r]);
  cudaFree(h->d_mailbox);
  cudaFree(h->d_gather_scratch);
  if (h->d_error_flag) cudaFreeHost(h->d_error_flag);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

int mppi_reset(MppiHandle* h, void* stream) {
  if (!h) return fail(MPPI_ERR_INVALID, "null handle");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(h->d_prev_action, 0, (size_t)h->E_pad * 4, st));
  CUDA_TRY(cudaMemsetAsync(h->d_history, 0, (size_t)std::max(1, (h->cfg.horizon - 1) * h->mi.du) * 4, st));
  return MPPI_OK;
}

int mppi_set_model_params(MppiHandle* h, const float* params, int32_t n) {
  if (!h || (!params && n > 0)) return fail(MPPI_ERR_INVALID, "null argument");
  if (n != h->mi.n_params) return fail(MPPI_ERR_INVALID, "model takes %d parameters, got %d", h->mi.n_params, n);
  ON_DEVICE(h->device);
  for (int i = 0; i < n; ++i) h->base.mp.v[i] = params[i];
  refresh_model_flags(h);
  refresh_launch_geometry(h);  // the flags decide between one and two samples per thread
  return MPPI_OK;
}

// (Re)allocate the packed grid of `slot` for a W x H map; *words / *bytes describe the bordered layout.
static int alloc_map_bits(MppiHandle* h, int slot, int W, int H, int* words, size_t* bytes) {
  *words = (H + 1 + 31) / 32;  // + the out-of-bounds border bit (see MapView)
  *bytes = pad16((size_t)(W + 1) * *words * 4);
  if (h->d_map[slot] && h->base.map_bytes[slot] == (unsigned)*bytes && h->base.map_W[slot] == W &&
      h->base.map_H[slot] == H)
    return MPPI_OK;  // same geometry as before (dynamic obstacles re-rasterised every control step): reuse
  cudaFree(h->d_map[slot]);
  h->d_map[slot] = nullptr;
  cudaError_t e = cudaMalloc((void**)&h->d_map[slot], *bytes);
  if (e == cudaSuccess) e = cudaMemset(h->d_map[slot], 0, *bytes);
  if (e != cudaSuccess) return fail(MPPI_ERR_CUDA, "map alloc: %s", cudaGetErrorString(e));
  return MPPI_OK;
}

// Everything after the packed bits of `slot` are in place: geometry, division proofs, flags, launch geometry.
static int finish_map_setup(MppiHandle* h, int slot, int W, int H, int words, size_t bytes, float cell, float ox,
                            float oy) {
  SolveParams& b = h->base;
  b.map_bits[slot] = h->d_map[slot];
  b.map_W[slot] = W;
  b.map_H[slot] = H;
  b.map_words[slot] = words;
  b.map_bytes[slot] = (unsigned)bytes;
  b.map_cell[slot] = cell;
  // exact fast division by the cell size: prove it for this divisor, or keep the true division. The proofs
  // depend on (cell, origin) only, so a re-rasterised map of the same geometry keeps them.
  if (!(h->proved[slot] && h->proved_cell[slot] == cell && h->proved_ox[slot] == ox && h->proved_oy[slot] == oy)) {
    float rcp = 0.0f;
    const bool ok = prove_exact_division(cell, &rcp);
    b.map_rcp[slot] = rcp;
    b.map_fastdiv[slot] = ok ? 1 : 0;
    h->fastdiv_mismatches[slot] = ok ? 0 : 1;
    // the paired loop's cell index carries no |x| < 1e-30 guard: prove it for this (cell, origin)
    h->tiny_quotient_ok[slot] = false;
    unsigned long long* d_bad = nullptr;
    unsigned long long bad = 1;
    if (ok && cudaMalloc((void**)&d_bad, 8) == cudaSuccess) {
      cudaMemset(d_bad, 0, 8);
      check_tiny_quotient_kernel<<<148 * 4, 256>>>(cell, rcp, ox, oy, d_bad);
      if (cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost) != cudaSuccess) bad = 1;
      cudaFree(d_bad);
      h->tiny_quotient_ok[slot] = bad == 0;
    }
    h->proved[slot] = true;
    h->proved_cell[slot] = cell;
    h->proved_ox[slot] = ox;
    h->proved_oy[slot] = oy;
  }
  b.map_ox[slot] = ox;
  b.map_oy[slot] = oy;
  h->map_set[slot] = true;
  refresh_model_flags(h);
  refresh_launch_geometry(h);
  if (std::max(h->geo[0].smem, h->geo[1].smem) > kMaxSmem)
    return fail(MPPI_ERR_UNSUPPORTED, "shared memory budget exceeded (%u B > %u)",
                std::max(h->geo[0].smem, h->geo[1].smem), kMaxSmem);
  return MPPI_OK;
}

int mppi_set_map(MppiHandle* h, int32_t slot, const float* grid, int32_t on_device, int32_t W, int32_t H,
                 float cell, float ox, float oy) {
  if (!h || !grid) return fail(MPPI_ERR_INVALID, "null argument");
  if (slot < 0 || slot >= h->mi.maps) return fail(MPPI_ERR_INVALID, "model has %d map slots, got slot %d", h->mi.maps, slot);
  if (W < 1 || H < 1 || !(cell > 0.0f)) return fail(MPPI_ERR_INVALID, "bad map geometry");
  ON_DEVICE(h->device);
  const float* d_grid = grid;
  float* tmp = nullptr;
  if (!on_device) {
    CUDA_TRY(cudaMalloc((void**)&tmp, (size_t)W * H * 4));
    cudaError_t e = cudaMemcpy(tmp, grid, (size_t)W * H * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(tmp);
      return fail(MPPI_ERR_CUDA, "map upload: %s", cudaGetErrorString(e));
    }
    d_grid = tmp;
  }
  int words = 0;
  size_t bytes = 0;
  int rc = alloc_map_bits(h, slot, W, H, &words, &bytes);
  if (rc) {
    cudaFree(tmp);
    return rc;
  }
  const long long n = (long long)(W + 1) * words;
  pack_map_kernel<<<(unsigned)((n + 255) / 256), 256>>>(d_grid, W, H, words, h->d_map[slot]);
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(tmp);
  if (e != cudaSuccess) return fail(MPPI_ERR_CUDA, "map pack: %s", cudaGetErrorString(e));
  return finish_map_setup(h, slot, W, H, words, bytes, cell, ox, oy);
}

int mppi_raster_map(MppiHandle* h, int32_t slot, int32_t mode, int32_t W, int32_t H, float cell, float ox, float oy,
                    const int32_t* h_discs, int32_t n_discs, const int32_t* h_rects, int32_t n_rects,
                    float* d_grid_out) {
  if (!h) return fail(MPPI_ERR_INVALID, "null handle");
  if (slot < 0 || slot >= h->mi.maps) return fail(MPPI_ERR_INVALID, "model has %d map slots, got slot %d", h->mi.maps, slot);
  if (W < 1 || H < 1 || !(cell > 0.0f)) return fail(MPPI_ERR_INVALID, "bad map geometry");
  if (mode != MPPI_RASTER_OBSTACLE && mode != MPPI_RASTER_LANE) return fail(MPPI_ERR_INVALID, "unknown raster mode %d", mode);
  if (n_discs < 0 || n_rects < 0 || (n_discs > 0 && !h_discs) || (n_rects > 0 && !h_rects))
    return fail(MPPI_ERR_INVALID, "bad shape list");
  if (mode == MPPI_RASTER_LANE && n_rects > 0) return fail(MPPI_ERR_INVALID, "lane maps are painted from discs only");
  ON_DEVICE(h->device);
  std::vector<RasterShape> shapes((size_t)n_discs + n_rects);
  for (int i = 0; i < n_discs; ++i) {
    if (h_discs[3 * i + 2] < 0) return fail(MPPI_ERR_INVALID, "disc %d has a negative squared radius", i);
    shapes[i] = RasterShape{h_discs[3 * i], h_discs[3 * i + 1], h_discs[3 * i + 2], 0};
  }
  for (int i = 0; i < n_rects; ++i)
    shapes[n_discs + i] = RasterShape{h_rects[4 * i], h_rects[4 * i + 1], h_rects[4 * i + 2], h_rects[4 * i + 3]};
  if (shapes.size() > h->shapes_cap) {
    cudaFree(h->d_shapes);
    h->d_shapes = nullptr;
    h->shapes_cap = 0;
    CUDA_TRY(cudaMalloc((void**)&h->d_shapes, shapes.size() * sizeof(RasterShape)));
    h->shapes_cap = shapes.size();
  }
  if (!shapes.empty())
    CUDA_TRY(cudaMemcpy(h->d_shapes, shapes.data(), shapes.size() * sizeof(RasterShape), cudaMemcpyHostToDevice));
  int words = 0;
  size_t bytes = 0;
  int rc = alloc_map_bits(h, slot, W, H, &words, &bytes);
  if (rc) return rc;
  const long long n = (long long)(W + 1) * words;
  const int threads = 256;
  raster_map_kernel<<<(unsigned)((n + threads - 1) / threads), threads, threads * sizeof(RasterShape)>>>(
      h->d_shapes, n_discs, h->d_shapes + n_discs, n_rects, mode == MPPI_RASTER_LANE ? 1 : 0, W, H, words,
      h->d_map[slot], d_grid_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(MPPI_ERR_CUDA, "map raster: %s", cudaGetErrorString(e));
  return finish_map_setup(h, slot, W, H, words, bytes, cell, ox, oy);
}

static int solve_impl(MppiHandle* h, const float* d_state, const float* d_refpath, const float* d_noise,
                      float* d_action, float* d_state_seq, cudaStream_t st) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!d_action || !d_state_seq) return fail(MPPI_ERR_INVALID, "output pointer is null");
  ON_DEVICE(h->device);
  SolveParams p;
  const bool lam_search = h->cfg.lambda_mode == MPPI_LAMBDA_LBPS || h->cfg.lambda_mode == MPPI_LAMBDA_ESSPS;
  const bool fused_shards = h->p2p_world > 1 && !lam_search;
  if (h->cfg.total_samples != h->cfg.num_samples && !fused_shards)
    return fail(MPPI_ERR_STATE, "this handle owns a shard: use mppi_shard_* (or connect the peers with mppi_p2p_connect)");
  rc = make_params(h, d_state, d_refpath, d_noise, d_action, d_state_seq, fused_shards ? h->p2p_world : 1, &p,
                   fused_shards);
  if (rc) return rc;
  h->last_launches = 0;
  const bool inject = d_noise != nullptr;
  if (lam_search) {
    if ((rc = dispatch_solve(h, p, kCosts, inject, st))) return rc;
    if ((rc = launch_search(h, h->d_costs, h->cfg.num_samples, st))) return rc;
    if ((rc = dispatch_solve(h, p, kReduce, inject, st))) return rc;
  } else {
    if ((rc = dispatch_solve(h, p, kFused, inject, st))) return rc;
  }
  h->last_noise = d_noise;
  h->solve_count++;
  h->solved = true;
  return MPPI_OK;
}

int mppi_solve(MppiHandle* h, const float* d_state, const float* d_refpath, const float* d_noise,
               float* d_action_seq, float* d_state_seq, void* stream) {
  if (!h) return fail(MPPI_ERR_INVALID, "null handle");
  return solve_impl(h, d_state, d_refpath, d_noise, d_action_seq, d_state_seq, (cudaStream_t)stream);
}

int mppi_solve_host(MppiHandle* h, const float* h_state, const float* h_refpath, float* h_action_seq,
                    float* h_state_seq) {
  if (!h || !h_state || !h_action_seq || !h_state_seq) return fail(MPPI_ERR_INVALID, "null argument");
  if (h->mi.refpath && !h_refpath) return fail(MPPI_ERR_INVALID, "this model needs a reference path [T+1,4]");
  ON_DEVICE(h->device);
  const int T = h->cfg.horizon, DS = h->mi.ds;
  const size_t n_state = 8, n_ref = (size_t)(T + 1) * 4, n_act = (size_t)h->E_pad, n_seq = (size_t)(T + 1) * DS;
  float* hp = h->h_pinned;
  cudaStream_t st = h->own_stream;
  // outputs: the finishing block stores straight into the pinned, device-mapped staging buffer
  // (UVA: the pinned host pointer is valid on the device), so no D2H copy is enqueued
  float* d_act = hp + n_state + n_ref;
  float* d_seq = d_act + n_act;
  int rc;
  // completion word: the finisher block stores the call's sequence number into mapped pinned memory after its
  // output stores (system-scope fence in between); the host spins on it - a stream synchronisation costs
  // wake-up latency on a ~60 us solve. A kernel fault never sets it: the bounded spin then falls back to
  // cudaStreamSynchronize, which reports the error.
  volatile unsigned* done = reinterpret_cast<volatile unsigned*>(d_seq + n_seq + 8);
  const unsigned seq = (unsigned)(h->solve_count + 1) | 0x80000000u;
  h->host_done = const_cast<unsigned*>(done);
  h->host_done_seq = seq;
  const bool inline_ok = (!h->mi.refpath || n_ref <= (size_t)kInlineRefFloats) &&
                         !(h->cfg.lambda_mode == MPPI_LAMBDA_LBPS || h->cfg.lambda_mode == MPPI_LAMBDA_ESSPS);
  if (inline_ok) {
    // inputs: state and reference path ride in the kernel parameter block - no H2D copy either
    h->inline_state = h_state;
    h->inline_ref = h->mi.refpath ? h_refpath : nullptr;
    rc = solve_impl(h, h->d_stage, h->mi.refpath ? h->d_stage + n_state : nullptr, nullptr, d_act, d_seq, st);
    h->inline_state = h->inline_ref = nullptr;
  } else {
    memcpy(hp, h_state, DS * 4);
    size_t in_floats = n_state;
    if (h->mi.refpath) {
      memcpy(hp + n_state, h_refpath, n_ref * 4);
      in_floats += n_ref;
    }
    cudaError_t e = cudaMemcpyAsync(h->d_stage, hp, in_floats * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
      h->host_done = nullptr;
      return fail(MPPI_ERR_CUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
    }
    rc = solve_impl(h, h->d_stage, h->mi.refpath ? h->d_stage + n_state : nullptr, nullptr, d_act, d_seq, st);
  }
  h->host_done = nullptr;
  if (rc) return rc;
  {
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (*done != seq) {
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
      if ((++spins & 0x3ffu) == 0 &&
          std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(200)) break;  // slow or faulted kernel
    }
    if (*done != seq) CUDA_TRY(cudaStreamSynchronize(st));  // (also surfaces a kernel fault)
    std::atomic_thread_fence(std::memory_order_acquire);
  }
  memcpy(h_action_seq, d_act, (size_t)h->E * 4);
  memcpy(h_state_seq, d_seq, n_seq * 4);
  return MPPI_OK;
}

// ---- sharded solve ---------------------------------------------------------------------------
int mppi_shard_rollout(MppiHandle* h, const float* d_state, const float* d_refpath, const float* d_noise,
                       void* stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  ON_DEVICE(h->device);
  SolveParams p;
  // outputs are written by mppi_shard_finish; n_shards = 2 only means "stop at the shard partial"
  rc = make_params(h, d_state, d_refpath, d_noise, nullptr, nullptr, 2, &p);
  if (rc) return rc;
  h->last_launches = 0;
  const bool lam_search = h->cfg.lambda_mode == MPPI_LAMBDA_LBPS || h->cfg.lambda_mode == MPPI_LAMBDA_ESSPS;
  rc = dispatch_solve(h, p, lam_search ? kCosts : kFused, d_noise != nullptr, (cudaStream_t)stream);
  if (rc) return rc;
  h->last_noise = d_noise;
  return MPPI_OK;
}

int mppi_shard_lambda(MppiHandle* h, const float* d_costs_all, void* stream) {
  if (!h || !d_costs_all) return fail(MPPI_ERR_INVALID, "null argument");
  if (!(h->cfg.lambda_mode == MPPI_LAMBDA_LBPS || h->cfg.lambda_mode == MPPI_LAMBDA_ESSPS))
    return fail(MPPI_ERR_STATE, "mppi_shard_lambda is for LBPS / ESSPS handles");
  ON_DEVICE(h->device);
  int rc = launch_search(h, d_costs_all, h->cfg.total_samples, (cudaStream_t)stream);
  if (rc) return rc;
  SolveParams p;
  // the reduce pass only touches costs, the warm start and the sampler; state is not read
  rc = make_params(h, h->d_state_snapshot, h->mi.refpath ? (const float*)h->d_state_snapshot : nullptr, h->last_noise,
                   nullptr, nullptr, 2, &p);
  if (rc) return rc;
  return dispatch_solve(h, p, kReduce, h->last_noise != nullptr, (cudaStream_t)stream);
}

int mppi_shard_finish(MppiHandle* h, const float* d_partials, int32_t n_shards, const float* d_state,
                      float* d_action_seq, float* d_state_seq, void* stream) {
  if (!h || !d_partials || !d_state || !d_action_seq || !d_state_seq || n_shards < 1)
    return fail(MPPI_ERR_INVALID, "bad argument");
  ON_DEVICE(h->device);
  SolveParams p;
  int rc = make_params(h, d_state, h->mi.refpath ? d_state : nullptr, nullptr, d_action_seq, d_state_seq, n_shards, &p);
  if (rc) return rc;
  rc = dispatch_finish(h, p, d_partials, n_shards, (cudaStream_t)stream);
  if (rc) return rc;
  h->solve_count++;
  h->solved = true;
  return MPPI_OK;
}

int mppi_p2p_export(MppiHandle* h, uint8_t handle_out[64]) {
  if (!h || !handle_out) return fail(MPPI_ERR_INVALID, "null argument");
  ON_DEVICE(h->device);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!h->d_mailbox) {
    const size_t bytes = mailbox_floats(h->P) * 4;
    CUDA_TRY(cudaMalloc((void**)&h->d_mailbox, bytes));
    CUDA_TRY(cudaMemset(h->d_mailbox, 0, bytes));
    CUDA_TRY(cudaMalloc((void**)&h->d_gather_scratch, (size_t)kMaxPeers * h->P * 4));
    // the timeout flag lives in mapped pinned host memory: the finishing block stores the sequence number of
    // the failed solve there and the host reads it without a copy or a synchronisation (mppi_p2p_status)
    CUDA_TRY(cudaHostAlloc((void**)&h->d_error_flag, 16, cudaHostAllocMapped));
    memset(h->d_error_flag, 0, 16);
    CUDA_TRY(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t ipc;
  CUDA_TRY(cudaIpcGetMemHandle(&ipc, h->d_mailbox));
  memcpy(handle_out, &ipc, 64);
  return MPPI_OK;
}

int mppi_p2p_connect(MppiHandle* h, const uint8_t* handles, int32_t world, int32_t rank) {
  if (!h || !handles) return fail(MPPI_ERR_INVALID, "null argument");
  if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(MPPI_ERR_INVALID, "p2p world must be in [2, %d]", kMaxPeers);
  if (!h->d_mailbox) return fail(MPPI_ERR_STATE, "call mppi_p2p_export first");
  ON_DEVICE(h->device);
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      h->peer_mailbox[r] = h->d_mailbox;
      continue;
    }
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, handles + (size_t)r * 64, 64);
    void* ptr = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->peer_mailbox[r] = (float*)ptr;
    h->peer_opened[r] = true;
  }
  h->p2p_world = world;
  h->p2p_rank = rank;
  return MPPI_OK;
}

int mppi_p2p_connect_local(MppiHandle* h, const uint64_t* mailbox_ptrs, int32_t world, int32_t rank) {
  if (!h || !mailbox_ptrs) return fail(MPPI_ERR_INVALID, "null argument");
  if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(MPPI_ERR_INVALID, "p2p world must be in [2, %d]", kMaxPeers);
  if (!h->d_mailbox || (uint64_t)(uintptr_t)h->d_mailbox != mailbox_ptrs[rank])
    return fail(MPPI_ERR_STATE, "mailbox_ptrs[rank] must be this handle's own mailbox (mppi_p2p_export first)");
  for (int r = 0; r < world; ++r) h->peer_mailbox[r] = (float*)(uintptr_t)mailbox_ptrs[r];
  h->p2p_world = world;
  h->p2p_rank = rank;
  return MPPI_OK;
}

int mppi_p2p_mailbox_ptr(MppiHandle* h, uint64_t* ptr) {
  if (!h || !ptr) return fail(MPPI_ERR_INVALID, "null argument");
  *ptr = (uint64_t)(uintptr_t)h->d_mailbox;
  return MPPI_OK;
}

int mppi_p2p_barrier(MppiHandle* h, void* stream) {
  if (!h) return fail(MPPI_ERR_INVALID, "null handle");
  if (h->p2p_world < 2) return fail(MPPI_ERR_STATE, "no peers connected (mppi_p2p_connect first)");
  ON_DEVICE(h->device);
  BarrierParams b{};
  for (int r = 0; r < h->p2p_world; ++r)
    b.peer_slots[r] = reinterpret_cast<unsigned*>(h->peer_mailbox[r] + mailbox_barrier_offset(h->P));
  b.world = h->p2p_world;
  b.rank = h->p2p_rank;
  b.seq = ++h->barrier_seq;
  b.error_flag = h->d_error_flag;
  p2p_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(b);
  CUDA_TRY(cudaGetLastError());
  return MPPI_OK;
}

int mppi_p2p_status(MppiHandle* h, int32_t* timed_out) {
  if (!h || !timed_out) return fail(MPPI_ERR_INVALID, "null argument");
  *timed_out = 0;
  if (h->d_error_flag) {  // mapped host memory, written by the kernel with a system-scope fence
    volatile int* f = h->d_error_flag;
    *timed_out = *f;  // sequence number (>= 1) of the solve whose exchange timed out, 0 if none
    *f = 0;           // reported once
  }
  return MPPI_OK;
}

int32_t mppi_partial_floats(const MppiHandle* h) { return h ? h->P : 0; }

int mppi_costs_ptr(MppiHandle* h, const float** d_costs) {
  if (!h || !d_costs) return fail(MPPI_ERR_INVALID, "null argument");
  *d_costs = h->d_costs;
  return MPPI_OK;
}
int mppi_partial_ptr(MppiHandle* h, const float** d_partial) {
  if (!h || !d_partial) return fail(MPPI_ERR_INVALID, "null argument");
  *d_partial = h->d_rank_partial;
  return MPPI_OK;
}
int mppi_prev_action_ptr(MppiHandle* h, const float** p) {
  if (!h || !p) return fail(MPPI_ERR_INVALID, "null argument");
  *p = h->d_prev_action;
  return MPPI_OK;
}

// ---- inspection ----------------------------------------------------------------------------------
int mppi_weights(MppiHandle* h, float* d_weights, void* stream) {
  if (!h || !d_weights) return fail(MPPI_ERR_INVALID, "null argument");
  if (!h->solved) return fail(MPPI_ERR_STATE, "no solve has run yet");
  ON_DEVICE(h->device);
  const int K = h->cfg.num_samples;
  weights_kernel<<<(K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->d_costs, K, h->d_sc, d_weights);
  CUDA_TRY(cudaGetLastError());
  return MPPI_OK;
}

}  // extern "C"

// Parameter block of the kernels that re-roll samples of the LAST solve (its state, warm start, sampler index).
static SolveParams last_solve_params(MppiHandle* h) {
  SolveParams p = h->base;
  p.state = h->d_state_snapshot;
  p.prev_action = h->d_nominal_snapshot;
  p.noise = h->last_noise;
  const uint64_t idx = h->solve_count - 1;
  p.key.solve_lo = (uint32_t)idx;
  p.key.solve_hi = (uint32_t)(idx >> 32);
  return p;
}

// Select tree: reduce `src` level by level (one CTA per kTopSlice candidates, n winners each) until a single
// CTA can finish; on return *src is what the final CTA selects from.
static int run_select_levels(MppiHandle* h, TopSource* src, int n, cudaStream_t st) {
  int side = 0;
  while (src->count > kTopSlice) {
    const long long blocks = ((long long)src->count + kTopSlice - 1) / kTopSlice;
    const size_t need = (size_t)blocks * n;
    if (need > (size_t)0x7fffffff) return fail(MPPI_ERR_UNSUPPORTED, "too many candidates for the select tree");
    if (need > h->cand_cap) {
      if (side != 0) return fail(MPPI_ERR_STATE, "select tree grew");  // later levels only shrink
      for (int i = 0; i < 2; ++i) {
        cudaFree(h->d_cand_cost[i]);
        cudaFree(h->d_cand_id[i]);
        h->d_cand_cost[i] = nullptr;
        h->d_cand_id[i] = nullptr;
      }
      h->cand_cap = 0;
      for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaMalloc((void**)&h->d_cand_cost[i], need * 4));
        CUDA_TRY(cudaMalloc((void**)&h->d_cand_id[i], need * 4));
      }
      h->cand_cap = need;
    }
    const int dst = side & 1;
    topn_select_kernel<<<(unsigned)blocks, kTopThreads, 0, st>>>(*src, n, 0, h->d_cand_cost[dst], h->d_cand_id[dst]);
    CUDA_TRY(cudaGetLastError());
    h->last_epilogue_launches++;
    src->costs = h->d_cand_cost[dst];
    src->ids = h->d_cand_id[dst];
    src->id_offset = 0;
    src->count = (int)need;
    ++side;
  }
  return MPPI_OK;
}

struct RerollArgs {
  long long noise_id_base;
  float* traj;
  float* w;
};

template <class M>
static int launch_epilogue(MppiHandle* h, const SolveParams& p, const EpilogueParams& e, const RerollArgs& r,
                           cudaStream_t st) {
  control_epilogue_kernel<M><<<1, kTopThreads, 0, st>>>(p, e);
  CUDA_TRY(cudaGetLastError());
  h->last_epilogue_launches++;
  if (e.top_n <= 0) return MPPI_OK;
  const int n = e.top_n;
  unsigned grid = (unsigned)((n + 31) / 32), block = 32, smem = 0;
  if constexpr (M::kParallelTail) {
    grid = (unsigned)n;
    block = 128;
    smem = ((unsigned)h->E_pad + (unsigned)tail_per_step<M>() * (unsigned)(h->cfg.horizon + 9) + 16u) * 4u;
  }
  if (p.noise)
    reroll_winners_kernel<M, true><<<grid, block, smem, st>>>(p, e.top_cost, e.top_id, r.noise_id_base, n, r.traj, r.w);
  else
    reroll_winners_kernel<M, false><<<grid, block, smem, st>>>(p, e.top_cost, e.top_id, r.noise_id_base, n, r.traj, r.w);
  CUDA_TRY(cudaGetLastError());
  h->last_epilogue_launches++;
  return MPPI_OK;
}

static int dispatch_epilogue(MppiHandle* h, const SolveParams& p, const EpilogueParams& e, const RerollArgs& r,
                             cudaStream_t st) {
  switch (h->cfg.model) {
    case MPPI_MODEL_PENDULUM: return launch_epilogue<Pendulum>(h, p, e, r, st);
    case MPPI_MODEL_CARTPOLE: return launch_epilogue<Cartpole>(h, p, e, r, st);
    case MPPI_MODEL_MOUNTAINCAR: return launch_epilogue<MountainCar>(h, p, e, r, st);
    case MPPI_MODEL_NAVIGATION2D: return launch_epilogue<Navigation2D>(h, p, e, r, st);
    case MPPI_MODEL_RACING: return launch_epilogue<Racing>(h, p, e, r, st);
    case MPPI_MODEL_CARTPOLE_CONTINUOUS: return launch_epilogue<CartpoleContinuous>(h, p, e, r, st);
    case MPPI_MODEL_GOAL_IN_DANGER_ZONE: return launch_epilogue<GoalInDangerZone>(h, p, e, r, st);
  }
  return fail(MPPI_ERR_INVALID, "unknown model");
}

extern "C" {

int mppi_top_candidates(MppiHandle* h, int32_t n, float* d_cand_cost, int32_t* d_cand_id, void* stream) {
  if (!h || !d_cand_cost || !d_cand_id) return fail(MPPI_ERR_INVALID, "null argument");
  if (n < 1 || n > kTopMax) return fail(MPPI_ERR_INVALID, "n must be in [1, %d]", kTopMax);
  if (!h->solved) return fail(MPPI_ERR_STATE, "no solve has run yet");
  if (h->cfg.sample_offset + (long long)h->cfg.num_samples > 0x7fffffffLL)
    return fail(MPPI_ERR_UNSUPPORTED, "global sample ids beyond 2^31");
  ON_DEVICE(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  h->last_epilogue_launches = 0;
  TopSource src{h->d_costs, nullptr, h->cfg.sample_offset, h->cfg.num_samples};
  int rc = run_select_levels(h, &src, n, st);
  if (rc) return rc;
  topn_select_kernel<<<1, kTopThreads, 0, st>>>(src, n, 1, d_cand_cost, d_cand_id);
  CUDA_TRY(cudaGetLastError());
  h->last_epilogue_launches++;
  return MPPI_OK;
}

int mppi_step_epilogue(MppiHandle* h, const MppiStepEpilogue* a, void* stream) {
  if (!h || !a) return fail(MPPI_ERR_INVALID, "null argument");
  if (!h->solved) return fail(MPPI_ERR_STATE, "no solve has run yet");
  const int K = h->cfg.num_samples;
  const bool want_step = a->d_next_state != nullptr, want_flags = a->d_flags != nullptr, want_top = a->top_n > 0;
  if (want_step && !a->d_action_seq) return fail(MPPI_ERR_INVALID, "env step needs d_action_seq");
  if (want_flags && !a->d_state_seq) return fail(MPPI_ERR_INVALID, "collision flags need d_state_seq");
  if (want_flags && !want_step) return fail(MPPI_ERR_INVALID, "d_flags[0] is the goal test of d_next_state: pass both");
  if (want_top) {
    if (!a->d_top_traj || !a->d_top_w) return fail(MPPI_ERR_INVALID, "top samples need d_top_traj and d_top_w");
    if (a->top_n > kTopMax) return fail(MPPI_ERR_UNSUPPORTED, "top_n > %d: use mppi_top_samples", kTopMax);
    if (a->top_n > h->cfg.total_samples) return fail(MPPI_ERR_INVALID, "num_samples must be in [1, %lld]", (long long)h->cfg.total_samples);
    if ((a->d_cand_cost == nullptr) != (a->d_cand_id == nullptr) || (a->d_cand_cost && a->n_cand < a->top_n))
      return fail(MPPI_ERR_INVALID, "candidate list must hold costs, ids and at least top_n entries");
    if (!a->d_cand_cost && a->top_n > K) return fail(MPPI_ERR_INVALID, "num_samples must be in [1, %d]", K);
    if (h->cfg.sample_offset + (long long)K > 0x7fffffffLL) return fail(MPPI_ERR_UNSUPPORTED, "global sample ids beyond 2^31");
  }
  ON_DEVICE(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  h->last_epilogue_launches = 0;
  SolveParams p = last_solve_params(h);
  EpilogueParams e{};
  e.state = a->d_state ? a->d_state : h->d_state_snapshot;
  e.action_seq = a->d_action_seq;
  e.state_seq = a->d_state_seq;
  e.goal_x = a->goal_x;
  e.goal_y = a->goal_y;
  e.goal_threshold = a->goal_threshold;
  e.next_state = a->d_next_state;
  e.flags = a->d_flags;
  e.top_n = want_top ? a->top_n : 0;
  RerollArgs r{h->cfg.sample_offset, a->d_top_traj, a->d_top_w};
  if (want_top) {
    if (!h->d_win_cost) {  // winners between the select and the re-roll launch (when the caller does not want them)
      CUDA_TRY(cudaMalloc((void**)&h->d_win_cost, kTopMax * 4));
      CUDA_TRY(cudaMalloc((void**)&h->d_win_id, kTopMax * 4));
    }
    e.top_cost = a->d_top_cost ? a->d_top_cost : h->d_win_cost;
    e.top_id = a->d_top_id ? a->d_top_id : h->d_win_id;
    if (a->d_noise_global) {  // injected noise indexed by GLOBAL sample id (winners of other ranks included)
      p.noise = a->d_noise_global;
      r.noise_id_base = 0;
    } else if (p.noise && a->d_cand_cost && h->cfg.total_samples != K) {
      return fail(MPPI_ERR_INVALID, "merged top samples of an injected-noise solve need d_noise_global");
    }
    e.src = a->d_cand_cost ? TopSource{a->d_cand_cost, a->d_cand_id, 0, a->n_cand}
                           : TopSource{h->d_costs, nullptr, h->cfg.sample_offset, K};
    int rc = run_select_levels(h, &e.src, e.top_n, st);
    if (rc) return rc;
  }
  return dispatch_epilogue(h, p, e, r, st);
}

int32_t mppi_last_epilogue_launches(const MppiHandle* h) { return h ? h->last_epilogue_launches : 0; }

int mppi_top_samples(MppiHandle* h, int32_t n, float* d_traj, float* d_w, void* stream) {
  if (!h || !d_traj || !d_w) return fail(MPPI_ERR_INVALID, "null argument");
  const int K = h->cfg.num_samples;
  if (n < 1 || n > K) return fail(MPPI_ERR_INVALID, "num_samples must be in [1, %d]", K);  // mppi.py:476
  if (!h->solved) return fail(MPPI_ERR_STATE, "no solve has run yet");
  if (n <= kTopMax && h->cfg.sample_offset + (long long)K <= 0x7fffffffLL) {
    // the usual case (the examples ask for 300): radix select + re-roll, no sort of all K costs
    MppiStepEpilogue a{};
    a.top_n = n;
    a.d_top_traj = d_traj;
    a.d_top_w = d_w;
    return mppi_step_epilogue(h, &a, stream);
  }
  ON_DEVICE(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (!h->d_idx_in) {
    CUDA_TRY(cudaMalloc((void**)&h->d_idx_in, (size_t)K * 4));
    CUDA_TRY(cudaMalloc((void**)&h->d_idx_out, (size_t)K * 4));
    CUDA_TRY(cudaMalloc((void**)&h->d_keys_out, (size_t)K * 4));
    iota_kernel<<<(K + 255) / 256, 256, 0, st>>>(h->d_idx_in, K);
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, h->sort_tmp_bytes, h->d_costs, h->d_keys_out, h->d_idx_in,
                                             h->d_idx_out, K, 0, 32, st));
    CUDA_TRY(cudaMalloc(&h->d_sort_tmp, h->sort_tmp_bytes));
  }
  // highest weight == lowest cost (mppi.py:479-485); ascending (stable) radix sort on the fp32 costs
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp, h->sort_tmp_bytes, h->d_costs, h->d_keys_out, h->d_idx_in,
                                           h->d_idx_out, K, 0, 32, st));
  SolveParams p = last_solve_params(h);
  switch (h->cfg.model) {
    case MPPI_MODEL_PENDULUM: return launch_reroll<Pendulum>(h, p, n, d_traj, d_w, st);
    case MPPI_MODEL_CARTPOLE: return launch_reroll<Cartpole>(h, p, n, d_traj, d_w, st);
    case MPPI_MODEL_MOUNTAINCAR: return launch_reroll<MountainCar>(h, p, n, d_traj, d_w, st);
    case MPPI_MODEL_NAVIGATION2D: return launch_reroll<Navigation2D>(h, p, n, d_traj, d_w, st);
    case MPPI_MODEL_RACING: return launch_reroll<Racing>(h, p, n, d_traj, d_w, st);
    case MPPI_MODEL_CARTPOLE_CONTINUOUS: return launch_reroll<CartpoleContinuous>(h, p, n, d_traj, d_w, st);
    case MPPI_MODEL_GOAL_IN_DANGER_ZONE: return launch_reroll<GoalInDangerZone>(h, p, n, d_traj, d_w, st);
  }
  return fail(MPPI_ERR_INVALID, "unknown model");
}

int mppi_rollout_actions(MppiHandle* h, const float* d_state, const float* d_actions, int32_t n, float* d_traj,
                         void* stream) {
  if (!h || !d_state || !d_actions || !d_traj || n < 1) return fail(MPPI_ERR_INVALID, "bad argument");
  ON_DEVICE(h->device);
  SolveParams p = h->base;
  p.state = d_state;
  cudaStream_t st = (cudaStream_t)stream;
  switch (h->cfg.model) {
    case MPPI_MODEL_PENDULUM: return launch_rollout_actions<Pendulum>(h, p, d_actions, n, d_traj, st);
    case MPPI_MODEL_CARTPOLE: return launch_rollout_actions<Cartpole>(h, p, d_actions, n, d_traj, st);
    case MPPI_MODEL_MOUNTAINCAR: return launch_rollout_actions<MountainCar>(h, p, d_actions, n, d_traj, st);
    case MPPI_MODEL_NAVIGATION2D: return launch_rollout_actions<Navigation2D>(h, p, d_actions, n, d_traj, st);
    case MPPI_MODEL_RACING: return launch_rollout_actions<Racing>(h, p, d_actions, n, d_traj, st);
    case MPPI_MODEL_CARTPOLE_CONTINUOUS: return launch_rollout_actions<CartpoleContinuous>(h, p, d_actions, n, d_traj, st);
    case MPPI_MODEL_GOAL_IN_DANGER_ZONE: return launch_rollout_actions<GoalInDangerZone>(h, p, d_actions, n, d_traj, st);
  }
  return fail(MPPI_ERR_INVALID, "unknown model");
}

int mppi_get_lambda(MppiHandle* h, double* lambda_used, double* lambda_next, void* stream) {
  if (!h) return fail(MPPI_ERR_INVALID, "null handle");
  ON_DEVICE(h->device);
  DeviceScalars sc;
  CUDA_TRY(cudaMemcpyAsync(&sc, h->d_sc, sizeof sc, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  if (lambda_used) *lambda_used = sc.lambda_used;
  if (lambda_next) *lambda_next = sc.lambda;
  return MPPI_OK;
}

int mppi_get_carry(MppiHandle* h, float* d_prev, float* d_hist, void* stream) {
  if (!h) return fail(MPPI_ERR_INVALID, "null handle");
  ON_DEVICE(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (d_prev) CUDA_TRY(cudaMemcpyAsync(d_prev, h->d_prev_action, (size_t)h->E * 4, cudaMemcpyDeviceToDevice, st));
  const size_t hb = (size_t)(h->cfg.horizon - 1) * h->mi.du * 4;
  if (d_hist && hb) CUDA_TRY(cudaMemcpyAsync(d_hist, h->d_history, hb, cudaMemcpyDeviceToDevice, st));
  return MPPI_OK;
}

int mppi_set_carry(MppiHandle* h, const float* d_prev, const float* d_hist, void* stream) {
  if (!h) return fail(MPPI_ERR_INVALID, "null handle");
  ON_DEVICE(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (d_prev) CUDA_TRY(cudaMemcpyAsync(h->d_prev_action, d_prev, (size_t)h->E * 4, cudaMemcpyDeviceToDevice, st));
  const size_t hb = (size_t)(h->cfg.horizon - 1) * h->mi.du * 4;
  if (d_hist && hb) CUDA_TRY(cudaMemcpyAsync(h->d_history, d_hist, hb, cudaMemcpyDeviceToDevice, st));
  return MPPI_OK;
}

int32_t mppi_last_launch_count(const MppiHandle* h) { return h ? h->last_launches : 0; }

int mppi_launch_info(const MppiHandle* h, int32_t* grid, int32_t* block, int32_t* smem_bytes) {
  if (!h) return fail(MPPI_ERR_INVALID, "null handle");
  if (grid) *grid = h->geo[0].grid + 1;  // workers + the finisher block (LBPS / ESSPS cost launches: workers only)
  if (block) *block = h->geo[0].block;
  if (smem_bytes) *smem_bytes = (int32_t)h->geo[0].smem;
  return MPPI_OK;
}

int mppi_map_info(const MppiHandle* h, int32_t slot, int32_t* fast_division, uint64_t* mismatches, int32_t* model_flags) {
  if (!h || slot < 0 || slot > 1) return fail(MPPI_ERR_INVALID, "bad argument");
  if (fast_division) *fast_division = h->base.map_fastdiv[slot];
  if (mismatches) *mismatches = h->fastdiv_mismatches[slot];
  if (model_flags) *model_flags = h->base.mp.flags;
  return MPPI_OK;
}

int mppi_block_trace(MppiHandle* h, int32_t enable, uint64_t* h_out, int32_t max_blocks) {
  if (!h) return fail(MPPI_ERR_INVALID, "null handle");
  ON_DEVICE(h->device);
  const size_t n = (size_t)((h->cfg.num_samples + 63) / 64 + 1) * kTraceSlots;  // every worker block + the finisher
  if (enable && !h->d_trace) {
    CUDA_TRY(cudaMalloc((void**)&h->d_trace, n * 8));
    CUDA_TRY(cudaMemset(h->d_trace, 0, n * 8));
  }
  if (h_out && h->d_trace) {
    CUDA_TRY(cudaDeviceSynchronize());
    size_t blocks = std::min<size_t>((size_t)max_blocks, (size_t)h->geo[0].grid + 1);
    CUDA_TRY(cudaMemcpy(h_out, h->d_trace, blocks * kTraceSlots * 8, cudaMemcpyDeviceToHost));
  }
  if (!enable && h->d_trace) {
    cudaFree(h->d_trace);
    h->d_trace = nullptr;
  }
  return MPPI_OK;
}

int mppi_selftest(int32_t device, uint64_t mismatches[4]) {
  if (!mismatches) return fail(MPPI_ERR_INVALID, "null argument");
  ON_DEVICE(device);
  unsigned long long* d = nullptr;
  CUDA_TRY(cudaMalloc((void**)&d, 32));
  CUDA_TRY(cudaMemset(d, 0, 32));
  selftest_kernel<<<148 * 8, 256>>>(d);
  unsigned long long hbad[4] = {1, 1, 1, 1};
  cudaError_t e = cudaMemcpy(hbad, d, 32, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(MPPI_ERR_CUDA, "selftest: %s", cudaGetErrorString(e));
  for (int i = 0; i < 4; ++i) mismatches[i] = hbad[i];
  return MPPI_OK;
}

// ---- racing reference path on the device ("next" row 1) -----------------------------
struct MppiRefPath {
  int device = 0, n = 0, rows = 0;
  float v_max = 0.f;
  float* d_path = nullptr;
  int* d_dind = nullptr;
  int* d_cind = nullptr;
};

int mppi_refpath_create(int32_t device, const float* h_path, int32_t n, const int32_t* h_index_offsets, int32_t rows,
                        float v_max, MppiRefPath** out) {
  if (!h_path || !h_index_offsets || !out || n < 1 || rows < 1) return fail(MPPI_ERR_INVALID, "bad argument");
  ON_DEVICE(device);
  MppiRefPath* r = new (std::nothrow) MppiRefPath();
  if (!r) return fail(MPPI_ERR_INVALID, "out of host memory");
  r->device = device;
  r->n = n;
  r->rows = rows;
  r->v_max = v_max;
  cudaError_t e = cudaMalloc((void**)&r->d_path, (size_t)n * 12);
  if (e == cudaSuccess) e = cudaMalloc((void**)&r->d_dind, (size_t)rows * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&r->d_cind, 16);
  if (e == cudaSuccess) e = cudaMemcpy(r->d_path, h_path, (size_t)n * 12, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(r->d_dind, h_index_offsets, (size_t)rows * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(r->d_cind, 0, 16);
  if (e != cudaSuccess) {
    cudaFree(r->d_path);
    cudaFree(r->d_dind);
    cudaFree(r->d_cind);
    delete r;
    return fail(MPPI_ERR_CUDA, "refpath create: %s", cudaGetErrorString(e));
  }
  *out = r;
  return MPPI_OK;
}

void mppi_refpath_destroy(MppiRefPath* r) {
  if (!r) return;
  DeviceGuard guard(r->device);
  cudaFree(r->d_path);
  cudaFree(r->d_dind);
  cudaFree(r->d_cind);
  delete r;
}

int mppi_refpath_update(MppiRefPath* r, const float* d_state, float* d_refpath_out, void* stream) {
  if (!r || !d_state || !d_refpath_out) return fail(MPPI_ERR_INVALID, "null argument");
  ON_DEVICE(r->device);
  racing_refpath_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(r->d_path, r->n, d_state, r->d_dind, r->rows, r->v_max,
                                                               r->d_cind, d_refpath_out);
  CUDA_TRY(cudaGetLastError());
  return MPPI_OK;
}

int mppi_refpath_index(MppiRefPath* r, int32_t set_value, int32_t* current, void* stream) {
  if (!r) return fail(MPPI_ERR_INVALID, "null argument");
  ON_DEVICE(r->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (set_value >= 0) CUDA_TRY(cudaMemcpyAsync(r->d_cind, &set_value, 4, cudaMemcpyHostToDevice, st));
  if (current) {
    CUDA_TRY(cudaMemcpyAsync(current, r->d_cind, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  } else if (set_value >= 0) {
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return MPPI_OK;
}

int mppi_kernel_timing(MppiHandle* h, int32_t enable) {
  if (!h) return fail(MPPI_ERR_INVALID, "null handle");
  h->timing = enable != 0;
  return MPPI_OK;
}

int mppi_kernel_time_ms(MppiHandle* h, double* mean_ms, int32_t* launches) {
  if (!h) return fail(MPPI_ERR_INVALID, "null handle");
  double total = 0;
  int n = 0;
  for (auto& ev : h->events) {
    CUDA_TRY(cudaEventSynchronize(ev.second));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, ev.first, ev.second));
    total += ms;
    ++n;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  h->events.clear();
  if (mean_ms) *mean_ms = n ? total / n : 0.0;
  if (launches) *launches = n;
  return MPPI_OK;
}

uint64_t mppi_solve_index(const MppiHandle* h) { return h ? h->solve_count : 0; }

int mppi_sample_noise(MppiHandle* h, uint64_t solve_index, float* d_noise_out, void* stream) {
  if (!h || !d_noise_out) return fail(MPPI_ERR_INVALID, "null argument");
  ON_DEVICE(h->device);
  SolveParams p = h->base;
  p.key.solve_lo = (uint32_t)solve_index;
  p.key.solve_hi = (uint32_t)(solve_index >> 32);
  const int K = h->cfg.num_samples;
  if (h->mi.du == 1)
    sample_noise_kernel<1><<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p, d_noise_out);
  else
    sample_noise_kernel<2><<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p, d_noise_out);
  CUDA_TRY(cudaGetLastError());
  return MPPI_OK;
}

}  // extern "C"
