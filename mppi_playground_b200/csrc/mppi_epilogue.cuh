// mppi_epilogue.cuh - what the reference's control loops run right AFTER a solve, on the device
// (SURVEY 8f "next" row 2; paths relative to the reference root):
//
//   state, is_goal_reached = env.step(action_seq[0, :])         example/racing.py:233, src/envs/racing_env.py:142-163
//   is_collisions = env.collision_check(state=state_seq)        example/racing.py:235, src/envs/racing_env.py:374-384
//   top_samples, top_weights = solver.get_top_samples(300)      example/racing.py:237, src/pi_mpc/mppi.py:462-487
//
// (same three calls in example/navigation2d.py:39-44 against src/envs/navigation_2d.py:97-117,281-291). In the
// reference these are ~100 tiny ATen launches per control step plus a topk over the K weights; here they are
// a few small launches: a radix SELECT of the n lowest costs (the n highest weights: the softmax is
// monotone) instead of a sort of all K, the winners' trajectories re-rolled from the sampler key, the executed
// action's dynamics step, the goal test and the occupancy flags of the predicted trajectory.
//
// Ordering contract of the select (deterministic, equal to a stable ascending sort of the costs): winners are the
// n smallest (ordered cost key, global sample id) pairs, returned in that order.
#pragma once
#include "mppi_kernels.cuh"

namespace mppi {

constexpr int kTopMax = 1024;      // largest n of the select path: one CTA sorts its winners in shared memory
constexpr int kTopThreads = 1024;  // threads of a select CTA
constexpr int kTopSlice = 8192;    // elements one CTA selects from: one batch of 8 per thread and pass. A single SM
                                   // needs ~50 us for four passes over 65536 candidates (issue bound), eight SMs ~6
constexpr int kTopBins = 2048;     // 11 bits per radix pass: 11 + 11 + 10

// fp32 -> u32 whose unsigned order is the float order (negative: flip all bits, else set the sign bit)
__device__ __forceinline__ uint32_t ordered_key(float c) {
  const uint32_t u = __float_as_uint(c);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct TopShared {
  unsigned hist[kTopBins];
  unsigned warp_tot[32];
  unsigned sel_bin, sel_before, sel_in_bin, count;
  unsigned long long win[kTopMax];  // (ordered key << 32) | global id; ~0 = empty
};

// The candidate set a CTA selects from: element i in [0, count) has cost costs[i] and id ids[i] (ids == null:
// id_offset + i, i.e. the handle's own samples by global id). Entries with id -1 are padding of an earlier level.
struct TopSource {
  const float* costs;
  const int* ids;
  long long id_offset;
  int count;
  __device__ __forceinline__ uint32_t key(int i) const { return ordered_key(costs[i]); }
  __device__ __forceinline__ uint32_t id(int i) const { return ids ? (uint32_t)ids[i] : (uint32_t)(id_offset + i); }
};

// Visit every element of src[lo, hi) with f(valid, key, id), the WHOLE warp in lock step (f may use warp
// collectives; `valid` is false on the lanes that ran off the end). Loads are issued in batches of kTopBatch
// before any is consumed: the candidates sit in L2, and one exposed ~600-cycle load per element and pass was
// the whole cost of the first version of this kernel.
constexpr int kTopBatch = 8;
template <class F>
__device__ __forceinline__ void for_each_candidate(const TopSource& src, int lo, int hi, F f) {
  const int tid = threadIdx.x;
  for (int b0 = lo; b0 < hi; b0 += kTopBatch * kTopThreads) {
    float c[kTopBatch];
    int id[kTopBatch];
#pragma unroll
    for (int j = 0; j < kTopBatch; ++j) {
      const int i = b0 + j * kTopThreads + tid;
      c[j] = (i < hi) ? __ldg(src.costs + i) : 0.0f;
      id[j] = (i < hi && src.ids) ? __ldg(src.ids + i) : 0;
    }
#pragma unroll
    for (int j = 0; j < kTopBatch; ++j) {
      const int i = b0 + j * kTopThreads + tid;
      f(i < hi, ordered_key(c[j]), src.ids ? (uint32_t)id[j] : (uint32_t)(src.id_offset + i));
    }
  }
}

// One radix-select pass over the elements of [lo, hi) whose `value` matches (value & pmask) == pval: histogram
// of the `bits` bits at `shift`, then the bin that holds the need-th smallest (1-based) of the matching elements.
// kOnIds: the value is the id of the elements whose KEY equals tie_key (tie break among equal costs), else the key.
// The histogram adds are aggregated per warp (one shared atomic per distinct bin and warp): MPPI costs crowd
// into a handful of bins.
template <bool kOnIds>
__device__ __forceinline__ void radix_pass(TopShared& sh, const TopSource& src, int lo, int hi, uint32_t tie_key,
                                           uint32_t pmask, uint32_t pval, int shift, int bits, unsigned need,
                                           unsigned* bin, unsigned* before, unsigned* in_bin) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bmask = (1u << bits) - 1u;
  for (int b = tid; b < kTopBins; b += kTopThreads) sh.hist[b] = 0u;
  __syncthreads();
  for_each_candidate(src, lo, hi, [&](bool valid, uint32_t k, uint32_t id) {
    const uint32_t v = kOnIds ? id : k;
    const bool counted = valid && (!kOnIds || k == tie_key) && ((v & pmask) == pval);
    const uint32_t b = counted ? ((v >> shift) & bmask) : 0xffffffffu;
    const unsigned same = __match_any_sync(kFullMask, b);
    if (counted && lane == __ffs(same) - 1) atomicAdd(&sh.hist[b], (unsigned)__popc(same));
  });
  __syncthreads();
  // exclusive scan over the 2048 bins, two bins per thread
  const unsigned a = sh.hist[2 * tid], b = sh.hist[2 * tid + 1], s = a + b;
  unsigned incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned up = __shfl_up_sync(kFullMask, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) sh.warp_tot[warp] = incl;
  __syncthreads();
  unsigned base = (lane < warp) ? sh.warp_tot[lane] : 0u;  // sum of the lower warps' totals (kTopThreads / 32 = 32 warps)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(kFullMask, base, o);
  const unsigned excl = base + incl - s;
  if (need > excl && need <= excl + a) {
    sh.sel_bin = 2u * tid;
    sh.sel_before = excl;
    sh.sel_in_bin = a;
  } else if (need > excl + a && need <= excl + s) {
    sh.sel_bin = 2u * tid + 1u;
    sh.sel_before = excl + a;
    sh.sel_in_bin = b;
  }
  __syncthreads();
  *bin = sh.sel_bin;
  *before = sh.sel_before;
  *in_bin = sh.sel_in_bin;
  __syncthreads();
}

// The n smallest (key, id) pairs of src[lo, hi), ascending, into sh.win[0 .. n) (entries beyond the number of
// elements stay ~0). Whole CTA of kTopThreads threads; 1 <= n <= kTopMax.
__device__ __forceinline__ void block_select_topn(TopShared& sh, const TopSource& src, int lo, int hi, int n,
                                                  bool sorted = true) {
  const int tid = threadIdx.x;
  const int count = hi - lo;
  uint32_t kth = 0xffffffffu, id_th = 0xffffffffu;  // select key < kth, or key == kth and id <= id_th
  if (count > n) {
    unsigned bin, before, in_bin, need = (unsigned)n;
    radix_pass<false>(sh, src, lo, hi, 0u, 0u, 0u, 21, 11, need, &bin, &before, &in_bin);
    uint32_t prefix = bin << 21;
    need -= before;
    radix_pass<false>(sh, src, lo, hi, 0u, 0xffe00000u, prefix, 10, 11, need, &bin, &before, &in_bin);
    prefix |= bin << 10;
    need -= before;
    radix_pass<false>(sh, src, lo, hi, 0u, 0xfffffc00u, prefix, 0, 10, need, &bin, &before, &in_bin);
    kth = prefix | bin;
    need -= before;  // how many of the in_bin samples whose key is exactly kth are taken: the lowest ids
    if (need < in_bin) {
      radix_pass<true>(sh, src, lo, hi, kth, 0u, 0u, 21, 11, need, &bin, &before, &in_bin);
      uint32_t ip = bin << 21;
      need -= before;
      radix_pass<true>(sh, src, lo, hi, kth, 0xffe00000u, ip, 10, 11, need, &bin, &before, &in_bin);
      ip |= bin << 10;
      need -= before;
      radix_pass<true>(sh, src, lo, hi, kth, 0xfffffc00u, ip, 0, 10, need, &bin, &before, &in_bin);
      id_th = ip | bin;
    }
  }
  for (int i = tid; i < kTopMax; i += kTopThreads) sh.win[i] = ~0ull;
  if (tid == 0) sh.count = 0u;
  __syncthreads();
  for_each_candidate(src, lo, hi, [&](bool valid, uint32_t k, uint32_t id) {
    if (valid && (k < kth || (k == kth && id <= id_th))) {
      const unsigned pos = atomicAdd(&sh.count, 1u);
      if (pos < (unsigned)kTopMax) sh.win[pos] = ((unsigned long long)k << 32) | id;
    }
  });
  __syncthreads();
  if (!sorted) return;  // an inner level of the select tree: the next level only needs the SET of winners
  // bitonic sort of the next power of two >= n entries (the pairs are unique: ids differ)
  int N = 2;
  while (N < n) N <<= 1;
  for (int k = 2; k <= N; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int i = tid, ixj = i ^ j;
      if (i < N && ixj > i) {
        const unsigned long long x = sh.win[i], y = sh.win[ixj];
        const bool up = (i & k) == 0;
        if ((x > y) == up) {
          sh.win[i] = y;
          sh.win[ixj] = x;
        }
      }
      __syncthreads();
    }
}

// Level of the select tree: CTA b reduces slice b of the candidates to its n best, [gridDim.x, n] pairs out
// (cost +inf / id -1 where a slice holds fewer than n elements).
__global__ void __launch_bounds__(kTopThreads, 1) topn_select_kernel(const __grid_constant__ TopSource src, int n,
                                                                     int sorted, float* __restrict__ out_cost,
                                                                     int* __restrict__ out_id) {
  __shared__ TopShared sh;
  const int lo = (int)min((long long)blockIdx.x * kTopSlice, (long long)src.count);
  const int hi = (int)min((long long)lo + kTopSlice, (long long)src.count);
  block_select_topn(sh, src, lo, hi, n, sorted != 0);
  for (int i = threadIdx.x; i < n; i += kTopThreads) {
    const unsigned long long e = sh.win[i];
    const uint32_t id = (uint32_t)e;
    const bool empty = id == 0xffffffffu;
    out_cost[(size_t)blockIdx.x * n + i] = empty ? INFINITY : key_to_float((uint32_t)(e >> 32));
    out_id[(size_t)blockIdx.x * n + i] = empty ? -1 : (int)id;
  }
}

// One sample of the last solve rolled again from its sampler key (or the injected noise), states stored:
// what the reference keeps as _state_seq_batch[k] (mppi.py:280-286). `k_noise` indexes p.noise when kInject.
template <class M, bool kInject>
__device__ __forceinline__ void reroll_sample(const SolveParams& p, long long kg, long long k_noise,
                                              float* __restrict__ out) {
  constexpr int DS = M::DS, DU = M::DU, SPC = Chunking<DU>::kStepsPerChunk;
  const uint32_t k_lo = (uint32_t)kg, k_hi = (uint32_t)((unsigned long long)kg >> 32);
  const bool zero_mean = kg >= p.explore_threshold;
  typename M::Ctx ctx{};
  if constexpr (uses_params<M>()) ctx.p = &p.mp;
  float s[DS], seen[DS], u[DU];
  for (int d = 0; d < DS; ++d) s[d] = p.state[d];
  const float* nz = kInject ? (p.noise + (size_t)k_noise * p.T * DU) : nullptr;
  for (int t0 = 0, chunk = 0; t0 < p.T; t0 += SPC, ++chunk) {
    float z[4];
    if (!kInject) normal4(p.key, k_lo, k_hi, (uint32_t)chunk, z);
    for (int j = 0; j < SPC; ++j) {
      const int t = t0 + j;
      if (t >= p.T) break;
      for (int d = 0; d < DU; ++d) {
        const float eps = kInject ? nz[t * DU + d] : p.sigma[d] * z[j * DU + d];
        // the warm start the solve used is gone (the carried state was overwritten); p.prev_action here
        // points at the snapshot taken before the solve
        u[d] = perturbed_entry<DU>(p, p.prev_action, zero_mean, t, d, eps);
      }
      M::step(ctx, s, u, seen);
      for (int d = 0; d < DS; ++d) out[t * DS + d] = seen[d];
    }
  }
  for (int d = 0; d < DS; ++d) out[p.T * DS + d] = s[d];
}

struct EpilogueParams {
  // env.step / collision_check
  const float* state;       // [ds] env._robot_state (the state the solve started from)
  const float* action_seq;  // [T,du]; row 0 is executed
  const float* state_seq;   // [T+1,ds] predicted trajectory
  float goal_x, goal_y, goal_threshold;  // threshold <= 0: no goal test
  float* next_state;  // [ds] or null
  float* flags;       // [1 + T+1]: is_goal_reached, then the occupancy value (0/1) of every predicted position; or null
  // get_top_samples
  TopSource src;
  int top_n;                // 0: none
  float* top_cost;          // [top_n] winners' costs (ascending) / global sample ids, re-rolled by
  int* top_id;              // reroll_winners_kernel
};

// One CTA: env step, goal test, collision flags and the top-n select (alone for K <= kTopSlice, else after
// topn_select_kernel levels); reroll_winners_kernel follows when top samples were asked for.
template <class M>
__global__ void __launch_bounds__(kTopThreads, 1) control_epilogue_kernel(const __grid_constant__ SolveParams p,
                                                                          const __grid_constant__ EpilogueParams e) {
  constexpr int DS = M::DS, DU = M::DU;
  __shared__ TopShared sh;
  const int tid = threadIdx.x;
  // ---- env.step (racing_env.py:142-163, navigation_2d.py:97-117): clamp to the env bounds, one dynamics step,
  //      goal test. The env's clamp is the first operation of the model's own dynamics (same bounds), so
  //      M::step alone reproduces clamp-then-dynamics.
  if (tid == kTopThreads - 1 && e.next_state) {
    typename M::Ctx ctx{};
    if constexpr (uses_params<M>()) ctx.p = &p.mp;
    float s[DS], seen[DS], u[DU];
    for (int d = 0; d < DS; ++d) s[d] = e.state[d];
    for (int d = 0; d < DU; ++d) u[d] = e.action_seq[d];
    M::step(ctx, s, u, seen);
    for (int d = 0; d < DS; ++d) e.next_state[d] = s[d];
    if (e.flags) {
      float reached = 0.0f;
      if (e.goal_threshold > 0.0f && DS >= 2) {
        const float dx = s[0] - e.goal_x, dy = s[DS >= 2 ? 1 : 0] - e.goal_y;
        reached = (sqrtf(dx * dx + dy * dy) < e.goal_threshold) ? 1.0f : 0.0f;  // torch.norm(...) < threshold
      }
      e.flags[0] = reached;
    }
  }
  // ---- collision_check (racing_env.py:374-384): obstacle-map lookup of every predicted position
  if (e.flags && e.state_seq) {
    for (int t = tid; t <= p.T; t += kTopThreads) {
      float occ = 0.0f;
      if constexpr (M::kMaps >= 1) {
        const MapView m{p.map_bits[0], p.map_W[0], p.map_H[0], p.map_words[0],
                        ExactDiv{p.map_cell[0], p.map_rcp[0], p.map_fastdiv[0]}, p.map_ox[0], p.map_oy[0], 0u};
        occ = map_lookup(m, e.state_seq[t * DS + 0], e.state_seq[t * DS + 1]);
      }
      e.flags[1 + t] = occ;
    }
  }
  if (e.top_n <= 0) return;
  // ---- get_top_samples (mppi.py:462-487), first half: WHICH samples. Their trajectories are re-rolled by
  //      reroll_winners_kernel on as many SMs as there are winners (one SM would spend ~100 us on 300 of them).
  block_select_topn(sh, e.src, 0, e.src.count, e.top_n);
  if (tid < e.top_n) {
    const unsigned long long w = sh.win[tid];
    e.top_cost[tid] = key_to_float((uint32_t)(w >> 32));
    e.top_id[tid] = (int)(uint32_t)w;
  }
}

// get_top_samples, second half: winner i (cost[i], global sample id[i]) rolled again from its sampler key (or the
// injected noise) - what the reference keeps as _state_seq_batch[k] (mppi.py:280-286) - and its weight.
// Models with a block-parallel rollout (racing, navigation2d: M::rollout_block, bit-identical to T serial step()
// calls, see finish_solve) get one block per winner: the controls of all stages are regenerated in parallel, then
// only the short recurrences run serially. The other models roll one winner per thread, 32 winners per block.
template <class M, bool kInject>
__global__ void __launch_bounds__(128, 1) reroll_winners_kernel(const __grid_constant__ SolveParams p,
                                                                const float* __restrict__ cost,
                                                                const int* __restrict__ id, long long noise_id_base,
                                                                int n, float* __restrict__ traj,
                                                                float* __restrict__ wout) {
  constexpr int DS = M::DS, DU = M::DU;
  const int tid = threadIdx.x;
  const float lam = (float)p.sc->lambda_used;
  if constexpr (M::kParallelTail) {
    extern __shared__ __align__(16) float rr_smem[];
    float* opt = rr_smem;                 // [E_pad] the winner's clamped controls
    float* scratch = rr_smem + p.E_pad;   // rollout_block scratch
    const int i = blockIdx.x;
    const long long kg = (long long)(uint32_t)id[i];
    const uint32_t k_lo = (uint32_t)kg, k_hi = (uint32_t)((unsigned long long)kg >> 32);
    const bool zero_mean = kg >= p.explore_threshold;
    const float* nz = kInject ? (p.noise + (size_t)(kg - noise_id_base) * p.T * DU) : nullptr;
    for (int c = tid; c * 4 < p.E; c += blockDim.x) {  // one sampler chunk = 4 consecutive entries of [T*DU]
      float z[4];
      if (!kInject) normal4(p.key, k_lo, k_hi, (uint32_t)c, z);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int e = c * 4 + j;
        if (e < p.E) {
          const int t = e / DU, d = e - t * DU;
          const float eps = kInject ? nz[e] : p.sigma[d] * z[j];
          opt[e] = perturbed_entry<DU>(p, p.prev_action, zero_mean, t, d, eps);
        }
      }
    }
    __syncthreads();
    typename M::Ctx ctx{};
    ctx.p = &p.mp;
    M::rollout_block(ctx, p.state, opt, p.T, traj + (size_t)i * (p.T + 1) * DS, scratch, nullptr, [](int, int) {});
    if (tid == 0) wout[i] = expf((-cost[i]) / lam - p.sc->xmax) / (float)p.sc->S;
  } else {
    const int i = blockIdx.x * blockDim.x + tid;
    if (i >= n) return;
    const long long kg = (long long)(uint32_t)id[i];
    reroll_sample<M, kInject>(p, kg, kg - noise_id_base, traj + (size_t)i * (p.T + 1) * DS);
    wout[i] = expf((-cost[i]) / lam - p.sc->xmax) / (float)p.sc->S;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Device rasteriser for the occupancy grids (SURVEY 8f "next" row 4): ObstacleMap.add_circle_obstacle /
// add_rectangle_obstacle (src/envs/obstacle_map_2d.py:103-160) and LaneMap.populate_map
// (src/envs/lane_map_2d.py:68-88), emitting the bordered bit-packed layout the solve kernels stage (MapView)
// and, optionally, the reference's [W,H] fp32 grid (_map_torch). The world -> cell conversions of the shapes
// are the reference's fp64 numpy expressions and stay on the host (mppi_playground_b200/maps.py); the device
// paints cells:
//   obstacle mode (background 0, paint 1): disc (cx, cy, r): every (i, j) with i^2 + j^2 <= r^2 paints cell
//     (clip(cx + i), clip(cy + j)) - the reference clips INDICES, so a disc that pokes out of the map smears
//     onto the border row / column (:118-123). Per cell: the smallest |i| that lands on ix is |ix - cx| inside
//     the map, max(cx, 0) on row 0 and max(W-1 - cx, 0) on row W-1 (same for j); painted iff di^2 + dj^2 <= r^2.
//     rectangle (x0, x1, y0, y1): the half-open, already clipped slice map[x0:x1, y0:y1] = 1 (:150-159).
//   lane mode (background 1, paint 0): populate_map zeroes the centre-line cells that fall inside the map and
//     keeps every cell whose Euclidean distance transform is <= (lane_width / 2) / cell: the union of the discs
//     d2 <= r2 around those cells, r2 = the largest integer squared distance the host's fp64 test accepts.
// ---------------------------------------------------------------------------------------------------------
struct RasterShape {
  int a, b, c, d;  // disc: cx, cy, r2, unused; rectangle: x0, x1, y0, y1
};

__global__ void raster_map_kernel(const RasterShape* __restrict__ discs, int n_discs,
                                  const RasterShape* __restrict__ rects, int n_rects, int lane_mode, int W, int H,
                                  int words, uint32_t* __restrict__ bits, float* __restrict__ grid) {
  extern __shared__ RasterShape s_shapes[];  // chunks of discs, then rectangles
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = idx < (W + 1) * words;
  const int ix = live ? idx / words : 0, wj = live ? idx - ix * words : 0;
  const int iy0 = wj * 32;
  uint32_t painted = 0u;  // bit b: some shape paints cell (ix, iy0 + b)
  const int chunk = blockDim.x;
  for (int base = 0; base < n_discs; base += chunk) {
    __syncthreads();
    if (base + (int)threadIdx.x < n_discs) s_shapes[threadIdx.x] = discs[base + threadIdx.x];
    __syncthreads();
    const int m = min(chunk, n_discs - base);
    if (!live || ix >= W) continue;
    for (int s = 0; s < m; ++s) {
      const RasterShape q = s_shapes[s];
      long long di;
      if (lane_mode) {
        di = ix - q.a;
      } else {
        di = (ix == 0) ? max(q.a, 0) : (ix == W - 1) ? max(W - 1 - q.a, 0) : abs(ix - q.a);
        if (ix == 0 && W == 1) di = 0;  // a one-row map: every index clips onto it
      }
      const long long rem = (long long)q.c - di * di;
      if (rem < 0) continue;
      for (int b = 0; b < 32; ++b) {
        const int iy = iy0 + b;
        if (iy >= H) break;
        long long dj;
        if (lane_mode) {
          dj = iy - q.b;
        } else {
          dj = (iy == 0) ? max(q.b, 0) : (iy == H - 1) ? max(H - 1 - q.b, 0) : abs(iy - q.b);
          if (iy == 0 && H == 1) dj = 0;
        }
        if (dj * dj <= rem) painted |= 1u << b;
      }
    }
  }
  for (int base = 0; base < n_rects; base += chunk) {
    __syncthreads();
    if (base + (int)threadIdx.x < n_rects) s_shapes[threadIdx.x] = rects[base + threadIdx.x];
    __syncthreads();
    const int m = min(chunk, n_rects - base);
    if (!live || ix >= W) continue;
    for (int s = 0; s < m; ++s) {
      const RasterShape q = s_shapes[s];
      if (ix < q.a || ix >= q.b) continue;
      for (int b = 0; b < 32; ++b) {
        const int iy = iy0 + b;
        if (iy >= q.c && iy < q.d && iy < H) painted |= 1u << b;
      }
    }
  }
  if (!live) return;
  uint32_t v = 0u;
  for (int b = 0; b < 32; ++b) {
    const int iy = iy0 + b;
    const bool inside = ix < W && iy < H;
    const bool hit = (painted >> b) & 1u;
    const bool one = (iy == H) || (iy < H && ix == W) || (inside && (lane_mode ? !hit : hit));  // border of ones
    if (one) v |= 1u << b;
    if (inside && grid) grid[(size_t)ix * H + iy] = (lane_mode ? !hit : hit) ? 1.0f : 0.0f;
  }
  bits[idx] = v;
}

}  // namespace mppi
