// mppi_kernels.cuh - the solve kernels of the B200 MPPI engine (sm_100a).
//
// One solve (reference src/pi_mpc/mppi.py:255-458) is ONE launch of
// solve_kernel<Model, kFused> when lambda is known up front (fixed / MPO):
//
//   pass 1  per thread = one sample k: T-step rollout with state in registers;
//           Gaussian perturbation from the counter-based sampler (or injected
//           noise), clamp, dynamics, stage cost; terminal cost -> costs[k]
//   block   baseline-subtracted softmax with a block-local baseline
//           (online-softmax form), exp weights
//   pass 2  the perturbed controls are regenerated (never stored) and reduced
//           with a transposed warp reduction into sum_k w_k u_k[t,d]
//   last    the last block to finish combines the per-block partials, divides,
//           applies the Savitzky-Golay stencil, rolls the optimal trajectory,
//           carries the warm start / history, updates MPO's temperature.
//
// LBPS / ESSPS need every cost before lambda exists, so they run the same
// kernel split in two (kCosts, kReduce) around lambda_search_kernel.
#pragma once
#include <cooperative_groups.h>
#include <math.h>

#include "mppi_models.cuh"

namespace mppi {

enum SolveMode : int { kFused = 0, kCosts = 1, kReduce = 2 };
enum LambdaMode : int { kLamFixed = 0, kLamMPO = 1, kLamLBPS = 2, kLamESSPS = 3 };

constexpr int kInlineRefFloats = 512;
constexpr int kMaxPeers = 8;
constexpr int kMaxSegments = 8;    // combine_partials: threads = segments x float4 columns
constexpr int kPartialHeader = 8;  // xmax S xmax_tau S_tau Sc_tau cmin cmax pad
constexpr int kMaxSgWindow = 33;
constexpr unsigned kRedBytes = 1536;  // block reduction scratch
constexpr int kTraceSlots = 24;  // %globaltimer stamps per block (mppi_block_trace)

// Scalars carried on the device between solves.
struct DeviceScalars {
  double lambda;       // lambda the NEXT solve's weights use
  double lambda_used;  // lambda the LAST solve's weights used
  double S;            // sum_k exp(x_k - xmax) of the last solve
  float xmax;          // max_k(-c_k / lambda) of the last solve
  float rho;           // MPO log_temperature
  double adam_m, adam_v;
  int adam_step;
  int search_evals;  // objective evaluations of the last LBPS / ESSPS search
  float cmin, cmax;
};

struct SolveParams {
  int K, T;
  long long k_offset;           // global id of this shard's first sample
  long long explore_threshold;  // global ids >= threshold are zero-mean (mppi.py:266)
  float u_min[4], u_max[4], sigma[4];
  ModelParams mp;
  const uint32_t* map_bits[2];
  int map_W[2], map_H[2], map_words[2];
  unsigned map_bytes[2];  // padded to 16 B
  float map_cell[2], map_rcp[2], map_ox[2], map_oy[2];
  int map_fastdiv[2];  // (cell, rcp) passed the exhaustive exact-division check
  const float* state;    // [ds]
  const float* refpath;  // [T+1,4] or null
  const float* noise;    // [K,T,du] or null
  int ref_bulk_ok;       // refpath is 16 B aligned -> bulk copy
  float* prev_action;    // [E] (allocation padded to 16 B)
  unsigned prev_action_bytes;
  float* history;  // [(T-1)*du]
  float* nominal_snapshot;  // [E] warm start the last solve sampled around (for get_top_samples)
  float* state_snapshot;    // [ds] state of the last solve
  DeviceScalars* sc;
  float* costs;           // [K]
  float* block_partials;  // [grid, P]
  float* rank_partial;    // [P]
  unsigned int* counter;
  float* action_out;     // [T,du]
  float* state_seq_out;  // [T+1,ds]
  SamplerKey key;
  int lambda_mode;
  float mpo_epsilon;
  int use_sg, sg_window;
  float sg_coeffs[kMaxSgWindow];
  // host-call path (mppi_solve_host): the solve's inputs travel inside the kernel parameter block
  int inline_inputs;
  float state_inline[8];
  float ref_inline[kInlineRefFloats];  // [T+1,4] when (T+1)*4 <= kInlineRefFloats
  unsigned long long* trace;  // optional [grid, kTraceSlots] %globaltimer stamps per block (profiling aid), else null
  int n_shards;  // 1: finish inside the kernel
  // fused peer exchange (NVLink P2P, one launch per solve on every GPU): each rank owns a mailbox
  // [2 parities][kMaxPeers][P] of 8-byte (payload, sequence number) words; peer_mailbox[r] is rank
  // r's mailbox mapped into this process (CUDA IPC). 0 ranks = staged path (rank_partial + finish_kernel).
  int p2p_world, p2p_rank;
  unsigned p2p_seq;  // sequence number of this solve (>= 1)
  float* peer_mailbox[kMaxPeers];
  float* gather_scratch;  // [kMaxPeers, P] private copy of the gathered partials
  int* error_flag;        // set when the exchange times out (mapped pinned host memory: the host polls it)
  unsigned stage_bytes;   // shared-memory landing zone the host reserved for the block partials (0: none)
  float* dry_scratch;     // global dummy targets of the finisher block's warm-up pass (dry_scratch_floats())
  unsigned* host_done;    // mppi_solve_host: mapped pinned word the finisher sets to host_done_seq when the outputs
  unsigned host_done_seq; // are in host memory (the host spins on it instead of a stream synchronisation), else null
  int E, E_pad, P;
};

// shared-memory carve-up (host and device agree through this helper)
struct SmemLayout {
  unsigned map_off[2];
  unsigned stage_off, stage_cap;  // landing zone of the block partials in the last block (overlays the grids)
  unsigned nominal_off, zero_off, refraw_off, ref4_off, refv_off, warpacc_off, list_off, red_off, misc_off, total;
};

__host__ __device__ inline unsigned align_up(unsigned x, unsigned a) { return (x + a - 1) / a * a; }

// floats of scratch finish_solve needs: N (doubles) | opt | y | scales | Combined | tail rollout
__host__ __device__ inline unsigned finish_scratch_core(int E_pad, int T, int tail_per_step) {
  return align_up((unsigned)E_pad * 20 + 256 * 4 + 64 + (unsigned)E_pad * 8 * kMaxSegments +
                      (unsigned)tail_per_step * (unsigned)(T + 9) * 4,
                  16);
}

__host__ __device__ inline SmemLayout make_layout(int n_maps, const unsigned* map_bytes, int T, int E_pad,
                                                   unsigned prev_action_bytes, bool refpath, int n_warps,
                                                   int tail_per_step, int spt, unsigned stage_bytes) {
  SmemLayout L;
  unsigned o = 16;  // mbarrier
  L.stage_off = o;
  for (int i = 0; i < 2; ++i) {
    L.map_off[i] = o;
    if (i < n_maps) o += align_up(map_bytes[i], 16);
  }
  if (o - 16 < stage_bytes) o = 16 + align_up(stage_bytes, 16);
  L.stage_cap = o - 16;
  L.nominal_off = o;
  o += align_up(prev_action_bytes, 16);
  L.zero_off = o;
  o += align_up((unsigned)E_pad * 4, 16);
  L.refraw_off = o;
  if (refpath) o += (unsigned)(T + 1) * 16;
  L.ref4_off = o;
  if (refpath) o += (unsigned)T * 16;
  L.refv_off = o;
  if (refpath) o += align_up((unsigned)T * 4, 16);
  L.warpacc_off = align_up(o, 16);
  o = L.warpacc_off;
  {  // per-warp accumulators; the last block reuses the region as finish scratch
    unsigned acc = (unsigned)n_warps * (unsigned)E_pad * 4;
    unsigned fin = finish_scratch_core(E_pad, T, tail_per_step);
    o += align_up(acc > fin ? acc : fin, 16);
  }
  L.list_off = o;
  o += (unsigned)n_warps * 32 * 8 * (unsigned)spt;  // (local sample, weight) of the samples with non-zero weight
  L.red_off = o;
  o += kRedBytes;  // block reduction scratch (up to 4 x 32 doubles; combine_partials_small: 4 x 32 floats + 3 x 32 doubles)
  L.misc_off = o;
  o += 256;
  L.total = o;
  return L;
}

// ---------------------------------------------------------------------------
// block reductions (deterministic: fixed shuffle tree, then warp 0 in order)
// ---------------------------------------------------------------------------
template <class T, class Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T identity, void* scratch) {
  T* s = reinterpret_cast<T*>(scratch);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(kFullMask, v, o));
  __syncthreads();  // scratch may still be read from a previous reduction
  if (lane == 0) s[warp] = v;
  __syncthreads();
  T r = (lane < nw) ? s[lane] : identity;
  if (nw > 32) {  // not reachable: blockDim <= 1024
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(kFullMask, r, o));
  return r;  // every thread holds the result
}
struct OpMax {
  __device__ float operator()(float a, float b) const { return fmaxf(a, b); }
};
struct OpMin {
  __device__ float operator()(float a, float b) const { return fminf(a, b); }
};
struct OpAddF {
  __device__ float operator()(float a, float b) const { return a + b; }
};
struct OpAddD {
  __device__ double operator()(double a, double b) const { return a + b; }
};

// ---------------------------------------------------------------------------
// sampling of one chunk (4 normals) -> controls of the timesteps it covers
// ---------------------------------------------------------------------------
template <int DU>
struct Chunking {
  static_assert(DU == 1 || DU == 2, "built-in models have 1 or 2 controls");
  static constexpr int kStepsPerChunk = 4 / DU;
};

// clamp(mean + sigma * eps) for entry e = t * DU + d (mppi.py:266-275)
template <int DU>
__device__ __forceinline__ float perturbed_entry(const SolveParams& p, const float* nominal, bool zero_mean, int t,
                                                 int d, float eps_scaled) {
  float m = zero_mean ? 0.0f : nominal[t * DU + d];
  return clampf(m + eps_scaled, p.u_min[d], p.u_max[d]);
}

// ---------------------------------------------------------------------------
// finish: combine partials -> optimal sequence -> SG filter -> trajectory, carry
// ---------------------------------------------------------------------------
__device__ __forceinline__ void stamp(const SolveParams& p, int slot);

// models whose Ctx carries the parameter block pointer `p`
template <class M>
__host__ __device__ constexpr bool uses_params() {
  return M::kUsesParams;
}

// state / reference path of this solve: device buffers, or the copies inside the parameter block
__device__ __forceinline__ const float* state_of(const SolveParams& p) {
  return p.inline_inputs ? p.state_inline : p.state;
}
__device__ __forceinline__ const float* refpath_of(const SolveParams& p) {
  return p.inline_inputs ? p.ref_inline : p.refpath;
}

struct Combined {
  float xmax, xmax_tau, cmin, cmax;
  double S, S_tau, Sc_tau;
};

// Reduce up to 4 values per thread across the block in one pass (two barriers in total).
template <class T, class Op, int N>
__device__ __forceinline__ void block_reduce_n(T (&v)[N], Op op, T identity, void* scratch) {
  T* s = reinterpret_cast<T*>(scratch);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] = op(v[i], __shfl_xor_sync(kFullMask, v[i], o));
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < N; ++i) s[i * 32 + warp] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    T r = (lane < nw) ? s[i * 32 + lane] : identity;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(kFullMask, r, o));
    v[i] = r;
  }
}

constexpr int kCombineChunk = 256;  // partials whose rescale factors are staged at once

// combine_partials for the common case (n <= blockDim, n <= kCombineChunk: one partial per thread in the header
// stage). Built for latency - this runs in ONE block while every other SM waits: four barriers in all, the
// numerators accumulate in fp32 (fmaf, <= ~n/segments terms per thread, fixed order) and only the handful of
// segment sums and the softmax denominators are added in fp64. Deterministic.
__device__ __forceinline__ void combine_partials_small(const float* __restrict__ parts, int n, int P, int E_pad,
                                                       Combined* out, double* N, float* scale_buf, void* red,
                                                       double* seg_buf, unsigned long long* trace_row) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = (nt + 31) >> 5;
  float* redf = reinterpret_cast<float*>(red);        // [4][32] maxima
  double* redd = reinterpret_cast<double*>(red) + 64;  // [3][32] sums (bytes 512..1279 of kRedBytes)
  float4 h0 = make_float4(-INFINITY, 0.f, -INFINITY, 0.f), h1 = make_float4(0.f, INFINITY, -INFINITY, 0.f);
  if (tid < n) {
    const float4* q = reinterpret_cast<const float4*>(parts + (size_t)tid * P);
    h0 = q[0];
    h1 = q[1];
  }
  float mx[4] = {h0.x, h0.z, -h1.y, h1.z};  // xmax, xmax_tau, -cmin, cmax
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx[i] = fmaxf(mx[i], __shfl_xor_sync(kFullMask, mx[i], o));
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < 4; ++i) redf[i * 32 + warp] = mx[i];
  __syncthreads();  // (1)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float r = (lane < nw) ? redf[i * 32 + lane] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(kFullMask, r, o));
    mx[i] = r;
  }
  const float xm = mx[0], xmt = mx[1];
  double sums[3] = {0.0, 0.0, 0.0};
  if (tid < n) {
    const float sc = expf(h0.x - xm);
    scale_buf[tid] = sc;
    sums[0] = (double)h0.y * (double)sc;
    if (xmt > -INFINITY) {
      const double st = (double)expf(h0.z - xmt);
      sums[1] = (double)h0.w * st;
      sums[2] = (double)h1.x * st;
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sums[i] += __shfl_xor_sync(kFullMask, sums[i], o);
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < 3; ++i) redd[i * 32 + warp] = sums[i];
  __syncthreads();  // (2) rescale factors and the per-warp sums are published
  stamp_row(trace_row, 8);
  const int n_vec = E_pad / 4;
  const int n_seg = max(1, min(kMaxSegments, nt / n_vec));
  const int seg = tid / n_vec, vc = tid - seg * n_vec;
  if (seg < n_seg) {
    const int per = (n + n_seg - 1) / n_seg, lo = seg * per, hi = min(n, lo + per);
    const float* col = parts + kPartialHeader + 4 * vc;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;  // even / odd rows: two independent chains
    int b = lo;
    for (; b + 2 <= hi; b += 2) {
      const float4 x0 = *reinterpret_cast<const float4*>(col + (size_t)b * P);
      const float4 x1 = *reinterpret_cast<const float4*>(col + (size_t)(b + 1) * P);
      const float s0 = scale_buf[b], s1 = scale_buf[b + 1];
      a0.x = fmaf(x0.x, s0, a0.x);
      a0.y = fmaf(x0.y, s0, a0.y);
      a0.z = fmaf(x0.z, s0, a0.z);
      a0.w = fmaf(x0.w, s0, a0.w);
      a1.x = fmaf(x1.x, s1, a1.x);
      a1.y = fmaf(x1.y, s1, a1.y);
      a1.z = fmaf(x1.z, s1, a1.z);
      a1.w = fmaf(x1.w, s1, a1.w);
    }
    if (b < hi) {
      const float4 x0 = *reinterpret_cast<const float4*>(col + (size_t)b * P);
      const float s0 = scale_buf[b];
      a0.x = fmaf(x0.x, s0, a0.x);
      a0.y = fmaf(x0.y, s0, a0.y);
      a0.z = fmaf(x0.z, s0, a0.z);
      a0.w = fmaf(x0.w, s0, a0.w);
    }
    double* dst = seg_buf + (size_t)seg * E_pad + 4 * vc;
    dst[0] = (double)a0.x + (double)a1.x;
    dst[1] = (double)a0.y + (double)a1.y;
    dst[2] = (double)a0.z + (double)a1.z;
    dst[3] = (double)a0.w + (double)a1.w;
  }
  __syncthreads();  // (3)
  stamp_row(trace_row, 9);
  for (int e = tid; e < E_pad; e += nt) {
    double a = 0.0;
    for (int sgm = 0; sgm < n_seg; ++sgm) a += seg_buf[(size_t)sgm * E_pad + e];
    N[e] = a;
  }
  if (tid == 0) {
    double t[3] = {0.0, 0.0, 0.0};
    for (int w = 0; w < nw; ++w) {
      t[0] += redd[w];
      t[1] += redd[32 + w];
      t[2] += redd[64 + w];
    }
    out->xmax = xm;
    out->xmax_tau = xmt;
    out->cmin = -mx[2];
    out->cmax = mx[3];
    out->S = t[0];
    out->S_tau = t[1];
    out->Sc_tau = t[2];
  }
  __syncthreads();  // (4)
}

// Combine n partials (stride P, rows 32 B aligned) into `out` + N[E_pad] (shared, doubles).
// Deterministic: fixed traversal order, independent of which block runs it.
// Two block reductions for the header (maxima, then rescaled sums); the numerators are summed by
// (segment, float4 column) threads with four 16 B loads in flight each, segments added in order.
__device__ __noinline__ void combine_partials(const float* __restrict__ parts, int n, int P, int E_pad, Combined* out,
                                        double* N, float* scale_buf /*[kCombineChunk]*/, void* red,
                                        double* seg_buf /*[kMaxSegments, E_pad]*/,
                                        unsigned long long* trace_row = nullptr) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (n <= nt && n <= kCombineChunk && E_pad / 4 <= nt) {
    combine_partials_small(parts, n, P, E_pad, out, N, scale_buf, red, seg_buf, trace_row);
    return;
  }
  // header pass: every thread keeps the header of its first partial in registers so that the common
  // case n <= blockDim costs ONE global round trip for both reductions and the rescale factors
  float4 h0 = make_float4(-INFINITY, 0.f, -INFINITY, 0.f), h1 = make_float4(0.f, INFINITY, -INFINITY, 0.f);
  if (tid < n) {
    const float4* q = reinterpret_cast<const float4*>(parts + (size_t)tid * P);
    h0 = q[0];
    h1 = q[1];
  }
  float mx[4] = {h0.x, h0.z, -h1.y, h1.z};  // xmax, xmax_tau, -cmin, cmax
  for (int b = tid + nt; b < n; b += nt) {
    const float4* q = reinterpret_cast<const float4*>(parts + (size_t)b * P);
    const float4 g0 = q[0], g1 = q[1];
    mx[0] = fmaxf(mx[0], g0.x);
    mx[1] = fmaxf(mx[1], g0.z);
    mx[2] = fmaxf(mx[2], -g1.y);
    mx[3] = fmaxf(mx[3], g1.z);
  }
  block_reduce_n(mx, OpMax(), -INFINITY, red);
  const float xm = mx[0], xmt = mx[1];
  double sums[3] = {0.0, 0.0, 0.0};
  const float my_scale = (tid < n) ? expf(h0.x - xm) : 0.0f;
  if (tid < n) {
    sums[0] = (double)h0.y * (double)my_scale;
    if (xmt > -INFINITY) {
      double sc = (double)expf(h0.z - xmt);
      sums[1] = (double)h0.w * sc;
      sums[2] = (double)h1.x * sc;
    }
  }
  for (int b = tid + nt; b < n; b += nt) {
    const float4* q = reinterpret_cast<const float4*>(parts + (size_t)b * P);
    const float4 g0 = q[0], g1 = q[1];
    sums[0] += (double)g0.y * (double)expf(g0.x - xm);
    if (xmt > -INFINITY) {
      double sc = (double)expf(g0.z - xmt);
      sums[1] += (double)g0.w * sc;
      sums[2] += (double)g1.x * sc;
    }
  }
  if (tid < kCombineChunk) scale_buf[tid] = my_scale;  // rescale factors of the first chunk
  block_reduce_n(sums, OpAddD(), 0.0, red);            // (its barriers also publish scale_buf)
  stamp_row(trace_row, 8);
  // weighted-sum numerators
  const int n_vec = E_pad / 4;
  const int n_seg = max(1, min(kMaxSegments, nt / n_vec));
  const int seg = tid / n_vec, vc = tid - seg * n_vec;
  if (n_vec <= nt) {
    const bool worker = seg < n_seg;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b0 = 0; b0 < n; b0 += kCombineChunk) {
      const int nb = min(kCombineChunk, n - b0);
      if (b0 > 0 || nt < kCombineChunk) {
        __syncthreads();
        for (int i = tid; i < nb; i += nt) scale_buf[i] = expf(parts[(size_t)(b0 + i) * P] - xm);
        __syncthreads();
      }
      if (worker) {
        const int per = (nb + n_seg - 1) / n_seg, lo = seg * per, hi = min(nb, lo + per);
        const float* col = parts + (size_t)b0 * P + kPartialHeader + 4 * vc;
        int b = lo;
        for (; b + 8 <= hi; b += 8) {  // 8 x 16 B in flight per thread
          float4 x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = *reinterpret_cast<const float4*>(col + (size_t)(b + j) * P);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const double sc = (double)scale_buf[b + j];
            acc[0] += (double)x[j].x * sc;
            acc[1] += (double)x[j].y * sc;
            acc[2] += (double)x[j].z * sc;
            acc[3] += (double)x[j].w * sc;
          }
        }
        if (b < hi) {  // remainder (< 8 rows): still issued together
          float4 x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            x[j] = (b + j < hi) ? *reinterpret_cast<const float4*>(col + (size_t)(b + j) * P)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (b + j < hi) {
              const double sc = (double)scale_buf[b + j];
              acc[0] += (double)x[j].x * sc;
              acc[1] += (double)x[j].y * sc;
              acc[2] += (double)x[j].z * sc;
              acc[3] += (double)x[j].w * sc;
            }
        }
      }
    }
    if (worker)
#pragma unroll
      for (int j = 0; j < 4; ++j) seg_buf[(size_t)seg * E_pad + 4 * vc + j] = acc[j];
    __syncthreads();
    stamp_row(trace_row, 9);
    for (int e = tid; e < E_pad; e += nt) {
      double a = 0.0;
      for (int sgm = 0; sgm < n_seg; ++sgm) a += seg_buf[(size_t)sgm * E_pad + e];
      N[e] = a;
    }
  } else {  // more float4 columns than threads (T*du > 4*blockDim): each thread walks its entries
    for (int e = tid; e < E_pad; e += nt) N[e] = 0.0;
    for (int b0 = 0; b0 < n; b0 += kCombineChunk) {
      const int nb = min(kCombineChunk, n - b0);
      __syncthreads();
      for (int i = tid; i < nb; i += nt) scale_buf[i] = expf(parts[(size_t)(b0 + i) * P] - xm);
      __syncthreads();
      for (int e = tid; e < E_pad; e += nt) {
        double a = N[e];
        for (int b = 0; b < nb; ++b)
          a += (double)parts[(size_t)(b0 + b) * P + kPartialHeader + e] * (double)scale_buf[b];
        N[e] = a;
      }
    }
  }
  if (tid == 0) {
    out->xmax = xm;
    out->xmax_tau = xmt;
    out->cmin = -mx[2];
    out->cmax = mx[3];
    out->S = sums[0];
    out->S_tau = sums[1];
    out->Sc_tau = sums[2];
  }
  __syncthreads();
}

// MPO temperature update (mppi.py:387-398) in the closed form that reproduces
// torch's fp32 autograd, see oracle/mppi_oracle.py:mpo_gradient_device_form.
__device__ inline void mpo_update(const SolveParams& p, const Combined& c, DeviceScalars* sc) {
  float rho = sc->rho;
  float tau = (rho > 20.0f) ? rho : log1pf(expf(rho));  // softplus, torch threshold 20
  float lse32 = __fadd_rn((float)log(c.S_tau), c.xmax_tau);
  double term1 = (double)__fadd_rn(p.mpo_epsilon, lse32);
  double term2 = exp((double)c.xmax_tau - (double)lse32) * c.Sc_tau / (double)tau;
  double g = (term1 + term2) / (1.0 + exp(-(double)rho));
  int step = sc->adam_step + 1;
  double m = sc->adam_m + (g - sc->adam_m) * (1.0 - 0.9);
  double v = 0.999 * sc->adam_v + (1.0 - 0.999) * g * g;
  double bc1 = 1.0 - pow(0.9, (double)step), bc2 = 1.0 - pow(0.999, (double)step);
  float nrho = (float)((double)rho - (0.2 / bc1) * m / (sqrt(v) / sqrt(bc2) + 1e-8));
  sc->adam_step = step;
  sc->adam_m = m;
  sc->adam_v = v;
  sc->rho = nrho;
  sc->lambda = (double)expf(nrho);  // torch.exp(log_temperature).item()
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}

// Where finish_solve stores its results: the solve's real outputs and carried state, or - for the finisher
// block's warm-up pass (see solve_kernel) - dummy targets of the same sizes in global memory.
struct FinishOut {
  float *action_out, *state_seq_out, *prev_action, *history, *nominal_snapshot, *state_snapshot;
  DeviceScalars* sc;
  bool dry;
};
__host__ __device__ inline size_t dry_scratch_floats(int E_pad, int T, int ds, int du) {
  return (size_t)3 * E_pad + (size_t)(T + 1) * ds + (size_t)(T > 1 ? (T - 1) * du : 1) + 16 +
         (sizeof(DeviceScalars) + 3) / 4 + 16;
}
template <class M>
__device__ __forceinline__ FinishOut real_outputs(const SolveParams& p) {
  return FinishOut{p.action_out, p.state_seq_out, p.prev_action, p.history, p.nominal_snapshot, p.state_snapshot,
                   p.sc, false};
}
template <class M>
__device__ __forceinline__ FinishOut dry_outputs(const SolveParams& p) {
  float* q = p.dry_scratch;
  FinishOut o;
  o.sc = reinterpret_cast<DeviceScalars*>(q);  // (cudaMalloc base: aligned for the doubles inside)
  q += (sizeof(DeviceScalars) + 3) / 4 + ((sizeof(DeviceScalars) + 3) / 4) % 2;
  o.action_out = q;
  q += p.E_pad;
  o.prev_action = q;
  q += p.E_pad;
  o.nominal_snapshot = q;
  q += p.E_pad;
  o.state_seq_out = q;
  q += (p.T + 1) * M::DS;
  o.history = q;
  q += (p.T > 1 ? (p.T - 1) * M::DU : 1);
  o.state_snapshot = q;
  o.dry = true;
  return o;
}

// Everything after the weighted sum (mppi.py:381-458). Runs in ONE block.
// smem: opt[E], y[(2T-1)*du] floats supplied by the caller. Dry pass only: `poll_counter` / `poll_need` /
// `poll_flag` - when the counter has reached `poll_need` the real work is ready and the remaining stages of the
// warm-up are skipped. (Plain arguments, not a callable: the dry and the real pass must run the SAME
// instantiation of this function, or the warm-up would warm a copy of the code the real pass never executes.)
// `state` / `mp`: the solve's initial state and the model parameters - the finisher block hands in copies it
// made in shared memory while the workers were rolling (`prepared`: it also took the nominal / state snapshots
// and loaded the SG history into ybuf), so that nothing here waits on global or parameter memory.
template <class M>
__device__ __noinline__ void finish_solve(const SolveParams& p, const FinishOut& o, const Combined& c, const double* N,
                                          float* opt, float* ybuf, float* tail, bool prepared, const float* state,
                                          const ModelParams* mp, const unsigned* poll_counter = nullptr,
                                          unsigned poll_need = 0, int* poll_flag = nullptr) {
  // the rollout of the optimal sequence only needs the model parameters (dynamics never read the maps
  // or the reference path); a local context keeps the caller's register-resident one from escaping
  typename M::Ctx ctx{};
  if constexpr (uses_params<M>()) ctx.p = mp;
  constexpr int DS = M::DS, DU = M::DU;
  const int tid = threadIdx.x, nt = blockDim.x, T = p.T, E = p.E;
  const int H = (T - 1) * DU;
  const double rS = 1.0 / c.S;  // one division per thread instead of one per entry
  for (int e = tid; e < E; e += nt) {
    float r = (float)(N[e] * rS);  // sum_k softmax_k * u_k  (mppi.py:381-384)
    opt[e] = r;
    ybuf[H + e] = r;
  }
  if (!prepared)
    for (int i = tid; i < H; i += nt) ybuf[i] = p.history[i];
  __syncthreads();
  if (p.use_sg) {  // mppi.py:423-443, 598-620
    const int W = p.sg_window, pad = W / 2, n = 2 * T - 1;
    for (int e = tid; e < E; e += nt) {
      int t = e / DU, d = e - t * DU;
      int o2 = (T - 1) + t;  // position in the prolonged sequence
      float acc = 0.0f;
      for (int j = 0; j < W; ++j) {
        int q = o2 + j;  // index into the padded signal
        int i = (q < pad) ? (pad - 1 - q) : ((q < pad + n) ? (q - pad) : (n - 1 - (q - pad - n)));
        acc = acc + p.sg_coeffs[j] * ybuf[i * DU + d];
      }
      opt[e] = acc;
    }
    __syncthreads();
  }
  // optimal-trajectory rollout (mppi.py:448-449, 508-524) first: it is the long pole; the stores of the
  // carried state below are issued by warps that have no part in it (or after it)
  if (!o.dry) stamp(p, 7);
  if (o.dry && poll_counter) {  // uniform over the block
    __syncthreads();
    if (tid == 0) *poll_flag = ld_acquire_gpu(poll_counter) >= poll_need ? 1 : 0;
    __syncthreads();
    if (*poll_flag) return;
  }
  // the carried state: plain stores (and two scalar reads), done by threads (first, first + stride, ...)
  auto carry = [&](int first, int stride) {
    for (int e = first; e < E; e += stride) {
      float a = opt[e];
      o.action_out[e] = a;
      if (!prepared) o.nominal_snapshot[e] = p.prev_action[e];
      o.prev_action[e] = a;  // warm start, no time shift (mppi.py:452)
    }
    for (int i = first; i < H; i += stride)  // history = cat(history[1:], opt[0]) (mppi.py:455-458)
      o.history[i] = (i < H - DU) ? ybuf[i + DU] : opt[i - (H - DU)];
    if (first == 0) {
      DeviceScalars* sc = o.sc;
      sc->lambda_used = sc->lambda;
      sc->S = c.S;
      sc->xmax = c.xmax;
      sc->cmin = c.cmin;
      sc->cmax = c.cmax;
      if (p.lambda_mode == kLamMPO) mpo_update(p, c, sc);
    }
    if (!prepared && first < DS) o.state_snapshot[first] = state[first];
  };
  if constexpr (M::kParallelTail) {
    // the rollout's spare warps run carry() beside its roles (nothing waits on those stores)
    M::rollout_block(ctx, state, opt, T, o.state_seq_out, tail,
                     (p.trace && !o.dry) ? p.trace + (size_t)blockIdx.x * kTraceSlots : nullptr, carry);
  } else {
    if (tid == 0) {
      float s[DS], seen[DS], u[DU];
#pragma unroll
      for (int i = 0; i < DS; ++i) s[i] = state[i];
      for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int d = 0; d < DU; ++d) u[d] = opt[t * DU + d];
        M::step(ctx, s, u, seen);
#pragma unroll
        for (int i = 0; i < DS; ++i) o.state_seq_out[t * DS + i] = seen[i];
      }
#pragma unroll
      for (int i = 0; i < DS; ++i) o.state_seq_out[T * DS + i] = s[i];
    }
    carry(tid, nt);
  }
}

// ---------------------------------------------------------------------------
// fused shard exchange over peer memory
// ---------------------------------------------------------------------------
// Mailbox of one rank: [2 parities][kMaxPeers][P] 8-byte words (payload bits, sequence flag), then the
// [2 parities][kMaxPeers] arrival words of the device-side rank barrier (mppi_p2p_barrier).
__host__ __device__ inline size_t mailbox_barrier_offset(int P) { return (size_t)2 * kMaxPeers * P * 2; }
__host__ __device__ inline size_t mailbox_floats(int P) { return mailbox_barrier_offset(P) + 2 * kMaxPeers; }

// One 8-byte store / load that cannot tear (the LL scheme NCCL's low-latency protocol is built on): the payload
// word and the sequence number it belongs to travel together, so the receiver needs no separate flag, the sender
// no system-scope fence between data and flag.
__device__ __forceinline__ void st_ll(uint2* ptr, unsigned data, unsigned flag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(ptr), "r"(data), "r"(flag) : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const uint2* ptr) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(ptr) : "memory");
  return v;
}

// Finisher block of every rank: push this shard's partial into every rank's mailbox (peer stores over NVLink /
// NVSwitch), every word tagged with the solve's sequence number, then poll the own mailbox until every rank's
// words carry that number and leave the gathered partials in gather_scratch. No fence, no flag round trip: one
// NVLink one-way latency. Double-buffered by the parity of the sequence number: a rank can only reach solve s+2
// after every rank's words of s+1 arrived, i.e. after every rank finished reading solve s.
__device__ inline bool exchange_partials(const SolveParams& p, const Combined& c, const double* N) {
  const int tid = threadIdx.x, nt = blockDim.x, G = p.p2p_world, P = p.P;
  stamp(p, 16);
  const unsigned seq = p.p2p_seq, parity = seq & 1u;
  const size_t slot = ((size_t)parity * kMaxPeers + p.p2p_rank) * P;
  for (int e = tid; e < P; e += nt) {
    float v;
    switch (e) {
      case 0: v = c.xmax; break;
      case 1: v = (float)c.S; break;
      case 2: v = c.xmax_tau; break;
      case 3: v = (float)c.S_tau; break;
      case 4: v = (float)c.Sc_tau; break;
      case 5: v = c.cmin; break;
      case 6: v = c.cmax; break;
      case 7: v = 0.0f; break;
      default: v = (e - kPartialHeader < p.E) ? (float)N[e - kPartialHeader] : 0.0f;
    }
    const unsigned bits = __float_as_uint(v);
    for (int r = 0; r < G; ++r) st_ll(reinterpret_cast<uint2*>(p.peer_mailbox[r]) + slot + e, bits, seq);
  }
  __shared__ int timed_out;
  if (tid == 0) timed_out = 0;
  __syncthreads();
  stamp(p, 17);  // this rank's partial is on its way to every peer
  const uint2* mine = reinterpret_cast<const uint2*>(p.peer_mailbox[p.p2p_rank]) + (size_t)parity * kMaxPeers * P;
  const long long t0 = clock64();
  for (int i = tid; i < G * P; i += nt) {
    uint2 w = ld_ll(mine + i);
    while (w.y != seq) {
      if (clock64() - t0 > 4000000000LL) {  // ~2 s: a peer never arrived - do not hang the GPU
        timed_out = 1;
        break;
      }
      __nanosleep(40);
      w = ld_ll(mine + i);
    }
    p.gather_scratch[i] = __uint_as_float(w.x);
  }
  __threadfence();
  __syncthreads();
  if (timed_out) {
    if (tid == 0 && p.error_flag) {
      *reinterpret_cast<volatile int*>(p.error_flag) = (int)seq;  // mapped host memory: which solve failed
      __threadfence_system();
    }
    return false;
  }
  stamp(p, 18);  // every peer's words have arrived (and are gathered)
  stamp(p, 19);
  return true;
}

// Device-side barrier across the ranks of a fused sharded solver: rank i stores the barrier's sequence number into
// slot [parity][i] of EVERY rank's mailbox and waits until all slots of its own mailbox carry it. All ranks leave
// within one NVLink latency of each other (a collective's completion times differ by several hops), which is
// what a benchmark needs to start a timed step on every GPU together. Bounded wait like the exchange.
struct BarrierParams {
  unsigned* peer_slots[kMaxPeers];  // rank r's [2][kMaxPeers] arrival words
  int world, rank;
  unsigned seq;
  int* error_flag;
};
__global__ void p2p_barrier_kernel(const __grid_constant__ BarrierParams b) {
  const int r = threadIdx.x;
  if (r >= b.world) return;
  const unsigned parity = b.seq & 1u;
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(b.peer_slots[r] + parity * kMaxPeers + b.rank), "r"(b.seq)
               : "memory");
  const unsigned* mine = b.peer_slots[b.rank] + parity * kMaxPeers + r;
  const long long t0 = clock64();
  unsigned v;
  do {
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (v != b.seq && clock64() - t0 > 4000000000LL) {
      if (b.error_flag) *reinterpret_cast<volatile int*>(b.error_flag) = -(int)(b.seq & 0x7fffffffu) - 1;
      break;
    }
  } while (v != b.seq);
}

// ---------------------------------------------------------------------------
// the solve kernel
// ---------------------------------------------------------------------------
// A sharded solve whose peer exchange timed out has no result: the outputs become NaN (never stale or
// uninitialised memory) and the error flag - mapped host memory - tells the host which solve failed.
template <class M>
__device__ __forceinline__ void poison_outputs(const SolveParams& p) {
  const float nan = __int_as_float(0x7fc00000);
  for (int e = threadIdx.x; e < p.E; e += blockDim.x) p.action_out[e] = nan;
  for (int i = threadIdx.x; i < (p.T + 1) * M::DS; i += blockDim.x) p.state_seq_out[i] = nan;
}

// Loop constants held in ordinary registers: ptxas otherwise re-materialises kernel parameters from the
// constant bank inside the pass-1 loop (one LDCU issue slot each, ~17 per step on the racing model).
// A warp shuffle of the (warp-uniform) value is opaque to ptxas, so it has to stay live in a register;
// values are unchanged.
__device__ __forceinline__ float pin(float v) {  // all lanes hold the same v; must be called by the full warp
  return __shfl_sync(kFullMask, v, 0);
}
__device__ __forceinline__ int pin(int v) {
  return __shfl_sync(kFullMask, v, 0);
}
__device__ __forceinline__ void pin_map(MapView& m) {
  m.saddr = (unsigned)pin((int)__cvta_generic_to_shared(m.bits));
  m.words = pin(m.words);
  m.cell.c = pin(m.cell.c);
  m.cell.r = pin(m.cell.r);
  m.ox = pin(m.ox);
  m.oy = pin(m.oy);
}

// Loop constants of one thread, pinned by the whole warp before the per-sample branch.
template <class M>
struct LoopConsts {
  typename M::Ctx ctx;
  float sigma[M::DU], lo[M::DU], hi[M::DU];
  int T;
};
template <class M>
__device__ __forceinline__ void pin_loop_consts(const SolveParams& p, const typename M::Ctx& ctx, LoopConsts<M>& lc) {
  lc.ctx = ctx;
  constexpr int kHot = sizeof(lc.ctx.hv) / sizeof(float);
#pragma unroll
  for (int i = 0; i < kHot; ++i) lc.ctx.hv[i] = pin(ctx.p->v[i]);
  if constexpr (M::kHasHotFlags) lc.ctx.hflags = pin(ctx.p->flags);
  if constexpr (M::kMaps == 1) pin_map(lc.ctx.map);
  if constexpr (M::kMaps == 2) {
    pin_map(lc.ctx.obstacle);
    pin_map(lc.ctx.lane);
  }
#pragma unroll
  for (int d = 0; d < M::DU; ++d) {
    lc.sigma[d] = pin(p.sigma[d]);
    lc.lo[d] = pin(p.u_min[d]);
    lc.hi[d] = pin(p.u_max[d]);
  }
  lc.T = pin(p.T);
}

// One sample per thread, the literal transcription (general loop): rollout + stage costs + terminal cost
// (mppi.py:280-336). `nominal` = this sample's mean sequence (the warm start, or zeros for the exploration
// tail, mppi.py:266-275).
template <class M, bool kInject>
__device__ __forceinline__ float rollout_cost(const SolveParams& p, const typename M::Ctx& ctx, const float* nominal,
                                              uint32_t k_lo, uint32_t k_hi, long long k_local) {
  constexpr int DS = M::DS, DU = M::DU, SPC = Chunking<DU>::kStepsPerChunk;
  const int T = p.T;
  float s[DS], seen[DS];
  {
    const float* state = state_of(p);
#pragma unroll
    for (int i = 0; i < DS; ++i) s[i] = state[i];
  }
  float u[DU], up[DU], upp[DU];
#pragma unroll
  for (int d = 0; d < DU; ++d) up[d] = upp[d] = 0.0f;
  float total = 0.0f;
  const float* nz = kInject ? (p.noise + (size_t)k_local * T * DU) : nullptr;
  for (int t0 = 0, chunk = 0; t0 < T; t0 += SPC, ++chunk) {
    float z[4];
    if (!kInject) normal4(p.key, k_lo, k_hi, (uint32_t)chunk, z);
#pragma unroll
    for (int j = 0; j < SPC; ++j) {
      const int t = t0 + j;
      if (t < T) {
#pragma unroll
        for (int d = 0; d < DU; ++d) {
          float eps = kInject ? nz[t * DU + d] : p.sigma[d] * z[j * DU + d];
          u[d] = clampf(nominal[t * DU + d] + eps, p.u_min[d], p.u_max[d]);  // == perturbed_entry
        }
        float pu[DU];  // info["prev_action"]: U[:, max(t-1, 0)]  (mppi.py:299-304)
#pragma unroll
        for (int d = 0; d < DU; ++d) pu[d] = (t == 0) ? u[d] : up[d];
        M::step(ctx, s, u, seen);                      // S[:, t+1] = dynamics(S[:, t], U[:, t])   (mppi.py:282-286)
        total = total + M::cost(ctx, seen, u, pu, t);  // stage cost on S[:, t]  (mppi.py:307-311)
#pragma unroll
        for (int d = 0; d < DU; ++d) {
          upp[d] = up[d];
          up[d] = u[d];
        }
      }
    }
  }
  // terminal: state S[:, T], zero action, stale t = T-1 and prev_action = U[:, T-2] (mppi.py:318-328)
  float zero[DU], pa[DU];
#pragma unroll
  for (int d = 0; d < DU; ++d) {
    zero[d] = 0.0f;
    pa[d] = (T >= 2) ? upp[d] : up[d];
  }
  return total + M::cost(ctx, s, zero, pa, T - 1);  // mppi.py:333-336
}

// One sample per thread, the bounded loop (racing / navigation2d): rollout_cost<> with the bounded helpers
// substituted (each proven bit-identical over its whole input range, see mppi_selftest), every loop constant
// taken from the pinned register copies and whole sampler chunks run without the per-step `t < T` guard.
// Chosen by the host when the samples are too few to keep two paired warps per scheduler busy (a sharded solve
// on several GPUs, the K ~ 4000 of the reference's examples): half the dependent chain per warp, twice the warps.
template <class M>
__device__ __forceinline__ float rollout_cost_bounded(const SolveParams& p, const LoopConsts<M>& lc,
                                                      const float* nominal, uint32_t k_lo, uint32_t k_hi) {
  constexpr int DS = M::DS, DU = M::DU, SPC = Chunking<DU>::kStepsPerChunk;
  const int T = lc.T;
  float s[DS], seen[DS];
  {
    const float* state = state_of(p);
#pragma unroll
    for (int i = 0; i < DS; ++i) s[i] = state[i];
  }
  float u[DU], up[DU], upp[DU];
#pragma unroll
  for (int d = 0; d < DU; ++d) up[d] = upp[d] = 0.0f;
  float total = 0.0f;
  auto one_step = [&](int t, const float* ez) {
#pragma unroll
    for (int d = 0; d < DU; ++d) u[d] = clampf(nominal[t * DU + d] + lc.sigma[d] * ez[d], lc.lo[d], lc.hi[d]);
    float pu[DU];  // info["prev_action"]: U[:, max(t-1, 0)]  (mppi.py:299-304)
#pragma unroll
    for (int d = 0; d < DU; ++d) pu[d] = (t == 0) ? u[d] : up[d];
    M::template step<true>(lc.ctx, s, u, seen);
    total = total + M::template cost<true>(lc.ctx, seen, u, pu, t);
#pragma unroll
    for (int d = 0; d < DU; ++d) {
      upp[d] = up[d];
      up[d] = u[d];
    }
  };
  int t0 = 0, chunk = 0;
  for (; t0 + SPC <= T; t0 += SPC, ++chunk) {
    float z[4];
    normal4(p.key, k_lo, k_hi, (uint32_t)chunk, z);
#pragma unroll
    for (int j = 0; j < SPC; ++j) one_step(t0 + j, z + j * DU);
  }
  if (t0 < T) {
    float z[4];
    normal4(p.key, k_lo, k_hi, (uint32_t)chunk, z);
#pragma unroll
    for (int j = 0; j < SPC; ++j)
      if (t0 + j < T) one_step(t0 + j, z + j * DU);
  }
  float zero[DU], pa[DU];
#pragma unroll
  for (int d = 0; d < DU; ++d) {
    zero[d] = 0.0f;
    pa[d] = (T >= 2) ? upp[d] : up[d];
  }
  return total + M::template cost<true>(lc.ctx, s, zero, pa, T - 1);  // mppi.py:333-336
}

// Two samples per thread, the bounded loop (racing / navigation2d): every fp32 add / mul / fma of the two
// rollouts is ONE packed instruction (P2; sm_100 FADD2 / FMUL2 / FFMA2, each lane rounded like the scalar
// op), the operations and their order per sample are those of the general loop with the bounded helpers
// substituted (each proven bit-identical over its whole input range, see mppi_selftest), so costs are
// bit-identical to rollout_cost<>. All loop constants come from the pinned register copies (lc); whole
// sampler chunks run without the per-step `t < T` guard (a trailing partial chunk keeps it); the
// `prev_action` of step 0 is the action itself (mppi.py:299-304), so `up` starts as u_0 instead of a per-step
// select.
template <class M>
struct PairRoll {  // two samples' rollout state, one P2 per quantity
  P2 s[M::DS], u[M::DU], up[M::DU], upp[M::DU];
  P2 total;
  const float *nom_a, *nom_b;  // each sample's mean sequence (the warm start, or zeros for the exploration tail)
};
// clamp(mean + sigma * eps) of both samples for step t (mppi.py:266-275)
template <class M>
__device__ __forceinline__ void pair_controls(const LoopConsts<M>& lc, const PairRoll<M>& r, int t, P2 e0, P2 e1,
                                              P2 (&out)[2]) {
  const float2 na = *reinterpret_cast<const float2*>(r.nom_a + 2 * t);
  const float2 nb = *reinterpret_cast<const float2*>(r.nom_b + 2 * t);
  out[0] = clamp2(P2(na.x, nb.x) + lc.sigma[0] * e0, lc.lo[0], lc.hi[0]);
  out[1] = clamp2(P2(na.y, nb.y) + lc.sigma[1] * e1, lc.lo[1], lc.hi[1]);
}
// S[:, t+1] = dynamics(S[:, t], U[:, t]); stage cost on S[:, t] with prev_action = U[:, t-1]  (mppi.py:282-311)
template <class M>
__device__ __forceinline__ void pair_step(const LoopConsts<M>& lc, PairRoll<M>& r, int t, P2 e0, P2 e1) {
  P2 seen[M::DS];
  pair_controls<M>(lc, r, t, e0, e1, r.u);
  M::step_pair(lc.ctx, r.s, r.u, seen);
  r.total = r.total + M::cost_pair(lc.ctx, seen, r.u, r.up, t);
#pragma unroll
  for (int d = 0; d < M::DU; ++d) {
    r.upp[d] = r.up[d];
    r.up[d] = r.u[d];
  }
}
template <class M>
__device__ __forceinline__ void rollout_cost_pair(const SolveParams& p, const LoopConsts<M>& lc, const float* nom_a,
                                                  const float* nom_b, uint32_t ka_lo, uint32_t ka_hi, uint32_t kb_lo,
                                                  uint32_t kb_hi, float (&cost_out)[2]) {
  constexpr int DS = M::DS, DU = M::DU;
  static_assert(DU == 2, "the paired loop is written for two controls per step (one sampler chunk = two steps)");
  const int T = lc.T;
  PairRoll<M> r;
  r.nom_a = nom_a;
  r.nom_b = nom_b;
  r.total = P2(0.0f);
  {
    const float* state = state_of(p);
#pragma unroll
    for (int i = 0; i < DS; ++i) r.s[i] = P2(state[i]);
  }
  // one sampler chunk = two steps; the loop body exists once (the instruction footprint matters: bench.py
  // starts every solve with a cold L2, i.e. cold instruction fetches)
  for (int t0 = 0, chunk = 0; t0 < T; t0 += 2, ++chunk) {
    P2 z[4];
    normal4_pair(p.key, ka_lo, ka_hi, kb_lo, kb_hi, (uint32_t)chunk, z);
    if (chunk == 0) {  // prev_action at t = 0 is the action itself (mppi.py:299-304)
      pair_controls<M>(lc, r, 0, z[0], z[1], r.up);
#pragma unroll
      for (int d = 0; d < DU; ++d) r.upp[d] = r.up[d];
    }
    pair_step<M>(lc, r, t0, z[0], z[1]);
    if (t0 + 1 < T) pair_step<M>(lc, r, t0 + 1, z[2], z[3]);  // odd horizon: the last chunk covers one step
  }
  // terminal: state S[:, T], zero action, stale t = T-1 and prev_action = U[:, T-2] (mppi.py:318-328)
  P2 zero[DU], pa[DU];
#pragma unroll
  for (int d = 0; d < DU; ++d) {
    zero[d] = P2(0.0f);
    pa[d] = (T >= 2) ? r.upp[d] : r.up[d];
  }
  r.total = r.total + M::cost_pair(lc.ctx, r.s, zero, pa, T - 1);  // mppi.py:333-336
  cost_out[0] = r.total.v.x;
  cost_out[1] = r.total.v.y;
}

__device__ __forceinline__ void stamp(const SolveParams& p, int slot) {
  if (p.trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[(size_t)blockIdx.x * kTraceSlots + slot] = t;
  }
}

template <class M>
__host__ __device__ constexpr int tail_per_step() {
  if constexpr (M::kParallelTail)
    return M::kTailScratchPerStep;
  else
    return 0;
}

// The finisher block of solve_kernel (block 0; the workers are blocks 1 .. gridDim.x - 1).
//  1. warm-up: combine + finish + optimal-trajectory rollout on dummy data with dummy targets. The epilogue is
//     ~25 kB of code that runs ONCE per solve in one block; executed cold (bench.py flushes L2 between solves;
//     a real control loop runs other work in between) every taken branch is an instruction-cache miss that
//     goes to DRAM, which made the single-block tail 31 us of a 70 us solve. The warm-up overlaps pass 1 of the
//     workers and is abandoned between stages as soon as every worker has delivered its partial.
//  2. wait for the workers' tickets (ld.acquire.gpu on the counter; bounded).
//  3. the real epilogue: partials -> shared memory (one bulk copy), combine, [peer exchange], finish.
template <class M>
__device__ __noinline__ void finisher_block(const SolveParams& p, const SmemLayout& L, unsigned char* smem) {
  constexpr int DU = M::DU;
  const int tid = threadIdx.x;
  const unsigned n_workers = gridDim.x - 1;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  void* red = smem + L.red_off;
  int* misc = reinterpret_cast<int*>(smem + L.misc_off);
  // the group-accumulator region of the workers is the finish scratch here (make_layout sizes it for both):
  // N[E_pad] doubles | opt[E_pad] | y[2 E_pad] | rescale factors[256] | Combined | segment sums | tail rollout
  double* Nbuf = reinterpret_cast<double*>(smem + L.warpacc_off);
  float* opt = reinterpret_cast<float*>(Nbuf + p.E_pad);
  float* ybuf = opt + p.E_pad;
  float* scale_buf = ybuf + 2 * p.E_pad;
  Combined* comb = reinterpret_cast<Combined*>(scale_buf + 256);
  double* seg_buf = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(comb) + 64);
  float* tail = reinterpret_cast<float*>(seg_buf + (size_t)p.E_pad * kMaxSegments);
  float* stage = reinterpret_cast<float*>(smem + L.stage_off);
  const unsigned part_bytes = n_workers * (unsigned)p.P * 4u;
  const bool staged = part_bytes <= L.stage_cap && part_bytes >= 2048u;

  stamp(p, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  // ---- 0. everything the epilogue reads besides the partials, fetched now: the carried SG history (finish_solve
  // reads it from ybuf; nobody writes it before the real finish), the solve's state and the model parameters
  // (shared-memory copies), and the snapshots get_top_samples re-rolls from (warm start, state)
  for (int i = tid; i < (p.T - 1) * DU; i += blockDim.x) ybuf[i] = p.history[i];
  ModelParams* mp_s = reinterpret_cast<ModelParams*>(misc + 16);
  float* state_s = reinterpret_cast<float*>(misc + 52);
  {
    const float* st = state_of(p);
    if (tid < M::DS) {
      const float v = st[tid];
      state_s[tid] = v;
      p.state_snapshot[tid] = v;
    }
    if (tid < 32) mp_s->v[tid] = p.mp.v[tid];
    if (tid == 32) mp_s->flags = p.mp.flags;
    for (int e = tid; e < p.E; e += blockDim.x) p.nominal_snapshot[e] = p.prev_action[e];
  }
  auto workers_done = [&]() -> bool {  // uniform over the block
    __syncthreads();
    if (tid == 0) misc[0] = ld_acquire_gpu(p.counter) >= n_workers ? 1 : 0;
    __syncthreads();
    return misc[0] != 0;
  };
  // ---- 1. warm-up pass
  if (p.dry_scratch && !workers_done()) {
    const float* parts = p.block_partials;  // (too many partials for shared memory: whatever global holds)
    if (staged) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (unsigned i = tid; i < part_bytes / 16u; i += blockDim.x) reinterpret_cast<float4*>(stage)[i] = z;
      __syncthreads();
      if (tid == 0) stage[1] = 1.0f;  // S of partial 0: a finite weight sum
      parts = stage;
    }
    __syncthreads();
    combine_partials(parts, (int)n_workers, p.P, p.E_pad, comb, Nbuf, scale_buf, red, seg_buf);
    if (!workers_done()) {
      const FinishOut dry = dry_outputs<M>(p);
      finish_solve<M>(p, dry, *comb, Nbuf, opt, ybuf, tail, true, state_s, mp_s, p.counter, n_workers, misc);
    }
  }
  stamp(p, 1);
  // ---- 2. every worker's partial is in global memory
  if (tid == 0) {
    const long long t0 = clock64();
    while (ld_acquire_gpu(p.counter) < n_workers) {
      __nanosleep(40);
      if (clock64() - t0 > 20000000000LL) __trap();  // ~10 s: a worker block died - fail instead of hanging
    }
  }
  __syncthreads();
  stamp(p, 2);
  // ---- 3. the real epilogue. The partials of all workers ([n_workers, P] floats, contiguous) come into
  // shared memory with ONE bulk copy when they fit (the landing zone overlays the workers' grid region).
  const float* parts = p.block_partials;
  if (staged) {
    if (tid == 0) {
      fence_proxy_async_all();  // generic-proxy writes (the workers' partials, made visible by their fences and
                                // the acquire above; this block's warm-up stores into the zone) before the
                                // async-proxy copy
      mbar_expect_tx(bar, part_bytes);
      bulk_g2s(stage, p.block_partials, part_bytes, bar);
    }
    mbar_wait(bar, 0);
    parts = stage;
  }
  __syncthreads();
  stamp(p, 3);
  combine_partials(parts, (int)n_workers, p.P, p.E_pad, comb, Nbuf, scale_buf, red, seg_buf,
                   p.trace ? p.trace + (size_t)blockIdx.x * kTraceSlots : nullptr);
  stamp(p, 5);
  const FinishOut out = real_outputs<M>(p);
  if (p.n_shards == 1) {
    finish_solve<M>(p, out, *comb, Nbuf, opt, ybuf, tail, true, state_s, mp_s);
  } else if (p.p2p_world > 0) {
    if (exchange_partials(p, *comb, Nbuf)) {
      combine_partials(p.gather_scratch, p.p2p_world, p.P, p.E_pad, comb, Nbuf, scale_buf, red, seg_buf);
      finish_solve<M>(p, out, *comb, Nbuf, opt, ybuf, tail, true, state_s, mp_s);
    } else {
      poison_outputs<M>(p);  // a peer never arrived: NaN outputs + the error flag, never stale memory
    }
  } else {
    for (int e = tid; e < p.E; e += blockDim.x) p.rank_partial[kPartialHeader + e] = (float)Nbuf[e];
    if (tid == 0) {
      float* q = p.rank_partial;
      q[0] = comb->xmax;
      q[1] = (float)comb->S;
      q[2] = comb->xmax_tau;
      q[3] = (float)comb->S_tau;
      q[4] = (float)comb->Sc_tau;
      q[5] = comb->cmin;
      q[6] = comb->cmax;
      q[7] = 0.0f;
    }
  }
  __syncthreads();
  stamp(p, 6);
  if (tid == 0) {
    *p.counter = 0u;
    if (p.host_done) {  // every output store of this block happened before the barrier above
      __threadfence_system();
      *reinterpret_cast<volatile unsigned*>(p.host_done) = p.host_done_seq;
    }
  }
}

// SPT = samples per thread. 2: the launch geometry of the paired bounded loop (host-selected when the model
// flags allow it and no noise is injected); thread `tid` of block `b` owns local samples
// b * 2 * blockDim + {tid, blockDim + tid}. If the solve's initial state fails the kernel-side range check the
// same launch runs the general loop once per sample (same results, slower).
//
// kGlobalMaps: occupancy grids too large for shared memory (the reference's lookup has no size limit,
// src/envs/obstacle_map_2d.py:168-200) stay in global memory (bit-packed, L2 resident) and are read by the general
// loop; a separate instantiation, so the staged kernels' shared-memory loads are untouched.
template <class M, bool kInject, int kMode, int SPT, bool kGlobalMaps = false>
__global__ void __launch_bounds__(SPT == 2 ? 256 : 512, 1) solve_kernel(const __grid_constant__ SolveParams p) {
  constexpr int DU = M::DU;
  static_assert(SPT == 1 || (M::kHasBounded && !kInject), "two samples per thread is the bounded sampler loop");
  static_assert(!kGlobalMaps || (M::kMaps > 0 && SPT == 1), "global-memory grids: map models, one sample per thread");
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
  const unsigned no_map_bytes[2] = {0u, 0u};
  const SmemLayout L = make_layout(M::kMaps, kGlobalMaps ? no_map_bytes : p.map_bytes, p.T, p.E_pad,
                                   p.prev_action_bytes, M::kRefPath, n_warps, tail_per_step<M>(), SPT, p.stage_bytes);
  // Block 0 is the FINISHER (kFused / kReduce): it never rolls samples. While the workers (blocks 1..) roll, it
  // runs the whole epilogue once on dummy data - which pulls the epilogue's code and tables into its SM's
  // caches - then waits for the workers' tickets and combines / finishes for real. kCosts has no epilogue.
  constexpr bool kHasFinisher = kMode != kCosts;
  if (kHasFinisher && blockIdx.x == 0) {
    finisher_block<M>(p, L, smem);
    return;
  }
  const int wid = (int)blockIdx.x - (kHasFinisher ? 1 : 0);  // worker id
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float* nominal = reinterpret_cast<float*>(smem + L.nominal_off);
  float* zero_nominal = reinterpret_cast<float*>(smem + L.zero_off);
  float* warp_acc = reinterpret_cast<float*>(smem + L.warpacc_off);
  void* red = smem + L.red_off;
  int* misc = reinterpret_cast<int*>(smem + L.misc_off);

  stamp(p, 0);
  // ---- stage the block's read-only inputs into shared memory (TMA bulk copies)
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    unsigned bytes = p.prev_action_bytes;
    if (kMode != kReduce) {
      if (!kGlobalMaps)
        for (int i = 0; i < M::kMaps; ++i) bytes += p.map_bytes[i];
      if (M::kRefPath && p.ref_bulk_ok && !p.inline_inputs) bytes += (unsigned)(p.T + 1) * 16;
    }
    mbar_expect_tx(bar, bytes);
    bulk_g2s(nominal, p.prev_action, p.prev_action_bytes, bar);
    if (kMode != kReduce) {
      if (!kGlobalMaps)
        for (int i = 0; i < M::kMaps; ++i) bulk_g2s(smem + L.map_off[i], p.map_bits[i], p.map_bytes[i], bar);
      if (M::kRefPath && p.ref_bulk_ok && !p.inline_inputs)
        bulk_g2s(smem + L.refraw_off, p.refpath, (unsigned)(p.T + 1) * 16, bar);
    }
  }
  if (M::kRefPath && kMode != kReduce && (!p.ref_bulk_ok || p.inline_inputs)) {
    float* raw = reinterpret_cast<float*>(smem + L.refraw_off);
    const float* src = refpath_of(p);
    for (int i = tid; i < (p.T + 1) * 4; i += blockDim.x) raw[i] = src[i];
  }
  for (int e = tid; e < p.E_pad; e += blockDim.x) zero_nominal[e] = 0.0f;  // mean of the exploration tail
  mbar_wait(bar, 0);
  __syncthreads();

  typename M::Ctx ctx;
  if constexpr (M::kMaps >= 1) {
    MapView mv[2];
    for (int i = 0; i < M::kMaps; ++i)
      mv[i] = MapView{kGlobalMaps ? p.map_bits[i] : reinterpret_cast<const uint32_t*>(smem + L.map_off[i]), p.map_W[i],
                      p.map_H[i], p.map_words[i], ExactDiv{p.map_cell[i], p.map_rcp[i], p.map_fastdiv[i]}, p.map_ox[i],
                      p.map_oy[i]};
    if constexpr (M::kMaps == 1) {
      ctx.map = mv[0];
    } else {
      ctx.obstacle = mv[0];
      ctx.lane = mv[1];
    }
  }
  if constexpr (uses_params<M>()) ctx.p = &p.mp;
  if constexpr (M::kRefPath) {
    // per-stage reference row: (x, y, sin yaw, cos yaw), v_target (racing.py:127-143)
    float4* ref4 = reinterpret_cast<float4*>(smem + L.ref4_off);
    float* refv = reinterpret_cast<float*>(smem + L.refv_off);
    if (kMode != kReduce) {
      const float* raw = reinterpret_cast<const float*>(smem + L.refraw_off);
      for (int t = tid; t < p.T; t += blockDim.x) {
        float sy, cy;
        sincosf(raw[t * 4 + 2], &sy, &cy);
        ref4[t] = make_float4(raw[t * 4 + 0], raw[t * 4 + 1], sy, cy);
        refv[t] = raw[t * 4 + 3];
      }
      __syncthreads();
    }
    ctx.ref = ref4;
    ctx.ref_v = refv;
  }

  stamp(p, 1);
  const long long block_k0 = (long long)wid * blockDim.x * SPT;
  long long k_local[SPT];
  bool active[SPT], zero_mean[SPT];
  uint32_t k_lo[SPT], k_hi[SPT];
#pragma unroll
  for (int j = 0; j < SPT; ++j) {
    k_local[j] = block_k0 + (long long)j * blockDim.x + tid;
    active[j] = k_local[j] < p.K;
    const long long k_global = p.k_offset + k_local[j];
    k_lo[j] = (uint32_t)k_global;
    k_hi[j] = (uint32_t)((unsigned long long)k_global >> 32);
    zero_mean[j] = k_global >= p.explore_threshold;
  }

  // ---- pass 1: rollout + cost -------------------------------------------------
  float cost[SPT];
#pragma unroll
  for (int j = 0; j < SPT; ++j) cost[j] = INFINITY;
  if (kMode == kReduce) {
#pragma unroll
    for (int j = 0; j < SPT; ++j)
      if (active[j]) cost[j] = p.costs[k_local[j]];
  } else {
    bool paired = false;  // uniform over the block: model flag + the solve's initial state
    if constexpr (SPT == 2) {
      paired = (p.mp.flags & kFlagBounded) && M::state_in_pair_bounds(ctx, state_of(p));
      if (paired) {
        LoopConsts<M> lc;
        pin_loop_consts<M>(p, ctx, lc);  // whole warps: before the per-sample branch
        if (active[0]) {                 // an inactive second sample (odd K) rolls a copy of the first
          const bool zb = active[1] ? zero_mean[1] : zero_mean[0];
          const uint32_t kb_lo = active[1] ? k_lo[1] : k_lo[0], kb_hi = active[1] ? k_hi[1] : k_hi[0];
          float c2[2];
          rollout_cost_pair<M>(p, lc, zero_mean[0] ? zero_nominal : nominal, zb ? zero_nominal : nominal, k_lo[0],
                               k_hi[0], kb_lo, kb_hi, c2);
          cost[0] = c2[0];
          if (active[1]) cost[1] = c2[1];
        }
      }
    }
    if constexpr (SPT == 1 && !kInject && M::kHasBounded && !kGlobalMaps) {
      // (uniform over the block) the bounded single-sample loop: same flag, the single loop's range check
      paired = (p.mp.flags & kFlagBounded) && M::state_in_bounds(ctx, state_of(p));
      if (paired) {
        LoopConsts<M> lc;
        pin_loop_consts<M>(p, ctx, lc);  // whole warps: before the per-sample branch
        if (active[0])
          cost[0] = rollout_cost_bounded<M>(p, lc, zero_mean[0] ? zero_nominal : nominal, k_lo[0], k_hi[0]);
      }
    }
    if (!paired) {
      // one sample after the other through the general loop; the per-sample values are recomputed from j
      // (no indexing of the register arrays by a run-time j, which would push them into local memory)
#pragma unroll 1
      for (int j = 0; j < SPT; ++j) {
        const long long kl = block_k0 + (long long)j * blockDim.x + tid, kg = p.k_offset + kl;
        float c = INFINITY;
        if (kl < p.K)
          c = rollout_cost<M, kInject>(p, ctx, (kg >= p.explore_threshold) ? zero_nominal : nominal, (uint32_t)kg,
                                       (uint32_t)((unsigned long long)kg >> 32), kl);
#pragma unroll
        for (int jj = 0; jj < SPT; ++jj) cost[jj] = (jj == j) ? c : cost[jj];
      }
    }
#pragma unroll
    for (int j = 0; j < SPT; ++j)
      if (active[j]) p.costs[k_local[j]] = cost[j];
  }
  if (kMode == kCosts) return;
  __syncthreads();
  stamp(p, 2);

  // ---- block-local softmax baseline (mppi.py:376; online-softmax form) -----------
  // x = -c / lambda is monotone in c, so max_k x_k = (-min_k c_k) / lambda exactly: one reduction gives the
  // baseline and the cost range, a second one the sums.
  const float lam = (float)p.sc->lambda;
  float mm[2] = {-INFINITY, -INFINITY};  // max(-c), max(c)
#pragma unroll
  for (int j = 0; j < SPT; ++j)
    if (active[j]) {
      mm[0] = fmaxf(mm[0], -cost[j]);
      mm[1] = fmaxf(mm[1], cost[j]);
    }
  block_reduce_n(mm, OpMax(), -INFINITY, red);
  const float cmin_b = -mm[0], cmax_b = mm[1];
  const float xmax_b = mm[0] / lam;
  float w[SPT];
  float sums[3] = {0.0f, 0.0f, 0.0f};  // S, S_tau, Sc_tau
  float xmt_b = -INFINITY;
  const bool mpo = p.lambda_mode == kLamMPO;
  float tau = 1.0f;
  if (mpo) {  // second softmax at tau = softplus(rho) for the MPO step
    const float rho = p.sc->rho;
    tau = (rho > 20.0f) ? rho : log1pf(expf(rho));
    xmt_b = mm[0] / tau;
  }
#pragma unroll
  for (int j = 0; j < SPT; ++j) {
    w[j] = active[j] ? expf((-cost[j]) / lam - xmax_b) : 0.0f;
    sums[0] += w[j];
    if (mpo && active[j]) {
      const float et = expf((-cost[j]) / tau - xmt_b);
      sums[1] += et;
      sums[2] += et * cost[j];
    }
  }
  block_reduce_n(sums, OpAddF(), 0.0f, red);
  const float S_b = sums[0], St_b = sums[1], Sct_b = sums[2];

  stamp(p, 3);
  // ---- pass 2: regenerate the perturbed controls of the samples that carry weight and reduce
  //      sum_k w_k u_k[t,d]. A zero weight contributes exactly nothing (the reference multiplies by
  //      it), so only samples with w != 0 are listed; the list is walked chunk-parallel: thread
  //      (g, c) owns the 4 entries of sampler chunk c and accumulates them over the listed samples
  //      i = g, g+G, ... in order, groups are then added in order -> deterministic, all warps busy,
  //      and the cost scales with the number of samples that matter, not with K.
  int2* list = reinterpret_cast<int2*>(smem + L.list_off);
  int* wcount = misc + 2;  // [SPT * n_warps]
  {
    unsigned m[SPT];
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      m[j] = __ballot_sync(kFullMask, w[j] != 0.0f);
      if (lane == 0) wcount[j * n_warps + warp] = __popc(m[j]);
    }
    __syncthreads();
    int off[SPT], n_active = 0;
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      off[j] = 0;
      for (int i = 0; i < n_warps; ++i) {
        const int c = wcount[j * n_warps + i];
        if (i < warp) off[j] += c;
        n_active += c;
      }
    }
    if (SPT == 2) {  // second-sample entries follow all first-sample entries
      int first = 0;
      for (int i = 0; i < n_warps; ++i) first += wcount[i];
      off[SPT - 1] += first;
    }
#pragma unroll
    for (int j = 0; j < SPT; ++j)
      if (w[j] != 0.0f)
        list[off[j] + __popc(m[j] & ((1u << lane) - 1u))] = make_int2(j * (int)blockDim.x + tid, __float_as_int(w[j]));
    __syncthreads();
    const int n_chunks = (p.E + 3) / 4;
    float* group_acc = warp_acc;  // [G, E_pad]
    const int G = (n_chunks <= (int)blockDim.x) ? max(1, min((int)blockDim.x / n_chunks, n_warps)) : 1;
    const int g = (n_chunks <= (int)blockDim.x) ? tid / n_chunks : 0;
    for (int c = (n_chunks <= (int)blockDim.x) ? tid - g * n_chunks : tid; c < n_chunks && g < G;
         c += (n_chunks <= (int)blockDim.x) ? n_chunks : (int)blockDim.x) {
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int i = g; i < n_active; i += G) {
        const int2 it = list[i];
        const float wi = __int_as_float(it.y);
        const long long kl = block_k0 + it.x, kg = p.k_offset + kl;
        const bool zm = kg >= p.explore_threshold;
        float z[4];
        if (!kInject) normal4(p.key, (uint32_t)kg, (uint32_t)((unsigned long long)kg >> 32), (uint32_t)c, z);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = c * 4 + j;
          if (e < p.E) {
            const int t = e / DU, d = e - t * DU;
            float eps = kInject ? p.noise[(size_t)kl * p.E + e] : p.sigma[d] * z[j];
            acc[j] = acc[j] + wi * perturbed_entry<DU>(p, nominal, zm, t, d, eps);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c * 4 + j < p.E_pad) group_acc[(size_t)g * p.E_pad + c * 4 + j] = acc[j];
    }
    __syncthreads();
    float* part = p.block_partials + (size_t)wid * p.P;
    for (int e = tid; e < p.E; e += blockDim.x) {
      float a = 0.0f;
      for (int gg = 0; gg < G; ++gg) a += group_acc[(size_t)gg * p.E_pad + e];
      part[kPartialHeader + e] = a;
    }
  }
  float* part = p.block_partials + (size_t)wid * p.P;
  if (tid == 0) {
    part[0] = xmax_b;
    part[1] = S_b;
    part[2] = xmt_b;
    part[3] = St_b;
    part[4] = Sct_b;
    part[5] = cmin_b;
    part[6] = cmax_b;
    part[7] = 0.0f;
  }

  // ---- done: publish the partial; the finisher block waits for every worker's ticket ----------
  __threadfence();
  __syncthreads();
  stamp(p, 4);
  if (tid == 0) atomicAdd(p.counter, 1u);
}

// Stage 3 of a sharded solve: combine the gathered shard partials and finish.
template <class M>
__global__ void __launch_bounds__(256, 1) finish_kernel(const __grid_constant__ SolveParams p, const float* parts,
                                                         int n) {
  extern __shared__ __align__(128) unsigned char smem[];
  double* Nbuf = reinterpret_cast<double*>(smem);
  float* opt = reinterpret_cast<float*>(Nbuf + p.E_pad);
  float* ybuf = opt + p.E_pad;
  float* scale_buf = ybuf + 2 * p.E_pad;
  Combined* comb = reinterpret_cast<Combined*>(scale_buf + 256);
  double* seg_buf = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(comb) + 64);
  float* tail = reinterpret_cast<float*>(seg_buf + (size_t)p.E_pad * kMaxSegments);
  void* red = smem + finish_scratch_core(p.E_pad, p.T, tail_per_step<M>());
  combine_partials(parts, n, p.P, p.E_pad, comb, Nbuf, scale_buf, red, seg_buf);
  finish_solve<M>(p, real_outputs<M>(p), *comb, Nbuf, opt, ybuf, tail, false, state_of(p), &p.mp);
}

__host__ __device__ inline unsigned finish_scratch_bytes(int E_pad, int T, int tail_per_step) {
  return finish_scratch_core(E_pad, T, tail_per_step) + kRedBytes;
}

// ---------------------------------------------------------------------------
// LBPS / ESSPS: lambda search on costs[K] (mppi.py:341-370, 526-566). One thread-block cluster; the control
// flow of scipy's bounded Brent / brentq runs in fp64 in warp 0 of every CTA (redundantly, identical values),
// which publishes each trial lambda to its block; every objective evaluation is a cluster-wide reduction
// (softmax_stats) in which all warps take part.
// ---------------------------------------------------------------------------
// ---- one objective evaluation of the lambda search -------------------------------------------------------
// w = softmax(-c / lambda) in fp32 like the reference; the three sums (sum e, sum e^2, sum e c with
// e = exp(x - xmax)) are formed in fp32 per thread and per warp, every warp stores its row straight into EVERY
// CTA's shared memory (distributed shared memory), ONE cluster barrier, and warp 0 of every CTA adds the
// kSearchCluster * warps rows in a fixed order - identical totals everywhere. fp64 is kept out of everything that
// runs in more than one warp: on this GPU a warp-wide fp64 instruction issues once per ~16 cycles per scheduler,
// and the round-1 form (fp64 accumulation per sample, the search bookkeeping redundantly in all 8192 threads)
// spent 3.5 us per evaluation there.
struct SearchStats {
  double S, S2, Sc;
};
constexpr int kSearchCluster = 8;  // CTAs of the lambda search (one thread-block cluster, portable size)
constexpr int kSearchMaxWarps = 32;

struct SearchShared {
  float4 rows[2][kSearchCluster * kSearchMaxWarps];  // [phase][cta * warps + warp] = (S, S2, Sc, -)
  float lam;                                         // next lambda, published by warp 0
  int done;
};

__device__ __forceinline__ void named_barrier_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

template <int kPerThread>
__device__ __forceinline__ SearchStats softmax_stats(const float* __restrict__ costs, long long n,
                                                     const float (&creg)[kPerThread > 0 ? kPerThread : 1], int n_mine,
                                                     float lam, float cmin, SearchShared* sh, int& phase) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank(), nblk = cluster.num_blocks();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const float xmax = (-cmin) / lam;
  // (-c) / lam per sample as q = x * r corrected by one residual step (r = RN(1 / lam)): the IEEE quotient in all
  // but rare last-bit cases, 3 instructions instead of the ~20 of a full division - the evaluation is issue
  // bound on the 8 SMs of the cluster, and a 1-ulp difference in a few exponents is far below the fp32
  // rounding of the sums the reference itself forms
  const float rl = 1.0f / lam;
  auto term = [&](float c, float& a0, float& a1, float& a2) {
    const float x = -c;
    float qd = x * rl;
    qd = fmaf(fmaf(-qd, lam, x), rl, qd);
    const float e = expf(qd - xmax);
    a0 += e;
    a1 = fmaf(e, e, a1);
    a2 = fmaf(e, c, a2);
  };
  float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f;
  if constexpr (kPerThread > 0) {
#pragma unroll
    for (int j = 0; j < kPerThread; ++j)
      if (j < n_mine) term(creg[j], v0, v1, v2);
  } else {
    const long long stride = (long long)nblk * blockDim.x;
    for (long long i = (long long)rank * blockDim.x + threadIdx.x; i < n; i += stride) term(costs[i], v0, v1, v2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v0 += __shfl_xor_sync(kFullMask, v0, o);
    v1 += __shfl_xor_sync(kFullMask, v1, o);
    v2 += __shfl_xor_sync(kFullMask, v2, o);
  }
  if (lane < (int)nblk) {
    float4* remote = cluster.map_shared_rank(&sh->rows[phase][0], lane);
    remote[rank * nw + warp] = make_float4(v0, v1, v2, 0.0f);
  }
  cluster.sync();
  SearchStats r{0.0, 0.0, 0.0};
  if (warp == 0) {  // (only warp 0 consumes the totals)
    // the top of the reduction tree is where fp32 rounding would show (a few large partial sums): the rows are
    // added as float-float pairs (error-free two-sum), i.e. to ~2^-45 - the per-thread and per-warp levels below
    // average their independent roundings out over thousands of partials
    const int n_rows = (int)nblk * nw;
    float hi[3] = {0.0f, 0.0f, 0.0f}, lo[3] = {0.0f, 0.0f, 0.0f};
    auto add = [](float& h, float& l, float xh, float xl) {  // (h, l) += (xh, xl)
      const float s = h + xh, bb = s - h;
      const float err = (h - (s - bb)) + (xh - bb);
      h = s;
      l += err + xl;
    };
    for (int i = lane; i < n_rows; i += 32) {
      const float4 t = sh->rows[phase][i];
      add(hi[0], lo[0], t.x, 0.0f);
      add(hi[1], lo[1], t.y, 0.0f);
      add(hi[2], lo[2], t.z, 0.0f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float oh = __shfl_xor_sync(kFullMask, hi[k], o), ol = __shfl_xor_sync(kFullMask, lo[k], o);
        add(hi[k], lo[k], oh, ol);
      }
    r = SearchStats{(double)hi[0] + (double)lo[0], (double)hi[1] + (double)lo[1], (double)hi[2] + (double)lo[2]};
  }
  phase ^= 1;
  return r;
}

struct SearchParams {
  const float* costs;
  long long n;
  int mode;  // kLamLBPS / kLamESSPS
  double lambda_min, lambda_max, lbps_delta, essps_target;
  DeviceScalars* sc;
};

__device__ inline double dsign(double x) { return (x > 0.0) - (x < 0.0); }

constexpr int kSearchThreadsReg = 512, kSearchPerThread = 16;  // register path: K <= 8 * 512 * 16 = 65536
constexpr int kSearchThreadsGlobal = 1024;

template <int kPerThread>
__global__ void __launch_bounds__(kPerThread > 0 ? kSearchThreadsReg : kSearchThreadsGlobal, 1)
    lambda_search_kernel(const SearchParams q) {
  // launched as ONE thread-block cluster of kSearchCluster CTAs (cudaLaunchAttributeClusterDimension)
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double red[2 * 32 * 3];
  __shared__ double xchg[2 * kSearchCluster * 4];
  __shared__ SearchShared sh;
  int phase = 0;
  float cmin = INFINITY, cmax = -INFINITY;
  float creg[kPerThread > 0 ? kPerThread : 1];
  int n_mine = 0;
  cluster.sync();  // every CTA of the cluster is resident before anyone touches remote shared memory
  {
    const long long stride = (long long)cluster.num_blocks() * blockDim.x;
    const long long first = (long long)cluster.block_rank() * blockDim.x + threadIdx.x;
    if constexpr (kPerThread > 0) {
#pragma unroll
      for (int j = 0; j < kPerThread; ++j) {
        const long long i = first + (long long)j * stride;
        creg[j] = (i < q.n) ? q.costs[i] : 0.0f;
        if (i < q.n) {
          n_mine = j + 1;
          cmin = fminf(cmin, creg[j]);
          cmax = fmaxf(cmax, creg[j]);
        }
      }
    } else {
      for (long long i = first; i < q.n; i += stride) {
        float c = q.costs[i];
        cmin = fminf(cmin, c);
        cmax = fmaxf(cmax, c);
      }
    }
    float mm[2] = {-cmin, cmax};
    block_reduce_n(mm, OpMax(), -INFINITY, red);
    if (threadIdx.x < cluster.num_blocks()) {
      double* remote = cluster.map_shared_rank(xchg, threadIdx.x) + (size_t)cluster.block_rank() * 4;
      remote[0] = (double)mm[0];
      remote[1] = (double)mm[1];
    }
    cluster.sync();
    float a = -INFINITY, b = -INFINITY;
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) {
      a = fmaxf(a, (float)xchg[r * 4 + 0]);
      b = fmaxf(b, (float)xchg[r * 4 + 1]);
    }
    cmin = -a;
    cmax = b;
  }
  const int n_threads = (int)blockDim.x;
  if (threadIdx.x >= 32) {
    // ---- helper warps: one cluster-wide evaluation per lambda that warp 0 publishes, until it is done
    for (;;) {
      named_barrier_sync(1, n_threads);
      if (sh.done) break;
      (void)softmax_stats<kPerThread>(q.costs, q.n, creg, n_mine, sh.lam, cmin, &sh, phase);
    }
    cluster.sync();  // no CTA may exit while a peer can still address its shared memory
    return;
  }
  // ---- warp 0: the search itself (all 32 lanes redundantly; only this warp touches the fp64 pipe)
  auto stats_at = [&](double lam) {
    if (threadIdx.x == 0) {
      sh.lam = (float)lam;
      sh.done = 0;
    }
    named_barrier_sync(1, n_threads);
    return softmax_stats<kPerThread>(q.costs, q.n, creg, n_mine, (float)lam, cmin, &sh, phase);
  };
  int evals = 0;
  double result;
  if (q.mode == kLamLBPS) {
    // J(lambda) = sum w c + (cmax - cmin) * sqrt((1-delta)/delta) / sqrt(ESS)   (mppi.py:534-557)
    const double range = (double)(float)(cmax - cmin);
    const double pen = sqrt((1.0 - q.lbps_delta) / q.lbps_delta);
    auto J = [&](double lam) {
      const SearchStats s = stats_at(lam);
      ++evals;
      // sum w^2 and sum w c as the fp32 values torch's .item() hands to Python, the rest in fp64 (mppi.py:526-557)
      double ess = 1.0 / (double)(float)(s.S2 / (s.S * s.S));
      double expected = (double)(float)(s.Sc / s.S);
      return expected + range * pen / sqrt(ess);
    };
    // scipy.optimize minimize_scalar(method="bounded"), xatol = 1e-5, maxiter = 500
    const double sqrt_eps = sqrt(2.2e-16), golden_mean = 0.5 * (3.0 - sqrt(5.0)), xatol = 1e-5;
    double a = q.lambda_min, b = q.lambda_max;
    double fulc = a + golden_mean * (b - a), nfc = fulc, xf = fulc, rat = 0.0, e = 0.0, x = xf;
    double fx = J(x);
    int num = 1;
    double ffulc = fx, fnfc = fx;
    double xm = 0.5 * (a + b), tol1 = sqrt_eps * fabs(xf) + xatol / 3.0, tol2 = 2.0 * tol1;
    while (fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
      bool golden = true;
      if (fabs(e) > tol1) {
        golden = false;
        double r = (xf - nfc) * (fx - ffulc);
        double qq = (xf - fulc) * (fx - fnfc);
        double pp = (xf - fulc) * qq - (xf - nfc) * r;
        qq = 2.0 * (qq - r);
        if (qq > 0.0) pp = -pp;
        qq = fabs(qq);
        r = e;
        e = rat;
        if (fabs(pp) < fabs(0.5 * qq * r) && pp > qq * (a - xf) && pp < qq * (b - xf)) {
          rat = (pp + 0.0) / qq;
          x = xf + rat;
          if ((x - a) < tol2 || (b - x) < tol2) {
            double si = dsign(xm - xf) + ((xm - xf) == 0.0 ? 1.0 : 0.0);
            rat = tol1 * si;
          }
        } else {
          golden = true;
        }
      }
      if (golden) {
        e = (xf >= xm) ? (a - xf) : (b - xf);
        rat = golden_mean * e;
      }
      double si = dsign(rat) + (rat == 0.0 ? 1.0 : 0.0);
      x = xf + si * fmax(fabs(rat), tol1);
      double fu = J(x);
      ++num;
      if (fu <= fx) {
        if (x >= xf)
          a = xf;
        else
          b = xf;
        fulc = nfc;
        ffulc = fnfc;
        nfc = xf;
        fnfc = fx;
        xf = x;
        fx = fu;
      } else {
        if (x < xf)
          a = x;
        else
          b = x;
        if (fu <= fnfc || nfc == xf) {
          fulc = nfc;
          ffulc = fnfc;
          nfc = x;
          fnfc = fu;
        } else if (fu <= ffulc || fulc == xf || fulc == nfc) {
          fulc = x;
          ffulc = fu;
        }
      }
      xm = 0.5 * (a + b);
      tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
      tol2 = 2.0 * tol1;
      if (num >= 500) break;
    }
    result = xf;
  } else {
    // ESS(lambda) = 1 / sum w^2  (mppi.py:526-532); root of ESS - target (mppi.py:351-370)
    auto F = [&](double lam) {
      const SearchStats s = stats_at(lam);
      ++evals;
      return 1.0 / (double)(float)(s.S2 / (s.S * s.S)) - q.essps_target;
    };
    double f_lo = F(q.lambda_min), f_hi = F(q.lambda_max);
    if (f_lo >= 0.0) {  // target <= ess_at_min
      result = q.lambda_min;
    } else if (f_hi <= 0.0) {  // target >= ess_at_max
      result = q.lambda_max;
    } else {
      // scipy.optimize.brentq, xtol = 2e-12, rtol = 4 eps, maxiter = 100
      const double xtol = 2e-12, rtol = 8.881784197001252e-16;
      // (brentq evaluates f(a), f(b) first: the very values the reference's own range check just formed)
      double xpre = q.lambda_min, xcur = q.lambda_max, fpre = f_lo, fcur = f_hi;
      double xblk = 0.0, fblk = 0.0, spre = 0.0, scur = 0.0;
      result = xcur;
      bool done = false;
      if (fpre == 0.0) {
        result = xpre;
        done = true;
      } else if (fcur == 0.0) {
        result = xcur;
        done = true;
      }
      for (int it = 0; it < 100 && !done; ++it) {
        if (fpre != 0.0 && fcur != 0.0 && ((fpre < 0.0) != (fcur < 0.0))) {
          xblk = xpre;
          fblk = fpre;
          spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
          xpre = xcur;
          xcur = xblk;
          xblk = xpre;
          fpre = fcur;
          fcur = fblk;
          fblk = fpre;
        }
        double delta = (xtol + rtol * fabs(xcur)) / 2.0;
        double sbis = (xblk - xcur) / 2.0;
        if (fcur == 0.0 || fabs(sbis) < delta) {
          result = xcur;
          done = true;
          break;
        }
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
          double stry;
          if (xpre == xblk) {
            stry = -fcur * (xcur - xpre) / (fcur - fpre);
          } else {
            double dpre = (fpre - fcur) / (xpre - xcur);
            double dblk = (fblk - fcur) / (xblk - xcur);
            stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
          }
          if (2.0 * fabs(stry) < fmin(fabs(spre), 3.0 * fabs(sbis) - delta)) {
            spre = scur;
            scur = stry;
          } else {
            spre = sbis;
            scur = sbis;
          }
        } else {
          spre = sbis;
          scur = sbis;
        }
        xpre = xcur;
        fpre = fcur;
        if (fabs(scur) > delta)
          xcur += scur;
        else
          xcur += (sbis > 0.0 ? delta : -delta);
        fcur = F(xcur);
        result = xcur;
      }
    }
  }
  if (threadIdx.x == 0) {
    sh.done = 1;
    if (cluster.block_rank() == 0) {
      q.sc->lambda = result;
      q.sc->search_evals = evals;
    }
  }
  named_barrier_sync(1, n_threads);  // releases the helper warps
  cluster.sync();  // no CTA may exit while a peer can still address its shared memory
}

// ---------------------------------------------------------------------------
// small utility kernels
// ---------------------------------------------------------------------------
__global__ void weights_kernel(const float* __restrict__ costs, int K, const DeviceScalars* sc,
                               float* __restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const float lam = (float)sc->lambda_used;
  out[k] = expf((-costs[k]) / lam - sc->xmax) / (float)sc->S;
}

__global__ void iota_kernel(int* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

// Exhaustive proof obligation of ExactDiv: count the x (all 2^32 bit patterns inside the guarded
// magnitude range of div_exact) for which the 3-instruction sequence differs from the IEEE x / c.
__global__ void check_fastdiv_kernel(float c, float r, unsigned long long* mismatches) {
  unsigned bad = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += stride) {
    float x = __uint_as_float((unsigned)i);
    float ax = fabsf(x);
    if (!(ax >= kFastDivMin && ax <= kFastDivMax)) continue;
    float q = x * r;
    float fast = fmaf(fmaf(-q, c, x), r, q);
    float ref = x / c;
    bad += (__float_as_uint(fast) == __float_as_uint(ref)) ? 0u : 1u;
  }
  bad = __reduce_add_sync(kFullMask, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, (unsigned long long)bad);
}

// Proof obligation of the guard-free cell index (map_cell_bounded2): for every x with |x| < kFastDivMin - zero and
// denormals included - the fast division sequence plus the origin rounds to the same cell as the origin alone.
__global__ void check_tiny_quotient_kernel(float c, float r, float ox, float oy, unsigned long long* mismatches) {
  unsigned bad = 0;
  const unsigned top = __float_as_uint(kFastDivMin);  // magnitudes [0, top) are the unproven range of div_exact
  const unsigned long long n = 2ull * top, stride = (unsigned long long)gridDim.x * blockDim.x;
  const int want_x = __float2int_rn(ox), want_y = __float2int_rn(oy);
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float x = __uint_as_float((unsigned)(i >> 1) | ((unsigned)(i & 1ull) << 31));
    float q = x * r;
    q = fmaf(fmaf(-q, c, x), r, q);
    bad += (__float2int_rn(q + ox) != want_x || __float2int_rn(q + oy) != want_y) ? 1u : 0u;
  }
  bad = __reduce_add_sync(kFullMask, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, (unsigned long long)bad);
}

// Exhaustive self-tests of the bounded helpers against the general ones (tests call these through
// mppi_selftest): every float in the claimed range, bit-for-bit.
__global__ void selftest_kernel(unsigned long long* bad /*[4]*/) {
  unsigned b_tan = 0, b_wrap = 0, b_rem = 0, b_sc = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += stride) {
    const float x = __uint_as_float((unsigned)i);
    const float ax = fabsf(x);
    if (ax <= 0.78f) b_tan += (__float_as_uint(tan_quarter(x)) != __float_as_uint(tanf(x))) ? 1u : 0u;
    if (ax <= 4.0f) {
      float s0, c0, s1, c1;
      sincos_bounded(x, &s0, &c0);
      sincosf(x, &s1, &c1);
      b_sc += (__float_as_uint(s0) != __float_as_uint(s1) || __float_as_uint(c0) != __float_as_uint(c1)) ? 1u : 0u;
    }
    if (ax < 9.0f) b_wrap += (__float_as_uint(wrap_angle_bounded(x)) != __float_as_uint(wrap_angle(x))) ? 1u : 0u;
    if (x > -9.4f && x < 9.0f) b_wrap += (__float_as_uint(wrap_angle_above(x)) != __float_as_uint(wrap_angle(x))) ? 1u : 0u;
    if (x >= -3.14159274101257324f && x < 9.0f)
      b_wrap += (__float_as_uint(wrap_angle_nonneg_fast(x)) != __float_as_uint(wrap_angle(x))) ? 1u : 0u;
    if (x >= -3.14159274101257324f && x < 9.0f)
      b_wrap += (__float_as_uint(wrap_angle_nonneg(x)) != __float_as_uint(wrap_angle(x))) ? 1u : 0u;
    {  // paired-sample (P2) forms against the scalar helpers, both lanes: lane 1 carries the neighbouring float
      const float y = __uint_as_float((unsigned)i ^ 1u);
      const float ay = fabsf(y);
      const P2 xy(x, y);
      if (ax <= 0.78f && ay <= 0.78f) {
        const P2 t = tan_quarter2(xy);
        b_tan += (__float_as_uint(t.v.x) != __float_as_uint(tan_quarter(x)) ||
                  __float_as_uint(t.v.y) != __float_as_uint(tan_quarter(y))) ? 1u : 0u;
      }
      if (ax <= 4.0f && ay <= 4.0f) {
        P2 s2, c2;
        float s0, c0, s1, c1;
        sincos_bounded2(xy, &s2, &c2);
        sincos_bounded(x, &s0, &c0);
        sincos_bounded(y, &s1, &c1);
        b_sc += (__float_as_uint(s2.v.x) != __float_as_uint(s0) || __float_as_uint(c2.v.x) != __float_as_uint(c0) ||
                 __float_as_uint(s2.v.y) != __float_as_uint(s1) || __float_as_uint(c2.v.y) != __float_as_uint(c1)) ? 1u : 0u;
      }
      if (ax < 9.0f && ay < 9.0f) {
        const P2 w = wrap_angle_bounded2(xy);
        b_wrap += (__float_as_uint(w.v.x) != __float_as_uint(wrap_angle(x)) ||
                   __float_as_uint(w.v.y) != __float_as_uint(wrap_angle(y))) ? 1u : 0u;
      }
      if (x >= -3.14159274101257324f && x < 9.0f && y >= -3.14159274101257324f && y < 9.0f) {
        const P2 w = wrap_angle_nonneg2(xy);
        b_wrap += (__float_as_uint(w.v.x) != __float_as_uint(wrap_angle(x)) ||
                   __float_as_uint(w.v.y) != __float_as_uint(wrap_angle(y))) ? 1u : 0u;
      }
      if (ax < 1e30f && ay < 1e30f) {  // the uncontractable packed add / multiply-then-add against scalar .rn
        const P2 sum = xy + P2(y, x), mad = xy * 0.1f + P2(y, x);
        b_rem += (__float_as_uint(sum.v.x) != __float_as_uint(__fadd_rn(x, y)) ||
                  __float_as_uint(mad.v.x) != __float_as_uint(__fadd_rn(__fmul_rn(x, 0.1f), y)) ||
                  __float_as_uint(mad.v.y) != __float_as_uint(__fadd_rn(__fmul_rn(y, 0.1f), x))) ? 1u : 0u;
      }
    }
    if (ax < 1e30f) {  // lean floored remainder vs the textbook fmodf form
      const float b = 6.28318548202514648f;
      float m = fmodf(x, b);
      if (m != 0.0f && (m < 0.0f)) m += b;
      b_rem += (__float_as_uint(m) != __float_as_uint(floored_remainder(x, b))) ? 1u : 0u;
    }
  }
  b_tan = __reduce_add_sync(kFullMask, b_tan);
  b_wrap = __reduce_add_sync(kFullMask, b_wrap);
  b_rem = __reduce_add_sync(kFullMask, b_rem);
  b_sc = __reduce_add_sync(kFullMask, b_sc);
  if ((threadIdx.x & 31) == 0) {
    if (b_sc) atomicAdd(bad + 3, (unsigned long long)b_sc);
    if (b_tan) atomicAdd(bad + 0, (unsigned long long)b_tan);
    if (b_wrap) atomicAdd(bad + 1, (unsigned long long)b_wrap);
    if (b_rem) atomicAdd(bad + 2, (unsigned long long)b_rem);
  }
}

__global__ void pack_map_kernel(const float* __restrict__ grid, int W, int H, int words, uint32_t* __restrict__ bits) {
  // (W + 1) rows x `words` words; row W and bit H of every row form the out-of-bounds border of ones
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (W + 1) * words) return;
  int ix = idx / words, wj = idx - ix * words;
  uint32_t v = 0;
  for (int b = 0; b < 32; ++b) {
    int iy = wj * 32 + b;
    bool one = (iy == H) || (iy < H && (ix == W || grid[(size_t)ix * H + iy] != 0.0f));
    if (one) v |= (1u << b);
  }
  bits[idx] = v;
}

// sigma * eps of the in-kernel sampler in the reference's [K,T,du] layout
// (what MultivariateNormal.rsample returns, mppi.py:261-263) - tests only.
template <int DU>
__global__ void sample_noise_kernel(const __grid_constant__ SolveParams p, float* __restrict__ out) {
  long long k_local = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k_local >= p.K) return;
  long long kg = p.k_offset + k_local;
  uint32_t k_lo = (uint32_t)kg, k_hi = (uint32_t)((unsigned long long)kg >> 32);
  for (int c = 0; c * 4 < p.E; ++c) {
    float z[4];
    normal4(p.key, k_lo, k_hi, (uint32_t)c, z);
    for (int j = 0; j < 4; ++j) {
      int e = c * 4 + j;
      if (e < p.E) out[(size_t)k_local * p.E + e] = p.sigma[e % DU] * z[j];
    }
  }
}

// get_top_samples (mppi.py:462-487): re-roll the selected samples, storing states.
template <class M, bool kInject>
__global__ void reroll_kernel(const __grid_constant__ SolveParams p, const int* __restrict__ order, int n, float* __restrict__ traj,
                              float* __restrict__ wout) {
  constexpr int DS = M::DS, DU = M::DU, SPC = Chunking<DU>::kStepsPerChunk;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long k_local = order[i];
  const long long kg = p.k_offset + k_local;
  const uint32_t k_lo = (uint32_t)kg, k_hi = (uint32_t)((unsigned long long)kg >> 32);
  const bool zero_mean = kg >= p.explore_threshold;
  typename M::Ctx ctx{};
  if constexpr (uses_params<M>()) ctx.p = &p.mp;
  float s[DS], seen[DS], u[DU];
  for (int d = 0; d < DS; ++d) s[d] = p.state[d];
  float* out = traj + (size_t)i * (p.T + 1) * DS;
  const float* nz = kInject ? (p.noise + (size_t)k_local * p.T * DU) : nullptr;
  for (int t0 = 0, chunk = 0; t0 < p.T; t0 += SPC, ++chunk) {
    float z[4];
    if (!kInject) normal4(p.key, k_lo, k_hi, (uint32_t)chunk, z);
    for (int j = 0; j < SPC; ++j) {
      int t = t0 + j;
      if (t >= p.T) break;
      for (int d = 0; d < DU; ++d) {
        float eps = kInject ? nz[t * DU + d] : p.sigma[d] * z[j * DU + d];
        // the warm start the solve used is gone (carried state was overwritten);
        // p.prev_action here points at a snapshot taken before the solve
        u[d] = perturbed_entry<DU>(p, p.prev_action, zero_mean, t, d, eps);
      }
      M::step(ctx, s, u, seen);
      for (int d = 0; d < DS; ++d) out[t * DS + d] = seen[d];
    }
  }
  for (int d = 0; d < DS; ++d) out[p.T * DS + d] = s[d];
  const float lam = (float)p.sc->lambda_used;
  wout[i] = expf((-p.costs[k_local]) / lam - p.sc->xmax) / (float)p.sc->S;
}

// _states_prediction (mppi.py:508-524): roll n given action sequences [n,T,du].
template <class M>
__global__ void rollout_actions_kernel(const __grid_constant__ SolveParams p, const float* __restrict__ actions, int n,
                                       float* __restrict__ traj) {
  constexpr int DS = M::DS, DU = M::DU;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename M::Ctx ctx{};
  if constexpr (uses_params<M>()) ctx.p = &p.mp;
  float s[DS], seen[DS], u[DU];
  for (int d = 0; d < DS; ++d) s[d] = p.state[d];
  float* out = traj + (size_t)i * (p.T + 1) * DS;
  const float* a = actions + (size_t)i * p.T * DU;
  for (int t = 0; t < p.T; ++t) {
    for (int d = 0; d < DU; ++d) u[d] = a[t * DU + d];
    M::step(ctx, s, u, seen);
    for (int d = 0; d < DS; ++d) out[t * DS + d] = seen[d];
  }
  for (int d = 0; d < DS; ++d) out[p.T * DS + d] = s[d];
}

// racing_controller.calc_ref_trajectory (example/racing.py:161-218) on the device: nearest centre-line
// point (first minimum of the fp32 distance, never behind the carried index), then one row per stage at the
// host-precomputed index offsets (the reference accumulates them in Python fp64, see
// models.racing_reference_path). One block. *cind is read and updated in place.
__global__ void __launch_bounds__(1024, 1) racing_refpath_kernel(const float* __restrict__ path, int n,
                                                                  const float* __restrict__ state,
                                                                  const int* __restrict__ dind, int rows, float v_max,
                                                                  int* cind, float* __restrict__ out) {
  __shared__ float s_d[32];
  __shared__ int s_i[32];
  __shared__ int s_ind, s_beyond;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float sx = state[0], sy = state[1];
  float best = INFINITY;
  int besti = 0x7fffffff;
  for (int i = tid; i < n; i += blockDim.x) {
    const float dx = path[3 * i] - sx, dy = path[3 * i + 1] - sy;
    float d = dx * dx + dy * dy;  // squared distance: same argmin as the hypot, formed like the host twin
    if (d < best) {  // strided scan visits indices in increasing order per thread
      best = d;
      besti = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float od = __shfl_xor_sync(kFullMask, best, o);
    int oi = __shfl_xor_sync(kFullMask, besti, o);
    if (od < best || (od == best && oi < besti)) {
      best = od;
      besti = oi;
    }
  }
  if (lane == 0) {
    s_d[warp] = best;
    s_i[warp] = besti;
  }
  if (tid == 0) s_beyond = 0;
  __syncthreads();
  if (warp == 0) {
    best = (lane < (int)(blockDim.x >> 5)) ? s_d[lane] : INFINITY;
    besti = (lane < (int)(blockDim.x >> 5)) ? s_i[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float od = __shfl_xor_sync(kFullMask, best, o);
      int oi = __shfl_xor_sync(kFullMask, besti, o);
      if (od < best || (od == best && oi < besti)) {
        best = od;
        besti = oi;
      }
    }
    if (lane == 0) {
      int ind = max(*cind, besti);  // racing.py:201-202
      s_ind = ind;
      *cind = ind;
    }
  }
  __syncthreads();
  const int ind = s_ind;
  for (int i = tid; i < rows; i += blockDim.x)
    if (ind + dind[i] >= n) s_beyond = 1;  // benign race: every writer stores 1
  __syncthreads();
  const float v = s_beyond ? 0.0f : v_max;  // racing.py:213-216: one row past the end zeroes every target speed
  for (int i = tid; i < rows; i += blockDim.x) {
    const int idx = min(ind + dind[i], n - 1);
    out[4 * i + 0] = path[3 * idx + 0];
    out[4 * i + 1] = path[3 * idx + 1];
    out[4 * i + 2] = path[3 * idx + 2];
    out[4 * i + 3] = v;
  }
}

}  // namespace mppi
