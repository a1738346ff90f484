// mppi_models.cuh - the reference's env models as __device__ functions.
//
// Each model restates, operation for operation, the fp32 ATen sequence of the
// reference callable it replaces (file:line cited per model; paths relative to
// the reference root). The translation unit is compiled with -fmad=false: the
// reference never fuses a multiply into an add (every ATen op is its own
// kernel), and the occupancy costs are discontinuous in the rolled-out
// position, so contraction is off here on purpose. Division is IEEE (`/`),
// rounding to cells is half-to-even (torch.round), sin/cos/tan are the
// precise CUDA libm versions.
//
// Interface (all static, state lives in registers of one thread):
//   step(ctx, s, u, seen)  s <- dynamics(s, u); `seen` = what the solver's
//                          stored S[:, t] holds when the cost loop reads it
//                          (== the input state, except MountainCar, whose
//                          dynamics writes through its input views).
//   cost(ctx, s, u, pu, t) stage cost with info = {prev_action: pu, t}.
#pragma once
#include "mppi_device.cuh"

namespace mppi {

// x / c for a fixed divisor, bit-identical to the IEEE quotient: q0 = x * r with r = RN(1/c), one
// exact residual (fma) and one correction (fma) - Markstein's sequence, 3 instructions instead of the
// ~10 + range check of a full division. `fast` is only set after an EXHAUSTIVE device-side comparison
// against `x / c` over every fp32 x with 1e-30 <= |x| <= 1e30 for this very (c, r) pair
// (check_fastdiv_kernel, run once in mppi_set_map); otherwise, and outside that range, the true
// division is used.
struct ExactDiv {
  float c, r;
  int fast;
};
// outside this magnitude range the residual fma can underflow / the product overflow: true division
constexpr float kFastDivMin = 1e-30f, kFastDivMax = 1e30f;
__device__ __forceinline__ float div_exact(float x, const ExactDiv& d) {
  const float ax = fabsf(x);
  if (d.fast && ax >= kFastDivMin && ax <= kFastDivMax) {
    float q = x * d.r;
    float rem = fmaf(-q, d.c, x);
    return fmaf(rem, d.r, q);
  }
  return x / d.c;
}

// Occupancy grid, bit-packed: bit (iy & 31) of word [ix * words + (iy >> 5)]; (W + 1) rows of
// words = ceil((H + 1) / 32) words: row W and bit H of every row are the out-of-bounds border (ones).
struct MapView {
  const uint32_t* bits;
  int W, H, words;
  ExactDiv cell;
  float ox, oy;
  unsigned saddr;  // shared-window address of `bits`, held in a register by the bounded loop (pin_map)
};

// cell index of a world coordinate: round(x / cell + origin) half-to-even, as integer
// (src/envs/obstacle_map_2d.py:179-180)
__device__ __forceinline__ int map_cell(float x, const ExactDiv& cell, float origin) {
  return __float2int_rn(div_exact(x, cell) + origin);
}

// Same index when the divisor passed the exhaustive check and |x| <= 1e30 (callers guarantee the upper
// bound: rolled-out positions are clamped to the map limits and the initial state is range-checked).
// For |x| < 1e-30 the quotient is below half an ulp of any origin >= 1e-22 (or the origin is 0), so the
// index is round(origin) exactly; this keeps the hot loop free of the division's slow-path call.
__device__ __forceinline__ int map_cell_bounded(float x, const ExactDiv& cell, float origin) {
  float q = x * cell.r;
  q = fmaf(fmaf(-q, cell.c, x), cell.r, q);
  return __float2int_rn(q + origin);  // (no |x| < 1e-30 guard: see map_cell_bounded2)
}

__device__ __forceinline__ float map_value(const MapView& m, int ix, int iy) {
  bool oob = (ix < 0) | (ix >= m.W) | (iy < 0) | (iy >= m.H);  // :183-190
  ix = min(max(ix, 0), m.W - 1);                                // :191-192
  iy = min(max(iy, 0), m.H - 1);
  uint32_t w = m.bits[ix * m.words + (iy >> 5)];  // :195
  float occ = (float)((w >> (iy & 31)) & 1u);
  return oob ? 1.0f : occ;  // :198
}

// The packed grid carries one extra row (ix == W) and one extra bit per row (iy == H) of ones, so a
// cell index in [0, W] x [0, H] reads the out-of-bounds value 1.0 without any bounds logic. Only used
// when the host proved that every rolled-out position maps into that range (kFlagBounded).
__device__ __forceinline__ float map_value_bordered(const MapView& m, int ix, int iy) {
  uint32_t w;  // the staged grid is read-only for the whole loop
  asm("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(m.saddr + 4u * (unsigned)(ix * m.words + (iy >> 5))));
  return (float)((w >> (iy & 31)) & 1u);
}

// map_cell_bounded for two samples: the division sequence is packed, the conversion runs per lane. No guard for
// |x| < 1e-30 (where the fast division is not proven): there the sequence yields some |q| <= ~|x| / cell, far
// below half an ulp of the origin, so q + origin == origin like the guarded form - checked for EVERY such x
// against this very (cell, origin) by check_tiny_quotient_kernel in mppi_set_map (kFlagBounded requires it).
__device__ __forceinline__ void map_cell_bounded2(P2 x, const ExactDiv& cell, float origin, int* i0, int* i1) {
  P2 q = x * cell.r;
  q = fma2(fma2(-q, cell.c, x), cell.r, q);
  const P2 qo = q + origin;
  *i0 = __float2int_rn(qo.v.x);
  *i1 = __float2int_rn(qo.v.y);
}

// src/envs/obstacle_map_2d.py:168-200 == src/envs/lane_map_2d.py:90-122.
__device__ __forceinline__ float map_lookup(const MapView& m, float x, float y) {
  return map_value(m, map_cell(x, m.cell, m.ox), map_cell(y, m.cell, m.oy));
}

// ---- serial recurrences of the optimal-trajectory rollout (Navigation2D / Racing rollout_block) --------------
// thw[t] = wrap(ths[t]); ths[t+1] = wrap(thw[t] + cdt[t]). kBounded: every rolled-out heading is itself a
// wrap output (>= -pi) and the host proved |cdt| < 6 < 2 pi, so the inner wrap only needs the upper fold and
// the outer one never sees an argument below -3 pi (wrap_angle_above).
template <bool kBounded>
__device__ __forceinline__ void heading_chain_flagged(float th, const float* cdt, float* thw, float* ths,
                                                             int T, const volatile int* wait_on,
                                                             volatile int* prog) {
  ths[0] = th;
  bool first = true;
  for (int t0 = 0; t0 < T; t0 += 8) {
    if (wait_on) wait_progress(wait_on, t0 + 8);
    float cc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cc[j] = cdt[t0 + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float w;
      if (kBounded) {
        w = wrap_angle_nonneg_fast(th);
        if (j == 0 && first) w = wrap_angle_bounded(th);  // the solve's initial heading may lie below -pi
      } else {
        w = wrap_angle(th);
      }
      thw[t0 + j] = w;
      th = kBounded ? wrap_angle_above(w + cc[j]) : wrap_angle(w + cc[j]);
      ths[t0 + j + 1] = th;
    }
    first = false;
    if (prog) publish_progress(prog, t0 + 8);
  }
}
// `aux(first, stride)` spread over the warps that have no role in a pipelined rollout (5, 7, 8, ...)
template <class Aux>
__device__ __forceinline__ void aux_by_spare_warps(int warp, int lane, int nt, Aux aux) {
  const int n_warps = nt >> 5, spare = 1 + (n_warps - 7);  // warp 5 and warps 7 .. n_warps - 1
  const int slot = warp == 5 ? 0 : warp - 6;
  aux(slot * 32 + lane, spare * 32);
}
// q[t+1] = clamp(q[t] + dq[t]) for the lanes `lane < 2` of one warp (x: lane 0, y: lane 1); wait_on[0] / [1]:
// progress of the even / odd groups of increments
__device__ __forceinline__ void position_chains(int lane, const float* state, const float* dx, const float* dy,
                                                       float* xs, float* ys, const float* lim, int T,
                                                       const volatile int* wait_on) {
  const bool isx = lane == 0;
  float q = isx ? state[0] : state[1];
  float* qs = isx ? xs : ys;
  const float* dq = isx ? dx : dy;
  const float lo = isx ? lim[0] : lim[2], hi = isx ? lim[1] : lim[3];
  qs[0] = q;
  for (int t0 = 0; t0 < T; t0 += 8) {
    if (wait_on) wait_progress(wait_on + ((t0 >> 3) & 1), t0 + 8);  // even / odd groups: two producer warps
    float d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = dq[t0 + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      q = clampf(q + d[j], lo, hi);
      qs[t0 + j + 1] = q;
    }
  }
}

struct ModelParams {
  float v[32];
  int flags;  // model specific, set by the host (see kFlag*)
};
constexpr int kFlagSameMapGeometry = 1;  // Racing: obstacle and lane grids share W, H, cell, origin
constexpr int kFlagUnitWheelbase = 2;    // Racing: L == 1.0f, so x / L == x exactly
// Host-verified bounds (mppi_engine.cu:refresh_model_flags) that make the branch-free helpers exact:
//   steering clamp within +-0.78 rad  -> tan_quarter == tanf
//   |yaw increment per step| < 6 rad  -> wrap_angle_bounded == wrap_angle on every rolled-out heading
//   solver bounds [u_min, u_max] inside the env's own clamp -> that second clamp of a sampled control is
//                                        the identity and is skipped (samples only; the SG-filtered optimal
//                                        sequence can overshoot, so the tail rollout keeps the clamp)
// The kernel additionally requires the solve's initial heading / speed to be in range (uniform check).
constexpr int kFlagBounded = 4;
constexpr int kFlagUnitL = 8;  // Racing: the wheelbase is exactly 1.0f: r / L is r itself, the division is skipped

// ---------------------------------------------------------------------------
struct Pendulum {  // example/pendulum.py:17-47
  static constexpr int DS = 2, DU = 1, kMaps = 0;
  static constexpr bool kRefPath = false, kParallelTail = false, kHasBounded = false, kUsesParams = false;
  struct Ctx {};
  __device__ static __forceinline__ void step(const Ctx&, float (&s)[DS], const float (&u)[DU], float (&seen)[DS]) {
    seen[0] = s[0];
    seen[1] = s[1];
    const float pi = 3.14159274101257324f;
    float tq = clampf(u[0], -2.0f, 2.0f);                       // :26-27
    float acc = (-15.0f * sinf(s[0] + pi) + 3.0f * tq) * 0.05f;  // :28-35
    float nthd = s[1] + acc;
    s[0] = s[0] + nthd * 0.05f;       // :36 (unclamped rate)
    s[1] = clampf(nthd, -8.0f, 8.0f);  // :37
  }
  __device__ static __forceinline__ float cost(const Ctx&, const float (&s)[DS], const float (&)[DU],
                                               const float (&)[DU], int) {
    float a = wrap_angle(s[0]);
    return a * a + 0.1f * (s[1] * s[1]);  // :42-47
  }
};

// ---------------------------------------------------------------------------
struct Cartpole {  // example/cartpole.py:17-81
  static constexpr int DS = 4, DU = 1, kMaps = 0;
  static constexpr bool kRefPath = false, kParallelTail = false, kHasBounded = false, kUsesParams = false;
  struct Ctx {};
  __device__ static __forceinline__ void step(const Ctx&, float (&s)[DS], const float (&u)[DU], float (&seen)[DS]) {
#pragma unroll
    for (int i = 0; i < DS; ++i) seen[i] = s[i];
    const float total_mass = 1.1f, pml = 0.05f;                              // :30-34
    float force = (u[0] >= 0.0f) ? 10.0f : ((u[0] < 0.0f) ? -10.0f : 0.0f);  // :41-44 bang-bang
    float st, ct;
    sincosf(s[2], &st, &ct);
    float temp = (force + pml * (s[3] * s[3]) * st) / total_mass;                                    // :49
    float thacc = (9.8f * st - ct * temp) / (0.5f * (1.33333337306976318f - 0.1f * (ct * ct) / total_mass));  // :50-52
    float xacc = temp - pml * thacc * ct / total_mass;                                               // :53
    float nx = s[0] + 0.02f * s[1];  // :55-58
    float nxd = s[1] + 0.02f * xacc;
    float nth = s[2] + 0.02f * s[3];
    float nthd = s[3] + 0.02f * thacc;
    s[0] = clampf(nx, -2.4f, 2.4f);                                  // :60-65
    s[2] = clampf(nth, -0.20943951606750488f, 0.20943951606750488f);  // float(12*2*pi/360)
    s[1] = nxd;
    s[3] = nthd;
  }
  __device__ static __forceinline__ float cost(const Ctx&, const float (&s)[DS], const float (&)[DU],
                                               const float (&)[DU], int) {
    float a = wrap_angle(s[2]);
    return a * a + 0.1f * (s[3] * s[3]) + 0.1f * (s[0] * s[0]);  // :71-81
  }
};

// ---------------------------------------------------------------------------
struct MountainCar {  // example/mountaincar.py:17-55
  static constexpr int DS = 2, DU = 1, kMaps = 0;
  static constexpr bool kRefPath = false, kParallelTail = false, kHasBounded = false, kUsesParams = false;
  struct Ctx {};
  __device__ static __forceinline__ void step(const Ctx&, float (&s)[DS], const float (&u)[DU], float (&seen)[DS]) {
    float force = clampf(u[0], -1.0f, 1.0f);                          // :32
    float v_raw = s[1] + (force * 0.0015f - 0.0025f * cosf(3.0f * s[0]));  // :34  `velocity +=` (in place)
    float v = clampf(v_raw, -0.07f, 0.07f);                            // :35
    float p_raw = s[0] + v;                                            // :36  `position +=` (in place)
    seen[0] = p_raw;  // the in-place writes land in the solver's S[:, t]
    seen[1] = v_raw;
    s[0] = clampf(p_raw, -1.2f, 0.6f);  // :37
    s[1] = v;
  }
  __device__ static __forceinline__ float cost(const Ctx&, const float (&s)[DS], const float (&)[DU],
                                               const float (&)[DU], int) {
    float d = 0.45f - s[0];
    return d * d;  // :45-55
  }
};

// ---------------------------------------------------------------------------
struct Navigation2D {  // src/envs/navigation_2d.py:218-279
  static constexpr int DS = 3, DU = 2, kMaps = 1;
  static constexpr bool kRefPath = false, kParallelTail = true, kHasBounded = true, kUsesParams = true;
  static constexpr bool kHasHotFlags = false;
  struct Ctx {
    MapView map;
    const ModelParams* p;  // v_min v_max w_min w_max goal_x goal_y x_lo x_hi y_lo y_hi dt w_obst
    float hv[12];          // register copy of p->v[0..11] for the bounded pass-1 loop (pin_loop_consts)
  };
  template <bool kBounded>
  __device__ static __forceinline__ const float* params(const Ctx& c) {
    if (kBounded) return c.hv;
    return c.p->v;
  }
  // kBounded: the host proved |omega dt| < 6 and the kernel checked the initial heading, so every
  // heading stays within the exact range of wrap_angle_bounded (same values, no slow path).
  template <bool kBounded = false>
  __device__ static __forceinline__ void step(const Ctx& c, float (&s)[DS], const float (&u)[DU], float (&seen)[DS]) {
    const float* p = params<kBounded>(c);
#pragma unroll
    for (int i = 0; i < DS; ++i) seen[i] = s[i];
    float v = kBounded ? u[0] : clampf(u[0], p[0], p[1]);  // :235-236 (identity under kFlagBounded)
    float w = kBounded ? u[1] : clampf(u[1], p[2], p[3]);
    float th = kBounded ? wrap_angle_bounded(s[2]) : wrap_angle(s[2]);  // :237
    float st, ct;
    if (kBounded)
      sincos_bounded(th, &st, &ct);
    else
      sincosf(th, &st, &ct);
    float nx = s[0] + v * ct * p[10];  // :239-241
    float ny = s[1] + v * st * p[10];
    float nth = kBounded ? wrap_angle_bounded(th + w * p[10]) : wrap_angle(th + w * p[10]);
    s[0] = clampf(nx, p[6], p[7]);  // :244-251
    s[1] = clampf(ny, p[8], p[9]);
    s[2] = nth;
  }
  __device__ static __forceinline__ bool state_in_bounds(const Ctx& c, const float* state) {
    const float* p = c.p->v;  // heading range, position inside the dynamics' clamp box
    return fabsf(state[2]) < 9.0f && state[0] >= p[6] && state[0] <= p[7] && state[1] >= p[8] && state[1] <= p[9];
  }
  template <bool kBounded = false>
  __device__ static __forceinline__ float cost(const Ctx& c, const float (&s)[DS], const float (&)[DU],
                                               const float (&)[DU], int) {
    const float* p = params<kBounded>(c);
    float dx = s[0] - p[4], dy = s[1] - p[5];
    float goal = sqrtf(dx * dx + dy * dy);  // :269
    float occ = kBounded ? map_value_bordered(c.map, map_cell_bounded(s[0], c.map.cell, c.map.ox),
                                              map_cell_bounded(s[1], c.map.cell, c.map.oy))
                         : map_lookup(c.map, s[0], s[1]);
    return goal + p[11] * occ;  // :271-277
  }
  // ---- two samples per thread (bounded loop only): step<true> / cost<true> with every fp32 add / mul / fma
  // issued once for both samples (P2). The initial heading is >= -pi (kernel-checked, state_in_pair_bounds),
  // so the first wrap of a step only needs the upper fold, like every later one.
  __device__ static __forceinline__ void step_pair(const Ctx& c, P2 (&s)[DS], const P2 (&u)[DU], P2 (&seen)[DS]) {
    const float* p = c.hv;
#pragma unroll
    for (int i = 0; i < DS; ++i) seen[i] = s[i];
    const P2 th = wrap_angle_nonneg2(s[2]);  // :237
    P2 st, ct;
    sincos_bounded2(th, &st, &ct);
    const P2 nx = s[0] + u[0] * ct * p[10];  // :239-241
    const P2 ny = s[1] + u[0] * st * p[10];
    s[2] = wrap_angle_bounded2(th + u[1] * p[10]);
    s[0] = clamp2(nx, p[6], p[7]);  // :244-251
    s[1] = clamp2(ny, p[8], p[9]);
  }
  __device__ static __forceinline__ bool state_in_pair_bounds(const Ctx& c, const float* state) {
    return state_in_bounds(c, state) && state[2] >= -3.14159274101257324f;
  }
  __device__ static __forceinline__ P2 cost_pair(const Ctx& c, const P2 (&s)[DS], const P2 (&)[DU], const P2 (&)[DU],
                                                 int) {
    const float* p = c.hv;
    const P2 dx = s[0] - p[4], dy = s[1] - p[5];
    const P2 d2 = dx * dx + dy * dy;
    const P2 goal(sqrtf(d2.v.x), sqrtf(d2.v.y));  // :269
    int ix0, ix1, iy0, iy1;
    map_cell_bounded2(s[0], c.map.cell, c.map.ox, &ix0, &ix1);
    map_cell_bounded2(s[1], c.map.cell, c.map.oy, &iy0, &iy1);
    const P2 occ(map_value_bordered(c.map, ix0, iy0), map_value_bordered(c.map, ix1, iy1));
    return goal + p[11] * occ;  // :271-277
  }
  // Optimal-trajectory rollout by one block (mppi.py:508-524). Same operations on the same values as
  // T calls of step(); only the schedule differs: the roles run concurrently in separate warps, each trailing
  // the one before through shared-memory progress flags (warp 1 heading recurrence | warps 2, 6 sin / cos +
  // position increments, alternating groups | warp 3 x and y recurrences | warps 5, 7.. `aux`), see
  // Racing::rollout_block. scratch: 9 * (T + 9) floats.
  template <class Aux>
  __device__ static __noinline__ void rollout_block(const Ctx& c, const float* state, const float* opt, int T, float* out,
                                       float* scratch, unsigned long long* trace_row, Aux aux) {
    const float* p = c.p->v;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int S = T + 9;  // the recurrences run in whole groups of 8 stages
    float *wdt = scratch, *ths = wdt + S, *thw = ths + S, *dx = thw + S, *dy = dx + S, *xs = dy + S, *ys = xs + S,
          *vc = ys + S;
    volatile int* prog = reinterpret_cast<volatile int*>(vc + S);  // [2] heading [3] increments even [4] odd groups
    const bool bounded = (c.p->flags & kFlagBounded) && state_in_bounds(c, state);
    const bool pipelined = nt >= 256;  // eight warps: the roles below plus two for `aux`
    for (int t = T + tid; t < S; t += nt) wdt[t] = dx[t] = dy[t] = 0.0f;
    if (tid < 5) prog[tid] = 0;
    for (int t = tid; t < T; t += nt) {
      vc[t] = clampf(opt[2 * t], p[0], p[1]);
      wdt[t] = clampf(opt[2 * t + 1], p[2], p[3]) * p[10];
    }
    __syncthreads();
    stamp_row(trace_row, 10);
    auto increments = [&](int t) {  // thw[t] is a wrap output (|thw| <= pi): sincos_bounded == sincosf there
      float st, ct;
      sincos_bounded(thw[t], &st, &ct);
      dx[t] = vc[t] * ct * p[10];
      dy[t] = vc[t] * st * p[10];
    };
    if (pipelined) {
      const int warp = tid >> 5, lane = tid & 31;
      if (warp == 1) {
        if (lane == 0) {  // heading: S[t+1].theta = wrap(wrap(S[t].theta) + c_t); thw keeps the inner wrap
          if (bounded)
            heading_chain_flagged<true>(state[2], wdt, thw, ths, T, nullptr, prog + 2);
          else
            heading_chain_flagged<false>(state[2], wdt, thw, ths, T, nullptr, prog + 2);
          stamp_row(trace_row, 12, 32);
        }
      } else if (warp == 2 || warp == 6) {
        const int parity = warp == 2 ? 0 : 1;
        for (int t0 = 8 * parity; t0 < T; t0 += 16) {
          if (lane == 0) wait_progress(prog + 2, t0 + 8);
          __syncwarp();
          if (lane < 8 && t0 + lane < T) increments(t0 + lane);
          __syncwarp();
          if (lane == 0) publish_progress(prog + 3 + parity, t0 + 8);
        }
      } else if (warp == 3) {
        if (lane < 2) position_chains(lane, state, dx, dy, xs, ys, p + 6, T, prog + 3);
      } else if (warp == 5 || warp >= 7) {
        aux_by_spare_warps(warp, lane, nt, aux);
      }
      __syncthreads();
      stamp_row(trace_row, 13);
    } else {
      if (tid == 0) {
        if (bounded)
          heading_chain_flagged<true>(state[2], wdt, thw, ths, T, nullptr, nullptr);
        else
          heading_chain_flagged<false>(state[2], wdt, thw, ths, T, nullptr, nullptr);
      }
      __syncthreads();
      for (int t = tid; t < T; t += nt) increments(t);
      __syncthreads();
      if (tid < 2) position_chains(tid, state, dx, dy, xs, ys, p + 6, T, nullptr);
      aux(tid, nt);
      __syncthreads();
    }
    for (int t = tid; t <= T; t += nt) {
      out[t * DS + 0] = xs[t];
      out[t * DS + 1] = ys[t];
      out[t * DS + 2] = ths[t];
    }
  }
  static constexpr int kTailScratchPerStep = 9;
};

// ---------------------------------------------------------------------------
struct Racing {  // src/envs/racing_env.py:327-372 + example/racing.py:110-159
  static constexpr int DS = 4, DU = 2, kMaps = 2;
  static constexpr bool kRefPath = true, kParallelTail = true, kHasBounded = true, kUsesParams = true;
  static constexpr bool kHasHotFlags = true;
  struct Ctx {
    MapView obstacle, lane;
    const ModelParams* p;  // a_min a_max s_min s_max L v_max x_lo x_hi y_lo y_hi dt Qc Ql Qv Qo Qin Qdin
    const float4* ref;     // per stage t: (x, y, sin yaw, cos yaw) of reference_path[t]
    const float* ref_v;    // per stage t: target speed reference_path[t, 3]
    float hv[18];          // register copy of p->v[0..17] for the bounded pass-1 loop (pin_loop_consts)
    int hflags;            // register copy of p->flags
  };
  template <bool kBounded>
  __device__ static __forceinline__ const float* params(const Ctx& c) {
    if (kBounded) return c.hv;
    return c.p->v;
  }
  template <bool kBounded = false>
  __device__ static __forceinline__ float yaw_rate(const ModelParams& mp, float v, float tan_steer) {
    // racing_env.py:352  v * tan(steer) / L ; the division by the wheelbase goes through the proven
    // exact 3-instruction form (trivially exact for L == 1) or the true division
    const float r = v * tan_steer;
    if (kBounded) {  // kFlagBounded implies the proven divisor; |r| <= v_max tan(0.78) << 1e30, and a
                     // quotient of |r| < 1e-30 only feeds `theta + q * dt`, where it vanishes like r itself
      float q = r * mp.v[17];
      q = fmaf(fmaf(-q, mp.v[4], r), mp.v[17], q);
      return (fabsf(r) >= kFastDivMin) ? q : r;
    }
    return div_exact(r, ExactDiv{mp.v[4], mp.v[17], mp.flags & kFlagUnitWheelbase});
  }
  // yaw_rate<true> on the register copy of the parameters
  __device__ static __forceinline__ float yaw_rate_hot(const float* pv, float v, float tan_steer) {
    const float r = v * tan_steer;
    float q = r * pv[17];
    q = fmaf(fmaf(-q, pv[4], r), pv[17], q);
    return (fabsf(r) >= kFastDivMin) ? q : r;
  }
  // kBounded: host-proved steering / yaw bounds + kernel-checked initial state, see kFlagBounded.
  template <bool kBounded = false>
  __device__ static __forceinline__ void step(const Ctx& c, float (&s)[DS], const float (&u)[DU], float (&seen)[DS]) {
    const float* p = params<kBounded>(c);
#pragma unroll
    for (int i = 0; i < DS; ++i) seen[i] = s[i];
    float accel = kBounded ? u[0] : clampf(u[0], p[0], p[1]);  // :345-346 (identity under kFlagBounded)
    float steer = kBounded ? u[1] : clampf(u[1], p[2], p[3]);
    float th = kBounded ? wrap_angle_bounded(s[2]) : wrap_angle(s[2]);  // :347
    float st, ct;
    if (kBounded)
      sincos_bounded(th, &st, &ct);
    else
      sincosf(th, &st, &ct);
    float dx = s[3] * ct;  // :349-352
    float dy = s[3] * st;
    float dth;
    if (kBounded)  // (L == 1.0f exactly: v tan(steer) / L is the product itself)
      dth = (c.hflags & kFlagUnitL) ? s[3] * tan_quarter(steer) : yaw_rate_hot(p, s[3], tan_quarter(steer));
    else
      dth = yaw_rate<false>(*c.p, s[3], tanf(steer));
    float nx = s[0] + dx * p[10];  // :354-357
    float ny = s[1] + dy * p[10];
    float nth = kBounded ? wrap_angle_bounded(th + dth * p[10]) : wrap_angle(th + dth * p[10]);
    float nv = s[3] + accel * p[10];
    s[0] = clampf(nx, p[6], p[7]);  // :360-368
    s[1] = clampf(ny, p[8], p[9]);
    s[2] = nth;
    s[3] = clampf(nv, -p[5], p[5]);
  }
  __device__ static __forceinline__ bool state_in_bounds(const Ctx& c, const float* state) {
    const float* p = c.p->v;  // heading range, |v| <= v_max, position inside the dynamics' clamp box
    return fabsf(state[2]) < 9.0f && fabsf(state[3]) <= p[5] && state[0] >= p[6] && state[0] <= p[7] &&
           state[1] >= p[8] && state[1] <= p[9];
  }
  template <bool kBounded = false>
  __device__ static __forceinline__ float cost(const Ctx& c, const float (&s)[DS], const float (&u)[DU],
                                               const float (&pu)[DU], int t) {
    const float* p = params<kBounded>(c);
    const float4 r = c.ref[t];
    float sx = s[0] - r.x, sy = s[1] - r.y;
    float ec = r.z * sx - r.w * sy;                         // racing.py:127-131
    float el = (-r.w) * sx - r.z * sy;                      // :132-136
    float path = p[11] * (ec * ec) + p[12] * (el * el);     // :138
    float dv = s[3] - c.ref_v[t];
    float vel = p[13] * (dv * dv);                           // :141-143
    float occ;                                               // :146-150
    if (kBounded) {  // kFlagBounded implies: one shared geometry, proven exact division
      int ix = map_cell_bounded(s[0], c.obstacle.cell, c.obstacle.ox);
      int iy = map_cell_bounded(s[1], c.obstacle.cell, c.obstacle.oy);
      occ = map_value_bordered(c.obstacle, ix, iy) + map_value_bordered(c.lane, ix, iy);
    } else if (c.p->flags & kFlagSameMapGeometry) {  // one cell index serves both grids
      int ix = map_cell(s[0], c.obstacle.cell, c.obstacle.ox), iy = map_cell(s[1], c.obstacle.cell, c.obstacle.oy);
      occ = map_value(c.obstacle, ix, iy) + map_value(c.lane, ix, iy);
    } else {
      occ = map_lookup(c.obstacle, s[0], s[1]);
      occ = occ + map_lookup(c.lane, s[0], s[1]);
    }
    occ = p[14] * occ;                                       // :151
    float in = p[15] * (u[0] * u[0] + u[1] * u[1]);          // :154
    float d0 = u[0] - pu[0], d1 = u[1] - pu[1];
    in = in + p[16] * (d0 * d0 + d1 * d1);                   // :155
    return path + vel + occ + in;                            // :157
  }
  // ---- two samples per thread (bounded loop only): step<true> / cost<true> with every fp32 add / mul / fma
  // issued once for both samples (P2); clamps, selects, conversions and the grid lookups stay per lane. The
  // initial heading is >= -pi (kernel-checked, state_in_pair_bounds), so the first wrap of a step only needs
  // the upper fold (wrap_angle_nonneg == wrap_angle there), like every later one.
  __device__ static __forceinline__ void step_pair(const Ctx& c, P2 (&s)[DS], const P2 (&u)[DU], P2 (&seen)[DS]) {
    const float* p = c.hv;
#pragma unroll
    for (int i = 0; i < DS; ++i) seen[i] = s[i];
    const P2 th = wrap_angle_nonneg2(s[2]);  // :347
    P2 st, ct;
    sincos_bounded2(th, &st, &ct);
    const P2 dx = s[3] * ct;  // :349-352
    const P2 dy = s[3] * st;
    const P2 r = s[3] * tan_quarter2(u[1]);
    P2 dth = r;  // v tan(steer) / L: for L == 1.0f exactly the quotient is r itself
    if (!(c.hflags & kFlagUnitL)) {  // (uniform over the launch) yaw_rate_hot per lane: exact r / L
      P2 q = r * p[17];
      q = fma2(fma2(-q, p[4], r), p[17], q);
      dth = P2((fabsf(r.v.x) >= kFastDivMin) ? q.v.x : r.v.x, (fabsf(r.v.y) >= kFastDivMin) ? q.v.y : r.v.y);
    }
    const P2 nx = s[0] + dx * p[10];  // :354-357
    const P2 ny = s[1] + dy * p[10];
    const P2 nv = s[3] + u[0] * p[10];
    s[2] = wrap_angle_bounded2(th + dth * p[10]);
    s[0] = clamp2(nx, p[6], p[7]);  // :360-368
    s[1] = clamp2(ny, p[8], p[9]);
    s[3] = clamp2(nv, -p[5], p[5]);
  }
  __device__ static __forceinline__ bool state_in_pair_bounds(const Ctx& c, const float* state) {
    return state_in_bounds(c, state) && state[2] >= -3.14159274101257324f;
  }
  __device__ static __forceinline__ P2 cost_pair(const Ctx& c, const P2 (&s)[DS], const P2 (&u)[DU],
                                                 const P2 (&pu)[DU], int t) {
    const float* p = c.hv;
    const float4 r = c.ref[t];
    const P2 sx = s[0] - r.x, sy = s[1] - r.y;
    const P2 ec = r.z * sx - r.w * sy;                         // racing.py:127-131
    const P2 el = (-r.w) * sx - r.z * sy;                      // :132-136
    const P2 path = p[11] * (ec * ec) + p[12] * (el * el);     // :138
    const P2 dv = s[3] - c.ref_v[t];
    const P2 vel = p[13] * (dv * dv);                          // :141-143
    int ix0, ix1, iy0, iy1;
    map_cell_bounded2(s[0], c.obstacle.cell, c.obstacle.ox, &ix0, &ix1);
    map_cell_bounded2(s[1], c.obstacle.cell, c.obstacle.oy, &iy0, &iy1);
    P2 occ = P2(map_value_bordered(c.obstacle, ix0, iy0), map_value_bordered(c.obstacle, ix1, iy1)) +
             P2(map_value_bordered(c.lane, ix0, iy0), map_value_bordered(c.lane, ix1, iy1));  // :146-150
    occ = p[14] * occ;                                         // :151
    P2 in = p[15] * (u[0] * u[0] + u[1] * u[1]);               // :154
    const P2 d0 = u[0] - pu[0], d1 = u[1] - pu[1];
    in = in + p[16] * (d0 * d0 + d1 * d1);                     // :155
    return path + vel + occ + in;                              // :157
  }
  // ---- optimal-trajectory rollout by one block (mppi.py:508-524): the same operations on the same values as
  // T calls of step(), rescheduled. tan of every stage first (parallel); then the roles below run CONCURRENTLY
  // in separate warps, each trailing the one before through shared-memory progress flags (groups of 8 stages):
  //   warp 0  speed recurrence v' = clamp(v + a dt) (3 dependent operations per stage)
  //   warp 4  yaw increments v tan(steer) / L * dt of a group (8 lanes)
  //   warp 1  heading recurrence (two angle wraps per stage) - the long pole, 9 dependent fp32 operations per
  //           stage in the bounded form (wrap_angle_nonneg_fast / wrap_angle_above)
  //   warps 2, 6  sin / cos of the wrapped headings and the position increments, alternating groups (8 lanes)
  //   warp 3  the x and y recurrences (add + clamp), two lanes
  //   warps 5, 7.. `aux` (the caller's stores of the carried state), off everybody's critical path
  // (warp w issues on scheduler w % 4: the heading warp shares its scheduler only with an `aux` warp, never
  // with a role that spins on a flag)
  // so the wall time is the heading chain plus a short fill / drain instead of the sum of the phases. Blocks
  // with fewer than eight warps run the same operations phase by phase. scratch: 11 * (T + 9) floats.
  __device__ static __forceinline__ void speed_chain(const ModelParams& mp, const float* state, const float* adt,
                                                     float* vs, int T, volatile int* prog) {
    const float vm = mp.v[5];
    float v = state[3];
    for (int t0 = 0; t0 < T; t0 += 8) {
      float a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = adt[t0 + j];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        vs[t0 + j] = v;
        v = clampf(v + a[j], -vm, vm);
      }
      if (prog) publish_progress(prog, t0 + 8);
    }
    if ((T & 7) == 0) vs[T] = v;  // otherwise stage T lies inside the last group and was stored there
  }
  template <class Aux>
  __device__ static __noinline__ void rollout_block(const Ctx& c, const float* state, const float* opt, int T, float* out,
                                       float* scratch, unsigned long long* trace_row, Aux aux) {
    const float* p = c.p->v;
    const int tid = threadIdx.x, nt = blockDim.x, S = T + 9;  // the recurrences run in whole groups of 8 stages
    float *adt = scratch, *tn = adt + S, *vs = tn + S, *cdt = vs + S, *ths = cdt + S, *thw = ths + S, *dx = thw + S,
          *dy = dx + S, *xs = dy + S, *ys = xs + S;
    // progress: [0] speed [1] yaw increments [2] heading [3] position increments, even groups [4] odd groups
    volatile int* prog = reinterpret_cast<volatile int*>(ys + S);
    const bool bounded = (c.p->flags & kFlagBounded) && state_in_bounds(c, state);
    const bool pipelined = nt >= 256;  // eight warps: the roles below plus two for `aux`
    for (int t = T + tid; t < S; t += nt) adt[t] = tn[t] = cdt[t] = dx[t] = dy[t] = 0.0f;
    if (tid < 5) prog[tid] = 0;
    for (int t = tid; t < T; t += nt) {
      adt[t] = clampf(opt[2 * t], p[0], p[1]) * p[10];
      const float st = clampf(opt[2 * t + 1], p[2], p[3]);
      tn[t] = bounded ? tan_quarter(st) : tanf(st);
    }
    __syncthreads();
    stamp_row(trace_row, 10);
    auto yaw_increment = [&](int t) {
      cdt[t] = (bounded ? yaw_rate<true>(*c.p, vs[t], tn[t]) : yaw_rate<false>(*c.p, vs[t], tn[t])) * p[10];
    };
    auto increments = [&](int t) {  // thw[t] is a wrap output (|thw| <= pi): sincos_bounded == sincosf there
      float st, ct;
      sincos_bounded(thw[t], &st, &ct);
      dx[t] = vs[t] * ct * p[10];
      dy[t] = vs[t] * st * p[10];
    };
    if (pipelined) {
      const int warp = tid >> 5, lane = tid & 31;
      if (warp == 0) {
        if (lane == 0) {
          speed_chain(*c.p, state, adt, vs, T, prog + 0);
          stamp_row(trace_row, 11, 0);
        }
      } else if (warp == 4) {
        for (int t0 = 0; t0 < T; t0 += 8) {
          if (lane == 0) wait_progress(prog + 0, t0 + 8);
          __syncwarp();
          if (lane < 8) yaw_increment(t0 + lane);  // (stages T.. of the last group: zero inputs, padding)
          __syncwarp();
          if (lane == 0) publish_progress(prog + 1, t0 + 8);
        }
      } else if (warp == 1) {
        if (lane == 0) {
          const long long c0 = clock64();
          if (bounded)
            heading_chain_flagged<true>(state[2], cdt, thw, ths, T, prog + 1, prog + 2);
          else
            heading_chain_flagged<false>(state[2], cdt, thw, ths, T, prog + 1, prog + 2);
          stamp_row(trace_row, 12, 32);
          if (trace_row) trace_row[15] = (unsigned long long)(clock64() - c0);  // SM cycles of the chain
        }
      } else if (warp == 2 || warp == 6) {
        const int parity = warp == 2 ? 0 : 1;
        for (int t0 = 8 * parity; t0 < T; t0 += 16) {
          if (lane == 0) wait_progress(prog + 2, t0 + 8);
          __syncwarp();
          if (lane < 8 && t0 + lane < T) increments(t0 + lane);
          __syncwarp();
          if (lane == 0) publish_progress(prog + 3 + parity, t0 + 8);
        }
        if (parity == 0) stamp_row(trace_row, 14, 64);
      } else if (warp == 3) {
        if (lane < 2) position_chains(lane, state, dx, dy, xs, ys, p + 6, T, prog + 3);
      } else if (warp == 5 || warp >= 7) {
        aux_by_spare_warps(warp, lane, nt, aux);
      }
      __syncthreads();
      stamp_row(trace_row, 13);
    } else {
      if (tid == 0) speed_chain(*c.p, state, adt, vs, T, nullptr);
      __syncthreads();
      for (int t = tid; t < T; t += nt) yaw_increment(t);
      __syncthreads();
      if (tid == 0) {
        if (bounded)
          heading_chain_flagged<true>(state[2], cdt, thw, ths, T, nullptr, nullptr);
        else
          heading_chain_flagged<false>(state[2], cdt, thw, ths, T, nullptr, nullptr);
      }
      __syncthreads();
      for (int t = tid; t < T; t += nt) increments(t);
      __syncthreads();
      if (tid < 2) position_chains(tid, state, dx, dy, xs, ys, p + 6, T, nullptr);
      aux(tid, nt);
      __syncthreads();
    }
    for (int t = tid; t <= T; t += nt) {
      out[t * DS + 0] = xs[t];
      out[t * DS + 1] = ys[t];
      out[t * DS + 2] = ths[t];
      out[t * DS + 3] = vs[t];
    }
  }
  static constexpr int kTailScratchPerStep = 11;
};

// ---------------------------------------------------------------------------
struct CartpoleContinuous {  // example/mujoco_cartpole.py:20-80 (pole mass 1.0, continuous force, |x| <= 1)
  static constexpr int DS = 4, DU = 1, kMaps = 0;
  static constexpr bool kRefPath = false, kParallelTail = false, kHasBounded = false, kUsesParams = false;
  struct Ctx {};
  __device__ static __forceinline__ void step(const Ctx&, float (&s)[DS], const float (&u)[DU], float (&seen)[DS]) {
#pragma unroll
    for (int i = 0; i < DS; ++i) seen[i] = s[i];
    const float total_mass = 2.0f, pml = 0.5f, masspole = 1.0f;  // :35-40
    const float force = u[0];                                    // :33 (no bang-bang here)
    float st, ct;
    sincosf(s[2], &st, &ct);
    float temp = (force + pml * (s[3] * s[3]) * st) / total_mass;  // :46
    float thacc = (9.8f * st - ct * temp) / (0.5f * (1.33333337306976318f - masspole * (ct * ct) / total_mass));  // :47-49
    float xacc = temp - pml * thacc * ct / total_mass;  // :50
    float nx = s[0] + 0.02f * s[1];                     // :52-55
    float nxd = s[1] + 0.02f * xacc;
    float nth = s[2] + 0.02f * s[3];
    float nthd = s[3] + 0.02f * thacc;
    s[0] = clampf(nx, -1.0f, 1.0f);                                  // :57-62
    s[2] = clampf(nth, -0.20943951606750488f, 0.20943951606750488f);
    s[1] = nxd;
    s[3] = nthd;
  }
  __device__ static __forceinline__ float cost(const Ctx&, const float (&s)[DS], const float (&)[DU],
                                               const float (&)[DU], int) {
    float a = wrap_angle(s[2]);
    return a * a + 0.1f * (s[3] * s[3]) + 0.1f * (s[0] * s[0]);  // :68-80
  }
};

// ---------------------------------------------------------------------------
struct GoalInDangerZone {  // src/envs/goal_in_danger_zone.py:113-156 (obs = x y theta, vec to goal, vec to centre)
  static constexpr int DS = 7, DU = 2, kMaps = 0;
  static constexpr bool kRefPath = false, kParallelTail = false, kHasBounded = false, kUsesParams = true;
  struct Ctx {
    const ModelParams* p;  // v_min v_max w_min w_max dt goal_x goal_y centre_x centre_y radius collision_cost
  };
  __device__ static __forceinline__ void step(const Ctx& c, float (&s)[DS], const float (&u)[DU], float (&seen)[DS]) {
    const float* p = c.p->v;
#pragma unroll
    for (int i = 0; i < DS; ++i) seen[i] = s[i];
    float v = clampf(u[0], p[0], p[1]);  // :118-119
    float w = clampf(u[1], p[2], p[3]);
    float th = wrap_angle(s[2] + w * p[4]);  // :124
    float st, ct;
    sincosf(th, &st, &ct);
    float nx = s[0] + v * ct * p[4];  // :126-127
    float ny = s[1] + v * st * p[4];
    s[0] = nx;
    s[1] = ny;
    s[2] = th;
    s[3] = p[5] - nx;  // :129-134
    s[4] = p[6] - ny;
    s[5] = p[7] - nx;
    s[6] = p[8] - ny;
  }
  __device__ static __forceinline__ float cost(const Ctx& c, const float (&s)[DS], const float (&)[DU],
                                               const float (&)[DU], int) {
    const float* p = c.p->v;
    float dist = sqrtf(s[3] * s[3] + s[4] * s[4]);                   // :145
    float collided = (sqrtf(s[5] * s[5] + s[6] * s[6]) < p[9]) ? 1.0f : 0.0f;  // :153
    return dist + collided * p[10];                                  // :154
  }
};

}  // namespace mppi
