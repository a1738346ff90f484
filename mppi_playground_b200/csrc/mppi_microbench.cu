// mppi_microbench.cu - measured fp32 denominators for the roofline of the solve kernel.
//
// MEASURED_PEAKS.json (driver-written) only holds HBM GB/s and dense bf16 TFLOP/s; the MPPI rollout is
// bound by the fp32 CUDA-core pipes, so the engine measures that peak itself:
//   * FFMA, 3-register form, 8 independent chains per thread, every SM full  -> the FMA peak
//   * FFMA2 (packed f32x2, sm_100)                                         -> the same peak in half the issue slots?
//   * FMUL + FADD alternating, never contracted (what -fmad=false leaves of the reference's a*b+c)
//   * FMUL2 + (FFMA2 by an opaque 1.0 = the uncontractable packed add) alternating: the paired-sample loop's form
//   * dependent-issue latency of FFMA / FFMA2 / FADD2 / FMNMX (one warp, one chain)
// bench.py calls mppi_fp32_microbench() once per run and reports achieved / measured next to the
// derived 148 x 128 x 2 x clock figure.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mppi_b200.h"

namespace {

constexpr int kChains = 8;

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// mode 0: FFMA, 1: FFMA2, 2: FMUL+FADD, 3: FMUL2+FADD2
template <int kMode>
__global__ void __launch_bounds__(1024, 2) throughput_kernel(const float* __restrict__ seed, float* __restrict__ sink,
                                                            int iters, unsigned long long* clk /*[4]*/) {
  const float b = seed[0], c = seed[1];
  unsigned long long t0 = 0, c0 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    t0 = globaltimer_ns();
    c0 = clock64();
  }
  if (kMode == 0 || kMode == 2) {
    float a[kChains];
#pragma unroll
    for (int j = 0; j < kChains; ++j) a[j] = seed[2 + j] + (float)threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < kChains; ++j) {
        if (kMode == 0)
          a[j] = fmaf(a[j], b, c);
        else
          a[j] = __fadd_rn(__fmul_rn(a[j], b), c);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kChains; ++j) s += a[j];
    if (s == 12345.678f) sink[threadIdx.x] = s;
  } else {
    float2 a[kChains];
    const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
    const float2 one2 = make_float2(seed[15], seed[15]);  // 1.0 the compiler cannot see (see P2 in mppi_device.cuh)
#pragma unroll
    for (int j = 0; j < kChains; ++j) a[j] = make_float2(seed[2 + j] + (float)threadIdx.x, seed[3 + j]);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < kChains; ++j) {
        if (kMode == 1)
          a[j] = __ffma2_rn(a[j], b2, c2);
        else
          a[j] = __ffma2_rn(__fmul2_rn(a[j], b2), one2, c2);  // product and sum rounded separately: ptxas fuses
                                                               // mul.rn.f32x2 + add.rn.f32x2 despite -fmad=false
      }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kChains; ++j) s += a[j].x + a[j].y;
    if (s == 12345.678f) sink[threadIdx.x] = s;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    clk[0] = globaltimer_ns() - t0;
    clk[1] = (unsigned long long)(clock64() - c0);
  }
}

// one warp, one dependent chain: cycles per operation. op 0 FFMA, 1 FFMA2, 2 FADD2, 3 FMNMX, 4 FADD, 5 FMUL2
template <int kOp>
__global__ void latency_kernel(const float* __restrict__ seed, float* __restrict__ sink, int iters,
                               unsigned long long* out) {
  const float b = seed[0], c = seed[1];
  float x = seed[2] + (float)threadIdx.x;
  float2 x2 = make_float2(x, seed[3]);
  const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
  const long long c0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i) {
    if (kOp == 0) x = fmaf(x, b, c);
    if (kOp == 1) x2 = __ffma2_rn(x2, b2, c2);
    if (kOp == 2) x2 = __fadd2_rn(x2, c2);
    if (kOp == 3) x = fmaxf(fminf(x, c), b);  // two dependent FMNMX
    if (kOp == 4) x = __fadd_rn(x, c);
    if (kOp == 5) x2 = __fmul2_rn(x2, b2);
  }
  const long long c1 = clock64();
  if (x + x2.x + x2.y == 12345.678f) sink[threadIdx.x] = x;
  if (threadIdx.x == 0) out[0] = (unsigned long long)(c1 - c0);
}

template <int kMode>
cudaError_t run_throughput(int sms, const float* d_seed, float* d_sink, unsigned long long* d_clk, double* tflops,
                           double* mhz) {
  const int iters = 4096, blocks = sms * 2, threads = 1024;
  throughput_kernel<kMode><<<blocks, threads>>>(d_seed, d_sink, 64, d_clk);  // warm-up (module load, clocks)
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  unsigned long long clk[2] = {0, 0};
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    throughput_kernel<kMode><<<blocks, threads>>>(d_seed, d_sink, iters, d_clk);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) return e;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) {
      best = ms;
      cudaMemcpy(clk, d_clk, 16, cudaMemcpyDeviceToHost);
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  // flop per thread-iteration: FFMA = 2, FFMA2 = 4, FMUL+FADD = 2, FMUL2+FADD2 = 4
  const double per = (kMode == 0 || kMode == 2) ? 2.0 : 4.0;
  const double flops = (double)blocks * threads * (double)iters * kChains * per;
  *tflops = flops / ((double)best * 1e-3) / 1e12;
  if (mhz && clk[0]) *mhz = (double)clk[1] / (double)clk[0] * 1e3;
  return cudaGetLastError();
}

template <int kOp>
cudaError_t run_latency(const float* d_seed, float* d_sink, unsigned long long* d_clk, double* cycles_per_op) {
  const int iters = 8192;
  latency_kernel<kOp><<<1, 32>>>(d_seed, d_sink, 256, d_clk);
  latency_kernel<kOp><<<1, 32>>>(d_seed, d_sink, iters, d_clk);
  unsigned long long c = 0;
  cudaError_t e = cudaMemcpy(&c, d_clk, 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return e;
  *cycles_per_op = (double)c / (double)iters / (kOp == 3 ? 2.0 : 1.0);
  return cudaSuccess;
}

}  // namespace

extern "C" int mppi_fp32_microbench(int32_t device, MppiFp32Report* out) {
  if (!out) return MPPI_ERR_INVALID;
  int prev = 0;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess) return MPPI_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MPPI_ERR_CUDA;
  float h_seed[16];
  for (int i = 0; i < 16; ++i) h_seed[i] = 0.5f + 0.03125f * (float)i;
  h_seed[0] = 0.999f;  // multiplier < 1: the chains stay finite
  h_seed[1] = 0.001f;
  h_seed[15] = 1.0f;
  float *d_seed = nullptr, *d_sink = nullptr;
  unsigned long long* d_clk = nullptr;
  cudaError_t e = cudaMalloc((void**)&d_seed, sizeof h_seed);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_sink, 4096);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_clk, 64);
  if (e == cudaSuccess) e = cudaMemcpy(d_seed, h_seed, sizeof h_seed, cudaMemcpyHostToDevice);
  MppiFp32Report r{};
  r.sms = prop.multiProcessorCount;
  if (e == cudaSuccess) e = run_throughput<0>(r.sms, d_seed, d_sink, d_clk, &r.ffma_tflops, &r.sm_clock_mhz);
  if (e == cudaSuccess) e = run_throughput<1>(r.sms, d_seed, d_sink, d_clk, &r.ffma2_tflops, nullptr);
  if (e == cudaSuccess) e = run_throughput<2>(r.sms, d_seed, d_sink, d_clk, &r.fmul_fadd_tflops, nullptr);
  if (e == cudaSuccess) e = run_throughput<3>(r.sms, d_seed, d_sink, d_clk, &r.fmul2_fadd2_tflops, nullptr);
  if (e == cudaSuccess) e = run_latency<0>(d_seed, d_sink, d_clk, &r.lat_ffma);
  if (e == cudaSuccess) e = run_latency<1>(d_seed, d_sink, d_clk, &r.lat_ffma2);
  if (e == cudaSuccess) e = run_latency<2>(d_seed, d_sink, d_clk, &r.lat_fadd2);
  if (e == cudaSuccess) e = run_latency<3>(d_seed, d_sink, d_clk, &r.lat_fmnmx);
  if (e == cudaSuccess) e = run_latency<4>(d_seed, d_sink, d_clk, &r.lat_fadd);
  if (e == cudaSuccess) e = run_latency<5>(d_seed, d_sink, d_clk, &r.lat_fmul2);
  cudaFree(d_seed);
  cudaFree(d_sink);
  cudaFree(d_clk);
  cudaSetDevice(prev);
  if (e != cudaSuccess) return MPPI_ERR_CUDA;
  *out = r;
  return MPPI_OK;
}
