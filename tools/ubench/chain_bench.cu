// Micro-benchmark of the optimal-trajectory heading recurrence (one thread, dependent chain):
//   th' = wrap(wrap(th) + c)     cycles per stage for several instruction-level forms of the same values.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o chain_bench chain_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float wrap_nonneg_sel(float x) {
  const float pi = 3.14159274101257324f, two_pi = 6.28318548202514648f;
  const float a = __fadd_rn(x, pi);
  const float am = __fsub_rn(a, two_pi);
  return __fsub_rn((a >= two_pi) ? am : a, pi);
}
__device__ __forceinline__ float wrap_above_sel(float x) {
  const float pi = 3.14159274101257324f, two_pi = 6.28318548202514648f;
  const float a = __fadd_rn(x, pi);
  const float am = __fsub_rn(a, two_pi), ap = __fadd_rn(a, two_pi);
  return __fsub_rn((a >= two_pi) ? am : ((a < 0.0f) ? ap : a), pi);
}
// arithmetic folds: fold factor as 0/1 float (FSET.BF), one fma each; no predicates on the chain
__device__ __forceinline__ float wrap_nonneg_fma(float x) {
  const float pi = 3.14159274101257324f, two_pi = 6.28318548202514648f;
  const float a = __fadd_rn(x, pi);
  const float f = (a >= two_pi) ? 1.0f : 0.0f;
  return __fsub_rn(fmaf(-two_pi, f, a), pi);
}
__device__ __forceinline__ float wrap_above_fma(float x) {
  const float pi = 3.14159274101257324f, two_pi = 6.28318548202514648f;
  const float a = __fadd_rn(x, pi);
  const float f = ((a >= two_pi) ? 1.0f : 0.0f) - ((a < 0.0f) ? 1.0f : 0.0f);
  return __fsub_rn(fmaf(-two_pi, f, a), pi);
}
template <int kForm, int kLanes>
__global__ void chain(const float* __restrict__ inc, float* out, long long* cyc, int T, int reps) {
  if ((int)threadIdx.x >= kLanes) return;
  float th = inc[0] + 0.001f * threadIdx.x;
  float acc = 0.f;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    for (int t = 0; t < T; t += 8) {
      float c[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) c[j] = inc[t + j];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float w = kForm == 0 ? wrap_nonneg_sel(th) : wrap_nonneg_fma(th);
        acc += w;
        th = kForm == 0 ? wrap_above_sel(w + c[j]) : wrap_above_fma(w + c[j]);
      }
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = th + acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// plain dependent FADD chain, same harness (calibration)
template <int kLanes>
__global__ void fadd_chain(const float* __restrict__ inc, float* out, long long* cyc, int T, int reps) {
  if ((int)threadIdx.x >= kLanes) return;
  float th = inc[0];
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r)
    for (int t = 0; t < T; t += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) th = __fadd_rn(th, inc[t + j]);
    }
  long long t1 = clock64();
  out[threadIdx.x] = th;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// the in-kernel structure: per group of 8 stages a flag wait, 8 LDS of the increments, the chain, 16 STS of the
// results, a release store of the progress flag. kPrefetch: the next group's wait + loads are issued before this
// group's chain. kForm as above.
__device__ __forceinline__ void wait_flag(const int* flag, int need) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(flag);
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  while (v < need) {
    __nanosleep(32);
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  }
}
__device__ __forceinline__ void publish_flag(int* flag, int value) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(flag);
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(a), "r"(value) : "memory");
}
template <int kForm, bool kPrefetch, bool kRelease, int kPark = 0>
__global__ void chain_smem(const float* __restrict__ inc, float* out, long long* cyc, int T, int reps) {
  __shared__ float cdt[96], thw[96], ths[97];
  __shared__ int flags[4];
  for (int i = threadIdx.x; i < 96; i += blockDim.x) cdt[i] = inc[i];
  if (threadIdx.x == 0) { flags[0] = 1 << 30; flags[1] = 0; flags[2] = 0; }
  __syncthreads();
  if (kPark == 0 && threadIdx.x != 0) return;
  if (kPark == 2 && threadIdx.x >= 32) {  // other warps: spin on a flag that is only set at the very end
    if ((threadIdx.x & 31) == 0) wait_flag(&flags[2], 1);
    __syncthreads();
    return;
  }
  if (threadIdx.x != 0) {  // parked lanes / warps
    __syncthreads();
    return;
  }
  float th = inc[0];
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    float c[8];
    if (kPrefetch) {
      wait_flag(&flags[0], 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) c[j] = cdt[j];
    }
    for (int t = 0; t < T; t += 8) {
      float cn[8];
      if (kPrefetch) {
        if (t + 8 < T) {
          wait_flag(&flags[0], t + 16);
#pragma unroll
          for (int j = 0; j < 8; ++j) cn[j] = cdt[t + 8 + j];
        }
      } else {
        wait_flag(&flags[0], t + 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = cdt[t + j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float w = kForm == 0 ? wrap_nonneg_sel(th) : wrap_nonneg_fma(th);
        thw[t + j] = w;
        th = kForm == 0 ? wrap_above_sel(w + c[j]) : wrap_above_fma(w + c[j]);
        ths[t + j + 1] = th;
      }
      if (kRelease) publish_flag(&flags[1], t + 8); else *(volatile int*)&flags[1] = t + 8;
      if (kPrefetch) {
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = cn[j];
      }
    }
  }
  long long t1 = clock64();
  out[0] = th + thw[5] + ths[7];
  cyc[0] = t1 - t0;
  if (kPark == 2) publish_flag(&flags[2], 1);
  if (kPark) __syncthreads();
}

int main() {
  const int T = 80, reps = 200;
  float h[96];
  for (int i = 0; i < 96; ++i) h[i] = 0.05f + 0.003f * (i % 7);
  float *d_inc, *d_out;
  long long* d_c;
  cudaMalloc(&d_inc, sizeof h);
  cudaMalloc(&d_out, 4096);
  cudaMalloc(&d_c, 8);
  cudaMemcpy(d_inc, h, sizeof h, cudaMemcpyHostToDevice);
  long long c;
  float o[2];
#define RUN(name, kern)                                                              \
  kern<<<1, 32>>>(d_inc, d_out, d_c, T, 2);                                          \
  kern<<<1, 32>>>(d_inc, d_out, d_c, T, reps);                                       \
  cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);                                    \
  cudaMemcpy(o, d_out, 8, cudaMemcpyDeviceToHost);                                   \
  printf("%-44s %7.1f cycles/stage  (check %.6f)\n", name, (double)c / (T * (double)reps), o[0]);
  RUN("select form, 1 lane", (chain<0, 1>));
  RUN("select form, 32 lanes", (chain<0, 32>));
  RUN("fma-fold form, 1 lane", (chain<1, 1>));
  RUN("fma-fold form, 32 lanes", (chain<1, 32>));
  RUN("dependent FADD x1, 1 lane", (fadd_chain<1>));
  RUN("dependent FADD x1, 32 lanes", (fadd_chain<32>));
  RUN("smem select, release flag", (chain_smem<0, false, true>));
  RUN("smem fma-fold, release flag", (chain_smem<1, false, true>));
  RUN("smem fma-fold, prefetch, release flag", (chain_smem<1, true, true>));
  RUN("smem fma-fold, prefetch, volatile flag", (chain_smem<1, true, false>));
  RUN("smem fma-fold, no prefetch, volatile flag", (chain_smem<1, false, false>));
#undef RUN
#define RUN(name, kern)                                                              \
  kern<<<1, 256>>>(d_inc, d_out, d_c, T, 2);                                         \
  kern<<<1, 256>>>(d_inc, d_out, d_c, T, reps);                                      \
  cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);                                    \
  cudaMemcpy(o, d_out, 8, cudaMemcpyDeviceToHost);                                   \
  printf("%-44s %7.1f cycles/stage  (check %.6f)\n", name, (double)c / (T * (double)reps), o[0]);
  RUN("smem select, 256 thr, others parked at bar", (chain_smem<0, false, true, 1>));
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
