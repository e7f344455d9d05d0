// Micro-benchmarks that decide the attention-kernel design (run on the GPU box; prints one line per probe):
//   * tcgen05.ld / tcgen05.st throughput per SM for 1..16 warps (is the softmax TMEM-read bound?)
//   * ex2.approx f32 vs f16x2 vs a degree-3 polynomial on the FMA pipe (what bounds exp for P?)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench.bin tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../avt_b200/csrc/ptx.cuh"

using namespace avt;

__global__ void tmem_ld_kernel(int reps, long long* cycles, uint32_t* sink, int do_store) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t a[32], b[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) { a[j] = j; b[j] = j; }
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (!do_store) {
    for (int r = 0; r < reps; ++r) {
      tmem_ld_32x32b_x32(tm + ((r * 64) & 255), a);
      tmem_ld_32x32b_x32(tm + ((r * 64 + 32) & 255) + 256 * ((warp >> 2) & 1), b);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += a[j] ^ b[j];
    }
  } else {
    uint32_t v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = j + threadIdx.x;
    for (int r = 0; r < reps; ++r) {
      tmem_st_32x32b_x16(tm + ((r * 64) & 255), v);
      tmem_st_32x32b_x16(tm + ((r * 64 + 16) & 255), v);
      tmem_st_32x32b_x16(tm + ((r * 64 + 32) & 255), v);
      tmem_st_32x32b_x16(tm + ((r * 64 + 48) & 255), v);
      tmem_st_wait();
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

// mode 0: ex2.approx.ftz.f32   1: ex2.approx.f16x2 (2 results per op)   2: polynomial exp2 on packed fp32x2
// 3: half the values on MUFU (f32), half on the polynomial
__device__ __forceinline__ float2 poly_exp2_2(float2 x) {
  // 2^x, x <= 0: round-to-nearest split x = n + f, f in [-0.5, 0.5]; 2^f by a cubic; exponent added as an integer
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  x.x = fmaxf(x.x, -125.0f); x.y = fmaxf(x.y, -125.0f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
  float2 p = __ffma2_rn(f, make_float2(0.05517167f, 0.05517167f), make_float2(0.24261113f, 0.24261113f));
  p = __ffma2_rn(p, f, make_float2(0.69326099f, 0.69326099f));
  p = __ffma2_rn(p, f, make_float2(0.99992807f, 0.99992807f));
  float2 r;
  r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return r;
}

__global__ void exp_kernel(int mode, int reps, long long* cycles, float* sink) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = -0.001f * (threadIdx.x + j);
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v[j]));
        v[j] = y - 1.5f;
      }
    } else if (mode == 1) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        __half2 h = __floats2half2_rn(v[j], v[j + 1]);
        uint32_t hi = *reinterpret_cast<uint32_t*>(&h), ho;
        asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(ho) : "r"(hi));
        const float2 f = __half22float2(*reinterpret_cast<__half2*>(&ho));
        v[j] = f.x - 1.5f;
        v[j + 1] = f.y - 1.5f;
      }
    } else if (mode == 2) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float2 y = poly_exp2_2(make_float2(v[j], v[j + 1]));
        v[j] = y.x - 1.5f;
        v[j + 1] = y.y - 1.5f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float2 y = poly_exp2_2(make_float2(v[j], v[j + 1]));
        v[j] = y.x - 1.5f;
        v[j + 1] = y.y - 1.5f;
        float y2, y3;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y2) : "f"(v[j + 2]));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y3) : "f"(v[j + 3]));
        v[j + 2] = y2 - 1.5f;
        v[j + 3] = y3 - 1.5f;
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += v[j];
  if (s == 123.456f) sink[0] = s;
}

__global__ void poly_err_kernel(float* out) {   // max relative error of poly_exp2_2 over [-30, 0]
  float worst = 0.f;
  for (int i = threadIdx.x; i < 3000000; i += blockDim.x) {
    const float x = -1e-5f * i;
    const float2 y = poly_exp2_2(make_float2(x, x));
    const float ref = exp2f(x);
    worst = fmaxf(worst, fabsf(y.x - ref) / ref);
  }
  atomicMax(reinterpret_cast<int*>(out), __float_as_int(worst));
}

int main() {
  long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, 1024 * sizeof(long long));
  cudaMalloc(&sink, 64);
  long long h[148];
  const int reps = 2048;
  for (int st = 0; st < 2; ++st)
    for (int warps : {1, 2, 4, 8, 16}) {
      for (int it = 0; it < 2; ++it) tmem_ld_kernel<<<148, warps * 32>>>(reps, cyc, sink, st);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("tmem kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes = (double)warps * reps * (st ? 4 * 16 : 2 * 32) * 32 * 4;
      printf("tmem_%s warps=%2d  cycles=%lld  bytes/clk/SM=%.1f\n", st ? "st" : "ld", warps, h[0], bytes / h[0]);
    }
  for (int mode = 0; mode < 4; ++mode)
    for (int warps : {4, 8, 16}) {
      for (int it = 0; it < 2; ++it) exp_kernel<<<148, warps * 32>>>(mode, 4096, cyc, reinterpret_cast<float*>(sink));
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("exp kernel failed\n"); return 1; }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double n = (double)warps * 32 * 16 * 4096;
      printf("exp mode=%d warps=%2d  cycles=%lld  results/clk/SM=%.2f\n", mode, warps, h[0], n / h[0]);
    }
  float* e;
  cudaMalloc(&e, 4);
  cudaMemset(e, 0, 4);
  poly_err_kernel<<<1, 256>>>(e);
  float he;
  cudaMemcpy(&he, e, 4, cudaMemcpyDeviceToHost);
  printf("poly exp2 max rel err %.3e\n", he);
  return 0;
}
