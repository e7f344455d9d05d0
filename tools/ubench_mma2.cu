// Does the rate at which ONE thread can push tcgen05.mma (M=128, N=64, K=16, A in TMEM: the P V of attention) depend on what
// the other warps of the SM are doing? 13 warps; warp 12 issues 64 x 13 MMAs, warps 0-11 run `mode` until it is done:
//   0 idle (wait on a flag)   1 FMA/ALU busy loop   2 MUFU ex2 loop   3 tcgen05.ld loop   4 tcgen05.ld + st loop
//   5 mbarrier try_wait spin  6 tcgen05.ld of OTHER columns than the MMA's accumulator
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_mma2 tools/ubench_mma2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../avt_b200/csrc/ptx.cuh"
using namespace avt;

__global__ void __launch_bounds__(416, 1) probe(int mode, int nbusy, int n_mma, long long* cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar_never;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 12) {
    if (lane == 0) { mbar_init(&bar, 1); mbar_init(&bar_never, 1); fence_mbar_init(); done = 0; }
    __syncwarp();
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = slot;
  if (warp == 12) {
    if (lane == 0) {
      constexpr uint64_t dMN1 = smem_desc_sw128(8192, 1024);
      constexpr uint32_t idesc = umma_idesc(1, 0, 1, 128, 64);
      const uint32_t aB = smem_u32(smem);
      const long long t0 = clock64();
      for (int r = 0; r < 64; ++r) {
#pragma unroll
        for (int k = 0; k < 13; ++k)
          if (k < n_mma) umma_f16_ts(tm + 416, tm + (r & 1) * 208 + (k < 6 ? 8 * k : 152 + 8 * (k - 6)), smem_desc_addr(dMN1, aB + k * 2048), idesc, k > 0);
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      cycles[blockIdx.x] = t1 - t0;
      done = 1;
    }
    __syncwarp();
  } else if (warp < nbusy) {
    const uint32_t tl = tm + (uint32_t((warp & 3) * 32) << 16);
    float acc = threadIdx.x;
    uint32_t v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = threadIdx.x + j;
    while (!done) {
      if (mode == 1) {
#pragma unroll
        for (int j = 0; j < 64; ++j) acc = fmaf(acc, 1.0001f, 0.5f);
      } else if (mode == 2 || mode == 7 || mode == 8) {
        if (mode == 7 && (warp & 3) == 0) { __nanosleep(100); continue; }   // keep the issuer's sub-partition free of MUFU work
        if (mode == 8 && (warp & 3) != 0) { __nanosleep(100); continue; }   // MUFU work ONLY on the issuer's sub-partition
        float w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = acc - j;
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
          for (int j = 0; j < 16; ++j) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(w[j])); w[j] = y - 1.5f; }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += w[j];
      } else if (mode == 3) {
        tmem_ld_32x32b_x16(tl + 416 + 16 * (warp >> 2), v);   // the accumulator columns
        tmem_ld_wait();
        acc += __uint_as_float(v[3]);
      } else if (mode == 4) {
        tmem_ld_32x32b_x16(tl + 16 * (warp >> 2), v);
        tmem_ld_wait();
        tmem_st_32x32b_x16(tl + 208 + 16 * (warp >> 2), v);
        tmem_st_wait();
      } else if (mode == 9 || mode == 10 || mode == 11) {
        uint32_t a[32], b[32];
        if (mode == 9 || (mode == 11 && warp < 6)) {           // back-to-back wide loads: ~300 B/clk/SM of TMEM read traffic
          tmem_ld_32x32b_x32(tl + 16 * (warp >> 2), a);
          tmem_ld_32x32b_x32(tl + 208 + 16 * (warp >> 2), b);
          tmem_ld_wait();
          acc += __uint_as_float(a[3] ^ b[5]);
        } else {                                               // back-to-back stores
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] = threadIdx.x + j;
          tmem_st_32x32b_x16(tl + 230 + 16 * (warp >> 2), v);
          tmem_st_32x32b_x16(tl + 300 + 16 * (warp >> 2), v);
          tmem_st_wait();
        }
      } else if (mode == 5) {
        mbar_try_wait(&bar_never, 0);
      } else if (mode == 6) {
        tmem_ld_32x32b_x16(tl + 300 + 16 * (warp >> 2), v);
        tmem_ld_wait();
        acc += __uint_as_float(v[3]);
      } else {
        __nanosleep(100);
      }
    }
    if (acc == 123.456f) sink[0] = acc;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tm, 512);
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long)); cudaMalloc(&sink, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const char* names[] = {"idle", "fma loop", "mufu saturating (16 chains)", "tmem ld (accumulator cols)", "tmem ld+st (P cols)", "mbarrier try_wait spin", "tmem ld (other cols)", "mufu, not on issuer SMSP", "mufu, only on issuer SMSP", "tmem wide loads back to back", "tmem stores back to back", "tmem loads + stores"};
  for (int mode : {0, 9, 10, 11})
    for (int nbusy : {12, 8}) {
      long long h[148];
      for (int it = 0; it < 2; ++it) probe<<<148, 416, 160 * 1024>>>(mode, nbusy, 13, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d failed: %s\n", mode, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      printf("others: %-28s x%2d warps : %.1f clk per PV MMA\n", names[mode], nbusy, (double)h[0] / (64 * 13));
    }
  return 0;
}
