// tcgen05.mma issue/throughput probe for the attention-sized shapes (one CTA per SM, one issuing thread).
// Prints cycles per MMA instruction for back-to-back instructions of one shape / operand source.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_mma tools/ubench_mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../avt_b200/csrc/ptx.cuh"

using namespace avt;


// A_SRC 0: smem K-major, 1: smem MN-major (2 blocks), 2: TMEM.  Fully unrolled k-steps, compile-time descriptors.
template <int N, int A_SRC, int B_MN, int KSTEPS, int CHAINS>
__global__ void __launch_bounds__(128, 1) mma_probe(int reps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint64_t dK = smem_desc_sw128(16, 1024), dMN1 = smem_desc_sw128(8192, 1024), dMN2 = smem_desc_sw128(16384, 1024);
    const uint32_t aA = smem_u32(smem), aB = smem_u32(smem) + 64 * 1024;
    constexpr uint32_t idesc = umma_idesc(1, A_SRC == 1 ? 1 : 0, B_MN, 128, N);
    umma_f16(tm, smem_desc_addr(dK, aA), smem_desc_addr(dK, aB), umma_idesc(1, 0, 0, 128, 64), 0);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) {
        const uint32_t d = tm + c * 256;
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) {
          const uint64_t bdesc = B_MN ? smem_desc_addr(dMN1, aB + (k & 7) * 2048) : smem_desc_addr(dK, aB + (k & 3) * 32);
          if (A_SRC == 2) umma_f16_ts(d, tm + 480 + (k & 3) * 8, bdesc, idesc, k > 0);
          else if (A_SRC == 1) umma_f16(d, smem_desc_addr(dMN2, aA + (k & 7) * 2048), bdesc, idesc, k > 0);
          else umma_f16(d, smem_desc_addr(dK, aA + (k & 3) * 32), bdesc, idesc, k > 0);
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    const long long t1 = clock64();
    cycles[blockIdx.x * 2] = t1 - t0;
    cycles[blockIdx.x * 2 + 1] = (long long)reps * CHAINS * KSTEPS;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, int A_SRC, int B_MN, int KSTEPS, int CHAINS>
void run(long long* cyc, const char* what) {
  auto kern = mma_probe<N, A_SRC, B_MN, KSTEPS, CHAINS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long h[2 * 148];
  for (int it = 0; it < 2; ++it) kern<<<148, 128, 200 * 1024>>>(64, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe %s failed: %s\n", what, cudaGetErrorString(e)); exit(1); }
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("mma M=128 N=%3d K=16 a_src=%d b_mn=%d chains=%d ksteps=%2d : %6.1f clk/instr (floor %3.0f)  %s\n", N, A_SRC, B_MN, CHAINS,
         KSTEPS, (double)h[0] / h[1], 128.0 * N / 256.0, what);
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 2 * 148 * sizeof(long long));
  run<64, 0, 1, 13, 1>(cyc, "PV-like, P in smem");
  run<64, 0, 1, 13, 2>(cyc, "PV-like, P in smem, 2 accumulators");
  run<64, 2, 1, 13, 1>(cyc, "PV-like, P in TMEM");
  run<64, 2, 1, 13, 2>(cyc, "PV-like, P in TMEM, 2 accumulators");
  run<208, 0, 0, 4, 1>(cyc, "S = Q K^T");
  run<208, 0, 0, 4, 2>(cyc, "S = Q K^T, 2 accumulators");
  run<64, 0, 0, 4, 1>(cyc, "bwd S^T chunk");
  run<64, 0, 0, 4, 2>(cyc, "bwd S^T chunk, 2 accumulators");
  run<128, 0, 0, 4, 1>(cyc, "N=128");
  run<256, 0, 0, 4, 1>(cyc, "GEMM-like");
  run<64, 1, 1, 8, 1>(cyc, "dQ-like (A MN-major)");
  run<208, 2, 0, 4, 1>(cyc, "A in TMEM, N=208");
  run<128, 2, 1, 8, 1>(cyc, "A in TMEM, N=128, B MN");
  run<16, 0, 0, 4, 1>(cyc, "N=16");
  run<32, 0, 1, 8, 1>(cyc, "N=32");
  return 0;
}
