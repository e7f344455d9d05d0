// bf16 GEMM with fused epilogues on tcgen05 tensor cores (sm_100a).
//
//   C[M,N] = epilogue(A[M,K] · B[N,K]^T)     fp32 accumulation in tensor memory (TMEM)
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0      TMA producer  : cp.async.bulk.tensor tiles of A and B into a 128B-swizzled smem ring
//   warp 1      MMA issuer    : one lane issues tcgen05.mma 128 x BN x 16 per 32-byte K step; tcgen05.commit
//                               releases smem slots and publishes the finished accumulator
//   warps 2..9  epilogue      : tcgen05.ld the accumulator (one row per thread), apply the fused
//                               epilogue (bias / GELU / dGELU / dropout / residual / pos-embed) and store
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// Either operand may be K-major (rows of the contraction dim contiguous) or MN-major (transposed in
// memory): forward, dgrad and wgrad of nn.Linear and HF Conv1D all map onto this one kernel without
// materialising a transpose. Split-K work units accumulate with fp32 atomics (weight gradients).
#include <cuda.h>
#include "common.cuh"
#include "ptx.cuh"

namespace avt {

constexpr int kBM = 128;        // tile rows  (UMMA M)
constexpr int kBK = 64;         // K per smem stage: 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;      // K per tcgen05.mma for 16-bit inputs
constexpr int kEpiWarps = 8;     // two warps per TMEM lane quarter, each takes half of the tile's columns
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN >= 256 ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // double-buffered accumulator
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_k_blocks, split_k, kb_per_split;
  avt_epilogue_t ep;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                       // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;   // [2]        MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int num_units = p.num_m_tiles * p.num_n_tiles * p.split_k;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int tile = unit / p.split_k, split = unit % p.split_k;
      const int m0 = (tile / p.num_n_tiles) * kBM, n0 = (tile % p.num_n_tiles) * BN;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sA = smem + stage * Cfg::kStageBytes;
          uint8_t* sB = sA + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if constexpr (!A_MN) {
            tma_load_2d(&tmA, &full_bar[stage], sA, kb * kBK, m0);  // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < kBM / 64; ++j)  // box {64 m, 64 k-rows} per 64-wide M block
              tma_load_2d(&tmA, &full_bar[stage], sA + j * 8192, m0 + j * 64, kb * kBK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(&tmB, &full_bar[stage], sB, kb * kBK, n0);  // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(&tmB, &full_bar[stage], sB + j * 8192, n0 + j * 64, kb * kBK);
          }
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc = umma_idesc(/*bf16*/ 1, A_MN ? 1 : 0, B_MN ? 1 : 0, kBM, BN);
    // K-major SW128: 8-row groups 1024 B apart (SBO); LBO unused.  MN-major SW128: 64-wide MN blocks
    // 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO).
    constexpr uint64_t descA = A_MN ? smem_desc_sw128(8192, 1024) : smem_desc_sw128(16, 1024);
    constexpr uint64_t descB = B_MN ? smem_desc_sw128(8192, 1024) : smem_desc_sw128(16, 1024);
    constexpr uint32_t kStepA = A_MN ? kUmmaK * 128 : kUmmaK * 2;  // bytes per 16-wide K step
    constexpr uint32_t kStepB = B_MN ? kUmmaK * 128 : kUmmaK * 2;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int split = unit % p.split_k;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        if (lane == 0) {
          const uint32_t sA = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sB = sA + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            umma_f16(d_tmem, smem_desc_addr(descA, sA + k * kStepA), smem_desc_addr(descB, sB + k * kStepB), idesc,
                     (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                    // smem slot reusable once these MMAs retire
          if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);   // accumulator complete
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ============================== epilogue (warps 2..9) ==============================
    const avt_epilogue_t& ep = p.ep;
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the only ones this warp may read
    const int chalf = (warp - 2) >> 2;  // which half of the tile's columns this warp handles
    int acc = 0;
    uint32_t acc_phase = 0;
    const float keep_scale = ep.drop_p > 0.f ? 1.0f / (1.0f - ep.drop_p) : 1.0f;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int tile = unit / p.split_k;
      const int m0 = (tile / p.num_n_tiles) * kBM, n0 = (tile % p.num_n_tiles) * BN;
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
      const int pos_t = ep.pos_period > 0 ? row % ep.pos_period : 0;
#pragma unroll 1
      for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + c0, r);
        tmem_ld_wait();
        const int col0 = n0 + c0;
        if (row_ok && col0 < p.N) {  // N is a multiple of 32 on every path that reaches here (checked on host)
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * ep.alpha;
          if (ep.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (ep.pos_period > 0) {
            if (pos_t == 0 && ep.cls) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(ep.cls + col0 + j));
                v[j] = b.x; v[j + 1] = b.y; v[j + 2] = b.z; v[j + 3] = b.w;
              }
            }
            const float* pp = ep.pos + (size_t)pos_t * p.N + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(pp + j));
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
          if (ep.aux_z) {
            uint4* zp = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.aux_z) + (size_t)row * ep.ldz + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              zp[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
          if (ep.act != AVT_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(ep.act, v[j]);
          }
          if (ep.dact_z) {
            const uint4* zp =
                reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(ep.dact_z) + (size_t)row * ep.ldz + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 z = __ldg(zp + j);
              const uint32_t zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[8 * j + 2 * q] *= apply_act_grad(ep.dact, bf16_lo(zz[q]));
                v[8 * j + 2 * q + 1] *= apply_act_grad(ep.dact, bf16_hi(zz[q]));
              }
            }
          }
          if (ep.drop_p > 0.f) {
            const uint64_t g0 = ((uint64_t)row * (uint64_t)p.N + (uint64_t)col0) >> 2;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t keep = dropout_keep4(ep.drop_seed, ep.drop_offset, g0 + j, ep.drop_p);
#pragma unroll
              for (int q = 0; q < 4; ++q) v[4 * j + q] = ((keep >> q) & 1u) ? v[4 * j + q] * keep_scale : 0.f;
            }
          }
          if (ep.residual) {
            const float4* rp = reinterpret_cast<const float4*>(ep.residual + (size_t)row * ep.ldr + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(rp + j);
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
          if (ep.out_fp32) {
            float* op = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo + col0;
            if (p.split_k > 1) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                atomicAdd(reinterpret_cast<float4*>(op + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
            } else if (ep.accumulate) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 o = *reinterpret_cast<float4*>(op + j);
                o.x += v[j]; o.y += v[j + 1]; o.z += v[j + 2]; o.w += v[j + 3];
                *reinterpret_cast<float4*>(op + j) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.out) + (size_t)row * ep.ldo + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              op[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows of `ld` elements; 128B swizzle.
int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                      uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled", "driver entry point not available", __FILE__, __LINE__);
    return AVT_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "CUresult %d (base %p inner %llu outer %llu ld %llu box %u x %u)", (int)r, base,
             (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
    set_last_error("cuTensorMapEncodeTiled", msg, __FILE__, __LINE__);
    return AVT_ERR_CUDA;
  }
  return AVT_OK;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    AVT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int units = p.num_m_tiles * p.num_n_tiles * p.split_k;
  const int grid = units < num_sms() ? units : num_sms();
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

template <int BN>
static int dispatch_major(int a_mn, int b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p,
                          cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false>(tmA, tmB, p, s);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true>(tmA, tmB, p, s);
  if (a_mn && !b_mn) return launch_gemm<BN, true, false>(tmA, tmB, p, s);
  return launch_gemm<BN, true, true>(tmA, tmB, p, s);
}

}  // namespace avt

using namespace avt;

extern "C" int avt_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int64_t M,
                             int64_t N, int64_t K, const avt_epilogue_t* ep, int split_k, int block_n, void* stream) {
  AVT_REQUIRE(A && B && ep && ep->out, "null pointer");
  AVT_REQUIRE(M > 0 && N > 0 && K > 0, "empty problem");
  AVT_REQUIRE(N % 32 == 0, "N must be a multiple of 32");
  AVT_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "leading dimensions must be multiples of 8 elements (16 bytes)");
  AVT_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
              "operands must be 16-byte aligned");
  AVT_REQUIRE(ep->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(ep->out) & 15) == 0, "output must be 16-byte aligned");
  if (block_n <= 0) block_n = (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : 64);
  AVT_REQUIRE(block_n == 64 || block_n == 128 || block_n == 256, "block_n must be 64, 128 or 256");
  if (split_k < 1) split_k = 1;
  if (split_k > 1) {
    AVT_REQUIRE(ep->out_fp32 && !ep->bias && !ep->aux_z && !ep->dact_z && !ep->residual && ep->act == AVT_ACT_NONE &&
                    ep->drop_p == 0.f && ep->pos_period == 0,
                "split_k > 1 supports only fp32 accumulation into out");
  }
  if (ep->residual) AVT_REQUIRE(ep->ldr % 4 == 0, "residual ld must be a multiple of 4");
  if (ep->aux_z || ep->dact_z) AVT_REQUIRE(ep->ldz % 8 == 0, "z ld must be a multiple of 8");

  GemmParams p;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.num_m_tiles = (int)((M + kBM - 1) / kBM);
  p.num_n_tiles = (int)((N + block_n - 1) / block_n);
  p.num_k_blocks = (int)((K + kBK - 1) / kBK);
  if (split_k > p.num_k_blocks) split_k = p.num_k_blocks;
  p.kb_per_split = (p.num_k_blocks + split_k - 1) / split_k;
  p.split_k = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.ep = *ep;
  if (p.ep.alpha == 0.f) p.ep.alpha = 1.0f;

  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, kBK, kBM);
  else       rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, kBK);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, kBK, (uint32_t)block_n);
  else       rc = make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, kBK);
  if (rc) return rc;

  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (block_n) {
    case 64: return dispatch_major<64>(a_mn, b_mn, tmA, tmB, p, s);
    case 128: return dispatch_major<128>(a_mn, b_mn, tmA, tmB, p, s);
    default: return dispatch_major<256>(a_mn, b_mn, tmA, tmB, p, s);
  }
}
