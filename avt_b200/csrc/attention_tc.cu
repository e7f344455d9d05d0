// ViT spatial attention on tcgen05 tensor cores: softmax(Q K^T * scale) V for N <= 208 tokens, head_dim 64.
//
// One CTA per (frame, head). The whole key range fits one MMA N extent (208 = 13*16 >= 197), so there is
// no online-softmax rescaling: S = Q K^T lands in TMEM in one shot, each softmax thread owns one query row
// (TMEM lane) and needs no shuffles, P is written as bf16 into 128B-swizzled smem (K-major A operand),
// O = P V accumulates in TMEM over the dead S columns.
//   warps 0-3 : softmax + epilogue for query rows   0..127 (TMEM lanes = rows, S at columns   0..207)
//   warps 4-7 : softmax + epilogue for query rows 128..255 (                   S at columns 256..463)
//   warp  8   : TMA loads (Q tiles, K, V straight out of the packed qkv matrix) and all tcgen05.mma issue
// Layout: qkv bf16 [F*N, 3*D], column = s*D + h*64 + d (timm Attention.qkv packing); out bf16 [F*N, D].
// Replaces timm Attention.forward: q@k^T*scale -> softmax -> attn@v (4 kernels + 80*12*197^2 score tensor).
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"

namespace avt {

int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                      uint32_t box_outer, int swizzle_bytes);

constexpr int kTcHd = 64;
constexpr int kTcKeys = 208;            // MMA N extent / PV contraction length (multiple of 16)
constexpr int kTcQBytes = 128 * 128;    // one Q tile: 128 rows x 64 bf16
constexpr int kTcKVBytes = kTcKeys * 128;
constexpr int kTcPBlk = 128 * 128;      // one P block: 128 rows x 64 keys bf16
constexpr int kTcPBytes = 4 * kTcPBlk;  // keys padded to 256 in smem addressing (only 208 are read)
constexpr int kTcSmem = 2 * kTcQBytes + 2 * kTcKVBytes + 2 * kTcPBytes + 256;
constexpr int kTcThreads = 288;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnTcParams {
  bf16* out;
  float* lse;
  int N, H, D, F;
  float scale;
};

__global__ void __launch_bounds__(kTcThreads, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnTcParams p) {
  // Persistent: each CTA walks (frame, head) items blockIdx.x, +gridDim.x, ... Single-buffered smem, but every
  // buffer is refilled as soon as its last reader retires (Q/K after both S MMAs, V after both PV MMAs), so the
  // next item's loads hide under the current item's softmax.
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                         // [2][128 x 64]   K-major SW128
  uint8_t* sK = sQ + 2 * kTcQBytes;           // [208 x 64]      K-major SW128 (B of S = Q K^T)
  uint8_t* sV = sK + kTcKVBytes;              // [208 x 64]      MN-major SW128 (B of O = P V)
  uint8_t* sP = sV + kTcKVBytes;              // [2][4][128 x 64] K-major SW128 (A of O = P V)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kTcPBytes);
  uint64_t* bar_qk = bars;          // TMA (Q tiles + K) landed
  uint64_t* bar_v = bars + 1;       // TMA (V) landed
  uint64_t* bar_s = bars + 2;       // [2] S ready                      (MMA -> softmax warps)
  uint64_t* bar_p = bars + 4;       // [2] P written, 128 arrivals      (softmax warps -> MMA)
  uint64_t* bar_o = bars + 6;       // [2] O ready                      (MMA -> softmax warps)
  uint64_t* bar_oread = bars + 8;   // [2] O drained from TMEM, 128 arrivals (S columns reusable)
  uint64_t* bar_qkfree = bars + 10; // both S MMAs retired: sQ / sK reusable
  uint64_t* bar_vfree = bars + 11;  // both PV MMAs retired: sV reusable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N;
  const int items = p.F * p.H;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      mbar_init(bar_qk, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_qkfree, 1);
      mbar_init(bar_vfree, 1);
      for (int t = 0; t < 2; ++t) {
        mbar_init(&bar_s[t], 1);
        mbar_init(&bar_p[t], 128);
        mbar_init(&bar_o[t], 1);
        mbar_init(&bar_oread[t], 128);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // prologue above overlapped the previous kernel's tail
  pdl_trigger();

  if (warp == 8) {
    // ------------------------------------------------------------- control warp: one lane issues TMA + MMA
    if (lane == 0) {
      auto issue_qk = [&](int item) {
        const int f = item / p.H, h = item % p.H;
        mbar_arrive_expect_tx(bar_qk, 2 * kTcQBytes + kTcKVBytes);
        tma_load_2d(&tmQ, bar_qk, sQ, h * kTcHd, f * N);
        tma_load_2d(&tmQ, bar_qk, sQ + kTcQBytes, h * kTcHd, f * N + 128);
        tma_load_2d(&tmKV, bar_qk, sK, p.D + h * kTcHd, f * N);
      };
      auto issue_v = [&](int item) {
        const int f = item / p.H, h = item % p.H;
        mbar_arrive_expect_tx(bar_v, kTcKVBytes);
        tma_load_2d(&tmKV, bar_v, sV, 2 * p.D + h * kTcHd, f * N);
      };
      constexpr uint32_t idesc_s = umma_idesc(1, 0, 0, 128, kTcKeys);
      constexpr uint32_t idesc_o = umma_idesc(1, 0, 1, 128, kTcHd);
      constexpr uint64_t desc_k = smem_desc_sw128(16, 1024);     // K-major
      constexpr uint64_t desc_v = smem_desc_sw128(8192, 1024);   // MN-major, one 64-wide block
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
      if ((int)blockIdx.x < items) {
        issue_qk(blockIdx.x);
        issue_v(blockIdx.x);
      }
      uint32_t n = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
        const uint32_t par = n & 1;
        const int next = item + gridDim.x;
        mbar_wait(bar_qk, par);
        for (int t = 0; t < 2; ++t) {
          if (n > 0) mbar_wait(&bar_oread[t], par ^ 1);   // previous item's O (aliases these S columns) was drained
          tc_fence_after_sync();
          const uint32_t aQ = smem_u32(sQ + t * kTcQBytes);
#pragma unroll
          for (int k = 0; k < kTcHd / 16; ++k)
            umma_f16(tmem_base + 256 * t, smem_desc_addr(desc_k, aQ + k * 32), smem_desc_addr(desc_k, aK + k * 32), idesc_s,
                     k > 0 ? 1u : 0u);
          umma_commit(&bar_s[t]);
        }
        umma_commit(bar_qkfree);
        if (next < items) {
          mbar_wait(bar_qkfree, par);
          issue_qk(next);
        }
        mbar_wait(bar_v, par);
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&bar_p[t], par);
          tc_fence_after_sync();
          const uint32_t aP = smem_u32(sP + t * kTcPBytes);
#pragma unroll
          for (int k = 0; k < kTcKeys / 16; ++k)
            umma_f16(tmem_base + 256 * t, smem_desc_addr(desc_k, aP + (k >> 2) * kTcPBlk + (k & 3) * 32),
                     smem_desc_addr(desc_v, aV + k * 2048), idesc_o, k > 0 ? 1u : 0u);
          umma_commit(&bar_o[t]);
        }
        umma_commit(bar_vfree);
        if (next < items) {
          mbar_wait(bar_vfree, par);
          issue_v(next);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------- softmax + epilogue warpgroups
    const int t = warp >> 2;                    // query tile
    const int r = (warp & 3) * 32 + lane;       // row within the tile == TMEM lane
    const int qrow = t * 128 + r;
    const uint32_t t_row = tmem_base + (uint32_t((warp & 3) * 32) << 16) + 256 * t;
    const float sl2 = p.scale * 1.4426950408889634f;
    uint8_t* prow = sP + t * kTcPBytes + r * 128;
    const int sw = r & 7;
    const bool ok = qrow < N;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
      const uint32_t par = n & 1;
      const int f = item / p.H, h = item % p.H;
      mbar_wait(&bar_s[t], par);
      tc_fence_after_sync();
      // pass 1: row max over the valid keys. Columns < nfull need no mask (N = 197: 6 of the 7 chunks); TMEM loads
      // are double-buffered so the next chunk is in flight while the current one is reduced.
      const int nfull = N >= 192 ? 192 : (N & ~31);
      float mx = -INFINITY;
      {
        uint32_t va[32], vb[32];
        tmem_ld_32x32b_x32(t_row, va);
#pragma unroll 1
        for (int c0 = 0; c0 < 192; c0 += 64) {
          tmem_ld_wait();
          tmem_ld_32x32b_x32(t_row + c0 + 32, vb);
          if (c0 < nfull) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(va[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < N) mx = fmaxf(mx, __uint_as_float(va[j]));
          }
          tmem_ld_wait();
          if (c0 + 64 < 192) tmem_ld_32x32b_x32(t_row + c0 + 64, va);
          if (c0 + 32 < nfull) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(vb[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + 32 + j < N) mx = fmaxf(mx, __uint_as_float(vb[j]));
          }
        }
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_row + 192, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (192 + j < N) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      // pass 2: e = exp(scale * (s - max)), row sum, P -> smem (bf16, swizzled K-major)
      float sum = 0.f;
      const float mxs = mx * sl2;
      auto emit = [&](const uint32_t (&v)[32], int c0) {
        float e[32];
        if (c0 < nfull) {
#pragma unroll
          for (int j = 0; j < 32; ++j) e[j] = fast_exp2(fmaf(__uint_as_float(v[j]), sl2, -mxs));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) e[j] = (c0 + j < N) ? fast_exp2(fmaf(__uint_as_float(v[j]), sl2, -mxs)) : 0.f;
        }
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) s4[j & 3] += e[j];
        sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
        uint8_t* blk = prow + (c0 >> 6) * kTcPBlk;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c0 & 63) >> 3) + q;
          *reinterpret_cast<uint4*>(blk + ((chunk ^ sw) << 4)) =
              make_uint4(pack_bf16x2(e[8 * q], e[8 * q + 1]), pack_bf16x2(e[8 * q + 2], e[8 * q + 3]),
                         pack_bf16x2(e[8 * q + 4], e[8 * q + 5]), pack_bf16x2(e[8 * q + 6], e[8 * q + 7]));
        }
      };
      {
        uint32_t va[32], vb[32];
        tmem_ld_32x32b_x32(t_row, va);
#pragma unroll 1
        for (int c0 = 0; c0 < 192; c0 += 64) {
          tmem_ld_wait();
          tmem_ld_32x32b_x32(t_row + c0 + 32, vb);
          emit(va, c0);
          tmem_ld_wait();
          if (c0 + 64 < 192) tmem_ld_32x32b_x32(t_row + c0 + 64, va);
          emit(vb, c0 + 32);
        }
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_row + 192, v);
        tmem_ld_wait();
        float e[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          e[j] = (192 + j < N) ? fast_exp2(fmaf(__uint_as_float(v[j]), sl2, -mxs)) : 0.f;
          sum += e[j];
        }
        uint8_t* blk = prow + 3 * kTcPBlk;
#pragma unroll
        for (int q = 0; q < 2; ++q)
          *reinterpret_cast<uint4*>(blk + ((q ^ sw) << 4)) =
              make_uint4(pack_bf16x2(e[8 * q], e[8 * q + 1]), pack_bf16x2(e[8 * q + 2], e[8 * q + 3]),
                         pack_bf16x2(e[8 * q + 4], e[8 * q + 5]), pack_bf16x2(e[8 * q + 6], e[8 * q + 7]));
      }
      fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before_sync();     // our tcgen05.ld of S are ordered before the MMA that overwrites those columns
      mbar_arrive(&bar_p[t]);
      // epilogue: O / sum -> bf16 -> global
      mbar_wait(&bar_o[t], par);
      tc_fence_after_sync();
      const float inv = 1.0f / sum;
      uint32_t o0[32], o1[32];
      tmem_ld_32x32b_x32(t_row, o0);
      tmem_ld_32x32b_x32(t_row + 32, o1);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&bar_oread[t]);   // the next item's S MMA may overwrite these columns now
      if (ok) {
        bf16* orow = p.out + ((size_t)f * N + qrow) * p.D + h * kTcHd;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<uint4*>(orow + 8 * q) = make_uint4(
              pack_bf16x2(__uint_as_float(o0[8 * q]) * inv, __uint_as_float(o0[8 * q + 1]) * inv),
              pack_bf16x2(__uint_as_float(o0[8 * q + 2]) * inv, __uint_as_float(o0[8 * q + 3]) * inv),
              pack_bf16x2(__uint_as_float(o0[8 * q + 4]) * inv, __uint_as_float(o0[8 * q + 5]) * inv),
              pack_bf16x2(__uint_as_float(o0[8 * q + 6]) * inv, __uint_as_float(o0[8 * q + 7]) * inv));
          *reinterpret_cast<uint4*>(orow + 32 + 8 * q) = make_uint4(
              pack_bf16x2(__uint_as_float(o1[8 * q]) * inv, __uint_as_float(o1[8 * q + 1]) * inv),
              pack_bf16x2(__uint_as_float(o1[8 * q + 2]) * inv, __uint_as_float(o1[8 * q + 3]) * inv),
              pack_bf16x2(__uint_as_float(o1[8 * q + 4]) * inv, __uint_as_float(o1[8 * q + 5]) * inv),
              pack_bf16x2(__uint_as_float(o1[8 * q + 6]) * inv, __uint_as_float(o1[8 * q + 7]) * inv));
        }
        if (p.lse) p.lse[((size_t)f * p.H + h) * N + qrow] = mx * p.scale + __logf(sum);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ backward
// One CTA per (frame, head), "keys on lanes": for key tile j (128 keys) and query half qh (128 / 80 queries)
//   S^T  = K_j Q_qh^T            dP^T = V_j dO_qh^T                         (tcgen05, fp32 in TMEM)
//   P^T  = exp(scale*S^T - lse)  dS^T = P^T o (dP^T - delta) * scale         (threads: lane = key row)
//   dV_j += P^T dO_qh            dK_j += dS^T Q_qh         dQ_qh += dS K_j   (tcgen05; P^T / dS^T from smem)
// The same swizzled smem block is read K-major (A of dV/dK) and MN-major (A of dQ): a [128 x 64] K-major
// SW128 tile is byte-identical to an MN-major tile whose contraction runs over the 128 rows.
// TMEM columns: S^T 0..127 | dP^T 128..255 | dV 256..319 | dK 320..383 | dQ tile0 384..447 | dQ tile1 448..511.
constexpr int kBwTile = kTcKeys * 128;            // 208 rows x 64 bf16
constexpr int kBwBlk = 128 * 128;                 // [128 x 64] bf16 block
constexpr int kBwSmem = 4 * kBwTile + 4 * kBwBlk + 2 * kTcKeys * 4 + 1024 + 128;

#ifdef AVT_ATTN_TRACE
#define TRACE_DECL __shared__ unsigned long long trace_t[256]; __shared__ int trace_id[256]; __shared__ int trace_n;
#define TRACE_INIT if (threadIdx.x == 0) trace_n = 0;
#define TRACE(id) do { if (blockIdx.x == 0) { int i_ = atomicAdd(&trace_n, 1); if (i_ < 256) { trace_t[i_] = global_timer_ns(); trace_id[i_] = (id); } } } while (0)
#define TRACE_DUMP if (blockIdx.x == 0 && threadIdx.x == 0) { for (int i_ = 0; i_ < trace_n && i_ < 256; ++i_) printf("trace %d %llu\n", trace_id[i_], trace_t[i_] - trace_t[0]); }
#else
#define TRACE_DECL
#define TRACE_INIT
#define TRACE(id)
#define TRACE_DUMP
#endif

struct AttnTcBwdParams {
  const bf16* out;    // forward output  [F*N, D]
  const bf16* dout;   // [F*N, D]
  const float* lse;   // [F*H, N]
  bf16* dqkv;         // [F*N, 3D]
  int N, H, D;
  float scale;
};

__global__ void __launch_bounds__(kTcThreads, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const AttnTcBwdParams p) {
  pdl_enter();   // this kernel starts its TMA loads in the prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kBwTile;
  uint8_t* sQ = sV + kBwTile;
  uint8_t* sG = sQ + kBwTile;            // dO
  uint8_t* sP = sG + kBwTile;            // [2 blocks][128 x 64]  P^T  (rows = keys, cols = queries of this half)
  uint8_t* sS = sP + 2 * kBwBlk;         // [2 blocks][128 x 64]  dS^T
  float* sLse = reinterpret_cast<float*>(sS + 2 * kBwBlk);   // [208] lse * log2(e)
  float* sDelta = sLse + kTcKeys;                            // [208]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + kTcKeys);
  uint64_t* bar_load = bars;
  uint64_t* bar_s = bars + 1;    // S^T / dP^T ready          (MMA -> threads), one phase per iteration
  uint64_t* bar_p = bars + 2;    // P^T / dS^T written        (256 arrivals)
  uint64_t* bar_d = bars + 3;    // dV / dK / dQ MMAs retired (MMA -> threads)
  uint64_t* bar_out = bars + 4;  // dV_0 / dK_0 read out      (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int N = p.N;
  TRACE_DECL
  TRACE_INIT

  if (warp == 8) {
    if (lane == 0) {
      TRACE(0);
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmDO);
      mbar_init(bar_load, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 256);
      mbar_init(bar_d, 1);
      mbar_init(bar_out, 256);
      fence_mbar_init();
      // the loads do not depend on TMEM: get them in flight before the allocation and the block-wide sync
      mbar_arrive_expect_tx(bar_load, 4 * kBwTile);
      tma_load_2d(&tmQKV, bar_load, sQ, h * kTcHd, f * N);
      tma_load_2d(&tmQKV, bar_load, sK, p.D + h * kTcHd, f * N);
      tma_load_2d(&tmQKV, bar_load, sV, 2 * p.D + h * kTcHd, f * N);
      tma_load_2d(&tmDO, bar_load, sG, h * kTcHd, f * N);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------- control warp
    mbar_wait(bar_load, 0);
    if (lane == 0) TRACE(1);
    tc_fence_after_sync();
    constexpr uint64_t dK_major = smem_desc_sw128(16, 1024);      // K-major operand
    constexpr uint64_t dMN_1blk = smem_desc_sw128(8192, 1024);    // MN-major, one 64-wide block (N = 64)
    constexpr uint64_t dMN_2blk = smem_desc_sw128(kBwBlk, 1024);  // MN-major, two 64-wide blocks 16 KB apart (M = 128)
    const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), aG = smem_u32(sG), aP = smem_u32(sP),
                   aS = smem_u32(sS);
    // S^T / dP^T of iteration it+1 are issued right behind the dV/dK/dQ MMAs of iteration it (the tensor pipe runs
    // them in order; the worker warps have already drained the S^T / dP^T columns when they signalled bar_p).
    auto issue_mma1 = [&](int it) {
      const int j = it >> 1, qh = it & 1;
      const int ncols = qh == 0 ? 128 : kTcKeys - 128;
      const uint32_t idesc1 = umma_idesc(1, 0, 0, 128, ncols);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(tm + 0, smem_desc_addr(dK_major, aK + j * kBwBlk + k * 32), smem_desc_addr(dK_major, aQ + qh * kBwBlk + k * 32),
                 idesc1, k > 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(tm + 128, smem_desc_addr(dK_major, aV + j * kBwBlk + k * 32), smem_desc_addr(dK_major, aG + qh * kBwBlk + k * 32),
                 idesc1, k > 0);
      umma_commit(bar_s);
    };
    if (lane == 0) issue_mma1(0);
    __syncwarp();
    for (int it = 0; it < 4; ++it) {
      const int j = it >> 1, qh = it & 1;
      const int ncols = qh == 0 ? 128 : kTcKeys - 128;
      mbar_wait(bar_p, it & 1);
      if (it == 2) mbar_wait(bar_out, 0);   // dV_0 / dK_0 have been read out of TMEM
      tc_fence_after_sync();
      if (lane == 0) {
        TRACE(100 + it);
        constexpr uint32_t idesc_kv = umma_idesc(1, 0, 1, 128, kTcHd);
        constexpr uint32_t idesc_q = umma_idesc(1, 1, 1, 128, kTcHd);
        const int ks = ncols / 16;
        for (int k = 0; k < ks; ++k) {
          const uint32_t a_off = (k >> 2) * kBwBlk + (k & 3) * 32;
          umma_f16(tm + 256, smem_desc_addr(dK_major, aP + a_off), smem_desc_addr(dMN_1blk, aG + qh * kBwBlk + k * 2048),
                   idesc_kv, (qh > 0 || k > 0) ? 1u : 0u);
          umma_f16(tm + 320, smem_desc_addr(dK_major, aS + a_off), smem_desc_addr(dMN_1blk, aQ + qh * kBwBlk + k * 2048),
                   idesc_kv, (qh > 0 || k > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tm + 384 + 64 * qh, smem_desc_addr(dMN_2blk, aS + k * 2048), smem_desc_addr(dMN_1blk, aK + j * kBwBlk + k * 2048),
                   idesc_q, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(bar_d);
        if (it < 3) issue_mma1(it + 1);
        TRACE(110 + it);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------- 8 worker warps
    const int quarter = warp & 3, ch = warp >> 2;
    const int kr = quarter * 32 + lane;                    // key row within the tile == TMEM lane
    const uint32_t t_lane = tm + (uint32_t(quarter * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    // prologue: delta[q] = dO[q,:] . O[q,:],  lse[q] * log2(e)
    {
      const int q = threadIdx.x;
      if (q < kTcKeys) {
        float d = 0.f, l = 0.f;
        if (q < N) {
          const uint4* po = reinterpret_cast<const uint4*>(p.out + ((size_t)f * N + q) * p.D + h * kTcHd);
          const uint4* pg = reinterpret_cast<const uint4*>(p.dout + ((size_t)f * N + q) * p.D + h * kTcHd);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 a = __ldg(po + i), b = __ldg(pg + i);
            d += bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) + bf16_hi(a.y) * bf16_hi(b.y) +
                 bf16_lo(a.z) * bf16_lo(b.z) + bf16_hi(a.z) * bf16_hi(b.z) + bf16_lo(a.w) * bf16_lo(b.w) + bf16_hi(a.w) * bf16_hi(b.w);
          }
          l = p.lse[((size_t)f * p.H + h) * N + q] * 1.4426950408889634f;
        }
        sDelta[q] = d;
        sLse[q] = l;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) TRACE(2);
    }
    bf16* dbase = p.dqkv + (size_t)f * N * 3 * p.D + h * kTcHd;
    for (int it = 0; it < 4; ++it) {
      const int j = it >> 1, qh = it & 1;
      const int key = j * 128 + kr;
      const bool key_ok = key < N;
      const int nchunk = (qh == 0 || ch == 0) ? 4 : 1;     // 16-column chunks handled by this warp
      mbar_wait(bar_s, it & 1);
      if (threadIdx.x == 0) TRACE(200 + it);
      if (it > 0) mbar_wait(bar_d, (it - 1) & 1);          // previous MMAs have finished reading sP / sS
      if (threadIdx.x == 0) TRACE(210 + it);
      tc_fence_after_sync();
      for (int c = 0; c < nchunk; ++c) {
        const int c0 = ch * 64 + c * 16;                   // column within this query half
        const int q0 = qh * 128 + c0;
        uint32_t sv[16], dv[16];
        tmem_ld_32x32b_x16(t_lane + c0, sv);
        tmem_ld_32x32b_x16(t_lane + 128 + c0, dv);
        float ls[16], dl[16];
#pragma unroll
        for (int jj = 0; jj < 16; jj += 4) {   // broadcast smem reads, issued while the TMEM loads are in flight
          const float4 a4 = *reinterpret_cast<const float4*>(sLse + q0 + jj);
          const float4 b4 = *reinterpret_cast<const float4*>(sDelta + q0 + jj);
          ls[jj] = a4.x; ls[jj + 1] = a4.y; ls[jj + 2] = a4.z; ls[jj + 3] = a4.w;
          dl[jj] = b4.x; dl[jj + 1] = b4.y; dl[jj + 2] = b4.z; dl[jj + 3] = b4.w;
        }
        tmem_ld_wait();
        float pv[16], ds[16];
        const bool qmask = q0 + 16 > N;        // only the chunk that straddles N needs per-column masking
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          float pr = fast_exp2(fmaf(__uint_as_float(sv[jj]), sl2, -ls[jj]));
          if (qmask && q0 + jj >= N) pr = 0.f;
          pr = key_ok ? pr : 0.f;
          pv[jj] = pr;
          ds[jj] = pr * (__uint_as_float(dv[jj]) - dl[jj]) * p.scale;
        }
        const uint32_t off = (c0 >> 6) * kBwBlk + kr * 128;
        const int chunk = (c0 & 63) >> 3;
#pragma unroll
        for (int q2 = 0; q2 < 2; ++q2) {
          const uint32_t o2 = off + (((chunk + q2) ^ (kr & 7)) << 4);
          *reinterpret_cast<uint4*>(sP + o2) =
              make_uint4(pack_bf16x2(pv[8 * q2], pv[8 * q2 + 1]), pack_bf16x2(pv[8 * q2 + 2], pv[8 * q2 + 3]),
                         pack_bf16x2(pv[8 * q2 + 4], pv[8 * q2 + 5]), pack_bf16x2(pv[8 * q2 + 6], pv[8 * q2 + 7]));
          *reinterpret_cast<uint4*>(sS + o2) =
              make_uint4(pack_bf16x2(ds[8 * q2], ds[8 * q2 + 1]), pack_bf16x2(ds[8 * q2 + 2], ds[8 * q2 + 3]),
                         pack_bf16x2(ds[8 * q2 + 4], ds[8 * q2 + 5]), pack_bf16x2(ds[8 * q2 + 6], ds[8 * q2 + 7]));
        }
      }
      fence_proxy_async_smem();
      tc_fence_before_sync();
      if (threadIdx.x == 0) TRACE(220 + it);
      mbar_arrive(bar_p);
      if (qh == 1) {
        // dV_j / dK_j are complete once this iteration's MMAs retire: read them out (32 columns per warp)
        mbar_wait(bar_d, it & 1);
        tc_fence_after_sync();
        uint32_t a[32], b[32];
        tmem_ld_32x32b_x32(t_lane + 256 + 32 * ch, a);
        tmem_ld_32x32b_x32(t_lane + 320 + 32 * ch, b);
        tmem_ld_wait();
        if (key_ok) {
          bf16* rk = dbase + (size_t)key * 3 * p.D + p.D + 32 * ch;
          bf16* rv = rk + p.D;
#pragma unroll
          for (int q2 = 0; q2 < 4; ++q2) {
            *reinterpret_cast<uint4*>(rv + 8 * q2) = make_uint4(
                pack_bf16x2(__uint_as_float(a[8 * q2]), __uint_as_float(a[8 * q2 + 1])),
                pack_bf16x2(__uint_as_float(a[8 * q2 + 2]), __uint_as_float(a[8 * q2 + 3])),
                pack_bf16x2(__uint_as_float(a[8 * q2 + 4]), __uint_as_float(a[8 * q2 + 5])),
                pack_bf16x2(__uint_as_float(a[8 * q2 + 6]), __uint_as_float(a[8 * q2 + 7])));
            *reinterpret_cast<uint4*>(rk + 8 * q2) = make_uint4(
                pack_bf16x2(__uint_as_float(b[8 * q2]), __uint_as_float(b[8 * q2 + 1])),
                pack_bf16x2(__uint_as_float(b[8 * q2 + 2]), __uint_as_float(b[8 * q2 + 3])),
                pack_bf16x2(__uint_as_float(b[8 * q2 + 4]), __uint_as_float(b[8 * q2 + 5])),
                pack_bf16x2(__uint_as_float(b[8 * q2 + 6]), __uint_as_float(b[8 * q2 + 7])));
          }
        }
        tc_fence_before_sync();
        if (it == 1) mbar_arrive(bar_out);
        if (threadIdx.x == 0) TRACE(230 + it);
      }
    }
    // dQ: warp (quarter, ch) reads query tile `ch`, rows quarter*32 + lane, all 64 columns
    {
      const int q = ch * 128 + kr;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t a[32];
        tmem_ld_32x32b_x32(t_lane + 384 + 64 * ch + c0, a);
        tmem_ld_wait();
        if (q < N) {
          bf16* rq = dbase + (size_t)q * 3 * p.D + c0;
#pragma unroll
          for (int q2 = 0; q2 < 4; ++q2)
            *reinterpret_cast<uint4*>(rq + 8 * q2) = make_uint4(
                pack_bf16x2(__uint_as_float(a[8 * q2]), __uint_as_float(a[8 * q2 + 1])),
                pack_bf16x2(__uint_as_float(a[8 * q2 + 2]), __uint_as_float(a[8 * q2 + 3])),
                pack_bf16x2(__uint_as_float(a[8 * q2 + 4]), __uint_as_float(a[8 * q2 + 5])),
                pack_bf16x2(__uint_as_float(a[8 * q2 + 6]), __uint_as_float(a[8 * q2 + 7])));
        }
      }
    }
  }

  if (threadIdx.x == 0) TRACE(999);
  tc_fence_before_sync();
  __syncthreads();
  TRACE_DUMP
  if (warp == 8) tmem_dealloc(tm, 512);
}


// ------------------------------------------------------------------------------------------------ backward, v2
// Persistent and software-pipelined. The v1 timeline (TRACE, one CTA = one (frame, head)): 4.7 us prologue (tile loads,
// delta), then per 128-query iteration  threads 1.0-1.3 us -> MMA 1.6-2.9 us -> threads ...  strictly in series
// (S^T/dP^T single-buffered in TMEM), 2.2 us read-out: 20.6 us per item, 7 rounds of CTAs per layer.
// v2:
//   * one CTA per SM walks items (frame, head) = blockIdx.x, +gridDim.x, ...; TMEM / barriers are set up once;
//   * queries are processed in chunks of 64 columns and S^T / dP^T are double-buffered in TMEM (2 x 128 columns), so the
//     tensor core computes chunk g+1's scores while the threads turn chunk g into P^T / dS^T, and chunk g's
//     dV / dK / dQ MMAs run under chunk g+1's thread work;
//   * P^T ping-pongs between 2 smem blocks, dS^T between 4 (the dQ MMA of a 128-query tile reads two of them);
//   * delta = rowsum(dO o O) and lse*log2(e) of the NEXT item are prepared by two otherwise idle warps;
//   * the read-out of dK / dV / dQ overlaps the TMA loads of the next item's tiles.
// TMEM columns: S^T[b] 128b..+63 | dP^T[b] 128b+64..+127 (b = 0,1) | dV 256 | dK 320 | dQ tile0 384 | dQ tile1 448.
// Measured (TRACE, B200): 113 us/layer vs 144 us for v1. What bounds it now: (1) the small MMAs are operand-fetch
// bound, not math bound - a K-major SW128 operand serves one 16-element k-step as 32 bytes out of every 128-byte row,
// so a 128 x 64 x 16 MMA (32 math cycles) spends ~130-190 cycles pulling 128 + 64 smem lines (interleaving independent
// accumulator chains did not help: 120 us); (2) the worker warps are bound by the TMEM read port (64 B/clk/SM):
// S^T + dP^T are 64 KB per chunk = ~0.55 us. Next step: 32-byte-slab (SWIZZLE_32B) operand tiles and P^T / dS^T as
// TMEM A operands, which cut the operand fetch per k-step to a quarter.
constexpr int kB2Workers = 8;                      // worker warps (2 per TMEM lane quarter)
constexpr int kB2Threads = 32 * (kB2Workers + 3);  // + control warp + 2 delta warps
constexpr int kB2Smem = 4 * kBwTile + 6 * kBwBlk + 4 * kTcKeys * 4 + 256 + 1024;

__global__ void __launch_bounds__(kB2Threads, 1)
attn_tc_bwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                    const AttnTcBwdParams p, int items) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kBwTile;
  uint8_t* sQ = sV + kBwTile;
  uint8_t* sG = sQ + kBwTile;            // dO
  uint8_t* sP = sG + kBwTile;            // [2][128 keys x 64 queries]  P^T
  uint8_t* sS = sP + 2 * kBwBlk;         // [4][128 keys x 64 queries]  dS^T
  float* sNLse = reinterpret_cast<float*>(sS + 4 * kBwBlk);   // [2][208]  -lse * log2(e)   (-inf for q >= N)
  float* sNDel = sNLse + 2 * kTcKeys;                          // [2][208]  -delta * scale
  uint64_t* bars = reinterpret_cast<uint64_t*>(sNDel + 2 * kTcKeys);
  uint64_t* bar_tiles = bars;        // TMA: the item's four tiles landed
  uint64_t* bar_s = bars + 1;        // [2] S^T/dP^T buffer b computed            (MMA -> workers)
  uint64_t* bar_p = bars + 3;        // [2] buffer b drained, P^T/dS^T in smem     (256 workers -> control)
  uint64_t* bar_m2 = bars + 5;       // [2] dV/dK/dQ MMAs of a chunk retired       (MMA -> workers, control)
  uint64_t* bar_vkfree = bars + 7;   // dV/dK read out of TMEM                     (256 workers -> control)
  uint64_t* bar_dqfree = bars + 8;   // dQ read out of TMEM                        (256 workers -> control)
  uint64_t* bar_dfull = bars + 9;    // [2] delta/lse buffer filled                (64 delta threads -> workers)
  uint64_t* bar_dfree = bars + 11;   // [2] delta/lse buffer no longer needed      (256 workers -> delta warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N;
  TRACE_DECL
  TRACE_INIT

  if (warp == kB2Workers) {
    if (lane == 0) {
      TRACE(0);
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmDO);
      mbar_init(bar_tiles, 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bar_s[b], 1);
        mbar_init(&bar_p[b], 32 * kB2Workers);
        mbar_init(&bar_m2[b], 1);
        mbar_init(&bar_dfull[b], 64);
        mbar_init(&bar_dfree[b], 32 * kB2Workers);
      }
      mbar_init(bar_vkfree, 32 * kB2Workers);
      mbar_init(bar_dqfree, 32 * kB2Workers);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = *tmem_slot;
  pdl_wait();   // prologue above overlapped the previous kernel's tail
  pdl_trigger();

  if (warp == kB2Workers) {
    // ------------------------------------------------------------- control warp: one lane issues TMA + MMA
    if (lane == 0) {
      constexpr uint64_t dK_major = smem_desc_sw128(16, 1024);      // K-major operand
      constexpr uint64_t dMN_1blk = smem_desc_sw128(8192, 1024);    // MN-major, one 64-wide block (N = 64)
      constexpr uint64_t dMN_2blk = smem_desc_sw128(kBwBlk, 1024);  // MN-major, two 64-wide blocks 16 KB apart (M = 128)
      constexpr uint32_t idesc_kv = umma_idesc(1, 0, 1, 128, kTcHd);
      constexpr uint32_t idesc_q = umma_idesc(1, 1, 1, 128, kTcHd);
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), aG = smem_u32(sG), aP = smem_u32(sP),
                     aS = smem_u32(sS);
      auto load_tiles = [&](int item) {
        const int f = item / p.H, h = item % p.H;
        mbar_arrive_expect_tx(bar_tiles, 4 * kBwTile);
        tma_load_2d(&tmQKV, bar_tiles, sQ, h * kTcHd, f * N);
        tma_load_2d(&tmQKV, bar_tiles, sK, p.D + h * kTcHd, f * N);
        tma_load_2d(&tmQKV, bar_tiles, sV, 2 * p.D + h * kTcHd, f * N);
        tma_load_2d(&tmDO, bar_tiles, sG, h * kTcHd, f * N);
      };
      // S^T[b] = K_j Q_c^T, dP^T[b] = V_j dO_c^T for chunk index lc = 4 j + c of the current item
      auto issue_mma1 = [&](int lc, int b) {
        const int j = lc >> 2, c = lc & 3;
        const uint32_t idesc1 = umma_idesc(1, 0, 0, 128, c == 3 ? 16 : 64);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tm + 128 * b, smem_desc_addr(dK_major, aK + j * kBwBlk + k * 32),
                   smem_desc_addr(dK_major, aQ + c * 8192 + k * 32), idesc1, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tm + 128 * b + 64, smem_desc_addr(dK_major, aV + j * kBwBlk + k * 32),
                   smem_desc_addr(dK_major, aG + c * 8192 + k * 32), idesc1, k > 0);
        umma_commit(&bar_s[b]);
      };
      uint32_t g = 0;      // chunks processed by this CTA so far (8 per item)
      uint32_t n = 0;      // items processed by this CTA so far
      if ((int)blockIdx.x < items) load_tiles(blockIdx.x);
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
        mbar_wait(bar_tiles, n & 1);
        if (n < 2) TRACE(1);
        tc_fence_after_sync();
        issue_mma1(0, g & 1);   // (the buffer was drained: bar_p of chunk g-1 was observed before that chunk's MMA2)
        for (int lc = 0; lc < 8; ++lc, ++g) {
          const int j = lc >> 2, c = lc & 3, b = g & 1;
          const uint32_t ph = (g >> 1) & 1;
          if (lc < 7) issue_mma1(lc + 1, b ^ 1);     // next chunk's scores run under this chunk's thread work
          if (n < 2) TRACE(90 + lc);
          mbar_wait(&bar_p[b], ph);                   // P^T / dS^T of this chunk are in smem, S^T/dP^T[b] drained
          if (n < 2) TRACE(100 + lc);
          if (c == 0) {
            const uint32_t v = 2 * n + j;             // dV/dK of the previous key tile must have left TMEM
            if (v > 0) mbar_wait(bar_vkfree, (v - 1) & 1);
          }
          if (lc == 1 && n > 0) mbar_wait(bar_dqfree, (n - 1) & 1);   // previous item's dQ was read out
          tc_fence_after_sync();
          const int ks = c == 3 ? 1 : 4;              // 16-query k-steps in this chunk
          const int sb = (((lc >> 1) & 1) << 1) | (c & 1);   // dS^T block: tile parity x chunk parity
          for (int k = 0; k < ks; ++k) {
            umma_f16(tm + 256, smem_desc_addr(dK_major, aP + b * kBwBlk + k * 32),
                     smem_desc_addr(dMN_1blk, aG + c * 8192 + k * 2048), idesc_kv, (c > 0 || k > 0) ? 1u : 0u);
            umma_f16(tm + 320, smem_desc_addr(dK_major, aS + sb * kBwBlk + k * 32),
                     smem_desc_addr(dMN_1blk, aQ + c * 8192 + k * 2048), idesc_kv, (c > 0 || k > 0) ? 1u : 0u);
          }
          if (c & 1) {   // both halves of query tile t = c/2 are in smem: dQ_t += dS K_j
            const int t = c >> 1;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_f16(tm + 384 + 64 * t, smem_desc_addr(dMN_2blk, aS + (sb & 2) * kBwBlk + k * 2048),
                       smem_desc_addr(dMN_1blk, aK + j * kBwBlk + k * 2048), idesc_q, (j > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&bar_m2[b]);
          if (n < 2) TRACE(110 + lc);
          if (lc == 7) {
            const int next = item + gridDim.x;
            if (next < items) {
              mbar_wait(&bar_m2[b], ph);   // every MMA of this item retired: the four tiles may be overwritten
              if (n < 2) TRACE(120);
              load_tiles(next);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp > kB2Workers) {
    // ------------------------------------------------------------- delta / lse warps (one item ahead)
    const int t = threadIdx.x - 32 * (kB2Workers + 1);   // 0..63
    uint32_t n = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
      const int f = item / p.H, h = item % p.H, ip = n & 1;
      if (n >= 2) mbar_wait(&bar_dfree[ip], ((n >> 1) - 1) & 1);
      for (int q = t; q < kTcKeys; q += 64) {
        float d = 0.f, l = -INFINITY;
        if (q < N) {
          const uint4* po = reinterpret_cast<const uint4*>(p.out + ((size_t)f * N + q) * p.D + h * kTcHd);
          const uint4* pg = reinterpret_cast<const uint4*>(p.dout + ((size_t)f * N + q) * p.D + h * kTcHd);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 a = __ldg(po + i), b = __ldg(pg + i);
            d += bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) + bf16_hi(a.y) * bf16_hi(b.y) +
                 bf16_lo(a.z) * bf16_lo(b.z) + bf16_hi(a.z) * bf16_hi(b.z) + bf16_lo(a.w) * bf16_lo(b.w) + bf16_hi(a.w) * bf16_hi(b.w);
          }
          l = -p.lse[((size_t)f * p.H + h) * N + q] * 1.4426950408889634f;
        }
        sNDel[ip * kTcKeys + q] = -d * p.scale;
        sNLse[ip * kTcKeys + q] = l;      // -inf for q >= N: exp2(s*c - inf) = 0, the column drops out
      }
      mbar_arrive(&bar_dfull[ip]);        // (release semantics: the smem writes above are visible to the waiters)
    }
  } else {
    // ------------------------------------------------------------- 8 worker warps
    const int quarter = warp & 3, ch = warp >> 2;
    const int kr = quarter * 32 + lane;                    // key row within the tile == TMEM lane
    const uint32_t t_lane = tm + (uint32_t(quarter * 32) << 16);
    const float2 sl2 = f2(p.scale * 1.4426950408889634f);
    const float2 sc2 = f2(p.scale);
    uint32_t g = 0, n = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
      const int f = item / p.H, h = item % p.H, ip = n & 1;
      bf16* dbase = p.dqkv + (size_t)f * N * 3 * p.D + h * kTcHd;
      const float* nlse = sNLse + ip * kTcKeys;
      const float* ndel = sNDel + ip * kTcKeys;
      mbar_wait(&bar_dfull[ip], (n >> 1) & 1);
      for (int lc = 0; lc < 8; ++lc, ++g) {
        const int j = lc >> 2, c = lc & 3, b = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        const int key = j * 128 + kr;
        const bool key_ok = key < N;
        mbar_wait(&bar_s[b], ph);
        if (threadIdx.x == 0 && n < 2) TRACE(200 + lc);
        if (g >= 2) mbar_wait(&bar_m2[b], ((g - 2) >> 1) & 1);   // chunk g-2's MMAs no longer read sP[b] (nor older dS^T blocks)
        if (threadIdx.x == 0 && n < 2) TRACE(210 + lc);
        tc_fence_after_sync();
        const int sb = (((lc >> 1) & 1) << 1) | (c & 1);
        const int nh = c == 3 ? 1 : 2;                          // 16-column halves of this warp's 32 columns (chunk 3: 16 columns, ch 0 only)
        if (c < 3 || ch == 0) {
          for (int hh = 0; hh < nh; ++hh) {
            const int c0 = ch * 32 + hh * 16;                    // column within the chunk
            const int q0 = c * 64 + c0;
            uint32_t sv[16], dv[16];
            tmem_ld_32x32b_x16(t_lane + 128 * b + c0, sv);
            tmem_ld_32x32b_x16(t_lane + 128 * b + 64 + c0, dv);
            float2 ls[8], dl[8];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {   // broadcast smem reads, issued while the TMEM loads are in flight
              const float4 a4 = *reinterpret_cast<const float4*>(nlse + q0 + 4 * jj);
              const float4 b4 = *reinterpret_cast<const float4*>(ndel + q0 + 4 * jj);
              ls[2 * jj] = make_float2(a4.x, a4.y); ls[2 * jj + 1] = make_float2(a4.z, a4.w);
              dl[2 * jj] = make_float2(b4.x, b4.y); dl[2 * jj + 1] = make_float2(b4.z, b4.w);
            }
            tmem_ld_wait();
            uint32_t pp[8], dd[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const float2 a = __ffma2_rn(make_float2(__uint_as_float(sv[2 * jj]), __uint_as_float(sv[2 * jj + 1])), sl2, ls[jj]);
              const float2 pr = make_float2(fast_exp2(a.x), fast_exp2(a.y));
              const float2 tt = __ffma2_rn(make_float2(__uint_as_float(dv[2 * jj]), __uint_as_float(dv[2 * jj + 1])), sc2, dl[jj]);
              const float2 ds = __fmul2_rn(pr, tt);
              pp[jj] = key_ok ? pack_bf16x2(pr.x, pr.y) : 0u;
              dd[jj] = key_ok ? pack_bf16x2(ds.x, ds.y) : 0u;
            }
            const uint32_t row = kr * 128;
            const int chunk = c0 >> 3;
#pragma unroll
            for (int q2 = 0; q2 < 2; ++q2) {
              const uint32_t o2 = row + (((chunk + q2) ^ (kr & 7)) << 4);
              *reinterpret_cast<uint4*>(sP + b * kBwBlk + o2) = make_uint4(pp[4 * q2], pp[4 * q2 + 1], pp[4 * q2 + 2], pp[4 * q2 + 3]);
              *reinterpret_cast<uint4*>(sS + sb * kBwBlk + o2) = make_uint4(dd[4 * q2], dd[4 * q2 + 1], dd[4 * q2 + 2], dd[4 * q2 + 3]);
            }
          }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(&bar_p[b]);
        if (threadIdx.x == 0 && n < 2) TRACE(220 + lc);
        if (c == 3) {
          // dV_j / dK_j are complete once this chunk's MMAs retire: read them out (32 columns per warp)
          mbar_wait(&bar_m2[b], ph);
          if (threadIdx.x == 0 && n < 2) TRACE(230 + lc);
          tc_fence_after_sync();
          uint32_t a[32], bb[32];
          tmem_ld_32x32b_x32(t_lane + 256 + 32 * ch, a);
          tmem_ld_32x32b_x32(t_lane + 320 + 32 * ch, bb);
          tmem_ld_wait();
          tc_fence_before_sync();
          mbar_arrive(bar_vkfree);
          if (key_ok) {
            bf16* rk = dbase + (size_t)key * 3 * p.D + p.D + 32 * ch;
            bf16* rv = rk + p.D;
#pragma unroll
            for (int q2 = 0; q2 < 4; ++q2) {
              *reinterpret_cast<uint4*>(rv + 8 * q2) = make_uint4(
                  pack_bf16x2(__uint_as_float(a[8 * q2]), __uint_as_float(a[8 * q2 + 1])),
                  pack_bf16x2(__uint_as_float(a[8 * q2 + 2]), __uint_as_float(a[8 * q2 + 3])),
                  pack_bf16x2(__uint_as_float(a[8 * q2 + 4]), __uint_as_float(a[8 * q2 + 5])),
                  pack_bf16x2(__uint_as_float(a[8 * q2 + 6]), __uint_as_float(a[8 * q2 + 7])));
              *reinterpret_cast<uint4*>(rk + 8 * q2) = make_uint4(
                  pack_bf16x2(__uint_as_float(bb[8 * q2]), __uint_as_float(bb[8 * q2 + 1])),
                  pack_bf16x2(__uint_as_float(bb[8 * q2 + 2]), __uint_as_float(bb[8 * q2 + 3])),
                  pack_bf16x2(__uint_as_float(bb[8 * q2 + 4]), __uint_as_float(bb[8 * q2 + 5])),
                  pack_bf16x2(__uint_as_float(bb[8 * q2 + 6]), __uint_as_float(bb[8 * q2 + 7])));
            }
          }
          if (j == 1) {
            // dQ: warp (quarter, ch) reads query tile `ch`, rows quarter*32 + lane, all 64 columns (the last MMA2 retired above)
            const int q = ch * 128 + kr;
            uint32_t x0[32], x1[32];
            tmem_ld_32x32b_x32(t_lane + 384 + 64 * ch, x0);
            tmem_ld_32x32b_x32(t_lane + 384 + 64 * ch + 32, x1);
            tmem_ld_wait();
            tc_fence_before_sync();
            mbar_arrive(bar_dqfree);
            mbar_arrive(&bar_dfree[ip]);
            if (q < N) {
              bf16* rq = dbase + (size_t)q * 3 * p.D;
#pragma unroll
              for (int q2 = 0; q2 < 4; ++q2) {
                *reinterpret_cast<uint4*>(rq + 8 * q2) = make_uint4(
                    pack_bf16x2(__uint_as_float(x0[8 * q2]), __uint_as_float(x0[8 * q2 + 1])),
                    pack_bf16x2(__uint_as_float(x0[8 * q2 + 2]), __uint_as_float(x0[8 * q2 + 3])),
                    pack_bf16x2(__uint_as_float(x0[8 * q2 + 4]), __uint_as_float(x0[8 * q2 + 5])),
                    pack_bf16x2(__uint_as_float(x0[8 * q2 + 6]), __uint_as_float(x0[8 * q2 + 7])));
                *reinterpret_cast<uint4*>(rq + 32 + 8 * q2) = make_uint4(
                    pack_bf16x2(__uint_as_float(x1[8 * q2]), __uint_as_float(x1[8 * q2 + 1])),
                    pack_bf16x2(__uint_as_float(x1[8 * q2 + 2]), __uint_as_float(x1[8 * q2 + 3])),
                    pack_bf16x2(__uint_as_float(x1[8 * q2 + 4]), __uint_as_float(x1[8 * q2 + 5])),
                    pack_bf16x2(__uint_as_float(x1[8 * q2 + 6]), __uint_as_float(x1[8 * q2 + 7])));
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  TRACE_DUMP
  if (warp == kB2Workers) tmem_dealloc(tm, 512);
}

}  // namespace avt

using namespace avt;

extern "C" int avt_attention_tc_fwd(const void* qkv, void* out, float* lse, int F, int H, int N, float scale, void* stream) {
  AVT_REQUIRE(qkv && out, "null pointer");
  AVT_REQUIRE(F > 0 && H > 0 && N > 0 && N <= kTcKeys, "tokens per frame must be in [1, 208]");
  const int D = H * kTcHd;
  CUtensorMap tmQ, tmKV;
  const uint64_t rows = (uint64_t)F * N;
  if (int rc = make_tmap_bf16_2d(&tmQ, qkv, 3ull * D, rows, 3ull * D, 64, 128, 128)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmKV, qkv, 3ull * D, rows, 3ull * D, 64, kTcKeys, 128)) return rc;
  static bool configured = false;
  if (!configured) {
    AVT_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
    configured = true;
  }
  AttnTcParams p;
  p.out = reinterpret_cast<bf16*>(out); p.lse = lse; p.N = N; p.H = H; p.D = D; p.F = F; p.scale = scale;
  const int grid = F * H < num_sms() ? F * H : num_sms();
  launch_kernel(attn_tc_fwd_kernel, dim3(grid), dim3(kTcThreads), kTcSmem, reinterpret_cast<cudaStream_t>(stream), tmQ, tmKV, p);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

extern "C" int avt_attention_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int F,
                                    int H, int N, float scale, void* stream) {
  AVT_REQUIRE(qkv && out && dout && lse && dqkv, "null pointer");
  AVT_REQUIRE(F > 0 && H > 0 && N > 0 && N <= kTcKeys, "tokens per frame must be in [1, 208]");
  const int D = H * kTcHd;
  CUtensorMap tmQKV, tmDO;
  const uint64_t rows = (uint64_t)F * N;
  if (int rc = make_tmap_bf16_2d(&tmQKV, qkv, 3ull * D, rows, 3ull * D, 64, kTcKeys, 128)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmDO, dout, (uint64_t)D, rows, (uint64_t)D, 64, kTcKeys, 128)) return rc;
  static bool configured = false;
  if (!configured) {
    AVT_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwSmem));
    configured = true;
  }
  AttnTcBwdParams p;
  p.out = reinterpret_cast<const bf16*>(out); p.dout = reinterpret_cast<const bf16*>(dout); p.lse = lse;
  p.dqkv = reinterpret_cast<bf16*>(dqkv); p.N = N; p.H = H; p.D = D; p.scale = scale;
  static const bool use_v1 = getenv("AVT_ATTN_BWD_V1") != nullptr;   // the one-CTA-per-item kernel, kept for A/B runs
  if (!use_v1) {
    static bool configured2 = false;
    if (!configured2) {
      AVT_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kB2Smem));
      configured2 = true;
    }
    const int items = F * H;
    const int grid = items < num_sms() ? items : num_sms();
    launch_kernel(attn_tc_bwd2_kernel, dim3(grid), dim3(kB2Threads), kB2Smem, reinterpret_cast<cudaStream_t>(stream), tmQKV, tmDO, p, items);
    AVT_CUDA_OK(cudaGetLastError());
    return AVT_OK;
  }
  launch_kernel(attn_tc_bwd_kernel, dim3(F * H), dim3(kTcThreads), kBwSmem, reinterpret_cast<cudaStream_t>(stream), tmQKV, tmDO, p);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}
