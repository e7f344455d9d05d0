// ViT spatial attention BACKWARD on tcgen05 tensor cores (N <= 208 tokens, head_dim 64); the forward kernel lives in
// attention_fwd.cu. Layout: qkv / dqkv bf16 [F*N, 3*D], column = s*D + h*64 + d (timm Attention.qkv packing); out, dout bf16
// [F*N, D]; lse fp32 [F*H, N] (log-sum-exp of the scaled scores, saved by the forward). Replaces the autograd backward of timm
// Attention.forward (q@k^T*scale -> softmax -> attn@v; models/video_classification.py:255-256 runs timm's ViT per frame)
// without ever materialising the 80*12*197^2 score / probability tensors.
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"

namespace avt {

int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                      uint32_t box_outer, int swizzle_bytes);
int make_tmap_bf16_3d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer1, uint64_t outer2, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer1, int swizzle_bytes);

constexpr int kTcHd = 64;
constexpr int kTcKeys = 208;            // MMA N extent / PV contraction length (multiple of 16)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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

#ifdef AVT_ATTN_TRACE   // per-phase timeline of CTA 0: one shared-memory slot per (item < 2, event id): a fire-and-forget clock store
#define TRACE_DECL __shared__ long long trace_t[512];
#define TRACE_INIT for (int i_ = threadIdx.x; i_ < 512; i_ += blockDim.x) trace_t[i_] = 0; __syncthreads();
#define TRACE(id) do { if (blockIdx.x == 0 && n < 2) trace_t[(n & 1) * 256 + (id)] = clock64(); } while (0)
#define TRACE_DUMP if (blockIdx.x == 0 && threadIdx.x == 0) { long long t0_ = 0; for (int i_ = 0; i_ < 512; ++i_) if (trace_t[i_] && (!t0_ || trace_t[i_] < t0_)) t0_ = trace_t[i_]; for (int i_ = 0; i_ < 512; ++i_) if (trace_t[i_]) printf("btrace %d %d %lld\n", i_ >> 8, i_ & 255, trace_t[i_] - t0_); }
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

// ------------------------------------------------------------------------------------------------ the kernel
// Persistent and software-pipelined (v1, one CTA per (frame, head) with threads -> MMA -> threads strictly in series, took
// 144 us per layer at 80 x 12 x 197):
//   * one CTA per SM walks items (frame, head) = blockIdx.x, +gridDim.x, ...; TMEM / barriers are set up once;
//   * queries are processed in chunks of 64 columns and S^T / dP^T are double-buffered in TMEM (2 x 128 columns): the
//     tensor core computes the scores of chunk lc+2 while the workers turn chunk lc+1 into P^T / dS^T, and chunk lc's
//     dV / dK / dQ MMAs are issued behind them (the pipe runs in order: the workers never wait for scores);
//   * P^T ping-pongs between 2 smem blocks, dS^T between 4 (the dQ MMA of a 128-query tile reads two of them);
//   * the issuing thread's chunk loop is fully unrolled: a chunk's buffers depend only on its index within the item, so
//     every operand descriptor is a base built once + an immediate (from run-time indices the ISSUE of a chunk's 20 MMAs
//     took 1800-2300 cycles, longer than the thread work of the chunk);
//   * delta = rowsum(dO o O) and lse*log2(e) of the NEXT item are prepared by two otherwise idle warps;
//   * dK / dV / dQ leave through swizzled staging tiles (P^T / dS^T blocks that are dead at that point) + TMA stores
//     whose 3-D map drops rows >= N; the tiles arrive through 3-D maps too (rows >= N zero-filled).
// TMEM columns: S^T[b] 128b..+63 | dP^T[b] 128b+64..+127 (b = 0,1) | dV 256 | dK 320 | dQ tile0 384 | dQ tile1 448.
// Measured (B200, phase trace AVT_ATTN_TRACE + ncu, profiles/r02_attn_ncu.txt): 100 us per layer, ~20 000 cycles per item.
// What bounds it: the 8 worker warps execute ~250 dependent instructions per chunk at ~0.12 IPC each (removing the
// exponentials, the TMEM loads, the proxy fences or the 256-thread barrier arrivals one at a time changes nothing; the
// broadcast ld.shared of lse / delta is worth 6 us), and the 148 small MMAs of an item take ~75 cycles each in situ
// (tools/ubench_mma.cu: 48 back to back, whatever N <= 64). Tried and dropped: k-steps spread over lanes with per-lane
// commits (130 us), 16 worker warps (109 us), one batched 32-column TMEM load per array (no change).
constexpr int kB2Workers = 8;                      // worker warps (2 per TMEM lane quarter)
constexpr int kB2Threads = 32 * (kB2Workers + 3);  // + control warp + 2 delta warps
constexpr int kB2Smem = 4 * kBwTile + 6 * kBwBlk + 4 * kTcKeys * 4 + 256 + 1024;

__global__ void __launch_bounds__(kB2Threads, 1)
attn_tc_bwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                    const __grid_constant__ CUtensorMap tmDQ, const AttnTcBwdParams p, int items) {
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
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmDO);
      tma_prefetch_desc(&tmDQ);
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
      // The trace of the rolled version (descriptors built from run-time chunk indices) showed the ISSUE of a chunk's 20 MMAs
      // taking 1800-2300 cycles (85-145 per instruction, tools/ubench_mma.cu: 48 when back to back) - longer than the thread
      // work of the chunk, with the worker warps idle 40 % of the time. The chunk loop is therefore fully unrolled: a chunk's
      // buffers depend only on its index within the item (8 chunks per item, so b = lc & 1 and the barrier parities are
      // compile-time too) and every descriptor is a base built once + an immediate.
      constexpr uint64_t dK_major = smem_desc_sw128(16, 1024);      // K-major operand
      constexpr uint64_t dMN_1blk = smem_desc_sw128(8192, 1024);    // MN-major, one 64-wide block (N = 64)
      constexpr uint64_t dMN_2blk = smem_desc_sw128(kBwBlk, 1024);  // MN-major, two 64-wide blocks 16 KB apart (M = 128)
      constexpr uint32_t idesc_kv = umma_idesc(1, 0, 1, 128, kTcHd);
      constexpr uint32_t idesc_q = umma_idesc(1, 1, 1, 128, kTcHd);
      // base descriptors (start address of the buffer folded in); a byte offset is added as (offset >> 4) - shared-memory
      // addresses stay below 2^18, so the 14-bit address field never carries
      const uint64_t kK = smem_desc_addr(dK_major, smem_u32(sK)), kV = smem_desc_addr(dK_major, smem_u32(sV)),
                     kQ = smem_desc_addr(dK_major, smem_u32(sQ)), kG = smem_desc_addr(dK_major, smem_u32(sG)),
                     kP = smem_desc_addr(dK_major, smem_u32(sP)), kS = smem_desc_addr(dK_major, smem_u32(sS)),
                     mG = smem_desc_addr(dMN_1blk, smem_u32(sG)), mQ = smem_desc_addr(dMN_1blk, smem_u32(sQ)),
                     mK = smem_desc_addr(dMN_1blk, smem_u32(sK)), mS = smem_desc_addr(dMN_2blk, smem_u32(sS));
      auto off = [](uint64_t d, uint32_t bytes) { return d + uint64_t(bytes >> 4); };
      auto load_tiles = [&](int item) {
        const int f = item / p.H, h = item % p.H;
        mbar_arrive_expect_tx(bar_tiles, 4 * kBwTile);
        tma_load_2d(&tmQKV, bar_tiles, sQ, h * kTcHd, f * N);
        tma_load_2d(&tmQKV, bar_tiles, sK, p.D + h * kTcHd, f * N);
        tma_load_2d(&tmQKV, bar_tiles, sV, 2 * p.D + h * kTcHd, f * N);
        tma_load_2d(&tmDO, bar_tiles, sG, h * kTcHd, f * N);
      };
      // S^T[b] = K_j Q_c^T, dP^T[b] = V_j dO_c^T for chunk index lc = 4 j + c of the current item (b = lc & 1)
      auto issue_mma1 = [&](const int lc) {
        const int j = lc >> 2, c = lc & 3, b = lc & 1;
        const uint32_t idesc1 = c == 3 ? umma_idesc(1, 0, 0, 128, 16) : umma_idesc(1, 0, 0, 128, 64);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tm + 128 * b, off(kK, j * kBwBlk + k * 32), off(kQ, c * 8192 + k * 32), idesc1, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tm + 128 * b + 64, off(kV, j * kBwBlk + k * 32), off(kG, c * 8192 + k * 32), idesc1, k > 0);
        umma_commit(&bar_s[b]);
      };
      uint32_t n = 0;      // items processed by this CTA so far
      if ((int)blockIdx.x < items) load_tiles(blockIdx.x);
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
        mbar_wait(bar_tiles, n & 1);
        TRACE(1);
        tc_fence_after_sync();
        // both score buffers were drained (bar_p of the previous item's last two chunks was observed): chunks 0 and 1
        issue_mma1(0);
        issue_mma1(1);
#pragma unroll
        for (int lc = 0; lc < 8; ++lc) {
          const int j = lc >> 2, c = lc & 3, b = lc & 1;
          const uint32_t ph = (lc >> 1) & 1;          // chunk g = 8 n + lc uses phase (g >> 1) & 1 of its buffer's barriers
          TRACE(130 + lc);
          mbar_wait(&bar_p[b], ph);                   // P^T / dS^T of this chunk are in smem, S^T/dP^T[b] drained
          TRACE(100 + lc);
          tc_fence_after_sync();
          // The scores of chunk lc + 2 go into the buffer just drained and are issued BEFORE this chunk's dV / dK / dQ MMAs:
          // the tensor pipe runs in order, and the workers (busy with chunk lc + 1 meanwhile) must never wait for scores.
          if (lc < 6) issue_mma1(lc + 2);
          TRACE(90 + lc);
          if (c == 0) {
            const uint32_t v = 2 * n + j;             // dV/dK of the previous key tile must have left TMEM
            if (v > 0) mbar_wait(bar_vkfree, (v - 1) & 1);
          }
          if (lc == 1 && n > 0) mbar_wait(bar_dqfree, (n - 1) & 1);   // previous item's dQ was read out
          tc_fence_after_sync();
          const int ks = c == 3 ? 1 : 4;              // 16-query k-steps in this chunk
          const int sb = (((lc >> 1) & 1) << 1) | (c & 1);   // dS^T block: tile parity x chunk parity
#pragma unroll
          for (int k = 0; k < ks; ++k) {
            umma_f16(tm + 256, off(kP, b * kBwBlk + k * 32), off(mG, c * 8192 + k * 2048), idesc_kv, (c > 0 || k > 0) ? 1u : 0u);
            umma_f16(tm + 320, off(kS, sb * kBwBlk + k * 32), off(mQ, c * 8192 + k * 2048), idesc_kv, (c > 0 || k > 0) ? 1u : 0u);
          }
          if (c & 1) {   // both halves of query tile t = c/2 are in smem: dQ_t += dS K_j
            const int t = c >> 1;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_f16(tm + 384 + 64 * t, off(mS, (sb & 2) * kBwBlk + k * 2048), off(mK, j * kBwBlk + k * 2048), idesc_q,
                       (j > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&bar_m2[b]);
          TRACE(110 + lc);
          if (lc == 7) {
            const int next = item + gridDim.x;
            if (next < items) {
              mbar_wait(&bar_m2[b], ph);   // every MMA of this item retired: the four tiles may be overwritten
              TRACE(120);
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
    bool store_pending = false;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
      const int f = item / p.H, h = item % p.H, ip = n & 1;
      const float* nlse = sNLse + ip * kTcKeys;
      const float* ndel = sNDel + ip * kTcKeys;
      mbar_wait(&bar_dfull[ip], (n >> 1) & 1);
      for (int lc = 0; lc < 8; ++lc, ++g) {
        const int j = lc >> 2, c = lc & 3, b = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        const int key = j * 128 + kr;
        const bool key_ok = key < N;
        mbar_wait(&bar_s[b], ph);
        if (threadIdx.x == 0) TRACE(200 + lc);
        if (g >= 2) mbar_wait(&bar_m2[b], ((g - 2) >> 1) & 1);   // chunk g-2's MMAs no longer read sP[b] (nor older dS^T blocks)
        if (threadIdx.x == 0) TRACE(210 + lc);
        tc_fence_after_sync();
        if (store_pending) {   // the TMA stores of the last read-out have finished READING their staging blocks
          if (threadIdx.x == 0) tma_store_wait_read<0>();
          named_bar_sync(2, 32 * kB2Workers);
          store_pending = false;
        }
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
        if (threadIdx.x == 0) TRACE(220 + lc);
        if (c == 3) {
          // dV_j / dK_j are complete once this chunk's MMAs retire: read them out (32 columns per warp). They leave through
          // swizzled staging tiles + TMA stores (row-per-thread global stores cost one L1 wavefront per row per instruction:
          // the read-outs of an item took 8 300 of its 29 000 cycles with the whole pipeline drained). Staging = P^T / dS^T
          // blocks nobody reads any more (every MMA of this key tile retired) and nobody writes before the chunk after next:
          // dV -> dS^T block 2, dK -> block 3, dQ tile 0 -> dS^T block 1, dQ tile 1 -> P^T block 1. The 3-D tensor map
          // [frame][token][column] drops rows >= N.
          mbar_wait(&bar_m2[b], ph);
          if (threadIdx.x == 0) TRACE(230 + lc);
          tc_fence_after_sync();
          {
            uint32_t a[32], bb[32];
            tmem_ld_32x32b_x32(t_lane + 256 + 32 * ch, a);
            tmem_ld_32x32b_x32(t_lane + 320 + 32 * ch, bb);
            tmem_ld_wait();
            tc_fence_before_sync();
            mbar_arrive(bar_vkfree);
            uint8_t* stv = sS + 2 * kBwBlk + kr * 128;
            uint8_t* stk = sS + 3 * kBwBlk + kr * 128;
#pragma unroll
            for (int q2 = 0; q2 < 4; ++q2) {
              const uint32_t o2 = ((4 * ch + q2) ^ (kr & 7)) << 4;
              *reinterpret_cast<uint4*>(stv + o2) = make_uint4(
                  pack_bf16x2(__uint_as_float(a[8 * q2]), __uint_as_float(a[8 * q2 + 1])),
                  pack_bf16x2(__uint_as_float(a[8 * q2 + 2]), __uint_as_float(a[8 * q2 + 3])),
                  pack_bf16x2(__uint_as_float(a[8 * q2 + 4]), __uint_as_float(a[8 * q2 + 5])),
                  pack_bf16x2(__uint_as_float(a[8 * q2 + 6]), __uint_as_float(a[8 * q2 + 7])));
              *reinterpret_cast<uint4*>(stk + o2) = make_uint4(
                  pack_bf16x2(__uint_as_float(bb[8 * q2]), __uint_as_float(bb[8 * q2 + 1])),
                  pack_bf16x2(__uint_as_float(bb[8 * q2 + 2]), __uint_as_float(bb[8 * q2 + 3])),
                  pack_bf16x2(__uint_as_float(bb[8 * q2 + 4]), __uint_as_float(bb[8 * q2 + 5])),
                  pack_bf16x2(__uint_as_float(bb[8 * q2 + 6]), __uint_as_float(bb[8 * q2 + 7])));
            }
          }
          if (j == 1) {
            // dQ: warp (quarter, ch) reads query tile `ch`, rows quarter*32 + lane, all 64 columns (the last MMA2 retired above)
            uint32_t x0[32], x1[32];
            tmem_ld_32x32b_x32(t_lane + 384 + 64 * ch, x0);
            tmem_ld_32x32b_x32(t_lane + 384 + 64 * ch + 32, x1);
            tmem_ld_wait();
            tc_fence_before_sync();
            mbar_arrive(bar_dqfree);
            mbar_arrive(&bar_dfree[ip]);
            uint8_t* stq = (ch == 0 ? sS + kBwBlk : sP + kBwBlk) + kr * 128;
#pragma unroll
            for (int q2 = 0; q2 < 4; ++q2) {
              *reinterpret_cast<uint4*>(stq + ((q2 ^ (kr & 7)) << 4)) = make_uint4(
                  pack_bf16x2(__uint_as_float(x0[8 * q2]), __uint_as_float(x0[8 * q2 + 1])),
                  pack_bf16x2(__uint_as_float(x0[8 * q2 + 2]), __uint_as_float(x0[8 * q2 + 3])),
                  pack_bf16x2(__uint_as_float(x0[8 * q2 + 4]), __uint_as_float(x0[8 * q2 + 5])),
                  pack_bf16x2(__uint_as_float(x0[8 * q2 + 6]), __uint_as_float(x0[8 * q2 + 7])));
              *reinterpret_cast<uint4*>(stq + (((4 + q2) ^ (kr & 7)) << 4)) = make_uint4(
                  pack_bf16x2(__uint_as_float(x1[8 * q2]), __uint_as_float(x1[8 * q2 + 1])),
                  pack_bf16x2(__uint_as_float(x1[8 * q2 + 2]), __uint_as_float(x1[8 * q2 + 3])),
                  pack_bf16x2(__uint_as_float(x1[8 * q2 + 4]), __uint_as_float(x1[8 * q2 + 5])),
                  pack_bf16x2(__uint_as_float(x1[8 * q2 + 6]), __uint_as_float(x1[8 * q2 + 7])));
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(2, 32 * kB2Workers);
          if (threadIdx.x == 0) {
            tma_store_3d(&tmDQ, sS + 2 * kBwBlk, 2 * p.D + h * kTcHd, j * 128, f);
            tma_store_3d(&tmDQ, sS + 3 * kBwBlk, p.D + h * kTcHd, j * 128, f);
            if (j == 1) {
              tma_store_3d(&tmDQ, sS + kBwBlk, h * kTcHd, 0, f);
              tma_store_3d(&tmDQ, sP + kBwBlk, h * kTcHd, 128, f);
            }
            tma_store_commit();
          }
          store_pending = true;
        }
        if (threadIdx.x == 0) TRACE(240 + lc);
      }
    }
    if (threadIdx.x == 0) tma_store_wait<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  TRACE_DUMP
  if (warp == kB2Workers) tmem_dealloc(tm, 512);
}

}  // namespace avt

using namespace avt;

extern "C" int avt_attention_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int F,
                                    int H, int N, float scale, void* stream) {
  AVT_REQUIRE(qkv && out && dout && lse && dqkv, "null pointer");
  AVT_REQUIRE(F > 0 && H > 0 && N > 0 && N <= kTcKeys, "tokens per frame must be in [1, 208]");
  const int D = H * kTcHd;
  CUtensorMap tmQKV, tmDO, tmDQ;
  const uint64_t rows = (uint64_t)F * N;
  if (int rc = make_tmap_bf16_3d(&tmDQ, dqkv, 3ull * D, (uint64_t)N, (uint64_t)F, 3ull * D, 64, 128, 128)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmQKV, qkv, 3ull * D, rows, 3ull * D, 64, kTcKeys, 128)) return rc;
  if (int rc = make_tmap_bf16_2d(&tmDO, dout, (uint64_t)D, rows, (uint64_t)D, 64, kTcKeys, 128)) return rc;
  AttnTcBwdParams p;
  p.out = reinterpret_cast<const bf16*>(out); p.dout = reinterpret_cast<const bf16*>(dout); p.lse = lse;
  p.dqkv = reinterpret_cast<bf16*>(dqkv); p.N = N; p.H = H; p.D = D; p.scale = scale;
    static bool configured2 = false;
    if (!configured2) {
      AVT_CUDA_OK(cudaFuncSetAttribute(attn_tc_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kB2Smem));
      configured2 = true;
    }
    const int items = F * H;
    const int grid = items < num_sms() ? items : num_sms();
    launch_kernel(attn_tc_bwd2_kernel, dim3(grid), dim3(kB2Threads), kB2Smem, reinterpret_cast<cudaStream_t>(stream), tmQKV, tmDO, tmDQ, p, items);
    AVT_CUDA_OK(cudaGetLastError());
    return AVT_OK;
}
