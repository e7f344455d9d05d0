// bf16 GEMM with fused epilogues on tcgen05 tensor cores (sm_100a).
//
//   C[M,N] = epilogue(A[M,K] · B[N,K]^T)     fp32 accumulation in tensor memory (TMEM)
//
// Structure (persistent, one CTA per SM, 320 threads):
//   warp 0      TMA producer  : cp.async.bulk.tensor tiles of A and B into a 128B-swizzled smem ring
//   warp 1      MMA issuer    : one lane issues tcgen05.mma (K = 16 per instruction); tcgen05.commit releases smem
//                               slots and publishes the finished accumulator
//   warps 2..3  column sums of the A operand (bias gradients riding on weight-gradient GEMMs), idle otherwise
//   warps 4..11 epilogue      : tcgen05.ld the accumulator (one row per thread, two warps per TMEM lane quarter),
//                               fused epilogue (bias / GELU (+ its derivative) / dropout / residual / pos-embed),
//                               bf16 results leave through per-warp swizzled smem slots + TMA stores
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// CG = 2 ("CTA pair", cta_group::2): two CTAs of a cluster share one 256 x BN tile. Each CTA loads its own 128 rows
// of A and HALF of the B tile; the leader CTA issues tcgen05.mma.cta_group::2 which reads both halves and writes
// 128 rows of the accumulator into each CTA's TMEM. Per output element this cuts the L2->SM operand traffic by a
// third (ncu: the 1-CTA 128x256 tile saturates the L2->SM fabric at ~10 TB/s with the tensor pipe 52 % active).
//
// Either operand may be K-major (contraction dim contiguous) or MN-major (transposed in memory): forward, dgrad and
// wgrad of nn.Linear and HF Conv1D all map onto this one kernel without materialising a transpose. Split-K work
// units accumulate with fp32 atomics (weight gradients).
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"

namespace avt {

constexpr int kBM = 128;        // accumulator rows per CTA (UMMA M = 128 per CTA, 256 per pair)
constexpr int kBK = 64;         // K per smem stage: 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;      // K per tcgen05.mma for 16-bit inputs
constexpr int kSumWarps = 2;     // column sums of the A operand (bias gradients fused into the weight-gradient GEMM)
// Epilogue warps: two per TMEM lane quarter, each takes half of the tile's columns. The kernel is written for 8 or 12
// (three per quarter, 3 / 3 / 2 of the eight 32-column chunks), but 12 bought nothing on B200: the fused erf-GELU epilogue
// is bound by the per-scheduler pipes (64 MUFU x 8 cycles + 240 packed FMA x 2 cycles + ~100 ALU x 2 per chunk, issued
// back to back: ~12 k cycles per 128 x 256 tile against a 9.4 k-cycle main loop), not by per-warp latency, and every
// scheduler sees the same eight chunks per tile whichever way they are dealt (78.1 us vs 78.5 us, tools/sweep.py fc1).
__host__ __device__ constexpr int epi_warps(int epi) { return epi >= 0 ? 8 : 12; }
__host__ __device__ constexpr int gemm_threads(int epi) { return 64 + 32 * kSumWarps + 32 * epi_warps(epi); }
constexpr int kSlotBytes = 2048;  // TMA-store staging slot: 32 rows x 32 bf16 (64-byte rows, SWIZZLE_64B)
constexpr int kSmemLimit = 227 * 1024;

constexpr int kMaxStages = 8;
static int g_epi_special = 1;   // avt_set_gemm_specialized_epilogues(0): always run the generic epilogue (A/B, tests)
template <int BN, int CG, int EW = 8>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = (BN / CG) * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = EW * 4 * kSlotBytes;  // 2 out + 2 in TMA staging slots per epilogue warp
  static constexpr int kBiasBytes = 2 * BN * 4;
  static constexpr int kBarBytes = 512;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // double-buffered accumulator
  // Operand ring depth. Epilogues that leave through plain stores / fp32 atomics (split-K weight gradients) do not need
  // the 64 KB of TMA staging: that memory becomes two more pipeline stages (256-wide pair tiles: 5 -> 7).
  static constexpr int stages(bool staging) {
    const int n = (kSmemLimit - (staging ? kStagingBytes : 0) - kBiasBytes - kBarBytes) / kStageBytes;
    return n > kMaxStages ? kMaxStages : n;
  }
  static constexpr int smem_bytes(bool staging) {
    return stages(staging) * kStageBytes + (staging ? kStagingBytes : 0) + kBiasBytes + kBarBytes;
  }
  static constexpr int kDeepStagingBytes = EW * 8 * kSlotBytes;  // 8 slots per epilogue warp (see GemmParams::deep)
  static constexpr int stages_deep() {
    const int n = (kSmemLimit - kDeepStagingBytes - kBiasBytes - kBarBytes) / kStageBytes;
    return n > kMaxStages ? kMaxStages : n;
  }
  static_assert(stages(true) >= 3, "not enough shared memory for a 3-stage pipeline");
};

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_k_blocks, split_k, kb_per_split;
  int tma_out;  // `out` (and bf16 aux_z) leave through TMA stores: 1 = bf16 out, 2 = fp32 out (plain overwrite)
  int tma_in;   // dact_z arrives through TMA loads into per-warp staging slots
  int split_slices;  // split-K partials go to out + split*M*ldo (deterministic two-pass) instead of atomics
  int cluster_reduce;  // split-K units of a tile form ONE thread-block cluster: partials meet through distributed shared memory
  int stream_k;      // atomics path only: the tiles' k-blocks are dealt out as ONE contiguous range per CTA group
  int stages;        // operand ring depth (GemmCfg::stages)
  int staging;       // TMA staging slots present in shared memory
  int staging_bytes; // size of the staging area: epilogue warps x slots per warp x kSlotBytes
  int deep;          // deep store staging (8 slots per warp, no TMA-in slots): GEMMs that are nothing but their epilogue
  int k_rotate;      // producer walks each tile's k-blocks from a tile-dependent start
  float* a_colsum;   // [M] += sum_k A[m, k] (A MN-major only): the bias gradient when A = dY^T of a weight-gradient GEMM
  avt_epilogue_t ep;
};

__device__ __forceinline__ uint4 pack8(const float2* v) {  // 4 float2 = 8 consecutive columns
  return make_uint4(pack_bf16x2(v[0].x, v[0].y), pack_bf16x2(v[1].x, v[1].y), pack_bf16x2(v[2].x, v[2].y),
                    pack_bf16x2(v[3].x, v[3].y));
}

// One unit of work of a CTA group: k-blocks [kb0, kb1) of output tile `tile`; `split` = slice index of uniform split-K.
struct Work {
  int tile, kb0, kb1, split;
  bool shared;   // other groups contribute to the same tile: partial sums meet in fp32 atomics / workspace slices
};
// Every warp role walks the same sequence. Uniform mode: units (tile, split) round-robin over the groups. Stream-K mode
// (weight gradients: few tiles, very long K, fp32 atomics into the output): the tiles x k-blocks space is cut into one
// contiguous range per group, so every group does the same number of k-blocks whatever the tile count and however many
// SMs the grid was given (72 units on 74 CTA pairs waste 3 %; on the 70 pairs left beside an NCCL all-reduce they would
// need two rounds).
struct WorkIter {
  int ngroups, unit, pos, end;
  __device__ WorkIter(const GemmParams& p, int group, int ngroups_) : ngroups(ngroups_), unit(group), pos(0), end(0) {
    if (p.stream_k) {
      const int total = p.num_m_tiles * p.num_n_tiles * p.num_k_blocks;
      const int q = (total + ngroups - 1) / ngroups;
      pos = min(group * q, total);
      end = min(pos + q, total);
    }
  }
  __device__ __forceinline__ bool next(const GemmParams& p, Work& w) {
    if (p.stream_k) {
      if (pos >= end) return false;
      const int nkb = p.num_k_blocks;
      w.tile = pos / nkb;
      w.kb0 = pos - w.tile * nkb;
      w.kb1 = min(nkb, w.kb0 + (end - pos));
      w.split = 0;
      w.shared = true;
      pos += w.kb1 - w.kb0;
      return true;
    }
    if (unit >= p.num_m_tiles * p.num_n_tiles * p.split_k) return false;
    w.tile = unit / p.split_k;
    w.split = unit % p.split_k;
    w.kb0 = w.split * p.kb_per_split;
    w.kb1 = min(w.kb0 + p.kb_per_split, p.num_k_blocks);
    w.shared = p.split_k > 1;
    unit += ngroups;
    return true;
  }
};

// The run-time epilogue applied to four consecutive columns of one output row (fp32 sums in a4): shared by the finishing pass
// of the two-pass split-K GEMMs and by the cluster split-K reduction inside the GEMM kernel.
__device__ __forceinline__ void apply_epilogue4(const avt_epilogue_t& ep, int row, int col, int N, float4 a4, float keep_scale,
                                                uint64_t drop_off) {
  float v[4] = {a4.x * ep.alpha, a4.y * ep.alpha, a4.z * ep.alpha, a4.w * ep.alpha};
  if (ep.bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (ep.pos_period > 0) {
    const int t = row % ep.pos_period;
    if (t == 0 && ep.cls) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.cls + col));
      v[0] = b.x; v[1] = b.y; v[2] = b.z; v[3] = b.w;
    }
    const float4 b = __ldg(reinterpret_cast<const float4*>(ep.pos + (size_t)t * N + col));
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (ep.aux_z) {
    float a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = ep.aux_mode == 1 ? apply_act_grad(ep.act, v[j]) : v[j];
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.aux_z) + (size_t)row * ep.ldz + col) =
        make_uint2(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]));
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = apply_act(ep.act, v[j]);
  if (ep.dact_z) {
    const uint2 z = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(ep.dact_z) + (size_t)row * ep.ldz + col);
    const float zz[4] = {bf16_lo(z.x), bf16_hi(z.x), bf16_lo(z.y), bf16_hi(z.y)};
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] *= ep.dact_mode == 1 ? zz[j] : apply_act_grad(ep.dact, zz[j]);
  }
  if (ep.drop_p > 0.f) {
    const uint32_t keep = dropout_keep4(ep.drop_seed, drop_off, ((uint64_t)row * (uint64_t)N + (uint64_t)col) >> 2, ep.drop_p);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * keep_scale : 0.f;
  }
  if (ep.residual) {
    const float4 b = *reinterpret_cast<const float4*>(ep.residual + (size_t)row * ep.ldr + col);
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (ep.out_fp32) {
    float* op = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo + col;
    float4 o = make_float4(v[0], v[1], v[2], v[3]);
    if (ep.accumulate) {
      const float4 c = *reinterpret_cast<const float4*>(op);
      o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
    }
    *reinterpret_cast<float4*>(op) = o;
  } else {
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out) + (size_t)row * ep.ldo + col) =
        make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  }
}

// Finishing pass of a split-K GEMM whose partial sums were written as fp32 slices into `acc`: applies the same
// epilogue as the fused path. Only used for the weight-streaming M <= 128 GEMMs of AVT-h (80 x N elements) when the
// split factor does not fit a thread-block cluster.
__global__ void __launch_bounds__(256)
epilogue_apply_kernel(const float* __restrict__ acc, int nslices, int M, int N, const avt_epilogue_t ep) {
  pdl_enter();
  const int64_t total4 = (int64_t)M * N / 4;
  const float keep_scale = ep.drop_p > 0.f ? 1.0f / (1.0f - ep.drop_p) : 1.0f;
  const uint64_t drop_off = ep.drop_offset + ((ep.drop_p > 0.f && ep.drop_offset_dev) ? __ldg(ep.drop_offset_dev) : 0ull);
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total4; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e0 = g * 4;
    const int row = (int)(e0 / N), col = (int)(e0 % N);
    float4 a4 = *reinterpret_cast<const float4*>(acc + e0);
    for (int sl = 1; sl < nslices; ++sl) {   // fixed summation order: bit-reproducible
      const float4 b4 = *reinterpret_cast<const float4*>(acc + (size_t)sl * M * N + e0);
      a4.x += b4.x; a4.y += b4.y; a4.z += b4.z; a4.w += b4.w;
    }
    apply_epilogue4(ep, row, col, N, a4, keep_scale, drop_off);
  }
}

// Epilogue classes (template parameter EPI)
constexpr int kEpiGeneric = 0;   // everything avt_epilogue_t can express, decided at run time
constexpr int kEpiStore = 1;     // [+ bias] -> bf16, TMA store                              (qkv / proj / fc2, plain dgrads)
constexpr int kEpiGeluAux = 2;   // [+ bias], erf-GELU and its derivative -> two bf16 TMA stores (timm Mlp.fc1 forward)
constexpr int kEpiMulZ = 3;      // x saved gelu' (TMA-loaded) -> bf16, TMA store               (fc2 dgrad through the GELU)
constexpr int kEpiAtomic = 4;    // fp32 atomics into the output (stream-K weight gradients, both operands MN-major)
constexpr int kEpiStoreF32 = 5;  // [+ bias] -> fp32, TMA store                              (AVT-h weight gradients, fp32 mode)
constexpr int kEpiResidual = 6;  // [+ bias] + fp32 residual (TMA-loaded) -> fp32, TMA store    (proj / fc2 closing a residual branch)

template <int BN, int CG, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(gemm_threads(EPI), 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAux,
                 const __grid_constant__ CUtensorMap tmIn, const GemmParams p) {
  constexpr int kEpiWarps = epi_warps(EPI);
  using Cfg = GemmCfg<BN, CG, kEpiWarps>;
  constexpr int BNL = BN / CG;  // B rows (N extent) held by this CTA
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nstages = p.stages;
  uint8_t* sStageOut = smem + nstages * Cfg::kStageBytes;                  // [kEpiWarps][2 out + 2 in][kSlotBytes] (if p.staging)
  float* sBias = reinterpret_cast<float*>(sStageOut + p.staging_bytes);  // [2][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + Cfg::kBiasBytes);
  uint64_t* full_bar = bars;                       // [kMaxStages]  TMA -> MMA          (CG=2: the leader's copy is used)
  uint64_t* empty_bar = bars + kMaxStages;         // [kMaxStages]  MMA -> TMA          (every CTA's own copy)
  uint64_t* tfull_bar = bars + 2 * kMaxStages;     // [2]        MMA -> epilogue     (every CTA's own copy)
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]        epilogue -> MMA     (CG=2: the leader's copy)
  uint64_t* tin_bar = tempty_bar + 2;              // [kEpiWarps][4] TMA -> epilogue warp (dact_z: 2 slots, residual: 4)
  uint64_t* sum_bar = tin_bar + 4 * kEpiWarps;     // [kStages]  MMA -> column-sum warps (stage consumed by the tensor core)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sum_bar + kMaxStages);
  const bool do_colsum = A_MN && p.a_colsum != nullptr;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;  // 0 = leader
  if ((smem_u32(smem) & 1023u) != 0) __trap();             // swizzled tiles need 1024-byte alignment

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_out) tma_prefetch_desc(&tmOut);
    if (p.tma_in) tma_prefetch_desc(&tmIn);
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], do_colsum ? 1 + kSumWarps : 1);   // a stage is free once the MMA and the column-sum warps left it
      mbar_init(&sum_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], CG * kEpiWarps);  // one arrive per epilogue warp of every CTA in the group
    }
    for (int s = 0; s < 4 * kEpiWarps; ++s) mbar_init(&tin_bar[s], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, TMEM and tensor-map prefetch are set up: only now wait for the previous kernel of the stream (its tail
  // overlapped all of the above), then let the next kernel start its own prologue
  pdl_wait();
  pdl_trigger();
  // Register re-allocation between the two warpgroup classes: the TMA / MMA / column-sum warps (warpgroup 0) need few
  // registers, the epilogue warpgroups hold 32-column accumulator chunks and evaluate GELU (+ its derivative) on them -
  // latency-bound code that only speeds up with more independent chains in flight. 128 x 56 + 256 x 224 = 384 x 168.
  // (The two setmaxnreg sit in branches that only re-join at the end of the kernel, so each side is compiled for its own budget.)
  const int group = blockIdx.x / CG, ngroups = gridDim.x / CG;
  Work wk;

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");   // 128 x 56 + 256 x 224 = 384 x 168;  128 x 56 + 384 x 152 = 512 x 128
  if (warp == 0) {
    // ============================== TMA producer ==============================
    int stage = 0;
    uint32_t phase = 0;
    for (WorkIter it(p, group, ngroups); it.next(p, wk);) {
      const int tile = wk.tile;
      const int m0 = (tile / p.num_n_tiles) * (kBM * CG) + (int)rank * kBM;
      const int n0 = (tile % p.num_n_tiles) * BN + (int)rank * BNL;
      const int kb0 = wk.kb0, kb1 = wk.kb1;
      // weight-streaming GEMMs (M <= 128): every CTA reads the SAME A k-blocks; starting each tile at a different
      // k offset keeps 100+ SMs from queueing on one L2 slice at the same moment (the sum order is irrelevant)
      const int nkb = kb1 - kb0;
      const int rot = p.k_rotate ? (tile * 5) % nkb : 0;
      for (int i = 0; i < nkb; ++i) {
        const int kb = kb0 + (i + rot >= nkb ? i + rot - nkb : i + rot);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sA = smem + stage * Cfg::kStageBytes;
          uint8_t* sB = sA + Cfg::kABytes;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], CG * Cfg::kStageBytes);
          auto load = [&](const CUtensorMap* tm, void* dst, int c0, int c1) {
            if constexpr (CG == 2) tma_load_2d_pair(tm, &full_bar[stage], dst, c0, c1);
            else tma_load_2d(tm, &full_bar[stage], dst, c0, c1);
          };
          if constexpr (!A_MN) {
            load(&tmA, sA, kb * kBK, m0);  // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < kBM / 64; ++j) load(&tmA, sA + j * 8192, m0 + j * 64, kb * kBK);  // box {64 m, 64 k}
          }
          if constexpr (!B_MN) {
            load(&tmB, sB, kb * kBK, n0);  // box {64 k, BNL rows}
          } else {
#pragma unroll
            for (int j = 0; j < BNL / 64; ++j) load(&tmB, sB + j * 8192, n0 + j * 64, kb * kBK);
          }
        }
        __syncwarp();
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer (leader CTA only when CG = 2) ==============================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc(/*bf16*/ 1, A_MN ? 1 : 0, B_MN ? 1 : 0, kBM * CG, BN);
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
      for (WorkIter it(p, group, ngroups); it.next(p, wk);) {
        const int kb0 = wk.kb0, kb1 = wk.kb1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // every epilogue warp has drained this accumulator
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
              const uint64_t da = smem_desc_addr(descA, sA + k * kStepA), db = smem_desc_addr(descB, sB + k * kStepB);
              const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
              if constexpr (CG == 2) umma_f16_pair(d_tmem, da, db, idesc, accum);
              else umma_f16(d_tmem, da, db, idesc, accum);
            }
            if constexpr (CG == 2) {
              umma_commit_pair(&empty_bar[stage], 3);                    // smem slot free in both CTAs
              if (do_colsum) umma_commit_pair(&sum_bar[stage], 3);
              if (kb == kb1 - 1) umma_commit_pair(&tfull_bar[acc], 3);   // accumulator complete in both CTAs
            } else {
              umma_commit(&empty_bar[stage]);
              if (do_colsum) umma_commit(&sum_bar[stage]);
              if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp < 2 + kSumWarps) {
    // ============================== column sums of A (warps 2..3; weight-gradient GEMMs with a bias) ==============
    // A = dY^T arrives MN-major: per stage two [64 k][64 m] boxes, 128-byte rows, SWIZZLE_128B. The MMA's commit on
    // sum_bar says the tensor core is done with the stage (so the bytes are there); these warps add the 64 k-rows of
    // every m column of the n-tile-0 units into registers, then release the stage. The reduction over the 15 760
    // activation rows rides on operand bytes that are in shared memory anyway: no separate pass over dY in HBM.
    if constexpr (A_MN) {
      if (do_colsum) {
        const int t = threadIdx.x - 64;                        // 0..63
        const int chunk = t & 15;                              // 8 m values (16 bytes): box = chunk / 8
        const int kgrp = t >> 4;                               // 16 k rows each
        const uint32_t box_off = (chunk >> 3) * 8192;
        int stage = 0;
        uint32_t phase = 0;
        for (WorkIter it(p, group, ngroups); it.next(p, wk);) {
          const int tile = wk.tile;
          const bool mine = (tile % p.num_n_tiles) == 0;
          const int m0 = (tile / p.num_n_tiles) * (kBM * CG) + (int)rank * kBM;
          const int kb0 = wk.kb0, kb1 = wk.kb1;
          float acc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&sum_bar[stage], phase);
            if (mine) {
              const uint8_t* sA = smem + stage * Cfg::kStageBytes + box_off;
#pragma unroll
              for (int r = 0; r < 16; ++r) {
                const int row = kgrp * 16 + r;
                const uint4 v = *reinterpret_cast<const uint4*>(sA + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
                acc8[0] += bf16_lo(v.x); acc8[1] += bf16_hi(v.x); acc8[2] += bf16_lo(v.y); acc8[3] += bf16_hi(v.y);
                acc8[4] += bf16_lo(v.z); acc8[5] += bf16_hi(v.z); acc8[6] += bf16_lo(v.w); acc8[7] += bf16_hi(v.w);
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == nstages) { stage = 0; phase ^= 1; }
          }
          if (mine) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {   // the four k-groups of a column meet in a shuffle, then one atomic per column
              float v = acc8[j];
              v += __shfl_xor_sync(0xffffffffu, v, 16);
              const int m = m0 + chunk * 8 + j;
              if (lane < 16 && m < p.M) atomicAdd(p.a_colsum + m, v);
            }
          }
        }
      }
    }
  }
  if constexpr (CG == 1 && EPI == kEpiGeneric) {
    if (p.cluster_reduce) {   // the epilogue warps' two cluster barriers (partials written / partials consumed)
      cluster_sync_all();
      cluster_sync_all();
    }
  }
  } else {
    if constexpr (kEpiWarps == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    // ============================== epilogue (warps 4..11) ==============================
    // Math runs on packed fp32x2 (FFMA2): the fused epilogues are issue-bound, two columns per instruction.
    const avt_epilogue_t& ep = p.ep;
    // EPI > 0: the hot ViT epilogues with everything but the bias decided at compile time (the generic body carries ~15
    // uniform branches, both GELU flavours in four modes and the dropout / pos-embed / residual paths: 600 instructions
    // per 32-column chunk where the specialised bodies need 300-450, and the fused-GELU epilogue is issue-bound)
    constexpr bool kGen = EPI == kEpiGeneric;
    const bool f_pos = kGen && ep.pos_period > 0;
    const bool f_aux = kGen ? ep.aux_z != nullptr : EPI == kEpiGeluAux;
    const bool f_dact = kGen ? ep.dact_z != nullptr : EPI == kEpiMulZ;
    const bool f_drop = kGen && ep.drop_p > 0.f;
    const bool f_res = kGen && ep.residual != nullptr;
    const int f_tma_out = kGen ? p.tma_out : (EPI == kEpiAtomic ? 0 : ((EPI == kEpiStoreF32 || EPI == kEpiResidual) ? 2 : 1));
    const bool f_out_fp32 = kGen ? ep.out_fp32 != 0 : (EPI == kEpiAtomic || EPI == kEpiStoreF32 || EPI == kEpiResidual);
    const bool f_slices = kGen && p.split_slices != 0;
    const bool f_tma_in = kGen ? p.tma_in != 0 : EPI == kEpiMulZ;
    const int f_act = kGen ? ep.act : (EPI == kEpiGeluAux ? AVT_ACT_GELU_ERF : AVT_ACT_NONE);
    const int f_aux_mode = kGen ? ep.aux_mode : 1;
    const int f_dact_mode = kGen ? ep.dact_mode : 1;
    const int ew = warp - (2 + kSumWarps);
    const int quarter = warp & 3;   // TMEM lanes [32*quarter, +32) are the only ones this warp may read
    const int cpart = ew >> 2;      // which part (half / third) of the tile's columns this warp handles
    constexpr int kParts = kEpiWarps / 4, kChunks = BN / 32;
    const int c_part_begin = 32 * (cpart * (kChunks / kParts) + min(cpart, kChunks % kParts));
    const int c_part_end = c_part_begin + 32 * (kChunks / kParts + (cpart < kChunks % kParts ? 1 : 0));
    const int etid = threadIdx.x - 32 * (2 + kSumWarps);
    // [2] staging for TMA stores + [2] for TMA loads (dact_z); deep mode: [8] for stores
    const bool f_deep = (kGen || EPI == kEpiStore || EPI == kEpiStoreF32) && p.deep != 0;
    uint8_t* out_slots = sStageOut + ew * (EPI == kEpiResidual ? 6 : (f_deep ? 8 : 4)) * kSlotBytes;
    uint8_t* in_slots = out_slots + 2 * kSlotBytes;
    uint64_t* in_bar = tin_bar + ew * 4;
    uint32_t n_st = 0, n_in_issued = 0, n_in_waited = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const float keep_scale = f_drop ? 1.0f / (1.0f - ep.drop_p) : 1.0f;
    // CUDA-graph friendly RNG: the per-step part of the Philox offset may live in device memory
    const uint64_t drop_off = ep.drop_offset + ((f_drop && ep.drop_offset_dev) ? __ldg(ep.drop_offset_dev) : 0ull);
    const bool scale_acc = kGen && ep.alpha != 1.0f;
    const int sw = (lane >> 1) & 3;  // SWIZZLE_64B: 16-byte chunk index ^= bits [7,9) of the byte offset (64 B rows)

    // Deep mode (weight gradients over the 80 rows of AVT-h: the kernel is nothing but this epilogue streaming its output):
    // with two slots per warp, 16 x 2-4 KB in flight per SM, the stores were bound by the TMA round trip (2.9 TB/s); eight
    // slots per warp (the operand ring needs only 3 stages there) quadruple the bytes in flight.
    auto tma_store_chunk = [&](const CUtensorMap* tm, const float2* v, int col0, int row0) {
      uint8_t* slot = out_slots + (f_deep ? (n_st & 7) : (n_st & 1)) * kSlotBytes;
      if (lane == 0) {  // the store that last used this slot no longer reads it
        if (f_deep) tma_store_wait_read<7>();
        else tma_store_wait_read<1>();
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(slot + lane * 64 + ((j ^ sw) << 4)) = pack8(v + 4 * j);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tm, slot, col0, row0);
        tma_store_commit();
      }
      ++n_st;
    };
    auto tma_store_chunk_f32 = [&](const CUtensorMap* tm, const float2* v, int col0, int row0) {
      // 32 rows x 32 fp32 = 4 KB, 128-byte rows, SWIZZLE_128B (chunk ^= row & 7). Two buffers when the dact_z staging
      // slots are free (both out slots / both in slots): the next chunk is staged while the TMA drains this one - the
      // weight gradients of AVT-h (contraction over 80 rows) are nothing but this epilogue streaming 67 MB to HBM.
      uint8_t* buf = out_slots + (f_deep ? (n_st & 3) * 2 * kSlotBytes : ((!f_tma_in && (n_st & 1)) ? 2 * kSlotBytes : 0));
      if (lane == 0) {
        if (f_deep) tma_store_wait_read<3>();
        else if (f_tma_in) tma_store_wait_read<0>();
        else tma_store_wait_read<1>();
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
            make_float4(v[2 * j].x, v[2 * j].y, v[2 * j + 1].x, v[2 * j + 1].y);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tm, buf, col0, row0);
        tma_store_commit();
      }
      ++n_st;
    };
    auto issue_in = [&](int col0, int row0) {   // prefetch a 32 x 32 bf16 block of dact_z into a staging slot
      if (lane == 0) {
        uint64_t* bar = &in_bar[n_in_issued & 1];
        fence_proxy_async_smem();
        mbar_arrive_expect_tx(bar, kSlotBytes);
        tma_load_2d(&tmIn, bar, in_slots + (n_in_issued & 1) * kSlotBytes, col0, row0);
      }
      ++n_in_issued;
    };

    bool cluster_done = false;
    if constexpr (CG == 1 && EPI == kEpiGeneric) {
      if (p.cluster_reduce) {
        // Split-K over a thread-block cluster (weight-streaming M <= 128 GEMMs of AVT-h): CTA `split` of the cluster holds the
        // partial sums of k-slab `split` of ONE output tile. Instead of fp32 slices in HBM + a finishing kernel, every CTA
        // parks its accumulator in its own (now idle) operand ring, and after a cluster barrier CTA r sums column slice r
        // of all the peers' copies through distributed shared memory - in a fixed order, so the result is bit-reproducible -
        // and applies the run-time epilogue to it. One unit per CTA (grid = tiles x split_k).
        constexpr int PS = BN + 4;                 // padded row pitch (floats): row-per-thread float4 stores stay conflict-free
        float* P = reinterpret_cast<float*>(smem);
        const int tile = blockIdx.x / p.split_k;
        const uint32_t crank = cluster_ctarank();  // == k-slab index (cluster CTAs are consecutive blocks)
        const int m0 = (tile / p.num_n_tiles) * kBM;
        const int n0 = (tile % p.num_n_tiles) * BN;
        const int rows_valid = min(kBM, p.M - m0);
        mbar_wait(&tfull_bar[0], 0);
        tc_fence_after_sync();
        const int prow = quarter * 32 + lane;
        if (quarter * 32 < rows_valid) {
          const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16);
#pragma unroll 1
          for (int c0 = c_part_begin; c0 < c_part_end; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_row + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(P + prow * PS + c0 + 4 * j) = make_float4(
                  __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          }
        }
        tc_fence_before_sync();
        cluster_sync_all();                        // every CTA's partial tile is visible cluster-wide
        {
          const int cols_per = BN / p.split_k;     // column slice of this CTA (multiple of 4: BN 64 / 128, split 2 / 4 / 8)
          const int f4_per_row = cols_per / 4;
          const int total = rows_valid * f4_per_row;
          const float keep_scale = ep.drop_p > 0.f ? 1.0f / (1.0f - ep.drop_p) : 1.0f;
          for (int idx = etid; idx < total; idx += 32 * kEpiWarps) {
            const int rr = idx / f4_per_row, cc = (int)crank * cols_per + 4 * (idx - rr * f4_per_row);
            if (n0 + cc >= p.N) continue;
            const uint32_t addr = smem_u32(P + rr * PS + cc);
            float4 a4 = ld_shared_cluster_f4(addr, 0);
            for (int sl = 1; sl < p.split_k; ++sl) {
              const float4 b4 = ld_shared_cluster_f4(addr, (uint32_t)sl);
              a4.x += b4.x; a4.y += b4.y; a4.z += b4.z; a4.w += b4.w;
            }
            apply_epilogue4(ep, m0 + rr, n0 + cc, p.N, a4, keep_scale, drop_off);
          }
        }
        cluster_sync_all();                        // nobody leaves while a peer still reads its partial tile
        cluster_done = true;
      }
    }
    for (WorkIter it(p, group, ngroups); !cluster_done && it.next(p, wk);) {
      const int tile = wk.tile;
      const int split = wk.split;
      const int m0 = (tile / p.num_n_tiles) * (kBM * CG) + (int)rank * kBM;
      const int n0 = (tile % p.num_n_tiles) * BN;
      const int row0 = m0 + quarter * 32;
      const int row = row0 + lane;
      const bool row_ok = row < p.M;
      const int c_begin = c_part_begin;
      const int c_end = min(c_part_end, p.N - n0);  // N is a multiple of 32 (checked on host)
      // stage this tile's bias slice in smem (one global read per column instead of one per row)
      float* bias_s = sBias + acc * BN;
      if (ep.bias) {
        for (int i = etid; i < BN; i += 32 * kEpiWarps) bias_s[i] = (n0 + i < p.N) ? __ldg(ep.bias + n0 + i) : 0.f;
      }
      if constexpr (EPI == kEpiResidual) {
        // x_out = x_in + (acc + bias): the branch-closing Linear does the residual add itself. The fp32 residual tile arrives by
        // TMA in 32-row x 16-column boxes (2 KB, 64-byte rows, SWIZZLE_64B) through a ring of FOUR slots per warp: the first
        // four of a tile's (up to) eight boxes are requested before the accumulator is even complete, so they land under
        // the main loop; the result leaves the same way through two slots. The extra 8 B per element ride on a tensor-bound
        // kernel instead of the HBM-bound LayerNorm that used to do the add.
        const int nhalf = (c_end - c_begin) / 16;
        auto issue_res = [&](int h) {
          if (lane == 0) {
            uint64_t* bar = &in_bar[n_in_issued & 3];
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(bar, kSlotBytes);
            tma_load_2d(&tmIn, bar, in_slots + (n_in_issued & 3) * kSlotBytes, n0 + c_begin + 16 * h, row0);
          }
          ++n_in_issued;
        };
        for (int h = 0; h < nhalf && h < 4; ++h) issue_res(h);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after_sync();
        const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_end; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int ch0 = c0 + 16 * hh, col0 = n0 + ch0;
            const uint32_t sl = n_in_waited & 3;
            mbar_wait(&in_bar[sl], (n_in_waited >> 2) & 1);
            const uint8_t* islot = in_slots + sl * kSlotBytes + lane * 64;
            float4 res[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) res[j] = *reinterpret_cast<const float4*>(islot + ((j ^ sw) << 4));
            __syncwarp();
            ++n_in_waited;
            const int hnext = (ch0 - c_begin) / 16 + 4;
            if (hnext < nhalf) issue_res(hnext);                     // into the slot just read
            uint8_t* oslot = out_slots + (n_st & 1) * kSlotBytes;
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (ep.bias) b = *reinterpret_cast<const float4*>(bias_s + ch0 + 4 * j);
              const int i = 16 * hh + 4 * j;
              *reinterpret_cast<float4*>(oslot + lane * 64 + ((j ^ sw) << 4)) =
                  make_float4(__uint_as_float(r[i]) + b.x + res[j].x, __uint_as_float(r[i + 1]) + b.y + res[j].y,
                              __uint_as_float(r[i + 2]) + b.z + res[j].z, __uint_as_float(r[i + 3]) + b.w + res[j].w);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmOut, oslot, col0, row0);
              tma_store_commit();
            }
            ++n_st;
          }
        }
      } else {
      if (f_tma_in && c_begin < c_end) issue_in(n0 + c_begin, row0);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
      const int pos_t = f_pos ? row % ep.pos_period : 0;
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        const int col0 = n0 + c0;
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + c0, r);
        uint4 zraw[4];
        if (f_dact) {
          if (f_tma_in) {
            if (c0 + 32 < c_end) issue_in(col0 + 32, row0);
            const uint32_t sl = n_in_waited & 1;
            mbar_wait(&in_bar[sl], (n_in_waited >> 1) & 1);
            const uint8_t* slot = in_slots + sl * kSlotBytes + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) zraw[j] = *reinterpret_cast<const uint4*>(slot + ((j ^ sw) << 4));
            __syncwarp();
            ++n_in_waited;
          } else if (row_ok) {
            const uint4* zp =
                reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(ep.dact_z) + (size_t)row * ep.ldz + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) zraw[j] = __ldg(zp + j);
          }
        }
        tmem_ld_wait();
        float2 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        if (scale_acc) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __fmul2_rn(v[j], f2(ep.alpha));
        }
        if (ep.bias) {
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float4 b = *reinterpret_cast<const float4*>(bias_s + c0 + 2 * j);
            v[j] = __fadd2_rn(v[j], make_float2(b.x, b.y));
            v[j + 1] = __fadd2_rn(v[j + 1], make_float2(b.z, b.w));
          }
        }
        if (f_pos && row_ok) {
          if (pos_t == 0 && ep.cls) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(ep.cls + col0 + 2 * j));
              v[j] = make_float2(b.x, b.y);
              v[j + 1] = make_float2(b.z, b.w);
            }
          }
          const float* pp = ep.pos + (size_t)pos_t * p.N + col0;
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(pp + 2 * j));
            v[j] = __fadd2_rn(v[j], make_float2(b.x, b.y));
            v[j + 1] = __fadd2_rn(v[j + 1], make_float2(b.z, b.w));
          }
        }
        if (f_aux) {
          if (f_tma_out == 1) {
            // act (and act') computed, packed and staged 16 bytes at a time; the saved tensor never sits in registers
            uint8_t* slot = out_slots + (n_st & 1) * kSlotBytes;
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
            if (f_aux_mode == 1) act_chunk_to_slot<2>(f_act, v, slot + lane * 64, sw);  // save act'(pre-activation)
            else act_chunk_to_slot<1>(f_act, v, slot + lane * 64, sw);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmAux, slot, col0, row0);
              tma_store_commit();
            }
            ++n_st;
          } else {
            float2 a[16];
            if (f_aux_mode == 1) act_chunk<2>(f_act, v, a);  // save act'(pre-activation): backward only multiplies
            else act_chunk<1>(f_act, v, a);
            if (row_ok) {
              uint4* zp = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.aux_z) + (size_t)row * ep.ldz + col0);
#pragma unroll
              for (int j = 0; j < 4; ++j) zp[j] = pack8(a + 4 * j);
            }
          }
        } else if (f_act != AVT_ACT_NONE) {
          float2 unused[16];
          act_chunk<0>(f_act, v, unused);
        }
        if (f_dact && (row_ok || f_tma_in)) {
          float2 z[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t zz[4] = {zraw[j].x, zraw[j].y, zraw[j].z, zraw[j].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) z[4 * j + q] = make_float2(bf16_lo(zz[q]), bf16_hi(zz[q]));
          }
          if (f_dact_mode != 1) {
            float2 unused[16];
            act_chunk<3>(ep.dact, z, unused);   // z <- dact'(z)
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __fmul2_rn(v[j], z[j]);
        }
        if (f_drop) {
          const uint64_t g0 = ((uint64_t)row * (uint64_t)p.N + (uint64_t)col0) >> 2;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t keep = dropout_keep4(ep.drop_seed, drop_off, g0 + j, ep.drop_p);
            v[2 * j].x = (keep & 1u) ? v[2 * j].x * keep_scale : 0.f;
            v[2 * j].y = (keep & 2u) ? v[2 * j].y * keep_scale : 0.f;
            v[2 * j + 1].x = (keep & 4u) ? v[2 * j + 1].x * keep_scale : 0.f;
            v[2 * j + 1].y = (keep & 8u) ? v[2 * j + 1].y * keep_scale : 0.f;
          }
        }
        if (f_res && row_ok) {
          const float4* rp = reinterpret_cast<const float4*>(ep.residual + (size_t)row * ep.ldr + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(rp + j);
            v[2 * j] = __fadd2_rn(v[2 * j], make_float2(b.x, b.y));
            v[2 * j + 1] = __fadd2_rn(v[2 * j + 1], make_float2(b.z, b.w));
          }
        }
        if (f_tma_out == 1) {
          tma_store_chunk(&tmOut, v, col0, row0);
        } else if (f_tma_out == 2) {
          tma_store_chunk_f32(&tmOut, v, col0, row0);
        } else if (row_ok) {
          if (f_out_fp32) {
            float* op = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo + col0;
            if (f_slices) {
              op += (size_t)split * p.M * ep.ldo;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(op + 4 * j) = make_float4(v[2 * j].x, v[2 * j].y, v[2 * j + 1].x, v[2 * j + 1].y);
            } else if (kGen ? wk.shared : true) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                atomicAdd(reinterpret_cast<float4*>(op + 4 * j), make_float4(v[2 * j].x, v[2 * j].y, v[2 * j + 1].x, v[2 * j + 1].y));
            } else if (ep.accumulate) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 o = *reinterpret_cast<float4*>(op + 4 * j);
                o.x += v[2 * j].x; o.y += v[2 * j].y; o.z += v[2 * j + 1].x; o.w += v[2 * j + 1].y;
                *reinterpret_cast<float4*>(op + 4 * j) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(op + 4 * j) = make_float4(v[2 * j].x, v[2 * j].y, v[2 * j + 1].x, v[2 * j + 1].y);
            }
          } else {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.out) + (size_t)row * ep.ldo + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) op[j] = pack8(v + 4 * j);
          }
        }
      }
      }  // EPI != kEpiResidual
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(&tempty_bar[acc], 0);  // the leader's barrier collects both CTAs
        else mbar_arrive(&tempty_bar[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) tma_store_wait<0>();  // all staged results have reached global memory
  }

  tc_fence_before_sync();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // neither CTA frees TMEM / exits while its peer still uses the pair
  if (warp == 1) {
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
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

// 2-D tensor map: `inner` contiguous elements, `outer` rows of `ld` elements; swizzle_bytes in {64, 128}.
static int make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                        uint32_t box_outer, int swizzle_bytes, bool fp32);
int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                      uint32_t box_outer, int swizzle_bytes) {
  return make_tmap_2d(tm, base, inner, outer, ld, box_inner, box_outer, swizzle_bytes, false);
}
static int make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                        uint32_t box_outer, int swizzle_bytes, bool fp32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled", "driver entry point not available", __FILE__, __LINE__);
    return AVT_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * (fp32 ? 4 : 2)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "CUresult %d (base %p inner %llu outer %llu ld %llu box %u x %u)", (int)r, base,
             (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
    set_last_error("cuTensorMapEncodeTiled", msg, __FILE__, __LINE__);
    return AVT_ERR_CUDA;
  }
  return AVT_OK;
}

// 3-D tensor map over a row-major bf16 matrix viewed as [outer2][outer1][inner] (e.g. [frame][token][column]): boxes never
// cross an outer2 boundary - rows past `outer1` are out of bounds, so loads zero-fill them and stores drop them.
int make_tmap_bf16_3d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer1, uint64_t outer2, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer1, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled", "driver entry point not available", __FILE__, __LINE__);
    return AVT_ERR_CUDA;
  }
  cuuint64_t dims[3] = {inner, outer1, outer2};
  cuuint64_t strides[2] = {ld * 2, ld * 2 * outer1};
  cuuint32_t box[3] = {box_inner, box_outer1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "CUresult %d (3-D, base %p inner %llu outer %llu x %llu ld %llu)", (int)r, base,
             (unsigned long long)inner, (unsigned long long)outer1, (unsigned long long)outer2, (unsigned long long)ld);
    set_last_error("cuTensorMapEncodeTiled", msg, __FILE__, __LINE__);
    return AVT_ERR_CUDA;
  }
  return AVT_OK;
}

struct GemmMaps {
  CUtensorMap a, b, out, aux, in;
};

template <int BN, int CG, bool A_MN, bool B_MN, int EPI = kEpiGeneric>
static int launch_gemm(const GemmMaps& tm, const GemmParams& p_in, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CG, epi_warps(EPI)>;
  auto kern = gemm_bf16_kernel<BN, CG, A_MN, B_MN, EPI>;
  static bool configured = false;
  if (!configured) {
    int mx = Cfg::smem_bytes(true) > Cfg::smem_bytes(false) ? Cfg::smem_bytes(true) : Cfg::smem_bytes(false);
    const int deep_bytes = Cfg::stages_deep() * Cfg::kStageBytes + Cfg::kDeepStagingBytes + Cfg::kBiasBytes + Cfg::kBarBytes;
    if (Cfg::stages_deep() >= 3 && deep_bytes > mx) mx = deep_bytes;
    AVT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    configured = true;
  }
  GemmParams p = p_in;
  p.staging = (p.tma_out || p.tma_in) ? 1 : 0;
  p.stages = Cfg::stages(p.staging != 0);
  p.staging_bytes = p.staging ? Cfg::kStagingBytes : 0;
  // deep store staging: generic-epilogue GEMMs with a very short contraction whose output leaves through TMA stores
  p.deep = ((EPI == kEpiGeneric || EPI == kEpiStore || EPI == kEpiStoreF32) && p.tma_out && !p.tma_in && !p.ep.aux_z &&
            p.kb_per_split <= 3 && Cfg::stages_deep() >= 3) ? 1 : 0;
  if (p.deep) {
    p.stages = Cfg::stages_deep();
    p.staging_bytes = Cfg::kDeepStagingBytes;
  }
  if (EPI == kEpiResidual) {   // 2 out + 4 in slots per epilogue warp: one operand stage less
    p.staging_bytes = epi_warps(EPI) * 6 * kSlotBytes;
    const int n = (kSmemLimit - p.staging_bytes - Cfg::kBiasBytes - Cfg::kBarBytes) / Cfg::kStageBytes;
    p.stages = n > kMaxStages ? kMaxStages : n;
  }
  const int units = p.num_m_tiles * p.num_n_tiles * (p.stream_k ? p.num_k_blocks : p.split_k);
  const int groups = num_sms() / CG;
  const int grid = p.cluster_reduce ? units : CG * (units < groups ? units : groups);   // cluster split-K: one unit per CTA
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(gemm_threads(EPI));
  cfg.dynamicSmemBytes = (p.deep || EPI == kEpiResidual) ? p.stages * Cfg::kStageBytes + p.staging_bytes + Cfg::kBiasBytes + Cfg::kBarBytes
                                                         : Cfg::smem_bytes(p.staging != 0);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster_reduce ? p.split_k : CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  count_launch();
  AVT_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tm.a, tm.b, tm.out, tm.aux, tm.in, p));
  return AVT_OK;
}

template <int BN, int CG>
static int dispatch_major(int a_mn, int b_mn, const GemmMaps& tm, const GemmParams& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_gemm<BN, CG, false, false>(tm, p, s);
  if (!a_mn && b_mn) return launch_gemm<BN, CG, false, true>(tm, p, s);
  if (a_mn && !b_mn) return launch_gemm<BN, CG, true, false>(tm, p, s);
  return launch_gemm<BN, CG, true, true>(tm, p, s);
}
// The specialised epilogues exist for the 256-wide CTA-pair kernel with a row-major A operand (every big ViT forward /
// dgrad GEMM); B is [N, K] (nn.Linear forward) or [K, N] (dgrad).
template <int EPI>
static int dispatch_epi(int b_mn, const GemmMaps& tm, const GemmParams& p, cudaStream_t s) {
  return b_mn ? launch_gemm<256, 2, false, true, EPI>(tm, p, s) : launch_gemm<256, 2, false, false, EPI>(tm, p, s);
}

}  // namespace avt

using namespace avt;

extern "C" int avt_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int64_t M,
                             int64_t N, int64_t K, const avt_epilogue_t* ep, int split_k, int block_n, int cta_group,
                             void* workspace, int64_t workspace_bytes, void* stream) {
  return avt_gemm_bf16_colsum(A, lda, a_mn, B, ldb, b_mn, M, N, K, ep, split_k, block_n, cta_group, workspace,
                              workspace_bytes, nullptr, stream);
}

extern "C" int avt_gemm_bf16_colsum(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int64_t M,
                                    int64_t N, int64_t K, const avt_epilogue_t* ep, int split_k, int block_n, int cta_group,
                                    void* workspace, int64_t workspace_bytes, float* a_colsum, void* stream) {
  AVT_REQUIRE(A && B && ep && ep->out, "null pointer");
  AVT_REQUIRE(M > 0 && N > 0 && K > 0, "empty problem");
  AVT_REQUIRE(N % 32 == 0, "N must be a multiple of 32");
  AVT_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "leading dimensions must be multiples of 8 elements (16 bytes)");
  AVT_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
              "operands must be 16-byte aligned");
  AVT_REQUIRE(ep->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(ep->out) & 15) == 0, "output must be 16-byte aligned");
  if (cta_group != 1 && cta_group != 2) cta_group = (M >= 512 && N >= 256) ? 2 : 1;  // pairs pay off on big tiles
  if (block_n <= 0) block_n = (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : 64);
  AVT_REQUIRE(block_n == 64 || block_n == 128 || block_n == 256, "block_n must be 64, 128 or 256");
  if (cta_group == 2 && block_n == 64) cta_group = 1;
  if (split_k < 1) split_k = 1;
  const bool plain_acc = ep->out_fp32 && !ep->bias && !ep->aux_z && !ep->dact_z && !ep->residual && ep->act == AVT_ACT_NONE &&
                         ep->drop_p == 0.f && ep->pos_period == 0 && (ep->alpha == 0.f || ep->alpha == 1.f);
  // partial sums into `workspace` slices + epilogue_apply_kernel whenever a workspace is supplied (deterministic);
  // fp32 atomics straight into `out` only for the plain accumulation without workspace (weight gradients)
  const bool two_pass = split_k > 1 && (!plain_acc || workspace != nullptr);
  if (two_pass) {
    AVT_REQUIRE(workspace && workspace_bytes >= (int64_t)split_k * M * N * 4,
                "split_k with a fused epilogue needs a split_k*M*N fp32 workspace");
    AVT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
  }
  if (ep->residual) AVT_REQUIRE(ep->ldr % 4 == 0, "residual ld must be a multiple of 4");
  if (ep->aux_z || ep->dact_z) AVT_REQUIRE(ep->ldz % 8 == 0, "z ld must be a multiple of 8");

  AVT_REQUIRE(!a_colsum || a_mn, "a_colsum needs the A operand stored transposed (a_mn = 1: A = dY^T of a weight gradient)");
  GemmParams p;
  p.a_colsum = a_colsum;
  p.k_rotate = M <= 128 ? 1 : 0;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.num_m_tiles = (int)((M + kBM * cta_group - 1) / (kBM * cta_group));
  p.num_n_tiles = (int)((N + block_n - 1) / block_n);
  p.num_k_blocks = (int)((K + kBK - 1) / kBK);
  if (split_k > p.num_k_blocks) split_k = p.num_k_blocks;
  p.kb_per_split = (p.num_k_blocks + split_k - 1) / split_k;
  p.split_k = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.ep = *ep;
  if (p.ep.alpha == 0.f) p.ep.alpha = 1.0f;
  avt_epilogue_t finish = p.ep;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (!two_pass && p.split_k > 1 && !p.ep.accumulate) {
    // plain fp32 output, overwrite semantics: the atomics need a zeroed destination
    AVT_CUDA_OK(cudaMemset2DAsync(p.ep.out, (size_t)p.ep.ldo * 4, 0, (size_t)N * 4, (size_t)M, s));
  }
  p.split_slices = 0;
  p.cluster_reduce = 0;
  p.stream_k = (!two_pass && p.split_k > 1) ? 1 : 0;   // fp32 atomics: balance the k-blocks over whatever grid there is
  {
    // ... unless uniform split-K units fill one round of the grid almost completely (fc1 / fc2 weight gradients: 36 tiles x 2
    // = 72 units on 74 CTA pairs). Uniform units walk their k-slabs in lockstep, so the pairs that share an operand block
    // fetch it at the same time and it comes out of HBM once; stream-K ranges start at unrelated k offsets and ncu saw every
    // operand byte read twice (243 MB instead of 121 MB, tensor pipe 61 % active against 75 % for the forward GEMM).
    const int units = p.num_m_tiles * p.num_n_tiles * p.split_k;
    const int groups = num_sms() / cta_group;
    static const bool no_uniform = getenv("AVT_WGRAD_STREAMK") != nullptr;
    if (p.stream_k && !no_uniform && units <= groups && units * 10 >= groups * 9) p.stream_k = 0;
  }
  // Split factors 2 / 4 / 8 of a single-CTA-tile GEMM run as ONE thread-block cluster per output tile: the partial sums meet in
  // distributed shared memory inside the kernel (no fp32 slices in HBM, no finishing launch - 52 launches per AVT-h step).
  static const bool no_cluster = getenv("AVT_NO_CLUSTER_SPLITK") != nullptr;
  const bool cluster_reduce = two_pass && !no_cluster && cta_group == 1 && !a_colsum &&
                              (p.split_k == 2 || p.split_k == 4 || p.split_k == 8) && block_n % (4 * p.split_k) == 0;
  if (cluster_reduce) {
    p.cluster_reduce = 1;
  } else if (two_pass && p.split_k > 1) {
    p.split_slices = 1;
    p.ep = avt_epilogue_t{};
    p.ep.alpha = 1.0f;
    p.ep.out = workspace;
    p.ep.ldo = N;
    p.ep.out_fp32 = 1;
  }
  p.tma_out = p.split_k > 1 ? 0 : (!p.ep.out_fp32 ? 1 : (p.ep.accumulate ? 0 : 2));
  p.tma_in = (p.ep.dact_z && !p.cluster_reduce) ? 1 : 0;

  GemmMaps tm;
  int rc;
  const uint32_t bnl = (uint32_t)(block_n / cta_group);
  if (!a_mn) rc = make_tmap_bf16_2d(&tm.a, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, kBK, kBM, 128);
  else       rc = make_tmap_bf16_2d(&tm.a, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, kBK, 128);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap_bf16_2d(&tm.b, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, kBK, bnl, 128);
  else       rc = make_tmap_bf16_2d(&tm.b, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, kBK, 128);
  if (rc) return rc;
  tm.out = tm.a;
  tm.aux = tm.a;
  tm.in = tm.a;
  if (p.tma_in && (rc = make_tmap_bf16_2d(&tm.in, p.ep.dact_z, (uint64_t)N, (uint64_t)M, (uint64_t)p.ep.ldz, 32, 32, 64))) return rc;
  if (p.tma_out) {
    if ((rc = make_tmap_2d(&tm.out, p.ep.out, (uint64_t)N, (uint64_t)M, (uint64_t)p.ep.ldo, 32, 32, p.tma_out == 2 ? 128 : 64,
                           p.tma_out == 2)))
      return rc;
    if (p.ep.aux_z && p.tma_out == 1 && (rc = make_tmap_bf16_2d(&tm.aux, p.ep.aux_z, (uint64_t)N, (uint64_t)M, (uint64_t)p.ep.ldz, 32, 32, 64)))
      return rc;
  }

  const avt_epilogue_t& e = p.ep;
  const bool simple = cta_group == 2 && block_n == 256 && !a_mn && p.split_k == 1 && p.tma_out == 1 && e.alpha == 1.0f &&
                      e.pos_period == 0 && e.drop_p == 0.f && !e.residual && !a_colsum && g_epi_special;
  int epi = kEpiGeneric;
  if (simple && !e.aux_z && !e.dact_z && e.act == AVT_ACT_NONE) epi = kEpiStore;
  else if (simple && e.aux_z && e.aux_mode == 1 && !e.dact_z && e.act == AVT_ACT_GELU_ERF) epi = kEpiGeluAux;
  else if (simple && !e.aux_z && e.dact_z && e.dact_mode == 1 && e.act == AVT_ACT_NONE && !e.bias) epi = kEpiMulZ;
  if (g_epi_special && cta_group == 2 && block_n == 256 && a_mn && b_mn && !two_pass && p.split_k > 1 && !p.split_slices && e.alpha == 1.0f &&
      !e.bias && !e.aux_z && !e.dact_z && e.act == AVT_ACT_NONE && e.drop_p == 0.f && !e.residual && e.pos_period == 0)
    epi = kEpiAtomic;
  // weight gradients over a short contraction (AVT-h: 80 rows): the kernel IS its epilogue (67 MB of output per launch), and
  // the generic body's ~600 instructions per 32-column chunk bounded it at 3.5 TB/s
  const bool plain_store = g_epi_special && cta_group == 2 && block_n == 256 && a_mn && b_mn && p.split_k == 1 && e.alpha == 1.0f &&
                           !e.bias && !e.aux_z && !e.dact_z && e.act == AVT_ACT_NONE && e.drop_p == 0.f && !e.residual &&
                           e.pos_period == 0 && !a_colsum;
  // branch-closing Linear with the residual add fused: fp32 residual in / fp32 out through 16-column TMA boxes
  const bool res_store = g_epi_special && cta_group == 2 && block_n == 256 && !a_mn && p.split_k == 1 && p.tma_out == 2 && e.residual &&
                         e.alpha == 1.0f && e.pos_period == 0 && e.drop_p == 0.f && !e.aux_z && !e.dact_z && e.act == AVT_ACT_NONE &&
                         !a_colsum && (reinterpret_cast<uintptr_t>(e.residual) & 15) == 0;
  if (res_store) {
    if ((rc = make_tmap_2d(&tm.in, e.residual, (uint64_t)N, (uint64_t)M, (uint64_t)e.ldr, 16, 32, 64, true))) return rc;
    if ((rc = make_tmap_2d(&tm.out, e.out, (uint64_t)N, (uint64_t)M, (uint64_t)e.ldo, 16, 32, 64, true))) return rc;
    rc = dispatch_epi<kEpiResidual>(b_mn, tm, p, s);
  } else if (plain_store && p.tma_out == 1) rc = launch_gemm<256, 2, true, true, kEpiStore>(tm, p, s);
  else if (plain_store && p.tma_out == 2) rc = launch_gemm<256, 2, true, true, kEpiStoreF32>(tm, p, s);
  else if (epi == kEpiAtomic) rc = launch_gemm<256, 2, true, true, kEpiAtomic>(tm, p, s);
  else if (epi == kEpiStore) rc = dispatch_epi<kEpiStore>(b_mn, tm, p, s);
  else if (epi == kEpiGeluAux) rc = dispatch_epi<kEpiGeluAux>(b_mn, tm, p, s);
  else if (epi == kEpiMulZ) rc = dispatch_epi<kEpiMulZ>(b_mn, tm, p, s);
  else if (cta_group == 2) {
    rc = block_n == 128 ? dispatch_major<128, 2>(a_mn, b_mn, tm, p, s) : dispatch_major<256, 2>(a_mn, b_mn, tm, p, s);
  } else {
    switch (block_n) {
      case 64: rc = dispatch_major<64, 1>(a_mn, b_mn, tm, p, s); break;
      case 128: rc = dispatch_major<128, 1>(a_mn, b_mn, tm, p, s); break;
      default: rc = dispatch_major<256, 1>(a_mn, b_mn, tm, p, s); break;
    }
  }
  if (rc) return rc;
  if (two_pass && p.split_k > 1 && !p.cluster_reduce) {
    const int64_t total4 = M * N / 4;
    int blocks = (int)((total4 + 255) / 256);
    if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
    launch_kernel(epilogue_apply_kernel, dim3(blocks), dim3(256), 0, s, reinterpret_cast<const float*>(workspace), p.split_k, (int)M, (int)N, finish);
    AVT_CUDA_OK(cudaGetLastError());
  }
  return AVT_OK;
}

extern "C" int avt_set_gemm_specialized_epilogues(int enable) {
  avt::g_epi_special = enable ? 1 : 0;
  return AVT_OK;
}
