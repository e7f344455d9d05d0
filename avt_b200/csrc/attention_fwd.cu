// ViT spatial attention forward on tcgen05: softmax(Q K^T * scale) V, N <= 208 tokens, head_dim 64.
// Replaces timm Attention.forward's q@k^T*scale -> softmax -> attn@v (models/video_classification.py:255-256 runs
// timm's VisionTransformer per frame).
//
// B200 probes (tools/ubench*.cu) and per-phase traces of the earlier versions (AVT_ATTN_TRACE) say what bounds this op:
// not the tensor pipe (S = Q K^T: 4 x 105 cycles per 128-query tile, P V: 13 x 48), not TMEM reads (~300 B/clk/SM), but
// MUFU: 16 ex2 / clk / SM = 1830 cycles per tile measured - IF the exponentials never wait for anything else. Earlier
// versions ran max pass -> exp pass -> P V -> O drain -> store of a tile in series on the same warps (MUFU busy 35 %),
// and their row-per-thread global stores (one L1 wavefront per row per instruction, 2500 cycles per tile) also starved
// the tensor core's shared-memory operand reads. This version is role-specialised; unit of work = one 128-query tile:
//   * 8 EXP warps (two threads per score row: keys 0..95 / 96..207) do nothing but the exp pass: chunked TMEM loads,
//     ex2, packed-bf16 P written back over the S columns (tcgen05.st) where it feeds the P V MMA as a TMEM A operand -
//     P never touches shared memory. Each thread overwrites only score columns it has already re-read: thread 0 walks
//     its chunks upwards and packs P into slot columns 0..47, thread 1 walks downwards and packs into 152..207 (the MMA
//     takes one A address per 16-key k-step, so P need not be contiguous);
//   * 4 AUX warps (one thread per row) run one tile AHEAD with the row-max pass, and one tile BEHIND with the O drain:
//     O / rowsum -> bf16 -> swizzled staging tile -> one TMA store per tile (3-D tensor map clips rows >= N);
//   * 2 issuing warps (one lane each, blocking waits on exactly the next event - a single thread polling several
//     barriers reacted in ~500 cycles): one issues the TMA loads and S of tile u+2 as soon as P V of tile u retired
//     (two TMEM slots); the other issues P V of tile u INCREMENTALLY, k-steps in four stages as the exp warps finish the
//     corresponding 16-key chunks, so it retires right behind the exp pass instead of adding its 13 MMAs to the slot's
//     turn-around time;
//   * the aux warps overwrite the score columns of keys >= N with -inf while they scan for the row max, so the exp pass
//     carries no masking code (it was half of that pass' instructions, and the kernel's 54 KB of SASS overflowed the
//     32 KB L1.5 instruction cache: 12 % stall_no_inst);
//   * Q/K and V of item n+2 are fetched as soon as item n's last S / last P V retired (3-D tensor maps
//     [frame][token][column]: rows past the frame's last token are zero-filled, never another frame's data).
// Warps: 0-7 exp (TMEM lane quarter = w & 3, key part = w >> 2), 8-11 aux (quarter = w & 3), 12 P V issue, 13 S issue + TMA.
// TMEM columns: slot t = tile parity at 208 t: S +0..+207, later P +0..+47 and +152..+207; O at 416..479.
#include <cuda.h>
#include "common.cuh"
#include "ptx.cuh"

namespace avt {

int make_tmap_bf16_3d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer1, uint64_t outer2, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer1, int swizzle_bytes);

namespace fwd3 {

#ifdef AVT_ATTN_TRACE   // per-phase timeline of CTA 0: every event id has its own shared-memory slot (one fire-and-forget store
                        // of the clock - an atomic slot counter cost ~250 cycles per event and distorted the picture)
#define FTRACE_DECL __shared__ long long ftrace_t[512];
#define FTRACE_INIT for (int i_ = threadIdx.x; i_ < 512; i_ += blockDim.x) ftrace_t[i_] = 0; __syncthreads();
#define FT(role, u, ph) FTRACE((role) * 96 + (u) * 16 + (ph))
#define FTRACE(id) do { if (blockIdx.x == 0 && (id) < 512) ftrace_t[(id)] = clock64(); } while (0)
#define FTRACE_DUMP if (blockIdx.x == 0 && threadIdx.x == 0) { long long t0_ = 0; for (int i_ = 0; i_ < 512; ++i_) if (ftrace_t[i_] && (!t0_ || ftrace_t[i_] < t0_)) t0_ = ftrace_t[i_]; for (int i_ = 0; i_ < 512; ++i_) if (ftrace_t[i_]) printf("ftrace %d %lld\n", i_, ftrace_t[i_] - t0_); }
#else
#define FTRACE_DECL
#define FTRACE_INIT
#define FTRACE(id)
#define FT(role, u, ph)
#define FTRACE_DUMP
#endif

constexpr int kHd = 64;
constexpr int kKeys = 208;               // MMA N extent of S / contraction length of P V (multiple of 16)
constexpr int kSplit = 96;               // keys 0..95 -> exp thread 0 of a row, 96..207 -> exp thread 1
constexpr int kP1Col = 152;              // packed P of keys 96..207 starts at this slot column
constexpr int kOCol = 2 * kKeys;         // O accumulator: TMEM columns 416..479
constexpr int kQBytes = 128 * 128;       // one Q tile: 128 rows x 64 bf16
constexpr int kKVBytes = kKeys * 128;    // K or V of one (frame, head)
constexpr int kItemBytes = 2 * kQBytes + 2 * kKVBytes;   // Q tile 0 | Q tile 1 | K | V
constexpr int kStageBytes = 128 * 128;   // one O tile, bf16 [128 x 64], SWIZZLE_128B, for the TMA store
constexpr int kExpWarps = 8, kAuxWarps = 4;
constexpr int kThreads = 32 * (kExpWarps + kAuxWarps + 2);
constexpr int kSmem = 1024 + 2 * kItemBytes + 2 * kStageBytes + (2 * 128 + 2 * 2 * 128) * 4 + 256;

struct Params {
  float* lse;
  int N, H, D, F;
  float scale;
};

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                const __grid_constant__ CUtensorMap tmO, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sStage = smem + 2 * kItemBytes;                             // [2][128 x 64 bf16]
  float* sMax = reinterpret_cast<float*>(sStage + 2 * kStageBytes);   // [slot][128] row max (raw scores)
  float* sSum = sMax + 2 * 128;                                        // [slot][key part][128] partial row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSum + 2 * 2 * 128);
  uint64_t* bar_qk = bars;          // [2] Q tiles + K of the item in smem slot i landed
  uint64_t* bar_v = bars + 2;       // [2] V landed
  uint64_t* bar_s = bars + 4;       // [2] S in TMEM slot t computed                       (MMA -> aux, control)
  uint64_t* bar_max = bars + 6;     // [2] row max of slot t in smem, 128 arrivals         (aux -> exp)
  uint64_t* bar_pst = bars + 8;     // [2][4] P of slot t, stage s written (k-steps 2s,2s+1 | 12-2s,11-2s; stage 3: k-step 6)
  uint64_t* bar_o = bars + 16;      // O of a tile computed, one phase per tile            (MMA -> aux, control)
  uint64_t* bar_ofree = bars + 17;  // O drained from TMEM, 128 arrivals                   (aux -> control)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N;
  const int items = p.F * p.H;
  FTRACE_DECL
  FTRACE_INIT
  const int my_items = (int)blockIdx.x < items ? (items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int tiles = 2 * my_items;

  if (warp == kExpWarps + kAuxWarps) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      tma_prefetch_desc(&tmO);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bar_qk[i], 1);
        mbar_init(&bar_v[i], 1);
        mbar_init(&bar_s[i], 4);                     // one tcgen05.commit per issuing lane
        mbar_init(&bar_max[i], 32 * kAuxWarps);
        for (int st = 0; st < 4; ++st) mbar_init(&bar_pst[i * 4 + st], st < 3 ? 32 * kExpWarps : 16 * kExpWarps);
      }
      mbar_init(bar_o, 4);
      mbar_init(bar_ofree, 32 * kAuxWarps);
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
  pdl_wait();   // the prologue above overlapped the previous kernel's tail
  pdl_trigger();

  if (warp >= kExpWarps + kAuxWarps) {
    // ------------------------------------------------------------- two issuing warps (lane 0 each), blocking waits only
    constexpr uint32_t idesc_s = umma_idesc(1, 0, 0, 128, kKeys);
    constexpr uint32_t idesc_o = umma_idesc(1, 0, 1, 128, kHd);
    constexpr uint64_t desc_k = smem_desc_sw128(16, 1024);     // K-major
    constexpr uint64_t desc_v = smem_desc_sw128(8192, 1024);   // MN-major, one 64-wide block
    // One thread computing 13 descriptors and issuing 13 MMAs is a ~40-instruction dependent chain per MMA (measured: 365
    // cycles per MMA for a rolled loop, 105-130 unrolled and falling out of the instruction cache). Here each k-step belongs
    // to its own LANE: the lanes compute their descriptors in parallel and execute the same tcgen05.mma instruction, which
    // the hardware issues once per active lane; every issuing lane commits its own MMAs (barrier counts = 4 lanes). Only
    // the accumulator-initialising MMA (k-step 0 of a tile) is issued on its own, ahead of the others.
    if (warp == kExpWarps + kAuxWarps) {
      // ---- P V issuer: O (+)= P (TMEM, slot u & 1) V, the k-steps of a stage as soon as the exp warps completed it:
      //      stage st -> k-steps 2st, 2st+1 (thread 0 of a row, upwards), 12-2st, 11-2st (thread 1, downwards); stage 3 -> 6
      for (int u = 0; u < tiles; ++u) {
        const int n = u >> 1, t = u & 1;
        const uint32_t aV = smem_u32(smem + (n & 1) * kItemBytes + 2 * kQBytes + kKVBytes);
        const uint32_t aP = tmem_base + kKeys * t, aO = tmem_base + kOCol;
        mbar_wait_ool(&bar_v[n & 1], (n >> 1) & 1);              // V landed
        if (u > 0) mbar_wait_ool(bar_ofree, (u - 1) & 1);        // O of the previous tile was drained
#pragma unroll 1
        for (int st = 0; st < 4; ++st) {
          mbar_wait_ool(&bar_pst[t * 4 + st], n & 1);
          tc_fence_after_sync();
          if (lane == 0 && u < 6) FT(3, u, st);
          const int j = lane & 3;
          const int k = st == 3 ? 6 : (j < 2 ? 2 * st + j : 14 - 2 * st - j);
          const uint32_t a_addr = aP + (k < kSplit / 16 ? 8 * k : kP1Col + 8 * (k - kSplit / 16));
          const uint64_t b_desc = smem_desc_addr(desc_v, aV + k * 2048);
#ifndef ABL_NOPV
          if (st == 0) {
            if (lane == 0) umma_f16_ts(aO, a_addr, b_desc, idesc_o, 0u);
            __syncwarp();
            if (lane >= 1 && lane < 4) umma_f16_ts(aO, a_addr, b_desc, idesc_o, 1u);
          } else if (lane < (st == 3 ? 1 : 4)) {
            umma_f16_ts(aO, a_addr, b_desc, idesc_o, 1u);
          }
#endif
          if ((st == 3 && lane == 0) || (st == 2 && lane >= 1 && lane < 4)) umma_commit(bar_o);   // each lane's last MMA
          __syncwarp();
          if (lane == 0 && u < 6) FT(3, u, 4 + st);
        }
      }
    } else if (my_items > 0) {
      // ---- S issuer + TMA loads
      auto item_fh = [&](int n, int& f, int& h) {
        const int item = blockIdx.x + n * gridDim.x;
        f = item / p.H;
        h = item % p.H;
      };
      auto load_qk = [&](int n) {   // n-th item of this CTA -> smem slot n & 1
        int f, h;
        item_fh(n, f, h);
        uint8_t* s = smem + (n & 1) * kItemBytes;
        if (lane == 0) {
          mbar_arrive_expect_tx(&bar_qk[n & 1], 2 * kQBytes + kKVBytes);
          tma_load_3d(&tmQ, &bar_qk[n & 1], s, h * kHd, 0, f);
          tma_load_3d(&tmKV, &bar_qk[n & 1], s + 2 * kQBytes, p.D + h * kHd, 0, f);
          tma_load_3d(&tmQ, &bar_qk[n & 1], s + kQBytes, h * kHd, 128, f);
        }
      };
      auto load_v = [&](int n) {
        int f, h;
        item_fh(n, f, h);
        if (lane == 0) {
          mbar_arrive_expect_tx(&bar_v[n & 1], kKVBytes);
          tma_load_3d(&tmKV, &bar_v[n & 1], smem + (n & 1) * kItemBytes + 2 * kQBytes + kKVBytes, 2 * p.D + h * kHd, 0, f);
        }
      };
      load_qk(0);
      load_v(0);
      if (my_items > 1) {
        load_qk(1);
        load_v(1);
      }
      for (int u = 0; u < tiles; ++u) {
        const int n = u >> 1, t = u & 1;
        if (u >= 2) {
          mbar_wait_ool(bar_o, (u - 2) & 1);                          // P V of tile u - 2 retired: TMEM slot t is free
          if (t == 1 && n + 1 < my_items) load_v(n + 1);          // ... and with it the V of item n - 1
        }
        if (t == 0) mbar_wait_ool(&bar_qk[n & 1], (n >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t aQ = smem_u32(smem + (n & 1) * kItemBytes + t * kQBytes);
        const uint32_t aK = smem_u32(smem + (n & 1) * kItemBytes + 2 * kQBytes);
        if (lane == 0 && u < 6) FT(4, u, 0);
        {   // k-step = lane (4 lanes); the accumulator-initialising one first
          const uint64_t a_desc = smem_desc_addr(desc_k, aQ + (lane & 3) * 32), b_desc = smem_desc_addr(desc_k, aK + (lane & 3) * 32);
#ifndef ABL_NOS
          if (lane == 0) umma_f16(tmem_base + kKeys * t, a_desc, b_desc, idesc_s, 0u);
          __syncwarp();
          if (lane >= 1 && lane < 4) umma_f16(tmem_base + kKeys * t, a_desc, b_desc, idesc_s, 1u);
#endif
          if (lane < 4) umma_commit(&bar_s[t]);
          __syncwarp();
        }
        if (lane == 0 && u < 6) FT(4, u, 1);
        if (t == 1 && n + 2 < my_items) {                         // both S of item n retired: its Q/K slot takes item n + 2
          mbar_wait_ool(&bar_s[0], n & 1);
          mbar_wait_ool(&bar_s[1], n & 1);
          load_qk(n + 2);
        }
      }
    }
    __syncwarp();
  } else if (warp >= kExpWarps) {
    // ------------------------------------------------------------- aux warps: row max one tile ahead, O drain one behind
    const int q = warp & 3;
    const int r = q * 32 + lane;                                   // row within the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
    const bool elected = warp == kExpWarps && lane == 0;
    const int tr_base = threadIdx.x == 32 * kExpWarps ? 1 : -1;
    (void)tr_base;
    float prev_mx = 0.f;

    auto drain = [&](int v, float mx) {
      const int n = v >> 1, t = v & 1;
      const int item = blockIdx.x + n * gridDim.x;
      const int f = item / p.H, h = item % p.H;
      mbar_wait_ool(&bar_pst[t * 4 + 2], n & 1);   // (acquire: the exp warps' partial row sums are visible)
      mbar_wait_ool(&bar_pst[t * 4 + 3], n & 1);
      mbar_wait_ool(bar_o, v & 1);
      if (tr_base >= 0 && v < 6) FT(1, v, 5);
      tc_fence_after_sync();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32b_x32(t_lane + kOCol, o0);
      tmem_ld_32x32b_x32(t_lane + kOCol + 32, o1);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(bar_ofree);
#ifdef ABL_NODRAIN
      if (tiles >= 0) return;
#endif
      const float sum = sSum[(t * 2 + 0) * 128 + r] + sSum[(t * 2 + 1) * 128 + r];
      const float inv = 1.0f / sum;       // (rows >= N: garbage in, clipped on the way out)
      uint8_t* srow = sStage + (v & 1) * kStageBytes + r * 128;
#pragma unroll
      for (int gq = 0; gq < 4; ++gq) {
        *reinterpret_cast<uint4*>(srow + ((gq ^ (r & 7)) << 4)) = make_uint4(
            pack_bf16x2(__uint_as_float(o0[8 * gq]) * inv, __uint_as_float(o0[8 * gq + 1]) * inv),
            pack_bf16x2(__uint_as_float(o0[8 * gq + 2]) * inv, __uint_as_float(o0[8 * gq + 3]) * inv),
            pack_bf16x2(__uint_as_float(o0[8 * gq + 4]) * inv, __uint_as_float(o0[8 * gq + 5]) * inv),
            pack_bf16x2(__uint_as_float(o0[8 * gq + 6]) * inv, __uint_as_float(o0[8 * gq + 7]) * inv));
        *reinterpret_cast<uint4*>(srow + (((4 + gq) ^ (r & 7)) << 4)) = make_uint4(
            pack_bf16x2(__uint_as_float(o1[8 * gq]) * inv, __uint_as_float(o1[8 * gq + 1]) * inv),
            pack_bf16x2(__uint_as_float(o1[8 * gq + 2]) * inv, __uint_as_float(o1[8 * gq + 3]) * inv),
            pack_bf16x2(__uint_as_float(o1[8 * gq + 4]) * inv, __uint_as_float(o1[8 * gq + 5]) * inv),
            pack_bf16x2(__uint_as_float(o1[8 * gq + 6]) * inv, __uint_as_float(o1[8 * gq + 7]) * inv));
      }
      const int qrow = t * 128 + r;
      if (qrow < N && p.lse) p.lse[((size_t)f * p.H + h) * N + qrow] = mx * p.scale + __logf(sum);
      fence_proxy_async_smem();
      // the thread that issued the previous tile's store makes sure it finished READING its staging tile before it joins
      // this barrier: after the barrier everybody may overwrite that tile (at the next drain)
      if (elected) tma_store_wait_read<0>();
      named_bar_sync(1, 32 * kAuxWarps);
      if (elected) {
        tma_store_3d(&tmO, sStage + (v & 1) * kStageBytes, h * kHd, t * 128, f);
        tma_store_commit();
      }
      if (tr_base >= 0 && v < 6) FT(1, v, 6);
    };

    for (int u = 0; u < tiles; ++u) {
      const int n = u >> 1, t = u & 1;
      const bool active = t * 128 + q * 32 < N;                    // warp-uniform: this warp owns at least one real query
      mbar_wait_ool(&bar_s[t], n & 1);
      if (tr_base >= 0 && u < 6) FT(1, u, 0);
      tc_fence_after_sync();
      float mx = -INFINITY;
      if (active) {
        // Columns of keys >= N (zero-filled K rows: S = 0 there) are first overwritten with -inf: neither this max pass nor
        // the exp pass then needs any masking (2^(-inf) = 0). [N, 208) is covered by one 1-, 2-, 4-column store + 8-column ones.
        const uint32_t s_addr = t_lane + kKeys * t;
        if (N < kKeys) {
          constexpr uint32_t kNegInf = 0xff800000u;
          const uint32_t ninf4[4] = {kNegInf, kNegInf, kNegInf, kNegInf};
          const uint32_t ninf8[8] = {kNegInf, kNegInf, kNegInf, kNegInf, kNegInf, kNegInf, kNegInf, kNegInf};
          int c = N;
          if (c & 1) { tmem_st_32x32b_x1(s_addr + c, kNegInf); c += 1; }
          if (c & 2) { tmem_st_32x32b_x2(s_addr + c, kNegInf, kNegInf); c += 2; }
          if (c & 4) { tmem_st_32x32b_x4(s_addr + c, ninf4); c += 4; }
#pragma unroll 1
          for (; c < kKeys; c += 8) tmem_st_32x32b_x8(s_addr + c, ninf8);
          tmem_st_wait();
        }
        // row max, two batches of TMEM loads
        auto cmax32 = [&](const uint32_t (&v)[32]) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        };
#ifndef ABL_NOMAX
        uint32_t va[32], vb[32], vc[32], vd[16];
        tmem_ld_32x32b_x32(s_addr, va);
        tmem_ld_32x32b_x32(s_addr + 32, vb);
        tmem_ld_32x32b_x32(s_addr + 64, vc);
        tmem_ld_wait();
        cmax32(va);
        cmax32(vb);
        cmax32(vc);
        tmem_ld_32x32b_x32(s_addr + 96, va);
        tmem_ld_32x32b_x32(s_addr + 128, vb);
        tmem_ld_32x32b_x32(s_addr + 160, vc);
        tmem_ld_32x32b_x16(s_addr + 192, vd);
        tmem_ld_wait();
        cmax32(va);
        cmax32(vb);
        cmax32(vc);
#pragma unroll
        for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(vd[j]));
#else
        mx = 0.f;
#endif
        sMax[t * 128 + r] = mx;
      }
      tc_fence_before_sync();          // our TMEM reads of S are ordered before the exp warps' P stores over those columns
      mbar_arrive(&bar_max[t]);
      if (tr_base >= 0 && u < 6) FT(1, u, 2);
      if (u > 0) drain(u - 1, prev_mx);
      prev_mx = mx;
    }
    if (tiles > 0) drain(tiles - 1, prev_mx);
    if (elected) tma_store_wait<0>();
  } else {
    // ------------------------------------------------------------- exp warps
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    const int tr_base = threadIdx.x == 0 ? 0 : (threadIdx.x == 128 ? 2 : -1);
    (void)tr_base;
    for (int u = 0; u < tiles; ++u) {
      const int n = u >> 1, t = u & 1;
      const bool active = t * 128 + q * 32 < N;
      const uint32_t t_slot = t_lane + kKeys * t;
      mbar_wait_ool(&bar_max[t], n & 1);
      if (tr_base >= 0 && u < 6) FT(1, u, 2);
      tc_fence_after_sync();
      uint64_t* stage = &bar_pst[t * 4];
      if (active) {
        const float mx = sMax[t * 128 + r];
        // e = 2^(scale*log2e*(s - max)) (invalid key columns hold -inf -> 0); partial row sum; P -> TMEM as packed bf16 over
        // columns this thread has re-read. A stage is announced one chunk late: by then its TMEM stores have long
        // completed, so tcgen05.wait::st never stalls the exponentials.
        const float2 sl2v = f2(sl2), nm = f2(-mx * sl2);
        float2 acc0 = f2(0.f), acc1 = f2(0.f);
        // One rolled loop over 16-key chunks (= P V k-steps), the SAME instructions for both key parts (thread 0: chunks
        // 0..5 upwards, thread 1: 12..6 downwards), ping-pong TMEM loads: the hot loop is ~100 instructions and stays in the
        // L0 instruction cache of a sub-partition whose four warps otherwise run four different code paths.
        auto compute16 = [&](const uint32_t (&v)[16], uint32_t (&pk)[8]) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 a = __ffma2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), sl2v, nm);
#ifdef ABL_NOEXP
            const float2 e = a;
#else
            const float2 e = make_float2(fast_ex2(a.x), fast_ex2(a.y));
#endif
            if (j & 1) acc1 = __fadd2_rn(acc1, e);
            else acc0 = __fadd2_rn(acc0, e);
            pk[j] = pack_bf16x2(e.x, e.y);
          }
        };
        auto announce = [&](int st) {
          tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive(&stage[st]);
        };
        const int nch = hf ? 7 : 6;
        const int scol0 = hf ? 16 * 12 : 0, sstep = hf ? -16 : 16;            // score columns of chunk i: scol0 + i * sstep
        const int pcol0 = hf ? kP1Col + 8 * 6 : 0, pstep = hf ? -8 : 8;      // packed P columns of chunk i
        uint32_t va[16], vb[16], pk[8];
        tmem_ld_32x32b_x16(t_slot + scol0, va);
        tmem_ld_wait();
#pragma unroll 1
        for (int i = 0; i < nch; i += 2) {
          const bool two = i + 1 < nch;
          if (two) tmem_ld_32x32b_x16(t_slot + scol0 + (i + 1) * sstep, vb);
          compute16(va, pk);
          if (i > 0) announce((i >> 1) - 1);      // the previous pair's P stores completed long ago: no stall
          tmem_st_32x32b_x8(t_slot + pcol0 + i * pstep, pk);
          if (two) {
            tmem_ld_wait();
            if (i + 2 < nch) tmem_ld_32x32b_x16(t_slot + scol0 + (i + 2) * sstep, va);
            compute16(vb, pk);
            tmem_st_32x32b_x8(t_slot + pcol0 + (i + 1) * pstep, pk);
            if (i + 2 < nch) tmem_ld_wait();
          }
        }
        const float2 acc = __fadd2_rn(acc0, acc1);
        sSum[(t * 2 + hf) * 128 + r] = acc.x + acc.y;
        announce(hf ? 3 : 2);
      } else {
        tc_fence_before_sync();
        for (int st = 0; st < (hf ? 4 : 3); ++st) mbar_arrive(&stage[st]);
      }
      if (tr_base >= 0 && u < 6) FT(tr_base, u, 3);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  FTRACE_DUMP
  if (warp == kExpWarps + kAuxWarps) tmem_dealloc(tmem_base, 512);
}

}  // namespace fwd3
}  // namespace avt

using namespace avt;

extern "C" int avt_attention_tc_fwd(const void* qkv, void* out, float* lse, int F, int H, int N, float scale, void* stream) {
  AVT_REQUIRE(qkv && out, "null pointer");
  AVT_REQUIRE(F > 0 && H > 0 && N > 0 && N <= fwd3::kKeys, "tokens per frame must be in [1, 208]");
  const int D = H * fwd3::kHd;
  CUtensorMap tmQ, tmKV;
  if (int rc = make_tmap_bf16_3d(&tmQ, qkv, 3ull * D, (uint64_t)N, (uint64_t)F, 3ull * D, 64, 128, 128)) return rc;
  if (int rc = make_tmap_bf16_3d(&tmKV, qkv, 3ull * D, (uint64_t)N, (uint64_t)F, 3ull * D, 64, fwd3::kKeys, 128)) return rc;
  CUtensorMap tmO;
  if (int rc = make_tmap_bf16_3d(&tmO, out, (uint64_t)D, (uint64_t)N, (uint64_t)F, (uint64_t)D, 64, 128, 128)) return rc;
  static bool configured = false;
  if (!configured) {
    AVT_CUDA_OK(cudaFuncSetAttribute(fwd3::attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd3::kSmem));
    configured = true;
  }
  fwd3::Params p;
  p.lse = lse; p.N = N; p.H = H; p.D = D; p.F = F; p.scale = scale;
  const int grid = F * H < num_sms() ? F * H : num_sms();
  AVT_CUDA_OK(launch_kernel(fwd3::attn_fwd_kernel, dim3(grid), dim3(fwd3::kThreads), fwd3::kSmem,
                            reinterpret_cast<cudaStream_t>(stream), tmQ, tmKV, tmO, p));
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}
