// Generic multi-head attention on CUDA cores (any sequence length that fits smem, head_dim <= 1024,
// optional causal mask, optional dropout on the probabilities), forward and backward.
//
// This is the AVT-h path: GPT-2 attention over T <= 16 frame tokens with head_dim 256..1024, i.e.
// B*H problems of 10x10x512 — a 128-row tcgen05 tile would be >= 84 % padding, so the dot products run
// on CUDA cores with warp-shuffle reductions. It also serves as the on-GPU reference for
// the tcgen05 ViT attention kernel in tests.
//
// Layout (timm Attention / HF GPT2Attention packing): qkv bf16 [B*N, 3*Dm], column = s*Dm + h*hd + d.
// Replaces: timm Attention.forward core; HF GPT2Attention._attn (modeling_gpt2.py) incl. attn_dropout.
#include "common.cuh"
#include "ptx.cuh"

namespace avt {

constexpr int kAttWarps = 8;

struct AttnParams {
  const bf16* qkv;   // [B*N, 3*Dm]
  bf16* out;         // fwd: [B*N, Dm]
  float* lse;        // [B*H, N]   log-sum-exp of the scaled scores (natural log)
  const bf16* o;     // bwd: forward output [B*N, Dm] (delta_i = dO_i . O_i)
  const bf16* dout;  // bwd: [B*N, Dm]
  bf16* dqkv;        // bwd: [B*N, 3*Dm]
  int B, H, N, hd, Dm;
  int causal;
  int q_row;         // fwd: >= 0 -> only this query row of every (batch, head) is computed (KV-cached decode step); -1 = all rows
  float scale;
  float drop_p;
  uint64_t seed, offset;
  const uint64_t* offset_dev;   // optional device-resident addend of `offset` (CUDA-graph replays draw fresh masks)
};

__device__ __forceinline__ uint64_t att_offset(const AttnParams& p) {
  return p.offset + ((p.drop_p > 0.f && p.offset_dev) ? __ldg(p.offset_dev) : 0ull);
}

__device__ __forceinline__ bool att_keep(const AttnParams& p, uint64_t off, int bh, int i, int j) {
  if (p.drop_p <= 0.f) return true;
  const uint64_t idx = ((uint64_t)bh * p.N + i) * p.N + j;
  return (dropout_keep4(p.seed, off, idx >> 2, p.drop_p) >> (idx & 3)) & 1u;
}

// smem row pitch in elements: rows stay 16-byte aligned for vector staging (a warp reads one row at a time, 128
// contiguous bytes per instruction, so no padding is needed against bank conflicts)
__device__ __host__ __forceinline__ int att_pitch(int hd) { return hd + 8; }

template <int NP>
__device__ __forceinline__ void load_row(const bf16* __restrict__ row, int hd, int lane, float2 (&r)[NP]) {
#pragma unroll
  for (int m = 0; m < NP; ++m) {
    const int d = lane * 2 + 64 * m;
    if (d < hd) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(row + d);
      r[m] = __bfloat1622float2(v);
    } else {
      r[m] = make_float2(0.f, 0.f);
    }
  }
}
template <int NP>
__device__ __forceinline__ float dot_row(const bf16* __restrict__ srow, int hd, int lane, const float2 (&q)[NP]) {
  float acc = 0.f;
#pragma unroll
  for (int m = 0; m < NP; ++m) {
    const int d = lane * 2 + 64 * m;
    if (d < hd) {
      const float2 k = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(srow + d));
      acc += q[m].x * k.x + q[m].y * k.y;
    }
  }
  return warp_sum(acc);
}

__device__ __forceinline__ void stage_tile(const bf16* __restrict__ g, int64_t ld, int N, int hd, bf16* s) {
  // copy [N, hd] (row stride ld) into smem with pitch att_pitch(hd): 16-byte chunks when the layout allows it
  const int pitch = att_pitch(hd);
  if (hd % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int hv = hd / 8;
    for (int idx = threadIdx.x; idx < N * hv; idx += blockDim.x) {
      const int r = idx / hv, c = idx % hv;
      reinterpret_cast<uint4*>(s + r * pitch)[c] = __ldg(reinterpret_cast<const uint4*>(g + r * ld) + c);
    }
  } else {
    const int hw = hd / 2;
    for (int idx = threadIdx.x; idx < N * hw; idx += blockDim.x) {
      const int r = idx / hw, c = idx % hw;
      reinterpret_cast<uint32_t*>(s + r * pitch)[c] = reinterpret_cast<const uint32_t*>(g + r * ld)[c];
    }
  }
}

template <int NP>
__global__ void __launch_bounds__(kAttWarps * 32) attn_fwd_simt_kernel(const AttnParams p) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int N = p.N, hd = p.hd, pitch = att_pitch(hd);
  bf16* sK = reinterpret_cast<bf16*>(smem_att);
  bf16* sV = sK + (size_t)N * pitch;
  float* sS = reinterpret_cast<float*>(sV + (size_t)N * pitch);  // [kAttWarps][N]
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ld = 3 * (int64_t)p.Dm;
  const bf16* base = p.qkv + (int64_t)b * N * ld + h * hd;
  stage_tile(base + p.Dm, ld, N, hd, sK);
  stage_tile(base + 2 * p.Dm, ld, N, hd, sV);
  __syncthreads();
  float* s = sS + warp * N;
  const float keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const uint64_t rng_off = att_offset(p);
  // decode step: the packed qkv buffer IS the KV cache (rows 0..q_row of every batch item are valid, later rows are not
  // read: causal); only the newest row queries it
  const int i_begin = p.q_row >= 0 ? p.q_row + (int)(blockIdx.y * kAttWarps + warp) * N : blockIdx.y * kAttWarps + warp;
  const int i_step = p.q_row >= 0 ? N : gridDim.y * kAttWarps;
  for (int i = i_begin; i < N; i += i_step) {
    float2 q[NP];
    load_row<NP>(base + (int64_t)i * ld, hd, lane, q);
    const int jmax = p.causal ? i + 1 : N;
    float mx = -INFINITY;
    for (int j = 0; j < jmax; ++j) {
      const float v = dot_row<NP>(sK + j * pitch, hd, lane, q) * p.scale;
      if (lane == 0) s[j] = v;
      mx = fmaxf(mx, v);
    }
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j < jmax; j += 32) {
      const float e = __expf(s[j] - mx);
      sum += e;
      s[j] = e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    if (lane == 0 && p.lse) p.lse[(int64_t)bh * N + i] = mx + __logf(sum);
    __syncwarp();
    float2 o[NP];
#pragma unroll
    for (int m = 0; m < NP; ++m) o[m] = make_float2(0.f, 0.f);
    for (int j = 0; j < jmax; ++j) {
      float pj = s[j] * inv;
      if (p.drop_p > 0.f) pj = att_keep(p, rng_off, bh, i, j) ? pj * keep_scale : 0.f;
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        const int d = lane * 2 + 64 * m;
        if (d < hd) {
          const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sV + j * pitch + d));
          o[m].x += pj * v.x;
          o[m].y += pj * v.y;
        }
      }
    }
    bf16* orow = p.out + ((int64_t)b * N + i) * p.Dm + h * hd;
#pragma unroll
    for (int m = 0; m < NP; ++m) {
      const int d = lane * 2 + 64 * m;
      if (d < hd) *reinterpret_cast<__nv_bfloat162*>(orow + d) = __floats2bfloat162_rn(o[m].x, o[m].y);
    }
    __syncwarp();
  }
}

// Backward. gridDim.y = 2*ny: the first ny slices own query rows (dq), the last ny own key rows (dk, dv); the two
// halves are independent because delta_i = sum_j P_ij dP_ij = dO_i . O_i comes from the saved forward output, so
// 32 (batch, head) problems of AVT-h spread over ~128 CTAs instead of 32.
template <int NP>
__global__ void __launch_bounds__(kAttWarps * 32) attn_bwd_simt_kernel(const AttnParams p) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t smem_att[];
  const int N = p.N, hd = p.hd, pitch = att_pitch(hd);
  bf16* sA = reinterpret_cast<bf16*>(smem_att);   // query slices: K     key slices: Q
  bf16* sB = sA + (size_t)N * pitch;              // query slices: V     key slices: dO
  float* sDelta = reinterpret_cast<float*>(sB + (size_t)N * pitch);  // [N]
  float* sLse = sDelta + N;                                          // [N]
  float* sS = sLse + N;                                              // [kAttWarps][2][N]
  const int ny = gridDim.y >> 1;
  const bool key_slice = (int)blockIdx.y >= ny;
  const int by = key_slice ? blockIdx.y - ny : blockIdx.y;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ld = 3 * (int64_t)p.Dm;
  const bf16* base = p.qkv + (int64_t)b * N * ld + h * hd;
  const bf16* gbase = p.dout + (int64_t)b * N * p.Dm + h * hd;
  const bf16* obase = p.o + (int64_t)b * N * p.Dm + h * hd;
  const float keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const uint64_t rng_off = att_offset(p);
  bf16* dbase = p.dqkv + (int64_t)b * N * ld + h * hd;

  if (!key_slice) {
    // ---- one warp per query row -> dq_i (delta_i recomputed exactly from P and dP)
    stage_tile(base + p.Dm, ld, N, hd, sA);
    stage_tile(base + 2 * p.Dm, ld, N, hd, sB);
    __syncthreads();
    float* s = sS + warp * 2 * N;
    float* a = s + N;
    for (int i = by * kAttWarps + warp; i < N; i += ny * kAttWarps) {
      float2 q[NP], g[NP];
      load_row<NP>(base + (int64_t)i * ld, hd, lane, q);
      load_row<NP>(gbase + (int64_t)i * p.Dm, hd, lane, g);
      const int jmax = p.causal ? i + 1 : N;
      const float lse = p.lse[(int64_t)bh * N + i];
      float delta = 0.f;
      for (int j = 0; j < jmax; ++j) {
        const float sv = dot_row<NP>(sA + j * pitch, hd, lane, q) * p.scale;
        float av = dot_row<NP>(sB + j * pitch, hd, lane, g);
        if (p.drop_p > 0.f) av = att_keep(p, rng_off, bh, i, j) ? av * keep_scale : 0.f;
        const float pj = __expf(sv - lse);
        delta += pj * av;
        if (lane == 0) { s[j] = pj; a[j] = av; }
      }
      __syncwarp();
      float2 dq[NP];
#pragma unroll
      for (int m = 0; m < NP; ++m) dq[m] = make_float2(0.f, 0.f);
      for (int j = 0; j < jmax; ++j) {
        const float ds = s[j] * (a[j] - delta) * p.scale;
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          const int d = lane * 2 + 64 * m;
          if (d < hd) {
            const float2 k = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sA + j * pitch + d));
            dq[m].x += ds * k.x;
            dq[m].y += ds * k.y;
          }
        }
      }
      bf16* drow = dbase + (int64_t)i * ld;
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        const int d = lane * 2 + 64 * m;
        if (d < hd) *reinterpret_cast<__nv_bfloat162*>(drow + d) = __floats2bfloat162_rn(dq[m].x, dq[m].y);
      }
      __syncwarp();
    }
    return;
  }
  // ---- one warp per key row -> dk_j, dv_j (scores recomputed)
  stage_tile(base, ld, N, hd, sA);
  stage_tile(gbase, p.Dm, N, hd, sB);
  for (int i = threadIdx.x; i < N; i += blockDim.x) sLse[i] = p.lse[(int64_t)bh * N + i];
  __syncthreads();
  const int j_first = by * kAttWarps;   // rows below the first key of this slice never contribute under a causal mask
  for (int i = (p.causal ? j_first : 0) + warp; i < N; i += kAttWarps) {
    float2 o[NP];
    load_row<NP>(obase + (int64_t)i * p.Dm, hd, lane, o);
    const float dl = dot_row<NP>(sB + i * pitch, hd, lane, o);
    if (lane == 0) sDelta[i] = dl;
  }
  __syncthreads();
  for (int j = by * kAttWarps + warp; j < N; j += ny * kAttWarps) {
    float2 k[NP], v[NP], dk[NP], dv[NP];
    load_row<NP>(base + p.Dm + (int64_t)j * ld, hd, lane, k);
    load_row<NP>(base + 2 * p.Dm + (int64_t)j * ld, hd, lane, v);
#pragma unroll
    for (int m = 0; m < NP; ++m) dk[m] = dv[m] = make_float2(0.f, 0.f);
    for (int i = p.causal ? j : 0; i < N; ++i) {
      const float sv = dot_row<NP>(sA + i * pitch, hd, lane, k) * p.scale;
      float av = dot_row<NP>(sB + i * pitch, hd, lane, v);
      float keepf = 1.f;
      if (p.drop_p > 0.f) keepf = att_keep(p, rng_off, bh, i, j) ? keep_scale : 0.f;
      av *= keepf;
      const float pj = __expf(sv - sLse[i]);
      const float ds = pj * (av - sDelta[i]) * p.scale;
      const float pd = pj * keepf;
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        const int d = lane * 2 + 64 * m;
        if (d < hd) {
          const float2 qq = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sA + i * pitch + d));
          const float2 gg = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sB + i * pitch + d));
          dk[m].x += ds * qq.x; dk[m].y += ds * qq.y;
          dv[m].x += pd * gg.x; dv[m].y += pd * gg.y;
        }
      }
    }
    bf16* drow = dbase + (int64_t)j * ld;
#pragma unroll
    for (int m = 0; m < NP; ++m) {
      const int d = lane * 2 + 64 * m;
      if (d < hd) {
        *reinterpret_cast<__nv_bfloat162*>(drow + p.Dm + d) = __floats2bfloat162_rn(dk[m].x, dk[m].y);
        *reinterpret_cast<__nv_bfloat162*>(drow + 2 * p.Dm + d) = __floats2bfloat162_rn(dv[m].x, dv[m].y);
      }
    }
  }
}

template <int NP>
static int launch_simt(const AttnParams& p, bool bwd, cudaStream_t st) {
  const size_t tile = (size_t)p.N * att_pitch(p.hd) * sizeof(bf16);
  size_t smem;
  if (!bwd) smem = 2 * tile + (size_t)kAttWarps * p.N * sizeof(float);
  else smem = 2 * tile + 2 * (size_t)p.N * sizeof(float) + (size_t)kAttWarps * 2 * p.N * sizeof(float);
  if (smem > 227 * 1024) {
    set_last_error("attention_simt", "sequence x head_dim does not fit shared memory", __FILE__, __LINE__);
    return AVT_ERR_INVALID;
  }
  if (!bwd) {
    AVT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_simt_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int gy = 1;
    const int bh = p.B * p.H;
    while (bh * gy < 2 * num_sms() && gy * kAttWarps < p.N) gy *= 2;
    launch_kernel(attn_fwd_simt_kernel<NP>, dim3(dim3(bh, gy)), dim3(kAttWarps * 32), smem, st, p);
  } else {
    AVT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_simt_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int bh = p.B * p.H;
    int ny = (num_sms() + 2 * bh - 1) / (2 * bh);
    const int max_ny = (p.N + kAttWarps - 1) / kAttWarps;
    if (ny > max_ny) ny = max_ny;
    if (ny < 1) ny = 1;
    launch_kernel(attn_bwd_simt_kernel<NP>, dim3(dim3(bh, 2 * ny)), dim3(kAttWarps * 32), smem, st, p);
  }
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

static int dispatch_simt(const AttnParams& p, bool bwd, cudaStream_t st) {
  const int np = (p.hd + 63) / 64;
  if (np <= 1) return launch_simt<1>(p, bwd, st);
  if (np <= 2) return launch_simt<2>(p, bwd, st);
  if (np <= 4) return launch_simt<4>(p, bwd, st);
  if (np <= 8) return launch_simt<8>(p, bwd, st);
  return launch_simt<16>(p, bwd, st);
}

}  // namespace avt

using namespace avt;

static int check_attn(const AttnParams& p) {
  AVT_REQUIRE(p.B > 0 && p.H > 0 && p.N > 0, "empty problem");
  AVT_REQUIRE(p.hd % 2 == 0 && p.hd >= 2 && p.hd <= 1024, "head_dim must be even and <= 1024");
  AVT_REQUIRE(p.Dm == p.H * p.hd, "model dim must equal heads * head_dim");
  AVT_REQUIRE(p.drop_p >= 0.f && p.drop_p < 1.f, "drop_p must be in [0, 1)");
  return AVT_OK;
}

extern "C" int avt_attention_simt_decode(const void* qkv_cache, void* out, int B, int H, int N, int hd, int q_row, float scale,
                                         void* stream) {
  AVT_REQUIRE(qkv_cache && out, "null pointer");
  AVT_REQUIRE(q_row >= 0 && q_row < N, "q_row must be a row of the cache");
  AttnParams p{};
  p.qkv = reinterpret_cast<const bf16*>(qkv_cache); p.out = reinterpret_cast<bf16*>(out); p.lse = nullptr;
  p.B = B; p.H = H; p.N = N; p.hd = hd; p.Dm = H * hd; p.causal = 1; p.scale = scale; p.q_row = q_row;
  if (int rc = check_attn(p)) return rc;
  return dispatch_simt(p, false, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int avt_attention_simt_fwd(const void* qkv, void* out, float* lse, int B, int H, int N, int hd, int causal,
                                      float scale, float drop_p, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                                      void* stream) {
  AVT_REQUIRE(qkv && out, "null pointer");
  AttnParams p{};
  p.q_row = -1;
  p.qkv = reinterpret_cast<const bf16*>(qkv); p.out = reinterpret_cast<bf16*>(out); p.lse = lse;
  p.B = B; p.H = H; p.N = N; p.hd = hd; p.Dm = H * hd; p.causal = causal; p.scale = scale;
  p.drop_p = drop_p; p.seed = seed; p.offset = offset; p.offset_dev = offset_dev;
  if (int rc = check_attn(p)) return rc;
  return dispatch_simt(p, false, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int avt_attention_simt_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B,
                                      int H, int N, int hd, int causal, float scale, float drop_p, uint64_t seed,
                                      uint64_t offset, const uint64_t* offset_dev, void* stream) {
  AVT_REQUIRE(qkv && out && dout && lse && dqkv, "null pointer");
  AttnParams p{};
  p.q_row = -1;
  p.qkv = reinterpret_cast<const bf16*>(qkv); p.dout = reinterpret_cast<const bf16*>(dout);
  p.o = reinterpret_cast<const bf16*>(out);
  p.lse = const_cast<float*>(lse); p.dqkv = reinterpret_cast<bf16*>(dqkv);
  p.B = B; p.H = H; p.N = N; p.hd = hd; p.Dm = H * hd; p.causal = causal; p.scale = scale;
  p.drop_p = drop_p; p.seed = seed; p.offset = offset; p.offset_dev = offset_dev;
  if (int rc = check_attn(p)) return rc;
  return dispatch_simt(p, true, reinterpret_cast<cudaStream_t>(stream));
}
