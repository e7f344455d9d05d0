// LayerNorm forward / backward: one warp per row, the row lives in registers (single HBM read),
// warp-shuffle reductions, 128-bit loads/stores. HBM-bound by design:
//   fwd  bytes/row = 4*D (x)  + 2*D (y bf16) [+8 stats]
//   bwd  bytes/row = 2*D (dy) + 4*D (x) + 4*D (dx in) + 4*D (dx out) + 2*D (dx bf16)
// Replaces torch.nn.LayerNorm in timm Block.norm1/norm2/VisionTransformer.norm (eps 1e-6) and
// HF GPT2Block.ln_1/ln_2/GPT2Model.ln_f (eps 1e-5), and their autograd backward.
#include "common.cuh"
#include "ptx.cuh"

namespace avt {

constexpr int kLnWarps = 8;

// VPT = float4 vectors per lane; covers D <= VPT*128 (columns >= D are masked).
template <int VPT>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_fwd_kernel(const float* __restrict__ x, int64_t x_stride, const bf16* __restrict__ add, int64_t add_stride,
              float* __restrict__ x_out, int64_t xo_stride, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, int64_t rows, int D, void* __restrict__ y, int y_fp32,
              int64_t y_stride, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * kLnWarps;
  const float inv_d = 1.0f / (float)D;
  for (int64_t r = warp_global; r < rows; r += warp_stride) {
    const float* xr = x + r * x_stride;
    float4 v[VPT];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * 32 + lane) * 4;
      v[i] = c < D ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (add && c < D) {  // residual stream update fused in: x_new = x + branch (bf16), x_new is what gets normalised
        const uint2 a = *reinterpret_cast<const uint2*>(add + r * add_stride + c);
        v[i].x += bf16_lo(a.x); v[i].y += bf16_hi(a.x); v[i].z += bf16_lo(a.y); v[i].w += bf16_hi(a.y);
      }
      if (x_out && c < D) *reinterpret_cast<float4*>(x_out + r * xo_stride + c) = v[i];   // x + add (or a copy of x)
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        q += a * a + b * b + cc * cc + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
        const float o0 = (v[i].x - mean) * rstd * g.x + b.x, o1 = (v[i].y - mean) * rstd * g.y + b.y;
        const float o2 = (v[i].z - mean) * rstd * g.z + b.z, o3 = (v[i].w - mean) * rstd * g.w + b.w;
        if (y_fp32) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + r * y_stride + c) = make_float4(o0, o1, o2, o3);
        } else {
          *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + r * y_stride + c) =
              make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
        }
      }
    }
  }
}

// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)) [+ dx_in];  partial dgamma/dbeta per CTA.
template <int VPT>
__global__ void __launch_bounds__(kLnWarps * 32, VPT <= 8 ? 2 : 1)
ln_bwd_kernel(const void* __restrict__ dy, int dy_fp32, int64_t dy_stride, const float* __restrict__ x, int64_t x_stride,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              int64_t rows, int D, const float* __restrict__ dx_in, float* __restrict__ dx_out, int64_t dx_stride,
              bf16* __restrict__ dx_bf16, int64_t dxb_stride, float* __restrict__ partial /*[grid][2][D]*/) {
  pdl_enter();
  __shared__ float red[kLnWarps][32 * 4 + 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * kLnWarps + warp;
  const int64_t warp_stride = (int64_t)gridDim.x * kLnWarps;
  const float inv_d = 1.0f / (float)D;
  float4 ag[VPT], ab[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    ag[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t r = warp_global; r < rows; r += warp_stride) {
    const float mu = mean[r], rs = rstd[r];
    float4 xh[VPT], d[VPT];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        const float4 xv = *reinterpret_cast<const float4*>(x + r * x_stride + c);
        if (dy_fp32) {
          d[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + r * dy_stride + c);
        } else {
          const uint2 p = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dy) + r * dy_stride + c);
          d[i] = make_float4(bf16_lo(p.x), bf16_hi(p.x), bf16_lo(p.y), bf16_hi(p.y));
        }
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        ab[i].x += d[i].x; ab[i].y += d[i].y; ab[i].z += d[i].z; ab[i].w += d[i].w;
        ag[i].x += d[i].x * xh[i].x; ag[i].y += d[i].y * xh[i].y; ag[i].z += d[i].z * xh[i].z; ag[i].w += d[i].w * xh[i].w;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));  // L1-resident; frees 4*VPT registers
        d[i].x *= g.x; d[i].y *= g.y; d[i].z *= g.z; d[i].w *= g.w;
        s1 += d[i].x + d[i].y + d[i].z + d[i].w;
        s2 += d[i].x * xh[i].x + d[i].y * xh[i].y + d[i].z * xh[i].z + d[i].w * xh[i].w;
      } else {
        xh[i] = d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float m1 = warp_sum(s1) * inv_d, m2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        float4 o = make_float4(rs * (d[i].x - m1 - xh[i].x * m2), rs * (d[i].y - m1 - xh[i].y * m2),
                               rs * (d[i].z - m1 - xh[i].z * m2), rs * (d[i].w - m1 - xh[i].w * m2));
        if (dx_in) {
          const float4 a = *reinterpret_cast<const float4*>(dx_in + r * dx_stride + c);
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        *reinterpret_cast<float4*>(dx_out + r * dx_stride + c) = o;
        if (dx_bf16)
          *reinterpret_cast<uint2*>(dx_bf16 + r * dxb_stride + c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      }
    }
  }
  // cross-warp reduction of the per-lane column sums, one float4 slot at a time
  float* pg = partial + (size_t)blockIdx.x * 2 * D;
  float* pb = pg + D;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    for (int which = 0; which < 2; ++which) {
      const float4 a = which == 0 ? ag[i] : ab[i];
      __syncthreads();
      red[warp][lane * 4 + 0] = a.x; red[warp][lane * 4 + 1] = a.y; red[warp][lane * 4 + 2] = a.z; red[warp][lane * 4 + 3] = a.w;
      __syncthreads();
      if (threadIdx.x < 128) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kLnWarps; ++w) s += red[w][threadIdx.x];
        const int c = i * 128 + threadIdx.x;
        if (c < D) (which == 0 ? pg : pb)[c] = s;
      }
    }
  }
}


// ----------------------------------------------------------------------------- bulk-async pipelined backward
// The register-resident kernel above is latency-bound at ViT sizes (ncu: 30 % of HBM peak, 16 warps/SM at 128
// registers, the dx_in reads serialised behind the row reductions). Here every warp streams its rows through a
// private shared-memory ring filled by 1-D bulk async copies (cp.async.bulk, mbarrier completion): the next row's
// x / dy / dx_in (7.7 KB at D = 768) is always in flight while the current one is reduced, independent of
// occupancy, and the per-lane column accumulators (dgamma, dbeta and the column sums of dx_out = the bias
// gradient of the Linear that produced this residual branch) stay in registers.
constexpr int kLnPipeWarps = 8;
constexpr int kLnPipeStages = 2;

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int VPT>
__global__ void __launch_bounds__(kLnPipeWarps * 32, 1)
ln_bwd_pipe_kernel(const void* __restrict__ dy, int dy_fp32, int64_t dy_stride, const float* __restrict__ x,
                   int64_t x_stride, const float* __restrict__ mean, const float* __restrict__ rstd,
                   const float* __restrict__ gamma, int64_t rows, int D, const float* dx_in, float* dx_out,
                   int64_t dx_stride, bf16* __restrict__ dx_bf16, int64_t dxb_stride, int ncols_out /*2 or 3*/,
                   float* __restrict__ partial /*[grid][ncols_out][D]*/, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   float* __restrict__ dcol, int atomic_out) {
  pdl_enter();
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t x_bytes = (uint32_t)D * 4u, dy_bytes = (uint32_t)D * (dy_fp32 ? 4u : 2u);
  const uint32_t in_bytes = dx_in ? x_bytes : 0u;
  const uint32_t stage_bytes = x_bytes + in_bytes + dy_bytes;   // [x | dx_in | dy], every part a multiple of 16 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem) + warp * kLnPipeStages;   // first 128 bytes: mbarriers
  uint8_t* data = ln_smem + 128;
  uint8_t* wbase = data + (size_t)warp * kLnPipeStages * stage_bytes;
  if (lane == 0) {
    for (int s = 0; s < kLnPipeStages; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int64_t warp_global = (int64_t)blockIdx.x * kLnPipeWarps + warp;
  const int64_t warp_stride = (int64_t)gridDim.x * kLnPipeWarps;
  const float inv_d = 1.0f / (float)D;

  auto issue = [&](int64_t r, int s) {   // lane 0 only
    uint8_t* dst = wbase + (size_t)s * stage_bytes;
    mbar_arrive_expect_tx(&bars[s], stage_bytes);
    bulk_load_1d(dst, x + r * x_stride, x_bytes, &bars[s]);
    if (dx_in) bulk_load_1d(dst + x_bytes, dx_in + r * dx_stride, x_bytes, &bars[s]);
    bulk_load_1d(dst + x_bytes + in_bytes,
                 dy_fp32 ? reinterpret_cast<const void*>(reinterpret_cast<const float*>(dy) + r * dy_stride)
                         : reinterpret_cast<const void*>(reinterpret_cast<const bf16*>(dy) + r * dy_stride),
                 dy_bytes, &bars[s]);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kLnPipeStages; ++s) {
      const int64_t r = warp_global + s * warp_stride;
      if (r < rows) issue(r, s);
    }
  }
  float4 gm[VPT], ag[VPT], ab[VPT], ao[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * 32 + lane) * 4;
    gm[i] = c < D ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    ag[i] = ab[i] = ao[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int stage = 0;
  uint32_t phase = 0;
  float mu = 0.f, rs = 0.f;
  if (warp_global < rows) { mu = mean[warp_global]; rs = rstd[warp_global]; }
  for (int64_t r = warp_global; r < rows; r += warp_stride) {
    const int64_t rn = r + warp_stride;
    float mu_n = 0.f, rs_n = 0.f;
    if (rn < rows) { mu_n = mean[rn]; rs_n = rstd[rn]; }   // next row's statistics: off the critical path
    mbar_wait(&bars[stage], phase);
    const uint8_t* sb = wbase + (size_t)stage * stage_bytes;
    const float* sx = reinterpret_cast<const float*>(sb);
    const float* sin = reinterpret_cast<const float*>(sb + x_bytes);
    const uint8_t* sdy = sb + x_bytes + in_bytes;
    float4 xh[VPT], d[VPT];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        const float4 xv = *reinterpret_cast<const float4*>(sx + c);
        if (dy_fp32) {
          d[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sdy) + c);
        } else {
          const uint2 p = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(sdy) + c);
          d[i] = make_float4(bf16_lo(p.x), bf16_hi(p.x), bf16_lo(p.y), bf16_hi(p.y));
        }
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        ab[i].x += d[i].x; ab[i].y += d[i].y; ab[i].z += d[i].z; ab[i].w += d[i].w;
        ag[i].x += d[i].x * xh[i].x; ag[i].y += d[i].y * xh[i].y; ag[i].z += d[i].z * xh[i].z; ag[i].w += d[i].w * xh[i].w;
        d[i].x *= gm[i].x; d[i].y *= gm[i].y; d[i].z *= gm[i].z; d[i].w *= gm[i].w;
        s1 += d[i].x + d[i].y + d[i].z + d[i].w;
        s2 += d[i].x * xh[i].x + d[i].y * xh[i].y + d[i].z * xh[i].z + d[i].w * xh[i].w;
      } else {
        xh[i] = d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float m1 = warp_sum(s1) * inv_d, m2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < D) {
        float4 o = make_float4(rs * (d[i].x - m1 - xh[i].x * m2), rs * (d[i].y - m1 - xh[i].y * m2),
                               rs * (d[i].z - m1 - xh[i].z * m2), rs * (d[i].w - m1 - xh[i].w * m2));
        if (dx_in) {
          const float4 a = *reinterpret_cast<const float4*>(sin + c);
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        ao[i].x += o.x; ao[i].y += o.y; ao[i].z += o.z; ao[i].w += o.w;
        *reinterpret_cast<float4*>(dx_out + r * dx_stride + c) = o;
        if (dx_bf16)
          *reinterpret_cast<uint2*>(dx_bf16 + r * dxb_stride + c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      }
    }
    __syncwarp();   // every lane is done reading this stage before the async proxy overwrites it
    const int64_t rnext = r + (int64_t)kLnPipeStages * warp_stride;
    if (lane == 0 && rnext < rows) issue(rnext, stage);
    if (++stage == kLnPipeStages) { stage = 0; phase ^= 1; }
    mu = mu_n; rs = rs_n;
  }
  // cross-warp reduction of the per-lane column accumulators through the (now idle) staging memory
  __syncthreads();
  constexpr int W = VPT * 128;
  float* red = reinterpret_cast<float*>(data);   // [kLnPipeWarps][3][W]
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = (i * 32 + lane) * 4;
    *reinterpret_cast<float4*>(red + ((size_t)warp * 3 + 0) * W + c) = ag[i];
    *reinterpret_cast<float4*>(red + ((size_t)warp * 3 + 1) * W + c) = ab[i];
    *reinterpret_cast<float4*>(red + ((size_t)warp * 3 + 2) * W + c) = ao[i];
  }
  __syncthreads();
  float* pout = partial + (size_t)blockIdx.x * ncols_out * D;
  for (int idx = threadIdx.x; idx < ncols_out * D; idx += kLnPipeWarps * 32) {
    const int which = idx / D, c = idx - which * D;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kLnPipeWarps; ++w) s += red[((size_t)w * 3 + which) * W + c];
    // accumulate semantics: the CTA's column sums go straight into the (zeroed or running) gradient vectors - 148 atomics per
    // address instead of a partials buffer + ln_reduce_partials_kernel (38 extra launches per ViT-B step)
    if (atomic_out) atomicAdd((which == 0 ? dgamma : (which == 1 ? dbeta : dcol)) + c, s);
    else pout[idx] = s;
  }
}

// ----------------------------------------------------------------------------- few rows: one CTA per row
// AVT-h (B*T = 80 rows of 2048) and the CLS-only final norm of the ViT (80 rows): one warp per row leaves 80 warps
// on the whole GPU chasing 8-24 KB each; with a CTA per row the row is one coalesced sweep of 256 threads.
constexpr int kLnRowThreads = 256;

__device__ __forceinline__ float2 block_sum2(float a, float b, float (*red)[2]) {   // red: [8][2] floats of smem
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { red[warp][0] = a; red[warp][1] = b; }
  __syncthreads();
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int w = 0; w < kLnRowThreads / 32; ++w) { sa += red[w][0]; sb += red[w][1]; }
  return make_float2(sa, sb);
}

template <int VB>   // float4 vectors per thread; D <= VB * 1024
__global__ void __launch_bounds__(kLnRowThreads)
ln_fwd_row_kernel(const float* __restrict__ x, int64_t x_stride, const bf16* __restrict__ add, int64_t add_stride,
                  float* __restrict__ x_out, int64_t xo_stride, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, int D, void* __restrict__ y, int y_fp32, int64_t y_stride,
                  float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  pdl_enter();
  __shared__ float red[kLnRowThreads / 32][2];
  const int64_t r = blockIdx.x;
  const float inv_d = 1.0f / (float)D;
  float4 v[VB], g[VB], b[VB];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VB; ++i) {
    const int c = (i * kLnRowThreads + threadIdx.x) * 4;
    v[i] = g[i] = b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D) {
      v[i] = *reinterpret_cast<const float4*>(x + r * x_stride + c);
      g[i] = __ldg(reinterpret_cast<const float4*>(gamma + c));
      b[i] = __ldg(reinterpret_cast<const float4*>(beta + c));
      if (add) {
        const uint2 a = *reinterpret_cast<const uint2*>(add + r * add_stride + c);
        v[i].x += bf16_lo(a.x); v[i].y += bf16_hi(a.x); v[i].z += bf16_lo(a.y); v[i].w += bf16_hi(a.y);
      }
      if (x_out) *reinterpret_cast<float4*>(x_out + r * xo_stride + c) = v[i];   // x + add, or a compact copy of strided rows
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = block_sum2(s, 0.f, red).x * inv_d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VB; ++i) {
    const int c = (i * kLnRowThreads + threadIdx.x) * 4;
    if (c < D) {
      const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
      q += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
    }
  }
  const float rstd = rsqrtf(block_sum2(q, 0.f, red).x * inv_d + eps);
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[r] = mean;
    if (rstd_out) rstd_out[r] = rstd;
  }
#pragma unroll
  for (int i = 0; i < VB; ++i) {
    const int c = (i * kLnRowThreads + threadIdx.x) * 4;
    if (c < D) {
      const float o0 = (v[i].x - mean) * rstd * g[i].x + b[i].x, o1 = (v[i].y - mean) * rstd * g[i].y + b[i].y;
      const float o2 = (v[i].z - mean) * rstd * g[i].z + b[i].z, o3 = (v[i].w - mean) * rstd * g[i].w + b[i].w;
      if (y_fp32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + r * y_stride + c) = make_float4(o0, o1, o2, o3);
      } else {
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + r * y_stride + c) =
            make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
      }
    }
  }
}

template <int VB>
__global__ void __launch_bounds__(kLnRowThreads)
ln_bwd_row_kernel(const void* __restrict__ dy, int dy_fp32, int64_t dy_stride, const float* __restrict__ x,
                  int64_t x_stride, const float* __restrict__ mean, const float* __restrict__ rstd,
                  const float* __restrict__ gamma, int D, const float* dx_in, float* dx_out, int64_t dx_stride,
                  bf16* __restrict__ dx_bf16, int64_t dxb_stride, int ncols_out, float* __restrict__ partial /*[rows][ncols_out][D]*/) {
  pdl_enter();
  __shared__ float red[kLnRowThreads / 32][2];
  const int64_t r = blockIdx.x;
  const float mu = mean[r], rs = rstd[r];
  const float inv_d = 1.0f / (float)D;
  float* pout = partial + (size_t)r * ncols_out * D;
  float4 xh[VB], d[VB], din[VB];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VB; ++i) {
    const int c = (i * kLnRowThreads + threadIdx.x) * 4;
    xh[i] = d[i] = din[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D) {
      const float4 xv = *reinterpret_cast<const float4*>(x + r * x_stride + c);
      if (dy_fp32) {
        d[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + r * dy_stride + c);
      } else {
        const uint2 p = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dy) + r * dy_stride + c);
        d[i] = make_float4(bf16_lo(p.x), bf16_hi(p.x), bf16_lo(p.y), bf16_hi(p.y));
      }
      if (dx_in) din[i] = *reinterpret_cast<const float4*>(dx_in + r * dx_stride + c);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      *reinterpret_cast<float4*>(pout + c) = make_float4(d[i].x * xh[i].x, d[i].y * xh[i].y, d[i].z * xh[i].z, d[i].w * xh[i].w);
      *reinterpret_cast<float4*>(pout + D + c) = d[i];
      d[i].x *= g.x; d[i].y *= g.y; d[i].z *= g.z; d[i].w *= g.w;
      s1 += d[i].x + d[i].y + d[i].z + d[i].w;
      s2 += d[i].x * xh[i].x + d[i].y * xh[i].y + d[i].z * xh[i].z + d[i].w * xh[i].w;
    }
  }
  const float2 m = block_sum2(s1, s2, red);
  const float m1 = m.x * inv_d, m2 = m.y * inv_d;
#pragma unroll
  for (int i = 0; i < VB; ++i) {
    const int c = (i * kLnRowThreads + threadIdx.x) * 4;
    if (c < D) {
      float4 o = make_float4(rs * (d[i].x - m1 - xh[i].x * m2), rs * (d[i].y - m1 - xh[i].y * m2),
                             rs * (d[i].z - m1 - xh[i].z * m2), rs * (d[i].w - m1 - xh[i].w * m2));
      o.x += din[i].x; o.y += din[i].y; o.z += din[i].z; o.w += din[i].w;
      *reinterpret_cast<float4*>(dx_out + r * dx_stride + c) = o;
      if (dx_bf16)
        *reinterpret_cast<uint2*>(dx_bf16 + r * dxb_stride + c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      if (ncols_out == 3) *reinterpret_cast<float4*>(pout + 2 * D + c) = o;
    }
  }
}

// out_k[c] (+)= sum_p partial[p][k][c] for k < ncols_out (dgamma, dbeta, column sums of dx_out);
// blockDim (32, 8): 32 columns per block, partials split 8 ways, 4 independent loads in flight per thread.
__global__ void __launch_bounds__(256)
ln_reduce_partials_kernel(const float* __restrict__ partial, int nparts, int D, int ncols_out, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, float* __restrict__ dcol, int accumulate) {
  pdl_enter();
  __shared__ float red[8][33];
  const int total = ncols_out * D;
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < total) {
    int p = threadIdx.y;
    for (; p + 24 < nparts; p += 32) {
      s0 += partial[(size_t)p * total + c];
      s1 += partial[(size_t)(p + 8) * total + c];
      s2 += partial[(size_t)(p + 16) * total + c];
      s3 += partial[(size_t)(p + 24) * total + c];
    }
    for (; p < nparts; p += 8) s0 += partial[(size_t)p * total + c];
  }
  red[threadIdx.y][threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (threadIdx.y == 0 && c < total) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    float* dst = c < D ? dgamma + c : (c < 2 * D ? dbeta + (c - D) : dcol + (c - 2 * D));
    *dst = accumulate ? *dst + t : t;
  }
}

// out[c] (+)= sum_r x[r, c] for fp32 x (only behind the register-resident fallback kernel)
__global__ void __launch_bounds__(256)
colsum_f32_kernel(const float* __restrict__ x, int64_t rows, int D, int64_t ld, float* __restrict__ out, int accumulate) {
  pdl_enter();
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < D)
    for (int64_t r = threadIdx.y; r < rows; r += 8) s += x[r * ld + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    out[c] = accumulate ? out[c] + t : t;
  }
}

template <int VPT>
static int launch_fwd(const float* x, int64_t xs, const bf16* add, int64_t as, float* xo, int64_t xos, const float* g,
                      const float* b, float eps, int64_t rows, int D, void* y, int y_fp32, int64_t ys, float* mean,
                      float* rstd, cudaStream_t st) {
  int64_t blocks = (rows + kLnWarps - 1) / kLnWarps;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  launch_kernel(ln_fwd_kernel<VPT>, dim3((int)blocks), dim3(kLnWarps * 32), 0, st, x, xs, add, as, xo, xos, g, b, eps, rows, D, y, y_fp32, ys, mean, rstd);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

static int ln_bwd_blocks(int64_t rows) {
  int64_t blocks = (rows + kLnWarps - 1) / kLnWarps;
  const int64_t cap = (int64_t)num_sms() * 2;
  return (int)(blocks > cap ? cap : blocks);
}

// rows <= this: one CTA per row (also the bound on the number of partial-sum slices of every backward variant)
static int64_t ln_row_kernel_max_rows() { return 2 * (int64_t)num_sms(); }

static size_t ln_pipe_smem(int D, int dy_fp32, bool has_in) {
  const size_t stage = (size_t)D * 4 + (has_in ? (size_t)D * 4 : 0) + (size_t)D * (dy_fp32 ? 4 : 2);
  const size_t ring = (size_t)kLnPipeWarps * kLnPipeStages * stage;
  const int vpt = D <= 128 ? 1 : D <= 256 ? 2 : D <= 512 ? 4 : D <= 768 ? 6 : D <= 1024 ? 8 : 16;   // AVT_LN_DISPATCH
  const size_t red = (size_t)kLnPipeWarps * 3 * vpt * 128 * sizeof(float);   // end-of-kernel reduction
  return 128 + (ring > red ? ring : red);
}

template <int VPT>
static int launch_bwd_pipe(const void* dy, int dy_fp32, int64_t dys, const float* x, int64_t xs, const float* mean,
                           const float* rstd, const float* gamma, int64_t rows, int D, const float* dx_in, float* dx_out,
                           int64_t dxs, bf16* dxb, int64_t dxbs, int ncols_out, float* partial, int blocks, size_t smem,
                           float* dgamma, float* dbeta, float* dcol, int atomic_out, cudaStream_t st) {
  auto kern = ln_bwd_pipe_kernel<VPT>;
  static size_t configured = 0;
  if (smem > configured) {
    AVT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  launch_kernel(kern, dim3(blocks), dim3(kLnPipeWarps * 32), smem, st, dy, dy_fp32, dys, x, xs, mean, rstd, gamma, rows, D, dx_in, dx_out, dxs, dxb,
                                                dxbs, ncols_out, partial, dgamma, dbeta, dcol, atomic_out);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

}  // namespace avt

using namespace avt;

#define AVT_LN_DISPATCH(D, CALL)                       \
  do {                                                 \
    if ((D) <= 128) { constexpr int V = 1; CALL; }     \
    else if ((D) <= 256) { constexpr int V = 2; CALL; }\
    else if ((D) <= 512) { constexpr int V = 4; CALL; }\
    else if ((D) <= 768) { constexpr int V = 6; CALL; }\
    else if ((D) <= 1024) { constexpr int V = 8; CALL; }\
    else { constexpr int V = 16; CALL; }               \
  } while (0)

extern "C" int avt_layernorm_fwd(const float* x, int64_t x_stride, const void* add_bf16, int64_t add_stride, float* x_out,
                                 int64_t x_out_stride, const float* gamma, const float* beta, float eps, int64_t rows,
                                 int D, void* y, int y_fp32, int64_t y_stride, float* mean, float* rstd, void* stream) {
  AVT_REQUIRE(x && gamma && beta && y, "null pointer");
  AVT_REQUIRE(D > 0 && D <= 2048 && D % 4 == 0, "D must be a multiple of 4 and <= 2048");
  AVT_REQUIRE(x_stride % 4 == 0 && y_stride % 4 == 0 && add_stride % 4 == 0 && x_out_stride % 4 == 0,
              "row strides must be multiples of 4");
  if (rows <= 0) return AVT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bf16* add = reinterpret_cast<const bf16*>(add_bf16);
  if (rows <= ln_row_kernel_max_rows()) {
    if (D <= 1024)
      launch_kernel(ln_fwd_row_kernel<1>, dim3((int)rows), dim3(kLnRowThreads), 0, st, x, x_stride, add, add_stride, x_out, x_out_stride, gamma, beta,
                                                                eps, D, y, y_fp32, y_stride, mean, rstd);
    else
      launch_kernel(ln_fwd_row_kernel<2>, dim3((int)rows), dim3(kLnRowThreads), 0, st, x, x_stride, add, add_stride, x_out, x_out_stride, gamma, beta,
                                                                eps, D, y, y_fp32, y_stride, mean, rstd);
    AVT_CUDA_OK(cudaGetLastError());
    return AVT_OK;
  }
  AVT_LN_DISPATCH(D, return launch_fwd<V>(x, x_stride, add, add_stride, x_out, x_out_stride, gamma, beta, eps, rows, D, y,
                                          y_fp32, y_stride, mean, rstd, st));
  return AVT_OK;
}

extern "C" int64_t avt_layernorm_bwd_workspace_bytes(int64_t rows, int D) {
  (void)rows;
  return ln_row_kernel_max_rows() * 3 * D * (int64_t)sizeof(float);
}

extern "C" int avt_layernorm_bwd(const void* dy, int dy_fp32, int64_t dy_stride, const float* x, int64_t x_stride,
                                 const float* mean, const float* rstd, const float* gamma, int64_t rows, int D,
                                 const float* dx_in, float* dx_out, int64_t dx_stride, void* dx_bf16, int64_t dxb_stride,
                                 float* dgamma, float* dbeta, float* dx_colsum, int accumulate, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
  AVT_REQUIRE(dy && x && mean && rstd && gamma && dx_out && dgamma && dbeta && workspace, "null pointer");
  AVT_REQUIRE(D > 0 && D <= 2048 && D % 4 == 0, "D must be a multiple of 4 and <= 2048");
  AVT_REQUIRE(dy_stride % 4 == 0 && x_stride % 4 == 0 && dx_stride % 4 == 0 && dxb_stride % 4 == 0,
              "row strides must be multiples of 4");
  AVT_REQUIRE(workspace_bytes >= avt_layernorm_bwd_workspace_bytes(rows, D), "workspace too small");
  if (rows <= 0) return AVT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* partial = reinterpret_cast<float*>(workspace);
  bf16* dxb = reinterpret_cast<bf16*>(dx_bf16);
  int ncols_out = dx_colsum ? 3 : 2;
  int nparts;
  const size_t pipe_smem = ln_pipe_smem(D, dy_fp32, dx_in != nullptr);
  const bool pipe_ok = D >= 256 && D <= 1024 && D % 8 == 0 && dy_stride % 8 == 0 && pipe_smem <= 227 * 1024 &&
                       (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(dx_in) & 15) == 0;
  if (rows <= ln_row_kernel_max_rows()) {
    nparts = (int)rows;
    if (D <= 1024)
      launch_kernel(ln_bwd_row_kernel<1>, dim3(nparts), dim3(kLnRowThreads), 0, st, dy, dy_fp32, dy_stride, x, x_stride, mean, rstd, gamma, D, dx_in,
                                                             dx_out, dx_stride, dxb, dxb_stride, ncols_out, partial);
    else
      launch_kernel(ln_bwd_row_kernel<2>, dim3(nparts), dim3(kLnRowThreads), 0, st, dy, dy_fp32, dy_stride, x, x_stride, mean, rstd, gamma, D, dx_in,
                                                             dx_out, dx_stride, dxb, dxb_stride, ncols_out, partial);
    AVT_CUDA_OK(cudaGetLastError());
  } else if (pipe_ok) {
    nparts = num_sms();
    AVT_LN_DISPATCH(D, {
      if (int rc = launch_bwd_pipe<V>(dy, dy_fp32, dy_stride, x, x_stride, mean, rstd, gamma, rows, D, dx_in, dx_out,
                                      dx_stride, dxb, dxb_stride, ncols_out, partial, nparts, pipe_smem, dgamma, dbeta, dx_colsum,
                                      accumulate ? 1 : 0, st))
        return rc;
    });
    if (accumulate) return AVT_OK;   // the pipe kernel added its column sums with atomics: no reduction pass
  } else {
    nparts = ln_bwd_blocks(rows);
    ncols_out = 2;
    AVT_LN_DISPATCH(D, (launch_kernel(ln_bwd_kernel<V>, dim3(nparts), dim3(kLnWarps * 32), 0, st, 
                           dy, dy_fp32, dy_stride, x, x_stride, mean, rstd, gamma, rows, D, dx_in, dx_out, dx_stride, dxb,
                           dxb_stride, partial)));
    AVT_CUDA_OK(cudaGetLastError());
    if (dx_colsum) {
      launch_kernel(colsum_f32_kernel, dim3((D + 31) / 32), dim3(dim3(32, 8)), 0, st, dx_out, rows, D, dx_stride, dx_colsum, accumulate);
      AVT_CUDA_OK(cudaGetLastError());
    }
  }
  launch_kernel(ln_reduce_partials_kernel, dim3((ncols_out * D + 31) / 32), dim3(dim3(32, 8)), 0, st, partial, nparts, D, ncols_out, dgamma, dbeta,
                                                                              dx_colsum, accumulate);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}
