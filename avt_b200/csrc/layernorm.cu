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
        if (x_out) *reinterpret_cast<float4*>(x_out + r * xo_stride + c) = v[i];
      }
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

// out[c] (+)= sum_p partial[p][c]; blockDim (32, 8): 32 columns per block, partials split 8 ways
__global__ void __launch_bounds__(256)
ln_reduce_partials_kernel(const float* __restrict__ partial, int nparts, int D, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, int accumulate) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < 2 * D)
    for (int p = threadIdx.y; p < nparts; p += 8) s += partial[(size_t)p * 2 * D + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < 2 * D) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    float* dst = c < D ? dgamma + c : dbeta + (c - D);
    *dst = accumulate ? *dst + t : t;
  }
}

template <int VPT>
static int launch_fwd(const float* x, int64_t xs, const bf16* add, int64_t as, float* xo, int64_t xos, const float* g,
                      const float* b, float eps, int64_t rows, int D, void* y, int y_fp32, int64_t ys, float* mean,
                      float* rstd, cudaStream_t st) {
  int64_t blocks = (rows + kLnWarps - 1) / kLnWarps;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  ln_fwd_kernel<VPT><<<(int)blocks, kLnWarps * 32, 0, st>>>(x, xs, add, as, xo, xos, g, b, eps, rows, D, y, y_fp32, ys, mean, rstd);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

static int ln_bwd_blocks(int64_t rows) {
  int64_t blocks = (rows + kLnWarps - 1) / kLnWarps;
  const int64_t cap = (int64_t)num_sms() * 2;
  return (int)(blocks > cap ? cap : blocks);
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
  AVT_LN_DISPATCH(D, return launch_fwd<V>(x, x_stride, reinterpret_cast<const bf16*>(add_bf16), add_stride, x_out,
                                          x_out_stride, gamma, beta, eps, rows, D, y, y_fp32, y_stride, mean, rstd, st));
  return AVT_OK;
}

extern "C" int64_t avt_layernorm_bwd_workspace_bytes(int64_t rows, int D) {
  return (int64_t)ln_bwd_blocks(rows) * 2 * D * (int64_t)sizeof(float);
}

extern "C" int avt_layernorm_bwd(const void* dy, int dy_fp32, int64_t dy_stride, const float* x, int64_t x_stride,
                                 const float* mean, const float* rstd, const float* gamma, int64_t rows, int D,
                                 const float* dx_in, float* dx_out, int64_t dx_stride, void* dx_bf16, int64_t dxb_stride,
                                 float* dgamma, float* dbeta, int accumulate, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
  AVT_REQUIRE(dy && x && mean && rstd && gamma && dx_out && dgamma && dbeta && workspace, "null pointer");
  AVT_REQUIRE(D > 0 && D <= 2048 && D % 4 == 0, "D must be a multiple of 4 and <= 2048");
  AVT_REQUIRE(dy_stride % 4 == 0 && x_stride % 4 == 0 && dx_stride % 4 == 0 && dxb_stride % 4 == 0,
              "row strides must be multiples of 4");
  AVT_REQUIRE(workspace_bytes >= avt_layernorm_bwd_workspace_bytes(rows, D), "workspace too small");
  if (rows <= 0) return AVT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = ln_bwd_blocks(rows);
  float* partial = reinterpret_cast<float*>(workspace);
  AVT_LN_DISPATCH(D, (ln_bwd_kernel<V><<<blocks, kLnWarps * 32, 0, st>>>(
                         dy, dy_fp32, dy_stride, x, x_stride, mean, rstd, gamma, rows, D, dx_in, dx_out, dx_stride,
                         reinterpret_cast<bf16*>(dx_bf16), dxb_stride, partial)));
  AVT_CUDA_OK(cudaGetLastError());
  ln_reduce_partials_kernel<<<(2 * D + 31) / 32, dim3(32, 8), 0, st>>>(partial, blocks, D, dgamma, dbeta, accumulate);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}
