// HBM-bound helper kernels around the GEMMs: casts, patch re-tiling, column / frame reductions,
// dropout backward. All use 128-bit accesses and grid sizes that are multiples of the SM count.
#include "common.cuh"
#include "ptx.cuh"

namespace avt {

// ----------------------------------------------------------------------------- fp32 -> bf16 cast
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n8) {
  pdl_enter();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    reinterpret_cast<uint4*>(dst)[i] =
        make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  }
}
__global__ void cast_f32_bf16_tail_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t begin, int64_t n) {
  pdl_enter();
  const int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

// ----------------------------------------------------------------------------- patchify
// video fp32 [F, C, H, W] -> A bf16 [F * (P + 1), C*ps*ps]; row f*(P+1) is zero (the CLS slot), row
// f*(P+1) + 1 + py*PW + px holds patch (py, px) flattened as (c, kh, kw) — the K order of
// Conv2d.weight.reshape(D, C*ps*ps). stride == kernel, so this is a pure re-tiling (no im2col blow-up).
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ video, bf16* __restrict__ out, int F, int C, int H, int W, int ps) {
  pdl_enter();
  const int PW = W / ps, PH = H / ps, P = PW * PH;
  const int K = C * ps * ps, K8 = K / 8;
  const int64_t total = (int64_t)F * (P + 1) * K8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int cpr = ps / 8;  // 8-wide chunks per patch row
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int ck = (int)(i % K8);
    const int64_t row = i / K8;
    const int t = (int)(row % (P + 1));
    const int f = (int)(row / (P + 1));
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (t > 0) {
      const int p = t - 1, py = p / PW, px = p % PW;
      const int c = ck / (ps * cpr), rem = ck % (ps * cpr), kh = rem / cpr, kw8 = rem % cpr;
      const float* src = video + (((int64_t)f * C + c) * H + (py * ps + kh)) * W + px * ps + kw8 * 8;
      const float4 a = __ldg(reinterpret_cast<const float4*>(src));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      o = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    }
    reinterpret_cast<uint4*>(out)[i] = o;
  }
}

// ----------------------------------------------------------------------------- column sums (bias grads)
// out[c] += sum_r x[r, c]   (x bf16 [R, ld]); blockDim (32, 8); each thread owns 8 columns.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const bf16* __restrict__ x, int64_t R, int C, int64_t ld, float* __restrict__ out) {
  pdl_enter();
  __shared__ float red[8][32 * 8 + 1];
  const int c0 = (blockIdx.x * 32 + threadIdx.x) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c0 < C) {
    for (int64_t r = (int64_t)blockIdx.y * 8 + threadIdx.y; r < R; r += (int64_t)gridDim.y * 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + r * ld + c0));
      acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
      acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.y][threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  const int tid = threadIdx.y * 32 + threadIdx.x;  // 256 threads <-> 256 columns of this block
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w][tid];
  const int c = blockIdx.x * 256 + tid;
  if (c < C) atomicAdd(out + c, s);
}

// ----------------------------------------------------------------------------- frame sums (pos / cls / wpe grads)
// s[t, c] = sum_f x[(f*period + t), c]   (x fp32 [F*period, D])
__global__ void __launch_bounds__(128)
framesum_kernel(const float* __restrict__ x, int F, int period, int D, float* __restrict__ s) {
  pdl_enter();
  const int t = blockIdx.y;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= D) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int f = 0; f < F; ++f) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((int64_t)f * period + t) * D + c));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(s + (int64_t)t * D + c) = acc;
}
// dpos (+)= s;  dcls (+)= s[0];  dbias (+)= sum_{t>=1} s[t]
__global__ void pos_cls_apply_kernel(const float* __restrict__ s, int period, int D, float* __restrict__ dpos,
                                     float* __restrict__ dcls, float* __restrict__ dbias, int accumulate) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float rest = 0.f;
  for (int t = 0; t < period; ++t) {
    const float v = s[(int64_t)t * D + c];
    if (dpos) dpos[(int64_t)t * D + c] = accumulate ? dpos[(int64_t)t * D + c] + v : v;
    if (t > 0) rest += v;
  }
  if (dcls) dcls[c] = accumulate ? dcls[c] + s[c] : s[c];
  if (dbias) dbias[c] = accumulate ? dbias[c] + rest : rest;
}

// ----------------------------------------------------------------------------- dropout backward / apply
// y = keep(seed, offset, i) ? x / (1-p) : 0 over a dense [n] tensor (i = linear index); fp32 in, bf16 and/or fp32 out.
__global__ void __launch_bounds__(256)
dropout_apply_kernel(const float* __restrict__ x, int64_t n4, float p, uint64_t seed, uint64_t offset,
                     const uint64_t* __restrict__ offset_dev, float* __restrict__ y32, bf16* __restrict__ y16) {
  pdl_enter();
  const float scale = 1.0f / (1.0f - p);
  if (offset_dev) offset += __ldg(offset_dev);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += stride) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + g);
    if (p > 0.f) {
      const uint32_t keep = dropout_keep4(seed, offset, (uint64_t)g, p);
      v.x = (keep & 1u) ? v.x * scale : 0.f; v.y = (keep & 2u) ? v.y * scale : 0.f;
      v.z = (keep & 4u) ? v.z * scale : 0.f; v.w = (keep & 8u) ? v.w * scale : 0.f;
    }
    if (y32) reinterpret_cast<float4*>(y32)[g] = v;
    if (y16) reinterpret_cast<uint2*>(y16)[g] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// ----------------------------------------------------------------------------- fused SGD step over a flat buffer
// torch.optim.SGD semantics (dampening 0): g += wd*p; m = first ? g : mom*m + g; g = nesterov ? g + mom*m : m;
// p -= lr*g; and the bf16 shadow of p used by the GEMMs is refreshed in the same pass (20 B/param instead of the
// 4 torch foreach passes + a separate down-cast). Elements [0, lo4*4) use wd_lo (the reference's bias / bn group,
// func/train.py:704-731), the rest wd. The gradients may be bf16 (the data-parallel payload), lr may live in device
// memory (a captured step then follows the reference's per-iteration lr schedule without re-capturing).
template <bool G_BF16>
__global__ void __launch_bounds__(256)
sgd_step_kernel(float* __restrict__ p, const void* __restrict__ g, float* __restrict__ m, bf16* __restrict__ shadow,
                int64_t n4, int64_t lo4, float lr, const float* __restrict__ lr_dev, float mom, float wd, float wd_lo,
                int nesterov, int first) {
  pdl_enter();
  if (lr_dev) lr = __ldg(lr_dev);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv;
    if constexpr (G_BF16) {
      const uint2 gb = __ldg(reinterpret_cast<const uint2*>(g) + i);
      gv = make_float4(bf16_lo(gb.x), bf16_hi(gb.x), bf16_lo(gb.y), bf16_hi(gb.y));
    } else {
      gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    }
    float4 mv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(m)[i];
    const float w = i < lo4 ? wd_lo : wd;
    float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w}, ma[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gg = fmaf(w, pa[k], ga[k]);
      ma[k] = first ? gg : fmaf(mom, ma[k], gg);
      gg = nesterov ? fmaf(mom, ma[k], gg) : ma[k];
      pa[k] = fmaf(-lr, gg, pa[k]);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    if (shadow) reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16x2(pa[0], pa[1]), pack_bf16x2(pa[2], pa[3]));
  }
}

// ----------------------------------------------------------------------------- input pipeline: uint8 frames -> normalised fp32 crop
// The reference's per-frame transform chain after decoding / resizing (func/train.py:550-584, common/transforms.py):
// ToTensorVideo (uint8 HWC -> float CHW / 255), RandomHorizontalFlipVideo, x * scale_pix_val, NormalizeVideo(mean, std),
// RandomCropVideo / a fixed crop - one pass on the GPU, so the step's host-to-device copy is the uint8 frames (4x fewer bytes)
// and the CPU workers stop at decode + resize. out[f, c, y, x] = (in[f, cy + y, cx + (flip ? w-1-x : x), c] / 255 * scale - mean[c]) / std[c]
// with the crop taken after the flip, as in the reference's order (flip, ..., normalize, crop).
__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ in, int F, int Hin, int Win, float* __restrict__ out, int h, int w, int cy, int cx,
                     const uint8_t* __restrict__ flip, float scale, float m0, float m1, float m2, float s0, float s1, float s2) {
  pdl_enter();
  const int64_t total = (int64_t)F * h * w;
  const float mean[3] = {m0, m1, m2}, inv[3] = {1.f / s0, 1.f / s1, 1.f / s2};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = i % w, y = (i / w) % h, f = i / ((int64_t)w * h);
    const int sx = (flip && flip[f]) ? Win - 1 - (cx + x) : cx + x;
    const uint8_t* px = in + (((int64_t)f * Hin + cy + y) * Win + sx) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out[(((int64_t)f * 3 + c) * h + y) * w + x] = (px[c] * (scale / 255.0f) - mean[c]) * inv[c];
  }
}

static int grid_for(int64_t work_items, int threads, int per_sm) {
  int64_t b = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace avt

using namespace avt;

extern "C" int avt_zero(void* dst, int64_t bytes, void* stream) {
  AVT_REQUIRE(dst && bytes >= 0, "null pointer");
  if (bytes == 0) return AVT_OK;
  AVT_CUDA_OK(cudaMemsetAsync(dst, 0, (size_t)bytes, reinterpret_cast<cudaStream_t>(stream)));
  return AVT_OK;
}

extern "C" int avt_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  AVT_REQUIRE(src && dst, "null pointer");
  AVT_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "16-byte alignment");
  if (n <= 0) return AVT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t n8 = n / 8;
  if (n8 > 0) launch_kernel(cast_f32_bf16_kernel, dim3(grid_for(n8, 256, 8)), dim3(256), 0, st, src, reinterpret_cast<bf16*>(dst), n8);
  if (n8 * 8 < n) launch_kernel(cast_f32_bf16_tail_kernel, dim3(1), dim3(32), 0, st, src, reinterpret_cast<bf16*>(dst), n8 * 8, n);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

extern "C" int avt_patchify_bf16(const float* video, void* out, int F, int C, int H, int W, int ps, void* stream) {
  AVT_REQUIRE(video && out, "null pointer");
  AVT_REQUIRE(ps % 8 == 0 && H % ps == 0 && W % ps == 0 && W % 4 == 0, "patch size must be a multiple of 8 dividing H and W");
  AVT_REQUIRE((reinterpret_cast<uintptr_t>(video) & 15) == 0, "video must be 16-byte aligned");
  const int64_t total = (int64_t)F * ((H / ps) * (W / ps) + 1) * (C * ps * ps / 8);
  launch_kernel(patchify_kernel, dim3(grid_for(total, 256, 8)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      video, reinterpret_cast<bf16*>(out), F, C, H, W, ps);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

extern "C" int avt_colsum_bf16(const void* x, int64_t rows, int cols, int64_t ld, float* out, void* stream) {
  AVT_REQUIRE(x && out, "null pointer");
  AVT_REQUIRE(cols % 8 == 0 && ld % 8 == 0, "cols and ld must be multiples of 8");
  if (rows <= 0) return AVT_OK;
  const int gx = (cols + 255) / 256;
  int gy = (num_sms() * 4 + gx - 1) / gx;
  const int64_t max_gy = (rows + 7) / 8;
  if (gy > max_gy) gy = (int)max_gy;
  launch_kernel(colsum_bf16_kernel, dim3(dim3(gx, gy)), dim3(dim3(32, 8)), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const bf16*>(x), rows, cols, ld, out);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

extern "C" int avt_frame_sum_grads(const float* dx, int F, int period, int D, float* dpos, float* dcls, float* dbias,
                                   int accumulate, float* workspace /* period*D floats */, void* stream) {
  AVT_REQUIRE(dx && workspace, "null pointer");
  AVT_REQUIRE(D % 4 == 0, "D must be a multiple of 4");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  launch_kernel(framesum_kernel, dim3(dim3((D / 4 + 127) / 128, period)), dim3(128), 0, st, dx, F, period, D, workspace);
  AVT_CUDA_OK(cudaGetLastError());
  launch_kernel(pos_cls_apply_kernel, dim3((D + 127) / 128), dim3(128), 0, st, workspace, period, D, dpos, dcls, dbias, accumulate);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

extern "C" int avt_dropout_apply(const float* x, int64_t n, float p, uint64_t seed, uint64_t offset,
                                 const uint64_t* offset_dev, float* y_f32, void* y_bf16, void* stream) {
  AVT_REQUIRE(x && (y_f32 || y_bf16), "null pointer");
  AVT_REQUIRE(n % 4 == 0, "n must be a multiple of 4");
  AVT_REQUIRE(p >= 0.f && p < 1.f, "p must be in [0, 1)");
  if (n <= 0) return AVT_OK;
  launch_kernel(dropout_apply_kernel, dim3(grid_for(n / 4, 256, 8)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, n / 4, p, seed, offset, offset_dev, y_f32, reinterpret_cast<bf16*>(y_bf16));
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

extern "C" int avt_sgd_step(float* p, const void* g, int g_is_bf16, float* m, void* p_bf16, int64_t n, float lr,
                            const float* lr_dev, float momentum, float weight_decay, float weight_decay_lo, int64_t lo_elems,
                            int nesterov, int first_step, void* stream) {
  AVT_REQUIRE(p && g && m, "null pointer");
  AVT_REQUIRE(n % 4 == 0 && lo_elems % 4 == 0, "n and lo_elems must be multiples of 4 (flat buffers are padded)");
  AVT_REQUIRE(lo_elems >= 0 && lo_elems <= n, "lo_elems out of range");
  if (n <= 0) return AVT_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid(grid_for(n / 4, 256, 8));
  if (g_is_bf16)
    launch_kernel(sgd_step_kernel<true>, grid, dim3(256), 0, st, p, g, m, reinterpret_cast<bf16*>(p_bf16), n / 4, lo_elems / 4, lr,
                  lr_dev, momentum, weight_decay, weight_decay_lo, nesterov, first_step);
  else
    launch_kernel(sgd_step_kernel<false>, grid, dim3(256), 0, st, p, g, m, reinterpret_cast<bf16*>(p_bf16), n / 4, lo_elems / 4, lr,
                  lr_dev, momentum, weight_decay, weight_decay_lo, nesterov, first_step);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

extern "C" int avt_preprocess_u8(const uint8_t* frames, int F, int Hin, int Win, float* out, int h, int w, int crop_y, int crop_x,
                                 const uint8_t* flip, float scale, const float* mean3, const float* std3, void* stream) {
  AVT_REQUIRE(frames && out && mean3 && std3, "null pointer");
  AVT_REQUIRE(crop_y >= 0 && crop_x >= 0 && crop_y + h <= Hin && crop_x + w <= Win, "crop outside the frame");
  if (F <= 0) return AVT_OK;
  launch_kernel(preprocess_u8_kernel, dim3(grid_for((int64_t)F * h * w, 256, 8)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                frames, F, Hin, Win, out, h, w, crop_y, crop_x, flip, scale, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}
