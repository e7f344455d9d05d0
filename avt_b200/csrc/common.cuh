// Shared device helpers: activations, Philox dropout, warp reductions, error plumbing.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../include/avt_b200.h"

namespace avt {

typedef __nv_bfloat16 bf16;

#define AVT_CUDA_OK(expr)                                                        \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) {                                                     \
      avt::set_last_error(#expr, cudaGetErrorString(_e), __FILE__, __LINE__);    \
      return AVT_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

#define AVT_REQUIRE(cond, msg)                                                   \
  do {                                                                           \
    if (!(cond)) {                                                               \
      avt::set_last_error(#cond, msg, __FILE__, __LINE__);                       \
      return AVT_ERR_INVALID;                                                    \
    }                                                                            \
  } while (0)

void set_last_error(const char* what, const char* detail, const char* file, int line);
int num_sms();

// ----------------------------------------------------------------------------- activations
// timm ViT uses nn.GELU (erf); HF GPT-2 uses gelu_new (tanh approximation).
// Both are evaluated branch-free (epilogue warps interleave 32 independent elements per thread; a branchy
// libm erff/tanhf serialises them and made the fused GEMM epilogue 4x slower than its main loop).
//   erf : Abramowitz & Stegun 7.1.28, |abs err| <= 3e-7:  erf(x) = 1 - (1 + a1 x + ... + a6 x^6)^-16, x >= 0
//   tanh: 1 - 2 / (1 + exp(2u)), rel err ~1e-6 (ex2.approx + rcp.approx)
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Phi(x) = 0.5 * (1 + erf(x / sqrt(2)))
__device__ __forceinline__ float normal_cdf(float x) {
  const float u = fabsf(x) * 0.70710678118654752f;
  float p = fmaf(u, 0.0000430638f, 0.0002765672f);
  p = fmaf(p, u, 0.0001520143f);
  p = fmaf(p, u, 0.0092705272f);
  p = fmaf(p, u, 0.0422820123f);
  p = fmaf(p, u, 0.0705230784f);
  p = fmaf(p, u, 1.0f);
  p = p * p; p = p * p; p = p * p; p = p * p;     // ^16 (overflows to +inf for |x| > ~25: rcp -> 0, erf -> 1)
  const float half_erf = 0.5f - 0.5f * fast_rcp(p);  // 0.5 * erf(|x|/sqrt2)
  return 0.5f + copysignf(half_erf, x);
}
__device__ __forceinline__ float gelu_erf(float x) { return x * normal_cdf(x); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float pdf = 0.39894228040143268f * fast_ex2(-0.72134752044448170f * x * x);  // exp(-x^2/2)/sqrt(2 pi)
  return fmaf(x, pdf, normal_cdf(x));
}
__device__ __forceinline__ float fast_tanh(float u) {
  const float e = fast_ex2(fminf(2.8853900817779268f * u, 80.0f));  // exp(2u), clamped (no inf/inf)
  return 1.0f - 2.0f * fast_rcp(1.0f + e);
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float u = 0.79788456080286536f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + fast_tanh(u));
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float x2 = x * x;
  const float u = 0.79788456080286536f * (x + 0.044715f * x * x2);
  const float t = fast_tanh(u);
  const float du = 0.79788456080286536f * (1.0f + 3.0f * 0.044715f * x2);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}
__device__ __forceinline__ float apply_act(int act, float x) {
  return act == AVT_ACT_GELU_ERF ? gelu_erf(x) : (act == AVT_ACT_GELU_TANH ? gelu_tanh(x) : x);
}
__device__ __forceinline__ float apply_act_grad(int act, float x) {
  return act == AVT_ACT_GELU_ERF ? gelu_erf_grad(x) : (act == AVT_ACT_GELU_TANH ? gelu_tanh_grad(x) : 1.0f);
}

// ---- packed fp32x2 variants (sm_100 FFMA2 / FMUL2 / FADD2: two lanes of work per issued instruction). The fused GEMM
// epilogues are issue-bound on 8 warps; evaluating two accumulator columns per instruction halves the FMA-pipe slots.
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 normal_cdf2(float2 x) {
  const float2 u = __fmul2_rn(make_float2(fabsf(x.x), fabsf(x.y)), f2(0.70710678118654752f));
  float2 p = __ffma2_rn(u, f2(0.0000430638f), f2(0.0002765672f));
  p = __ffma2_rn(p, u, f2(0.0001520143f));
  p = __ffma2_rn(p, u, f2(0.0092705272f));
  p = __ffma2_rn(p, u, f2(0.0422820123f));
  p = __ffma2_rn(p, u, f2(0.0705230784f));
  p = __ffma2_rn(p, u, f2(1.0f));
  p = __fmul2_rn(p, p); p = __fmul2_rn(p, p); p = __fmul2_rn(p, p); p = __fmul2_rn(p, p);
  const float2 r = make_float2(fast_rcp(p.x), fast_rcp(p.y));
  const float2 he = __ffma2_rn(r, f2(-0.5f), f2(0.5f));   // 0.5 * erf(|x|/sqrt2)
  return __fadd2_rn(f2(0.5f), make_float2(copysignf(he.x, x.x), copysignf(he.y, x.y)));
}
__device__ __forceinline__ float2 normal_pdf2(float2 x) {
  const float2 t = __fmul2_rn(__fmul2_rn(x, x), f2(-0.72134752044448170f));
  return __fmul2_rn(make_float2(fast_ex2(t.x), fast_ex2(t.y)), f2(0.39894228040143268f));
}
__device__ __forceinline__ float2 tanh2(float2 u) {
  const float2 t = __fmul2_rn(u, f2(2.8853900817779268f));
  const float2 e = make_float2(fast_ex2(fminf(t.x, 80.0f)), fast_ex2(fminf(t.y, 80.0f)));
  const float2 d = __fadd2_rn(e, f2(1.0f));
  return __ffma2_rn(make_float2(fast_rcp(d.x), fast_rcp(d.y)), f2(-2.0f), f2(1.0f));
}
// y = act(x), dy = act'(x) for two elements; ACT and GRAD are compile-time so the unrolled epilogue loops carry no branches
template <int ACT, bool GRAD>
__device__ __forceinline__ void act_and_grad2(float2 x, float2& y, float2& dy) {
  if constexpr (ACT == AVT_ACT_GELU_ERF) {
    const float2 cdf = normal_cdf2(x);
    y = __fmul2_rn(x, cdf);
    if constexpr (GRAD) dy = __ffma2_rn(x, normal_pdf2(x), cdf);
  } else if constexpr (ACT == AVT_ACT_GELU_TANH) {
    const float2 x2 = __fmul2_rn(x, x);
    const float2 inner = __fmul2_rn(__ffma2_rn(__fmul2_rn(x2, x), f2(0.044715f), x), f2(0.79788456080286536f));
    const float2 t = tanh2(inner);
    const float2 hx = __fmul2_rn(x, f2(0.5f));
    const float2 tp1 = __fadd2_rn(t, f2(1.0f));
    y = __fmul2_rn(hx, tp1);
    if constexpr (GRAD) {
      const float2 du = __ffma2_rn(x2, f2(3.0f * 0.044715f * 0.79788456080286536f), f2(0.79788456080286536f));
      const float2 omt2 = __ffma2_rn(t, make_float2(-t.x, -t.y), f2(1.0f));
      dy = __ffma2_rn(__fmul2_rn(hx, omt2), du, __fmul2_rn(tp1, f2(0.5f)));
    }
  } else {
    y = x;
    if constexpr (GRAD) dy = f2(1.0f);
  }
}
// 16 float2 = one 32-column chunk of a row. MODE 0: v = act(v); 1: a = v, v = act(v); 2: a = act'(v), v = act(v);
// 3: v = act'(v) (v holds pre-activations).
template <int ACT, int MODE>
__device__ __forceinline__ void act_chunk_t(float2 (&v)[16], float2 (&a)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float2 y, dy;
    if constexpr (MODE == 2 || MODE == 3) act_and_grad2<ACT, true>(v[j], y, dy);
    else act_and_grad2<ACT, false>(v[j], y, dy);
    if constexpr (MODE == 1) a[j] = v[j];
    if constexpr (MODE == 2) a[j] = dy;
    v[j] = MODE == 3 ? dy : y;
  }
}
template <int MODE>
__device__ __forceinline__ void act_chunk(int act, float2 (&v)[16], float2 (&a)[16]) {
  if (act == AVT_ACT_GELU_ERF) act_chunk_t<AVT_ACT_GELU_ERF, MODE>(v, a);
  else if (act == AVT_ACT_GELU_TANH) act_chunk_t<AVT_ACT_GELU_TANH, MODE>(v, a);
  else act_chunk_t<AVT_ACT_NONE, MODE>(v, a);
}

// ----------------------------------------------------------------------------- Philox4x32-10
// Counter-based RNG: the dropout mask of element i is a pure function of (seed, offset, i), so the
// backward pass regenerates it instead of storing it.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// Four consecutive elements [4*g, 4*g+4) share one Philox call. Returns keep-mask bits (bit j = keep elem j).
__device__ __forceinline__ uint32_t dropout_keep4(uint64_t seed, uint64_t offset, uint64_t group, float p) {
  const uint64_t c = offset + group;
  uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t thr = (uint32_t)fminf(p * 4294967296.0f, 4294967295.0f);
  return (r.x >= thr ? 1u : 0u) | (r.y >= thr ? 2u : 0u) | (r.z >= thr ? 4u : 0u) | (r.w >= thr ? 8u : 0u);
}

// ----------------------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace avt
