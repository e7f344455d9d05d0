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
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float u = 0.79788456080286536f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float x2 = x * x;
  const float u = 0.79788456080286536f * (x + 0.044715f * x * x2);
  const float t = tanhf(u);
  const float du = 0.79788456080286536f * (1.0f + 3.0f * 0.044715f * x2);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}
__device__ __forceinline__ float apply_act(int act, float x) {
  return act == AVT_ACT_GELU_ERF ? gelu_erf(x) : (act == AVT_ACT_GELU_TANH ? gelu_tanh(x) : x);
}
__device__ __forceinline__ float apply_act_grad(int act, float x) {
  return act == AVT_ACT_GELU_ERF ? gelu_erf_grad(x) : (act == AVT_ACT_GELU_TANH ? gelu_tanh_grad(x) : 1.0f);
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
