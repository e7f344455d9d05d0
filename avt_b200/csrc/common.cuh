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
bool pdl_enabled();

// ----------------------------------------------------------------------------- programmatic dependent launch
// Every kernel of the library is launched with programmatic stream serialization: its CTAs may become resident while
// the previous kernel of the stream is still draining (launch latency, barrier / TMEM set-up and tensor-map prefetch
// overlap that tail). pdl_wait() blocks until the previous grid has completed and its writes are visible; nothing
// before it may touch global memory written by earlier kernels. pdl_trigger() lets the NEXT kernel start launching.
// A training step is ~550 dependent launches; without this each boundary costs a full drain + launch bubble.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {   // kernels without a prologue worth overlapping
  pdl_wait();
  pdl_trigger();
}

void count_launch();   // core.cu: every kernel launch of the library is counted (avt_kernel_launch_count)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  count_launch();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);
}

// ----------------------------------------------------------------------------- activations
// timm ViT uses nn.GELU (erf); HF GPT-2 uses gelu_new (tanh approximation).
// Both are evaluated branch-free (epilogue warps interleave 32 independent elements per thread; a branchy
// libm erff/tanhf serialises them and made the fused GEMM epilogue 4x slower than its main loop).
//   erf : Abramowitz & Stegun 7.1.26 (rational-polynomial x exp, |abs err| <= 1.5e-7), sharing its exponential with gelu'
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
// Phi(x) = 0.5 * (1 + erf(x / sqrt(2))) and E = exp(-x^2 / 2) from ONE exponential (Abramowitz & Stegun 7.1.26):
//   erfc(u) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-u^2),  t = 1 / (1 + p u),  u = |x| / sqrt 2,  |err| <= 1.5e-7
// so that gelu(x) = x Phi(x) and gelu'(x) = Phi(x) + x E / sqrt(2 pi) share the ex2 and cost 14 FMA-pipe operations per
// element pair instead of the 19 of the 7.1.28 form (a 6th-degree polynomial raised to the 16th power plus a separate exp).
constexpr float kErfP = 0.3275911f * 0.70710678118654752f;
constexpr float kErfA1 = 0.5f * 0.254829592f, kErfA2 = 0.5f * -0.284496736f, kErfA3 = 0.5f * 1.421413741f,
                kErfA4 = 0.5f * -1.453152027f, kErfA5 = 0.5f * 1.061405429f;
constexpr float kNegHalfLog2e = -0.72134752044448170f;   // exp(-x^2/2) = 2^(x^2 * this)
constexpr float kInvSqrt2Pi = 0.39894228040143268f;
__device__ __forceinline__ float normal_cdf_e(float x, float& e) {
  const float t = fast_rcp(fmaf(fabsf(x), kErfP, 1.0f));
  e = fast_ex2(x * x * kNegHalfLog2e);
  float q = fmaf(t, kErfA5, kErfA4);
  q = fmaf(q, t, kErfA3);
  q = fmaf(q, t, kErfA2);
  q = fmaf(q, t, kErfA1);
  const float h = q * t * e;                       // 0.5 * erfc(|x| / sqrt 2)
  return 0.5f + copysignf(0.5f - h, x);
}
__device__ __forceinline__ float normal_cdf(float x) {
  float e;
  return normal_cdf_e(x, e);
}
__device__ __forceinline__ float gelu_erf(float x) { return x * normal_cdf(x); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;
  const float cdf = normal_cdf_e(x, e);
  return fmaf(x * kInvSqrt2Pi, e, cdf);
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
__device__ __forceinline__ float2 normal_cdf_e2(float2 x, float2& e) {
  const float2 d = __ffma2_rn(make_float2(fabsf(x.x), fabsf(x.y)), f2(kErfP), f2(1.0f));
  const float2 t = make_float2(fast_rcp(d.x), fast_rcp(d.y));
  const float2 arg = __fmul2_rn(__fmul2_rn(x, x), f2(kNegHalfLog2e));
  e = make_float2(fast_ex2(arg.x), fast_ex2(arg.y));
  float2 q = __ffma2_rn(t, f2(kErfA5), f2(kErfA4));
  q = __ffma2_rn(q, t, f2(kErfA3));
  q = __ffma2_rn(q, t, f2(kErfA2));
  q = __ffma2_rn(q, t, f2(kErfA1));
  const float2 h = __fmul2_rn(__fmul2_rn(q, t), e);            // 0.5 * erfc(|x| / sqrt 2)
  const float2 r = __ffma2_rn(h, f2(-1.0f), f2(0.5f));
  return __fadd2_rn(f2(0.5f), make_float2(copysignf(r.x, x.x), copysignf(r.y, x.y)));
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
    float2 e;
    const float2 cdf = normal_cdf_e2(x, e);
    y = __fmul2_rn(x, cdf);
    if constexpr (GRAD) dy = __ffma2_rn(__fmul2_rn(x, f2(kInvSqrt2Pi)), e, cdf);
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
// Same as MODE 1 / 2, but the saved tensor leaves straight for a TMA-store staging slot (32 rows x 64 B, SWIZZLE_64B; this
// lane's row): each 16-byte group is packed and stored as soon as its 4 pairs are done, so the 32 registers of `a` never
// exist and the compiler can interleave more of the 16 independent chains (the fused GELU epilogue is latency-bound).
template <int ACT, int MODE>
__device__ __forceinline__ void act_chunk_to_slot_t(float2 (&v)[16], uint8_t* slot_row, int sw) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 s4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 y, dy;
      act_and_grad2<ACT, MODE == 2>(v[4 * j + q], y, dy);
      s4[q] = MODE == 2 ? dy : v[4 * j + q];
      v[4 * j + q] = y;
    }
    __nv_bfloat162 b0 = __floats2bfloat162_rn(s4[0].x, s4[0].y), b1 = __floats2bfloat162_rn(s4[1].x, s4[1].y),
                   b2 = __floats2bfloat162_rn(s4[2].x, s4[2].y), b3 = __floats2bfloat162_rn(s4[3].x, s4[3].y);
    *reinterpret_cast<uint4*>(slot_row + ((j ^ sw) << 4)) =
        make_uint4(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1), *reinterpret_cast<uint32_t*>(&b2),
                   *reinterpret_cast<uint32_t*>(&b3));
  }
}
template <int MODE>
__device__ __forceinline__ void act_chunk_to_slot(int act, float2 (&v)[16], uint8_t* slot_row, int sw) {
  if (act == AVT_ACT_GELU_ERF) act_chunk_to_slot_t<AVT_ACT_GELU_ERF, MODE>(v, slot_row, sw);
  else if (act == AVT_ACT_GELU_TANH) act_chunk_to_slot_t<AVT_ACT_GELU_TANH, MODE>(v, slot_row, sw);
  else act_chunk_to_slot_t<AVT_ACT_NONE, MODE>(v, slot_row, sw);
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
