// fp32-accuracy mode of the hot path (inference): every contraction in fp32 on the CUDA cores, exact erf / tanh GELU.
// BASELINE.json's north star states two tolerances - 1e-3 relative for the bf16 tensor-core path and 1e-5 for fp32; the
// bf16 path cannot get closer than a few 1e-3 to an fp64 oracle end to end (every GEMM operand is stored in bf16), so this
// mode exists to pin the arithmetic of the restated model itself (operator order, LayerNorm eps, GELU flavour, qkv packing,
// causal mask, cls/pos handling) at 1e-5 on the GPU. It is a validation path, not the product path: plain tiled kernels,
// ~15 TF/s. Replaces the same reference code as the bf16 kernels: timm VisionTransformer / HF GPT2Model forward
// (models/video_classification.py:255-256, models/future_prediction.py:163-190).
#include "common.cuh"

namespace avt {

// ----------------------------------------------------------------------------- SGEMM with fused bias / activation / residual
// out[m, n] = act(sum_k A[m, k] * Bop[k, n] + bias[n]) + residual[m, n];  B stored [N, K] (nn.Linear) or [K, N] (HF Conv1D)
constexpr int kSgBM = 64, kSgBN = 64, kSgBK = 16;

template <bool B_KN>
__global__ void __launch_bounds__(256)
sgemm_f32_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, int M, int N, int K,
                 const float* __restrict__ bias, const float* __restrict__ residual, int64_t ldr, int act,
                 const float* __restrict__ pos, const float* __restrict__ cls, int pos_period, float* __restrict__ out, int64_t ldo) {
  pdl_enter();
  __shared__ float sA[kSgBK][kSgBM + 1];
  __shared__ float sB[kSgBK][kSgBN + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * kSgBM, n0 = blockIdx.x * kSgBN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += kSgBK) {
    for (int i = threadIdx.x; i < kSgBM * kSgBK; i += 256) {
      const int r = i / kSgBK, c = i % kSgBK;
      sA[c][r] = (m0 + r < M && k0 + c < K) ? A[(int64_t)(m0 + r) * lda + k0 + c] : 0.f;
    }
    for (int i = threadIdx.x; i < kSgBN * kSgBK; i += 256) {
      if (B_KN) {
        const int c = i / kSgBN, r = i % kSgBN;      // B[k, n]: n contiguous
        sB[c][r] = (n0 + r < N && k0 + c < K) ? B[(int64_t)(k0 + c) * ldb + n0 + r] : 0.f;
      } else {
        const int r = i / kSgBK, c = i % kSgBK;      // B[n, k]: k contiguous
        sB[c][r] = (n0 + r < N && k0 + c < K) ? B[(int64_t)(n0 + r) * ldb + k0 + c] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSgBK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (pos_period > 0) {          // patch / frame embedding: + pos[token]; the cls rows are cls + pos[0] (their A rows are zero)
        const int t = m % pos_period;
        if (cls && t == 0) v = cls[n] + pos[n];
        else v += pos[(int64_t)t * N + n];
      }
      if (act == AVT_ACT_GELU_ERF) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
      else if (act == AVT_ACT_GELU_TANH) v = 0.5f * v * (1.0f + tanhf(0.79788456080286536f * (v + 0.044715f * v * v * v)));
      if (residual) v += residual[(int64_t)m * ldr + n];
      out[(int64_t)m * ldo + n] = v;
    }
  }
}

// ----------------------------------------------------------------------------- attention, fp32, one CTA per (batch, head)
__global__ void __launch_bounds__(256)
attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out, int H, int N, int hd, int Dm, int causal, float scale) {
  pdl_enter();
  extern __shared__ float sm[];
  float* sK = sm;                       // [N][hd]
  float* sV = sK + (size_t)N * hd;      // [N][hd]
  float* sS = sV + (size_t)N * hd;      // [8 warps][N]
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ld = 3 * (int64_t)Dm;
  const float* base = qkv + (int64_t)b * N * ld + h * hd;
  for (int i = threadIdx.x; i < N * hd; i += blockDim.x) {
    const int r = i / hd, c = i % hd;
    sK[i] = base[(int64_t)r * ld + Dm + c];
    sV[i] = base[(int64_t)r * ld + 2 * Dm + c];
  }
  __syncthreads();
  float* s = sS + warp * N;
  for (int i = warp; i < N; i += 8) {
    const float* q = base + (int64_t)i * ld;
    const int jmax = causal ? i + 1 : N;
    float mx = -INFINITY;
    for (int j = 0; j < jmax; ++j) {
      float a = 0.f;
      for (int d = lane; d < hd; d += 32) a = fmaf(q[d], sK[j * hd + d], a);
      a = warp_sum(a) * scale;
      if (lane == 0) s[j] = a;
      mx = fmaxf(mx, a);
    }
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j < jmax; j += 32) {
      const float e = expf(s[j] - mx);
      s[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int d = lane; d < hd; d += 32) {
      float a = 0.f;
      for (int j = 0; j < jmax; ++j) a = fmaf(s[j], sV[j * hd + d], a);
      out[((int64_t)b * N + i) * Dm + h * hd + d] = a * inv;
    }
    __syncwarp();
  }
}

// patch rows for the fp32 patch-embedding GEMM: video fp32 [F, C, H, W] -> [F * (P + 1), C * p * p], zero rows in the cls slots
__global__ void __launch_bounds__(256)
patchify_f32_kernel(const float* __restrict__ video, float* __restrict__ out, int F, int C, int Hh, int W, int patch) {
  pdl_enter();
  const int gw = W / patch, gh = Hh / patch, P = gw * gh, Kp = C * patch * patch;
  const int64_t total = (int64_t)F * (P + 1) * Kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = i % Kp;
    const int64_t row = i / Kp;
    const int t = row % (P + 1), f = row / (P + 1);
    float v = 0.f;
    if (t > 0) {
      const int c = k / (patch * patch), py = (k / patch) % patch, px = k % patch;
      const int gy = (t - 1) / gw, gx = (t - 1) % gw;
      v = video[(((int64_t)f * C + c) * Hh + gy * patch + py) * W + gx * patch + px];
    }
    out[i] = v;
  }
}

}  // namespace avt

using namespace avt;

extern "C" int avt_sgemm_f32(const float* A, int64_t lda, const float* B, int64_t ldb, int b_kn, int M, int N, int K,
                             const float* bias, const float* residual, int64_t ldr, int act, const float* pos, const float* cls,
                             int pos_period, float* out, int64_t ldo, void* stream) {
  AVT_REQUIRE(A && B && out, "null pointer");
  AVT_REQUIRE(M > 0 && N > 0 && K > 0, "empty problem");
  AVT_REQUIRE(pos_period == 0 || pos, "pos_period needs pos");
  const dim3 grid((N + kSgBN - 1) / kSgBN, (M + kSgBM - 1) / kSgBM);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (b_kn) launch_kernel(sgemm_f32_kernel<true>, grid, dim3(256), 0, st, A, lda, B, ldb, M, N, K, bias, residual, ldr, act, pos, cls, pos_period, out, ldo);
  else launch_kernel(sgemm_f32_kernel<false>, grid, dim3(256), 0, st, A, lda, B, ldb, M, N, K, bias, residual, ldr, act, pos, cls, pos_period, out, ldo);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

extern "C" int avt_attention_f32_fwd(const float* qkv, float* out, int B, int H, int N, int hd, int causal, float scale,
                                     void* stream) {
  AVT_REQUIRE(qkv && out, "null pointer");
  const size_t smem = ((size_t)2 * N * hd + 8 * (size_t)N) * sizeof(float);
  AVT_REQUIRE(smem <= 220 * 1024, "sequence x head_dim does not fit shared memory");
  static size_t configured = 0;
  if (smem > configured) {
    AVT_CUDA_OK(cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  launch_kernel(attention_f32_kernel, dim3(B * H), dim3(256), smem, reinterpret_cast<cudaStream_t>(stream), qkv, out, H, N, hd,
                H * hd, causal, scale);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}

extern "C" int avt_patchify_f32(const float* video, float* out, int F, int C, int H, int W, int patch, void* stream) {
  AVT_REQUIRE(video && out, "null pointer");
  AVT_REQUIRE(patch > 0 && H % patch == 0 && W % patch == 0, "image size must be a multiple of the patch size");
  launch_kernel(patchify_f32_kernel, dim3(num_sms() * 8), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), video, out, F, C,
                H, W, patch);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}
