// Classifier loss head: row-wise softmax cross-entropy over the classifier logits, forward AND gradient in one pass.
// Reference: `models/base_model.py:203-216` (dropout -> Linear 768->3806 on the T past + 1 future rows of every clip),
// `func/train_eval_ops.py:57-85` + `loss_fn/multidim_xentropy.py:10-25` (CrossEntropyLoss(ignore_index=-1,
// reduction='none')) and `common/utils.py:17-44` (top-1 / top-5 accuracy of the future logits, every iteration).
// The reference runs ~40 small ATen / cutlass-simt launches for this per step (log_softmax fwd/bwd, nll fwd/bwd, topk,
// eq, sums ...); here the two GEMMs around it are the library's tcgen05 GEMM and everything between them is this kernel:
// one CTA per row keeps the row's logits in registers, and produces
//   loss[r]    = logsumexp(l) - l[target]                       (0 for target < 0: ignored)
//   dlogits[r] = (softmax(l) - onehot(target)) * row_scale[r]   bf16, the A operand of the weight / input gradient GEMMs
//   rank[r]    = #{c : l[c] > l[target]}                        (top-k correct <=> rank < k)
#include "common.cuh"

namespace avt {

constexpr int kXentThreads = 256;
constexpr int kXentPerThread = 16;   // classes per thread held in registers: C <= 4096

__global__ void __launch_bounds__(kXentThreads)
softmax_xent_kernel(const float* __restrict__ logits, int64_t ld, int C, const int64_t* __restrict__ target,
                    const float* __restrict__ row_scale, float* __restrict__ loss, int* __restrict__ rank,
                    bf16* __restrict__ dlogits, int64_t ldd, int Cpad) {
  pdl_enter();
  __shared__ float red[kXentThreads / 32];
  __shared__ int redi[kXentThreads / 32];
  const int r = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const float* row = logits + (size_t)r * ld;
  const int64_t tg = target[r];
  float v[kXentPerThread];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kXentPerThread; ++i) {
    const int c = t + i * kXentThreads;
    v[i] = c < C ? __ldg(row + c) : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < kXentThreads / 32; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  const float tl = (tg >= 0 && tg < C) ? __ldg(row + tg) : INFINITY;   // target logit (ignored rows: nothing is greater)
  float s = 0.f;
  int above = 0;
#pragma unroll
  for (int i = 0; i < kXentPerThread; ++i) {
    const float e = __expf(v[i] - mx);     // exp(-inf) = 0 for the padding
    s += e;
    above += v[i] > tl ? 1 : 0;
    v[i] = e;
  }
  s = warp_sum(s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
  if (lane == 0) {
    red[warp] = s;
    redi[warp] = above;
  }
  __syncthreads();
  s = 0.f;
  above = 0;
#pragma unroll
  for (int w = 0; w < kXentThreads / 32; ++w) {
    s += red[w];
    above += redi[w];
  }
  const bool valid = tg >= 0 && tg < C;
  if (t == 0) {
    loss[r] = valid ? (mx + __logf(s)) - tl : 0.f;
    if (rank) rank[r] = valid ? above : C;
  }
  if (dlogits) {
    const float g = valid ? row_scale[r] : 0.f;
    const float inv = g / s;
    bf16* drow = dlogits + (size_t)r * ldd;
#pragma unroll
    for (int i = 0; i < kXentPerThread; ++i) {
      const int c = t + i * kXentThreads;
      if (c < Cpad) drow[c] = __float2bfloat16(c < C ? (v[i] * inv - (c == tg ? g : 0.f)) : 0.f);
    }
  }
}

}  // namespace avt

using namespace avt;

extern "C" int avt_softmax_xent(const float* logits, int64_t ld, int rows, int classes, const int64_t* target,
                                const float* row_scale, float* loss, int* rank, void* dlogits_bf16, int64_t ldd,
                                int classes_padded, void* stream) {
  AVT_REQUIRE(logits && target && loss, "null pointer");
  AVT_REQUIRE(classes > 0 && classes <= kXentThreads * kXentPerThread, "1 <= classes <= 4096");
  AVT_REQUIRE(!dlogits_bf16 || (row_scale && classes_padded >= classes && classes_padded <= kXentThreads * kXentPerThread &&
                                ldd >= classes_padded),
              "dlogits needs row_scale and classes <= classes_padded <= min(ldd, 4096)");
  if (rows <= 0) return AVT_OK;
  launch_kernel(softmax_xent_kernel, dim3(rows), dim3(kXentThreads), 0, reinterpret_cast<cudaStream_t>(stream), logits, ld, classes,
                target, row_scale, loss, rank, reinterpret_cast<bf16*>(dlogits_bf16), ldd, classes_padded);
  AVT_CUDA_OK(cudaGetLastError());
  return AVT_OK;
}
