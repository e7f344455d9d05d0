/*
 * avt_b200 — C-ABI of the Blackwell-native AVT hot path (ViT backbone AVT-b + causal GPT-2 head AVT-h).
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (passed as void* so the
 * header needs no CUDA include), allocates nothing, launches hand-written sm_100a kernels on that
 * stream and returns 0 or a negative AVT_ERR_* code; avt_last_error() describes the failure.
 * The reference (facebookresearch/AVT) has no native code: the "interface each entry replaces" is
 * the PyTorch/timm/HF call cited beside it (paths relative to the reference checkout).
 *
 * Matrices are row-major. bf16 = IEEE bfloat16 (uint16 storage). Unless stated, fp32 tensors are
 * float and "rows" is the flattened (frame, token) or (clip, frame) index.
 */
#ifndef AVT_B200_H_
#define AVT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVT_OK 0
#define AVT_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define AVT_ERR_CUDA (-2)    /* CUDA runtime or driver error    */
#define AVT_ERR_NO_GPU (-3)  /* no sm_100 device visible        */

#define AVT_ACT_NONE 0
#define AVT_ACT_GELU_ERF 1  /* torch.nn.GELU(): timm Mlp.act (timm 0.4.12 vision_transformer.py, Mlp)      */
#define AVT_ACT_GELU_TANH 2 /* HF ACT2FN["gelu_new"]: GPT2MLP.act (transformers modeling_gpt2.py, GPT2MLP) */

/* Library / device --------------------------------------------------------------------------- */
int avt_abi_version(void);
const char* avt_last_error(void);
/* 0 if a compute-capability-10.x device is current, AVT_ERR_NO_GPU otherwise. */
int avt_check_device(void);

/* Fused-epilogue description for avt_gemm_bf16. All pointers may be NULL (feature off).
 * Per output element (r, c), in this order:
 *   v = alpha * acc
 *   v += bias[c]
 *   if pos_period > 0:  t = r % pos_period;  if (t == 0 && cls) v = cls[c];  v += pos[t * N + c]
 *   if aux_z:   aux_z[r * ldz + c] = bf16(v)              (pre-activation, saved for backward)
 *   v = act(v)
 *   if dact_z:  v *= dact'(dact_z[r * ldz + c])           (backward through activation kind `dact`)
 *   if drop_p > 0: v = keep(seed, drop_offset, r * N + c) ? v / (1 - drop_p) : 0
 *   if residual: v += residual[r * ldr + c]
 *   out[r * ldo + c] = out_fp32 ? v : bf16(v)
 * With split_k > 1 the only allowed epilogue is fp32 accumulation into `out` (atomic adds). */
typedef struct avt_epilogue {
  const float* bias;
  const float* residual;
  int64_t ldr;
  const void* dact_z; /* bf16 */
  void* aux_z;        /* bf16 */
  int64_t ldz;
  const float* pos;
  const float* cls;
  int32_t pos_period;
  int32_t act;
  int32_t dact;
  float alpha;
  float drop_p;
  uint64_t drop_seed;
  uint64_t drop_offset;
  void* out;
  int64_t ldo;
  int32_t out_fp32;
  int32_t accumulate; /* fp32 out only: out += v (used for weight gradients / split-K) */
} avt_epilogue_t;

/* C[M,N] = epilogue(A[M,K] * B[N,K]^T), bf16 operands, fp32 accumulation in tensor memory
 * (tcgen05.mma, TMA-fed 128B-swizzled smem ring, persistent tile loop).
 *   a_mn = 0: A stored [M rows][K cols] (ld = lda);  a_mn = 1: A stored transposed, [K rows][M cols].
 *   b_mn = 0: B stored [N rows][K cols] (ld = ldb);  b_mn = 1: B stored transposed, [K rows][N cols].
 * Replaces torch.nn.Linear / F.linear (timm Attention.qkv/proj, Mlp.fc1/fc2; reference
 * models/future_prediction.py:80-81 encoder/decoder), HF Conv1D (torch.addmm; GPT2Attention.c_attn /
 * c_proj, GPT2MLP.c_fc / c_proj) and their autograd dgrad / wgrad matmuls (func/train.py:222). */
int avt_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int64_t M, int64_t N,
                  int64_t K, const avt_epilogue_t* ep, int split_k, int block_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVT_B200_H_ */
