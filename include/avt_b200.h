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
/* Cap the number of SMs the persistent kernels (GEMM) occupy; 0 = all. Used while an NCCL all-reduce runs
 * concurrently: a persistent grid that assumes every SM is free is stretched 2x by the SMs the collective holds. */
int avt_set_sm_limit(int n);
/* Programmatic dependent launch (default on; environment AVT_PDL=0 or avt_set_pdl(0) turns it off). Every kernel of
 * the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and executes griddepcontrol.wait
 * before it touches global memory, so the ~550 dependent launches of a training step overlap their launch latency and
 * prologue (barrier init, TMEM allocation, tensor-map prefetch) with the tail of the previous kernel. */
int avt_set_pdl(int enable);

/* Fused-epilogue description for avt_gemm_bf16. All pointers may be NULL (feature off).
 * Per output element (r, c), in this order:
 *   v = alpha * acc
 *   v += bias[c]
 *   if pos_period > 0:  t = r % pos_period;  if (t == 0 && cls) v = cls[c];  v += pos[t * N + c]
 *   if aux_z:   aux_z[r * ldz + c] = bf16(aux_mode == 1 ? act'(v) : v)      (saved for backward)
 *   v = act(v)
 *   if dact_z:  v *= (dact_mode == 1 ? dact_z[r * ldz + c] : dact'(dact_z[r * ldz + c]))
 *               (backward through the activation: dact_z holds either the pre-activation, differentiated here
 *                with kind `dact`, or the derivative itself as saved by a forward with aux_mode = 1)
 *   if drop_p > 0: v = keep(seed, drop_offset [+ *drop_offset_dev], r * N + c) ? v / (1 - drop_p) : 0
 *               (drop_offset_dev: optional device-resident part of the Philox offset, so that a captured CUDA graph
 *                draws a fresh mask on every replay)
 *   if residual: v += residual[r * ldr + c]
 *   out[r * ldo + c] = out_fp32 ? v : bf16(v)
 * With split_k > 1 the partial sums are combined
 *  - with fp32 atomics directly into `out` when the epilogue is a plain fp32 output and no workspace is given (weight
 *    gradients; zero-filled first unless `accumulate` is set). The k-range is cut uniformly when tiles x split_k fill one
 *    round of the grid, otherwise into one contiguous tiles x k-blocks range per CTA pair (stream-K);
 *  - otherwise (fused epilogue; the weight-streaming M <= 128 GEMMs of AVT-h, where one tile per CTA would leave most SMs
 *    idle; the caller passes a split_k*M*N fp32 `workspace`): split factors 2 / 4 / 8 of single-CTA tiles run as ONE
 *    thread-block cluster per output tile and reduce through distributed shared memory inside the kernel, the epilogue
 *    applied in place (fixed summation order: bit-reproducible; the workspace stays untouched); any other factor
 *    writes one fp32 slice per split into the workspace and a small finishing kernel sums them and applies the
 *    epilogue. */
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
  int32_t aux_mode;
  int32_t dact_mode;
  float alpha;
  float drop_p;
  uint64_t drop_seed;
  uint64_t drop_offset;
  const uint64_t* drop_offset_dev; /* device pointer or NULL */
  void* out;
  int64_t ldo;
  int32_t out_fp32;
  int32_t accumulate; /* fp32 out only: out += v (used for weight gradients / split-K) */
} avt_epilogue_t;

/* C[M,N] = epilogue(A[M,K] * B[N,K]^T), bf16 operands, fp32 accumulation in tensor memory
 * (tcgen05.mma, TMA-fed 128B-swizzled smem ring, persistent tile loop). block_n in {0 = auto, 64, 128, 256};
 * cta_group in {0 = auto, 1, 2}: 2 runs CTA pairs (tcgen05 cta_group::2, 256 x block_n tiles) for large problems.
 *   a_mn = 0: A stored [M rows][K cols] (ld = lda);  a_mn = 1: A stored transposed, [K rows][M cols].
 *   b_mn = 0: B stored [N rows][K cols] (ld = ldb);  b_mn = 1: B stored transposed, [K rows][N cols].
 * Replaces torch.nn.Linear / F.linear (timm Attention.qkv/proj, Mlp.fc1/fc2; reference
 * models/future_prediction.py:80-81 encoder/decoder), HF Conv1D (torch.addmm; GPT2Attention.c_attn /
 * c_proj, GPT2MLP.c_fc / c_proj) and their autograd dgrad / wgrad matmuls (func/train.py:222). */
int avt_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int64_t M, int64_t N,
                  int64_t K, const avt_epilogue_t* ep, int split_k, int block_n, int cta_group, void* workspace,
                  int64_t workspace_bytes, void* stream);

/* The three hot ViT epilogues ([+bias] -> bf16; [+bias] + erf-GELU + saved derivative; x saved derivative) have
 * compile-time specialised kernels (no run-time epilogue branches); 0 forces the generic kernel everywhere (A/B runs,
 * and the parity tests run both). Default 1. */
int avt_set_gemm_specialized_epilogues(int enable);

/* Number of CUDA kernels this library has launched so far in this process (every launch goes through one helper): what
 * bench.py reports as `gpu_launches` (delta over the timed region; for a captured graph: launches per replay x replays). */
long long avt_kernel_launch_count(void);

/* Same GEMM, plus a_colsum[m] += sum_k A[m, k] (fp32 [M], atomics; NULL = off). Requires a_mn = 1. In a weight-gradient
 * GEMM dW = dY^T X the A operand is dY^T, so a_colsum is the bias gradient (column sums of dY): two extra warps add up
 * the A tiles that are in shared memory for the tensor core anyway, and the separate avt_colsum_bf16 pass over
 * dY (97 MB for timm Mlp.fc1 at the BASELINE shape) disappears. Replaces autograd's bias-gradient reduction of
 * torch.nn.Linear. */
int avt_gemm_bf16_colsum(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int64_t M, int64_t N,
                         int64_t K, const avt_epilogue_t* ep, int split_k, int block_n, int cta_group, void* workspace,
                         int64_t workspace_bytes, float* a_colsum, void* stream);

/* y[r,:] = LayerNorm(x'[r,:]) * gamma + beta over the last dim D (<= 2048, multiple of 4); one warp per row.
 * x' = x, or — residual update fused in — x' = x + add_bf16 (a bf16 branch output), with x' also written to
 * x_out (fp32, may be NULL): `x = x + drop_path(attn(...))` / `x = x + mlp(...)` of timm Block.forward and the
 * two residual adds of HF GPT2Block.forward happen here instead of in a GEMM epilogue, in coalesced rows.
 * x fp32 with row stride x_stride (elements) — a stride of tokens*D selects one token per frame (the CLS
 * row for timm VisionTransformer.norm + x[:, 0]). y is bf16 (y_fp32 = 0) or fp32. mean / rstd ([rows],
 * may be NULL) are saved for backward. Replaces torch.nn.LayerNorm in timm Block.norm1/norm2,
 * VisionTransformer.norm (eps 1e-6) and HF GPT2Block.ln_1/ln_2, GPT2Model.ln_f (eps 1e-5). */
int avt_layernorm_fwd(const float* x, int64_t x_stride, const void* add_bf16, int64_t add_stride, float* x_out,
                      int64_t x_out_stride, const float* gamma, const float* beta, float eps, int64_t rows, int D, void* y,
                      int y_fp32, int64_t y_stride, float* mean, float* rstd, void* stream);

/* LayerNorm backward. dx_out = (dx_in ? dx_in : 0) + dLN(dy); optional bf16 copy of dx_out (the A operand
 * of the next dgrad / wgrad GEMM); dgamma / dbeta are overwritten or accumulated. dx_colsum ([D], may be NULL)
 * receives the column sums of dx_out: dx_out is the gradient of the residual stream, i.e. of the output of the
 * Linear that closed the previous residual branch (timm Attention.proj / Mlp.fc2), so this IS that layer's bias
 * gradient and the separate reduction pass over [rows, D] disappears. `workspace` must hold
 * avt_layernorm_bwd_workspace_bytes(rows, D) bytes. Large row counts stream through a per-warp shared-memory
 * ring filled by bulk async copies; <= 2 x SM-count rows use one CTA per row. Replaces autograd of
 * torch.nn.LayerNorm + the residual-branch gradient add. */
int64_t avt_layernorm_bwd_workspace_bytes(int64_t rows, int D);
int avt_layernorm_bwd(const void* dy, int dy_fp32, int64_t dy_stride, const float* x, int64_t x_stride,
                      const float* mean, const float* rstd, const float* gamma, int64_t rows, int D, const float* dx_in,
                      float* dx_out, int64_t dx_stride, void* dx_bf16, int64_t dxb_stride, float* dgamma, float* dbeta,
                      float* dx_colsum, int accumulate, void* workspace, int64_t workspace_bytes, void* stream);

/* dst[i] = bf16(src[i]), i < n (weight down-cast of the fp32 master parameters). */
int avt_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);

/* Zero `bytes` bytes at dst (a memset node on the stream / in a captured graph): the flat gradient buffers are zeroed once per
 * backward because the split-K weight gradients and the bias gradients accumulate with atomics. */
int avt_zero(void* dst, int64_t bytes, void* stream);

/* Re-tile frames for the patch-embedding GEMM: video fp32 [F, C, H, W] -> out bf16 [F*(P+1), C*ps*ps],
 * P = (H/ps)*(W/ps); row f*(P+1) is zero (CLS slot), row f*(P+1)+1+p is patch p flattened as (c, kh, kw),
 * the K order of Conv2d.weight.view(D, -1). Replaces the im2col inside timm PatchEmbed.proj (Conv2d with
 * stride == kernel) + flatten(2).transpose(1, 2). */
int avt_patchify_bf16(const float* video, void* out, int F, int C, int H, int W, int ps, void* stream);

/* out[c] += sum_r x[r, c]  (x bf16 [rows, ld]): bias gradients of nn.Linear / Conv1D. */
int avt_colsum_bf16(const void* x, int64_t rows, int cols, int64_t ld, float* out, void* stream);

/* s[t, :] = sum_f dx[f*period + t, :] (dx fp32 [F*period, D]); then dpos (+)= s, dcls (+)= s[0],
 * dbias (+)= sum_{t>=1} s[t] (each may be NULL). Gradients of timm pos_embed / cls_token /
 * patch_embed.proj.bias, and of HF wpe rows (period = T). workspace: period*D floats. */
int avt_frame_sum_grads(const float* dx, int F, int period, int D, float* dpos, float* dcls, float* dbias,
                        int accumulate, float* workspace, void* stream);

/* y = keep(seed, offset, i) ? x[i] / (1-p) : 0 over a dense tensor of n (multiple of 4) elements; the same
 * (seed, offset) reproduces the mask a GEMM epilogue applied to a dense [M, N] output (i = r*N + c).
 * Backward of torch.nn.Dropout (HF embd / resid dropout). Either output may be NULL. offset_dev (device pointer,
 * may be NULL) is added to `offset`, as for avt_epilogue_t.drop_offset_dev. */
int avt_dropout_apply(const float* x, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                      float* y_f32, void* y_bf16, void* stream);

/* Multi-head attention on CUDA cores for short sequences / wide heads (AVT-h: T <= 16, head_dim 256..1024).
 * qkv bf16 [B*N, 3*H*hd], column = s*H*hd + h*hd + d (timm Attention.qkv and HF c_attn packing);
 * out bf16 [B*N, H*hd]; lse fp32 [B*H, N]. causal != 0 masks keys j > i. Dropout (drop_p) is applied to the
 * softmax probabilities with element index ((b*H + h)*N + i)*N + j.
 * Replaces HF GPT2Attention._attn (matmul, /sqrt(hd), causal mask, softmax, attn_dropout, matmul) and its
 * autograd backward; also usable for timm Attention (N = 197, hd = 64, causal = 0). */
int avt_attention_simt_fwd(const void* qkv, void* out, float* lse, int B, int H, int N, int hd, int causal, float scale,
                           float drop_p, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* stream);
/* Backward: `out` is the forward output (delta_i = dO_i . O_i, so query-row and key-row work items are independent). */
int avt_attention_simt_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int H,
                           int N, int hd, int causal, float scale, float drop_p, uint64_t seed, uint64_t offset,
                           const uint64_t* offset_dev, void* stream);

/* ViT spatial attention on tcgen05 tensor cores: out = softmax(Q K^T * scale) V per (frame, head), for
 * N <= 208 tokens per frame and head_dim 64 (ViT-B/16, ViT-L/16: N = 197). Same qkv / out / lse layout as
 * avt_attention_simt_*; no mask, no dropout (timm attn_drop = 0). S and O live in tensor memory, the
 * 197x197 score matrix never reaches HBM. Replaces timm Attention.forward's q@k^T -> softmax -> attn@v. */
int avt_attention_tc_fwd(const void* qkv, void* out, float* lse, int F, int H, int N, float scale, void* stream);
/* Backward of the above: dqkv bf16 [F*N, 3*H*64] from qkv, the saved forward output `out`, dout and lse.
 * Five tcgen05 contractions per (key tile, query half); P and dS are recomputed, never stored in HBM.
 * Replaces autograd through timm Attention's matmul / softmax / matmul. */
int avt_attention_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int F, int H,
                         int N, float scale, void* stream);

/* One KV-cached decode step of the causal attention (evaluation-time rollout, models/future_prediction.py:168-202: HF
 * GPT2Model called with past_key_values): qkv_cache bf16 [B*N, 3*H*hd] is the packed qkv buffer of the prefill pass with the
 * new token's q/k/v written into row q_row of every batch item; only that row queries keys 0..q_row. out bf16 [B*N, H*hd]
 * (row q_row of every item is written). */
int avt_attention_simt_decode(const void* qkv_cache, void* out, int B, int H, int N, int hd, int q_row, float scale, void* stream);

/* GPU input pipeline: decoded (and resized) uint8 frames [F, Hin, Win, 3] -> normalised fp32 crops [F, 3, h, w], the tail of the
 * reference's transform chain in one pass: ToTensorVideo, RandomHorizontalFlipVideo (flip[f] != 0; NULL = none), * scale_pix_val,
 * NormalizeVideo(mean, std), crop at (crop_y, crop_x) (func/train.py:550-584, common/transforms.py:124-191,327-396).
 * mean3 / std3 are HOST arrays of three floats. The step then copies uint8 frames to the device (4x fewer bytes). */
int avt_preprocess_u8(const uint8_t* frames, int F, int Hin, int Win, float* out, int h, int w, int crop_y, int crop_x,
                      const uint8_t* flip, float scale, const float* mean3, const float* std3, void* stream);

/* fp32-accuracy mode (inference; BASELINE.json north star: "within 1e-5 (fp32)"): the same operators on the CUDA cores in
 * fp32, exact erf / tanh GELU. A validation path for the restated arithmetic, not the product path.
 * avt_sgemm_f32: out[m,n] = act(sum_k A[m,k] * Bop[k,n] + bias[n]) + residual[m,n]; B stored [N,K] (b_kn = 0, nn.Linear)
 * or [K,N] (b_kn = 1, HF Conv1D); pos / cls / pos_period as in avt_epilogue_t (patch / frame embedding).
 * avt_attention_f32_fwd: softmax(q k^T * scale [causal]) v from packed fp32 qkv [B*N, 3*H*hd] -> out [B*N, H*hd].
 * avt_patchify_f32: video fp32 [F,C,H,W] -> patch rows [F*(P+1), C*patch*patch] (zero rows in the cls slots). */
int avt_sgemm_f32(const float* A, int64_t lda, const float* B, int64_t ldb, int b_kn, int M, int N, int K, const float* bias,
                  const float* residual, int64_t ldr, int act, const float* pos, const float* cls, int pos_period, float* out,
                  int64_t ldo, void* stream);
int avt_attention_f32_fwd(const float* qkv, float* out, int B, int H, int N, int hd, int causal, float scale, void* stream);
int avt_patchify_f32(const float* video, float* out, int F, int C, int H, int W, int patch, void* stream);

/* Row-wise softmax cross-entropy over classifier logits, forward and gradient in one pass (one CTA per row, the row in
 * registers): loss[r] = logsumexp(l) - l[target[r]] (0 when target < 0: nn.CrossEntropyLoss(ignore_index=-1,
 * reduction='none'), loss_fn/multidim_xentropy.py:10-25 via func/train_eval_ops.py:57-85), rank[r] = number of classes
 * with a larger logit than the target's (top-k correct <=> rank < k: common/utils.py:17-44; `classes` for ignored rows),
 * dlogits[r] (bf16, may be NULL; columns [classes, classes_padded) zeroed) = (softmax(l) - onehot) * row_scale[r].
 * logits fp32 [rows, classes] with row stride ld; classes <= 4096. Replaces ~40 ATen launches of the reference step. */
int avt_softmax_xent(const float* logits, int64_t ld, int rows, int classes, const int64_t* target, const float* row_scale,
                     float* loss, int* rank, void* dlogits_bf16, int64_t ldd, int classes_padded, void* stream);

/* One SGD-with-momentum step over a flat parameter buffer (or one rank's shard of it), torch.optim.SGD semantics with
 * dampening 0 (conf/opt/optimizer/sgd.yaml, expts/01:26-28: momentum 0.9, nesterov): p, m updated in place, and the bf16
 * copy of p read by the GEMMs (p_bf16, may be NULL) refreshed in the same pass.
 *   g / g_is_bf16      gradients, fp32 or bf16 (the data-parallel gradient payload is bf16)
 *   weight_decay_lo    weight decay of elements [0, lo_elems): the reference's bias / bn parameter group, whose decay is
 *                      scaled by opt.bias_bn_wd_scale (func/train.py:704-731); the flat buffers keep biases first
 *   lr_dev             optional device scalar overriding lr (per-iteration lr schedules under a captured CUDA graph)
 * Replaces optimizer.step() (func/train.py:233) for the flat AVT-b / AVT-h buffers. */
int avt_sgd_step(float* p, const void* g, int g_is_bf16, float* m, void* p_bf16, int64_t n, float lr, const float* lr_dev,
                 float momentum, float weight_decay, float weight_decay_lo, int64_t lo_elems, int nesterov, int first_step,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVT_B200_H_ */
