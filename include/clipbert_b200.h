/*
 * clipbert_b200.h — C ABI of libclipbert_sm100.so, the B200 (sm_100a) kernels behind the ClipBERT
 * forward/backward hot path.
 *
 * The reference (jayleicn/ClipBERT) has no FFI: its hot path is a stack of torch.nn modules whose
 * arithmetic runs in third-party CUDA libraries (cuDNN, cuBLAS, apex). Each entry point below
 * replaces the library kernels reached from one reference call site (cited as file:line relative to
 * the reference root). Conventions:
 *   - every pointer is a device pointer into caller-owned memory; the library never allocates
 *     device memory, never synchronises, and enqueues all work on the `stream` argument
 *     (a cudaStream_t passed as void*), so every call is CUDA-graph capturable;
 *   - activations / packed weights are bf16 (uint16 storage), statistics / gradients are fp32,
 *     token ids are int64; accumulation is always fp32;
 *   - return 0 on success, a negative cb_status on failure; cb_last_error() gives a thread-local
 *     message. Nothing throws across the ABI. Shape / alignment violations are errors, not UB.
 */
#ifndef CLIPBERT_B200_H_
#define CLIPBERT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cb_status {
  CB_OK = 0,
  CB_ERR_INVALID = -1, /* bad shape / alignment / null pointer */
  CB_ERR_CUDA = -2,    /* CUDA runtime or driver error */
  CB_ERR_UNSUPPORTED = -3
} cb_status;

const char* cb_last_error(void);
/* library version and the SM architecture it was compiled for (100 = sm_100a) */
int cb_version(void);
int cb_sm_arch(void);
/* number of kernels this library has launched since load (process-wide, all threads). */
int64_t cb_launch_count(void);
/* Programmatic dependent launch between this library's kernels (default OFF; env CB_PDL=1 or cb_set_pdl(1) enables):
 * each kernel's prologue overlaps the previous kernel's tail; every kernel issues griddepcontrol.wait before its first
 * global access, so results are identical to plain stream order. Measured on B200 (profiles/r01b_ab_runs.txt): +1.6 %
 * on a single-stream step, but it cancels the +9.6 % of running the wgrad GEMMs on a second stream, because early-
 * launched dependent CTAs hold their 200 KB of shared memory while they wait and keep the other stream off those SMs.
 * enable = 2: every kernel EXCEPT the persistent GEMMs (an LN / attention / column-sum CTA waiting early holds a few KB).
 * Returns the previous setting. Replaces nothing in the reference (its ~400 launches per clip are plain stream-
 * ordered cuDNN/cuBLAS/ATen kernels). */
int cb_set_pdl(int enable);

/* Dropout stream position in DEVICE memory. The reference draws a fresh mask at every nn.Dropout call
 * (src/modeling/transformers.py:170,222,295,375; modeling.py:57,552). Here a mask is a pure function of
 * (seed, element index) and every mask-drawing entry point takes the seed BY VALUE, so a captured CUDA graph would
 * replay the same masks. cb_dropout_offset_bind(word) makes every launch issued AFTERWARDS (process-wide, until the
 * next bind; NULL unbinds) read the 64-bit `word` on the device when it RUNS and use  seed + word * odd_constant
 * instead of seed: the forward and the backward of one step bind the same word and regenerate the same masks, the
 * next replay sees an advanced word and draws new ones. cb_dropout_offset_advance enqueues a one-thread kernel:
 * ++*counter; if (snapshot) *snapshot = *counter - a step advances the model's counter once and binds the per-step
 * snapshot, so a later step (or a second forward before this one's backward) cannot change the masks of this one.
 * Keep decision of element e: 16-bit lane (e & 3) of splitmix64(seed, e >> 2) >= round(p * 65536). */
int cb_dropout_offset_bind(const uint64_t* device_word);
int cb_dropout_offset_advance(uint64_t* counter, uint64_t* snapshot, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tensor-core contraction (tcgen05.mma, TMA operand staging, TMEM accumulators).
 *
 * One descriptor covers every dense contraction on the path:
 *   nn.Linear            src/modeling/transformers.py:218-220,292,357,372,467  modeling.py:534-539
 *   d2 Conv2d 1x1 / 3x3  src/modeling/grid_feat.py:95 (detectron2 ResNet), :43-48 (grid_encoder)
 *   and their autograd dgrad / wgrad.
 *
 * mode CB_GEMM_TN : out[M,N] = epi( sum_t A[m + shift_t, 0:K] . B[n, t*K : (t+1)*K] )
 *     A: bf16 [a_rows, K]  (row stride a_ld)   B: bf16 [N, ntaps*K] (row stride b_ld)
 *     ntaps = 1 for Linear / 1x1 conv. ntaps = 9 is a 3x3 pad-1 conv over a zero-bordered
 *     ("padded") NHWC activation whose rows are flat pixels p = (img*(H+2) + y)*(W+2) + x;
 *     shift_t = tap_sign * ((t/3 - 1)*(W+2) + (t%3 - 1)). tap_sign = -1 gives the dgrad conv.
 *     ntaps = 4 ("row taps", TN only): shift_t = tap_sign * t * tap_w - the space-to-depth stem (cb_stem_s2d).
 * mode CB_GEMM_NN : as TN but B is stored [K, ntaps*N] row-major (the forward weight [out=K, in=N]
 *     read "MN-major"): out[m, n] = epi( sum_t sum_k A[m + shift_t, k] * B[k, t*N + n] ). This is
 *     the dgrad of Linear / conv straight from the forward weight layout (no transposed copy).
 * mode CB_GEMM_WGRAD : out[m, t*N + n] += rowscale[m] * sum_p A[p, m] * B[p + shift_t, n]
 *     A: bf16 [P, M] (dY), B: bf16 [P, N] (X); both operands are read "MN-major" straight from
 *     the activation layout; fp32 red.global.add accumulation (split over P across grid.z).
 *
 * Epilogue (TN), applied in this order on the fp32 accumulator v of element (m, n):
 *     v = v * scale[n] + shift[n]        (FrozenBN affine / bias; either may be NULL)
 *     v = dropout(v)                     (if dropout_p > 0; counter RNG keyed by seed + out index)
 *     v += residual[m, n]                (bf16, optional)
 *     if out2: out2[row, n] = v          (bf16 pre-activation stash, optional; gelu'(v) for CB_ACT_GELU_STASH_GRAD)
 *     v = act(v)                         (none / relu / gelu(erf) / tanh)
 *     v *= auxfn(aux[m, n])              (backward masks: relu' , gelu', tanh' ; optional)
 *     out[row(m), n] = v                 (bf16 or fp32)
 * row(m) re-maps between compact NHWC pixel rows and zero-bordered rows (cb_rowmap).
 * ------------------------------------------------------------------------------------------ */
enum { CB_GEMM_TN = 0, CB_GEMM_WGRAD = 1, CB_GEMM_NN = 2 };
enum {
  CB_ACT_NONE = 0, CB_ACT_RELU = 1, CB_ACT_GELU = 2, CB_ACT_TANH = 3,
  CB_ACT_GELU_STASH_GRAD = 4 /* out = gelu(v) and, instead of the pre-activation, out2 = gelu'(v) (bf16): the erf and the
                                exp(-v^2/2) are shared, and the backward of BertIntermediate (transformers.py:363-366)
                                becomes CB_AUX_MUL - one multiply - instead of re-evaluating erf + exp per element      */
};
enum {
  CB_AUX_NONE = 0,
  CB_AUX_RELU_MASK = 1, /* v *= (aux > 0)              aux = forward output of the ReLU       */
  CB_AUX_GELU_GRAD = 2, /* v *= gelu'(aux)             aux = forward pre-activation            */
  CB_AUX_TANH_GRAD = 3, /* v *= 1 - aux^2              aux = forward tanh output               */
  CB_AUX_MUL = 4        /* v *= aux                    aux = stashed derivative (CB_ACT_GELU_STASH_GRAD) */
};
enum {
  CB_ROWMAP_NONE = 0,
  CB_ROWMAP_PAD = 1,  /* m indexes compact [img,H,W] pixels, output row is the padded pixel      */
  CB_ROWMAP_UNPAD = 2 /* m indexes padded [img,H+2,W+2] pixels, border rows are dropped          */
};

typedef struct cb_gemm_desc {
  int32_t mode;
  int32_t m, n, k; /* TN: rows, cols, per-tap K.  WGRAD: out rows (=A cols), out cols per tap, P */
  const void* a;
  int64_t a_rows, a_ld;
  const void* b;
  int64_t b_rows, b_ld;
  int32_t ntaps, tap_w, tap_sign; /* tap_w = W + 2 (padded row pitch in pixels) */
  int32_t split_k;                /* WGRAD only; 0 = let the library choose, else the number of K splits */
  /* epilogue */
  const float* scale;
  const float* shift;
  const void* residual;
  int64_t res_ld;
  const void* aux;
  int64_t aux_ld;
  int32_t aux_mode;
  int32_t act;
  void* out;
  int64_t out_ld;
  int32_t out_fp32;
  void* out2;
  int64_t out2_ld;
  int32_t rowmap, map_h, map_w; /* spatial size (unpadded) for cb_rowmap */
  float dropout_p;
  uint64_t dropout_seed;
  int32_t block_n;  /* 0 = let the library choose (64 / 128 / 256) */
  int32_t reserved; /* tuning / test knobs: bit0 force the staged (non-TMA) epilogue, bit5 ask for / bit6 forbid the
                       two-CTAs-per-SM instantiation (128 x <=128 tiles), bits 8-11 k-chunks per pipeline stage
                       (0 = automatic); other bits ignored */
} cb_gemm_desc;

int cb_gemm(const cb_gemm_desc* desc, void* stream);
/* n independent CB_GEMM_WGRAD problems in ONE persistent launch: the weight gradients of the four Linear layers of a BertLayer
 * (autograd of transformers.py:238-301,363-381) or of the convs of one bottleneck block. The problems should share their
 * reduction length k (tokens / pixels); 1 <= n <= 8. One prologue and tail instead of n, no K-split when the group fills the SMs.
 * Falls back to n cb_gemm launches for n == 1, n > 8 or very different k. Epilogue fields other than scale / out are ignored. */
int cb_gemm_wgrad_group(const cb_gemm_desc* descs, int n, void* stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm over rows of a bf16 [m, 768] matrix (fp32 statistics, warp-shuffle reductions).
 * Replaces apex FusedLayerNorm (src/modeling/transformers.py:32,165,293,373; modeling.py:12,58).
 * The residual add / dropout that precede it in BertSelfOutput / BertOutput
 * (transformers.py:297-301, 377-381) are fused into the producing cb_gemm epilogue.
 *   fwd : y = (x - mean) * rstd * gamma + beta ; stats[m] = (mean, rstd)
 *   bwd : dx (bf16) ; dx_drop = dx * dropout_mask(seed, element) (bf16, optional: the gradient
 *         entering the dense layer that fed this LN through dropout) ; dgamma / dbeta / dbias_drop
 *         (fp32, atomically accumulated; dbias_drop = column sums of dx_drop = that dense's bias grad)
 * ------------------------------------------------------------------------------------------ */
int cb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats, int m, int hidden,
                     float eps, void* stream);
int cb_layernorm_bwd(const void* dy, const void* x, const float* stats, const float* gamma, void* dx, void* dx_drop,
                     float* dgamma, float* dbeta, float* dbias_drop, int m, int hidden, float dropout_p,
                     uint64_t dropout_seed, void* stream);

/* ------------------------------------------------------------------------------------------
 * Embeddings. out is the bf16 [nseq * l, 768] encoder input; text rows go to positions [0, lt),
 * visual rows to [lt, l) of every sequence, i.e. the torch.cat([text; visual]) of
 * src/modeling/modeling.py:219-223 is a write offset, not a copy.
 *   text   : BertEmbeddings.forward (transformers.py:172-199): LN(word[id] + pos[t] + type[0]), dropout
 *   visual : VisualInputEmbedding.forward (modeling.py:62-101): mean over frames, + row/col position
 *            (modeling.py:124-153), + type[0], LN, dropout; the row gather of repeat_tensor_rows
 *            (src/datasets/data_utils.py:344-357) is fused: sequence s reads video seq2vid[s]
 *            (or s / n_ex when seq2vid is NULL). Tables / LN parameters are fp32.
 *   bwd    : scatter-adds into the fp32 table gradients; the visual backward also reduces over the
 *            sequences of a video and the frame mean, producing dgrid (bf16 [nvid, t, gh*gw, 768]).
 * ------------------------------------------------------------------------------------------ */
int cb_embed_text_fwd(const int64_t* ids, const float* word, const float* pos, const float* type0, const float* gamma,
                      const float* beta, void* out, float* stats, int nseq, int lt, int l, int vocab, int hidden,
                      float eps, float dropout_p, uint64_t seed, void* stream);
int cb_embed_text_bwd(const void* dh, const int64_t* ids, const float* word, const float* pos, const float* type0,
                      const float* gamma, const float* stats, float* dword, float* dpos, float* dtype0, float* dgamma,
                      float* dbeta, int nseq, int lt, int l, int vocab, int hidden, float dropout_p, uint64_t seed,
                      void* stream);
int cb_embed_visual_fwd(const void* grid, const int32_t* seq2vid, int n_ex, const float* rowemb, const float* colemb,
                        const float* type0, const float* gamma, const float* beta, void* out, float* stats, int nseq,
                        int t, int gh, int gw, int lt, int l, int hidden, float eps, float dropout_p, uint64_t seed,
                        void* stream);
int cb_embed_visual_bwd(const void* dh, const void* grid, const int32_t* seq2vid, const int32_t* vid_start, int n_ex,
                        const float* rowemb, const float* colemb, const float* type0, const float* gamma,
                        const float* stats, float* dv_tmp, void* dgrid, float* drow, float* dcol, float* dtype0,
                        float* dgamma, float* dbeta, int nseq, int nvid, int t, int gh, int gw, int lt, int l, int hidden,
                        float dropout_p, uint64_t seed, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused self-attention, BertSelfAttention.forward (transformers.py:230-286):
 *   S = Q K^T / 8 + (1 - mask) * -10000 ; P = softmax(S) (fp32) ; dropout(P) ; ctx = P V ; heads merged.
 * qkv: bf16 [nseq*l, 3*heads*64] (Q | K | V as produced by the fused N=2304 projection);
 * text_mask: int64 [nseq, lt] (visual tokens are always attendable, modeling.py:217-220);
 * ctx: bf16 [nseq*l, heads*64]; lse: fp32 [nseq, heads, l] log-sum-exp saved for the backward.
 * The backward recomputes P tile by tile from Q, K, lse and regenerates the dropout mask.
 * ------------------------------------------------------------------------------------------ */
int cb_attention_fwd(const void* qkv, int64_t ld_qkv, const int64_t* text_mask, void* ctx, int64_t ld_ctx, float* lse,
                     int nseq, int l, int lt, int heads, int head_dim, float dropout_p, uint64_t seed, void* stream);
int cb_attention_bwd(const void* qkv, int64_t ld_qkv, const int64_t* text_mask, const void* ctx, const void* dctx,
                     int64_t ld_ctx, const float* lse, void* dqkv, int64_t ld_dqkv, int nseq, int l, int lt, int heads,
                     int head_dim, float dropout_p, uint64_t seed, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small helpers on the transformer side.
 *   cb_colsum     out[n] += sum_m x[m, n]            bias gradients of nn.Linear (autograd of F.linear)
 *   cb_dropout    y = dropout(x)                     nn.Dropout before the classifier (modeling.py:552)
 *   cb_pad_cast   fp32 [rows, c] -> bf16 [rows, cpad] zero padded (d logits -> padded head gradient)
 *   cb_cast_scale fp32 -> bf16 operand packing, optional per-row scale (FrozenBN scale folded into the
 *                 conv weight: w'[o, :] = w[o, :] * gamma[o] * rsqrt(var[o] + 1e-5), d2 FrozenBatchNorm2d)
 * ------------------------------------------------------------------------------------------ */
int cb_colsum(const void* x, int64_t ld, float* out, int m, int n, void* stream);
int cb_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, void* stream);
/* dx = dy * gelu'(u): backward of BertPredictionHeadTransform's activation (transformers.py:486-495) */
int cb_gelu_bwd(const void* dy, const void* u, void* dx, int64_t n, void* stream);
int cb_pad_cast(const float* in, int64_t in_ld, void* out, int rows, int c, int cpad, void* stream);
int cb_cast_scale(const float* in, const float* rowscale, int64_t row_len, void* out, int64_t n, void* stream);
/* bf16 -> fp32: the averaged gradients come back from the bf16 wire format of the data-parallel exchange (the reference
 * all-reduces its gradients in fp16 under amp O2, src/tasks/run_video_retrieval.py:299-309,432); n elements, 16-byte aligned */
int cb_cast_bf16_f32(const void* in, float* out, int64_t n, void* stream);
/* the same for every conv of the backbone in one launch: segments is a device int64 [nseg][4] table
 * (element offset, element count, row length, offset into `scales` or -1); offsets are 64-element aligned */
int cb_cast_scale_segments(const float* master, void* packed, const int64_t* segments, int nseg, const float* scales, void* stream);

/* ------------------------------------------------------------------------------------------
 * CNN-side data movement (NHWC bf16, 8 channels per 128-bit access). Call sites replaced:
 * GridFeatBackbone.forward (src/modeling/grid_feat.py:89-105) -> detectron2 BasicStem /
 * BottleneckBlock (stride-in-1x1) / grid_encoder MaxPool2d+ReLU (grid_feat.py:43-48).
 *   cb_stem_im2col          7x7/s2/p3 patch gather of the NCHW RGB input into bf16 [n*ho*wo, kp] rows,
 *                           K = (r, s, c) with c in BGR order (the x[:, [2,1,0]] flip of grid_feat.py:92-94
 *                           is folded in); in_dtype 1 = uint8 frames with the ImageNorm mean subtraction
 *                           (src/datasets/data_utils.py:256-276) fused. The stem GEMM follows.
 *   cb_stem_s2d             the stem WITHOUT a patch matrix: space-to-depth(2) of the zero-padded BGR frame,
 *                           S[n, Y, X, (dy*2+dx)*4 + c] (c = 3 is a zero lane), Y < ho+3, X < wo+3. The 7x7/s2/p3 conv
 *                           (kernel zero-extended to 8x8) is then cb_gemm with ntaps = 4, k = 64, tap_w = wo+3 over
 *                           the matrix whose row m is the 64 contiguous bf16 starting at S pixel m: ld = 16 makes the
 *                           rows overlap (a_ld = 16 < k; 54 MB per 128 frames instead of 488 MB of patches), ld = 64
 *                           stores every 4-pixel window explicitly. Output rows follow the same (ho+3) x (wo+3) grid.
 *   cb_maxpool3x3s2         BasicStem max_pool2d(3, 2, 1); _strided reads an input whose pixel rows / images are
 *                           row_pitch / img_pitch pixels apart (the (ho+3) x (wo+3) grid of the s2d stem)
 *   cb_subsample2           input of a stride-2 1x1 conv (res3/4/5 block 0 conv1 + shortcut)
 *   cb_unsubsample2_mask    its backward fused with the ReLU mask of the producing block
 *   cb_maxpool2x2_relu_fwd  grid_encoder MaxPool2d(2,2) + ReLU (7x7 -> 3x3 at 224 px, 14x14 -> 7x7 at 448)
 *   cb_maxpool2x2_relu_bwd  its backward, written into the zero-bordered layout read by the 3x3 dgrad/wgrad
 *   cb_relu_mask            dx = dy * (act > 0)
 * ------------------------------------------------------------------------------------------ */
/* Frame resize + pad of the data pipeline (src/datasets/data_utils.py:202-234 ImageResize: F.interpolate(bilinear,
 * align_corners=False) to new_h x new_w, longer side = max_size; :136-160 ImagePad: zeros at the bottom / right up to
 * max_size x max_size; src/datasets/dataset_base.py:191-195). x: `planes` = n * c planes of h x w (in_dtype 0 fp32, 1 uint8),
 * y: fp32 [planes, max_size, max_size]. */
int cb_resize_pad(const void* x, int in_dtype, float* y, int planes, int h, int w, int new_h, int new_w, int max_size, void* stream);
int cb_stem_im2col(const void* x, int in_dtype, void* out, int n, int h, int w, int kp, float mean_r, float mean_g,
                   float mean_b, void* stream);
int cb_stem_s2d(const void* x, int in_dtype, void* out, int n, int h, int w, int ld, float mean_r, float mean_g, float mean_b,
                void* stream);
int cb_maxpool3x3s2(const void* x, void* y, int n, int h, int w, int c, void* stream);
int cb_maxpool3x3s2_strided(const void* x, void* y, int n, int h, int w, int c, int64_t row_pitch, int64_t img_pitch, void* stream);
int cb_subsample2(const void* x, void* y, int n, int h, int w, int c, void* stream);
int cb_unsubsample2_mask(const void* dsub, const void* act, void* dx, int n, int h, int w, int c, void* stream);
int cb_maxpool2x2_relu_fwd(const void* x, void* y, int n, int h, int w, int c, void* stream);
int cb_maxpool2x2_relu_bwd(const void* dy, const void* x, void* dx_pad, int n, int h, int w, int c, void* stream);
int cb_relu_mask(const void* dy, const void* act, void* dx, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Clip-level score aggregation + loss of the training loops, pool_method "lse"
 * (src/tasks/run_video_retrieval.py:404-422, src/tasks/run_video_qa.py:484-501), forward AND backward in one launch:
 *   loss[0]  = mean_b ( logsumexp_{k,c} z[k,b,c] - logsumexp_k z[k,b,y_b] )          z = logits, fp32 [n_clips, nseq, ncls]
 *   dlogits  = grad_scale * d loss / d z   (fp32, same shape; NULL = forward only)   y = labels, int64 [nseq]
 * Replaces torch.stack + permute + two torch.logsumexp + torch.gather + mean and their autograd nodes (~45 ATen launches).
 * ------------------------------------------------------------------------------------------ */
int cb_clip_lse_loss(const float* logits, const int64_t* labels, float* loss, float* dlogits, int n_clips, int nseq, int ncls,
                     float grad_scale, void* stream);
/* pool_method "mean" (pool = 1) / "max" (pool = 2) of the same loops (run_video_retrieval.py:405-408, run_video_qa.py:485-488):
 * logits.mean(0) / logits.max(0)[0] followed by F.cross_entropy(reduction="none").mean(); same tensors as cb_clip_lse_loss */
int cb_clip_pool_ce_loss(const float* logits, const int64_t* labels, float* loss, float* dlogits, int n_clips, int nseq, int ncls, int pool,
                         float grad_scale, void* stream);
/* F.cross_entropy(logits, labels, reduction="none") over `rows` rows of `ncls` fp32 logits (row pitch ld), labels int64 with
 * ignore_index (-100: loss 0, no gradient): the masked-LM loss over the vocabulary (src/modeling/modeling.py:286-299), the
 * ITM / multiple-choice / retrieval CE (:560-580, :430-436). fwd: loss[rows], lse[rows] (stash). bwd: dlogits[r, c] =
 * grad_loss[r] * (softmax(z)[c] - [c == y]) with row pitch dld. One block per row, one pass over the row each. */
int cb_cross_entropy_fwd(const float* logits, int64_t ld, const int64_t* labels, float* loss, float* lse, int64_t rows, int ncls,
                         int64_t ignore_index, void* stream);
int cb_cross_entropy_bwd(const float* logits, int64_t ld, const int64_t* labels, const float* lse, const float* grad_loss, float* dlogits,
                         int64_t dld, int64_t rows, int ncls, int64_t ignore_index, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange through the NVSwitch, replacing hvd.DistributedOptimizer.synchronize()
 * (src/tasks/run_video_retrieval.py:299-301,432) for a flat fp32 buffer that every rank has mapped at the same MULTICAST
 * address (symmetric memory): rank r reduces the r-th 1/world slice in the switch (multimem.ld_reduce.add), scales it
 * (1/world = average) and stores it into every rank's copy (multimem.st). n: elements (multiple of 4); max_ctas: CTAs this
 * launch may use (0 = 64; 128 threads each, small enough to share an SM with a GEMM CTA). The caller places a cross-rank barrier before (all gradients written) and after (all slices stored).
 * ------------------------------------------------------------------------------------------ */
int cb_nvls_allreduce_f32(void* multicast_ptr, int64_t n, int rank, int world, float scale, int max_ctas, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused optimizer step over the flat fp32 parameter buffers.
 *   cb_sumsq       out[0] += sum(x^2)  - the norm half of torch.nn.utils.clip_grad_norm_ as called at
 *                  src/tasks/run_video_retrieval.py:477-480 (one call per flat gradient buffer, caller zeroes out); with a
 *                  chunk table only the table's elements are summed (alignment padding between parameters and the
 *                  zero-padded classifier rows belong to no parameter), else x[0, n)
 *   cb_adamw_step  AdamW of src/optimization/adamw.py:40-103 (eps added to sqrt(v), bias-corrected step size,
 *                  decoupled decay p -= lr*wd*p AFTER the Adam update) on every element named by the chunk table,
 *                  with the clip coefficient min(1, max_norm / (sqrt(*grad_sumsq) + 1e-6)) applied to the gradient
 *                  on the fly (grad_sumsq NULL or max_norm <= 0: no clipping), optional zeroing of the gradient
 *                  (optimizer.zero_grad(), :486) and emission of the bf16 tensor-core operand copy (the apex amp O2
 *                  master->model copy, :307-309) with the FrozenBN scale folded in for conv weights.
 * chunks: int64 [nchunks][8] device = offset, numel (<= 65536), group, row_len, scale_off (-1 none), flags (bit 0: emit
 *         packed), elem0 (index of the chunk's first element inside its parameter), 0
 * hyper : fp32 [ngroups][8] device = lr, step_size (lr * sqrt(1-b2^t) / (1-b1^t) when correct_bias), weight_decay,
 *         beta1, beta2, eps, 0, 0  - refreshed by the caller each step, so the launch itself is graph-capturable.
 * ------------------------------------------------------------------------------------------ */
int cb_sumsq(const float* x, int64_t n, const int64_t* chunks, int nchunks, float* out, void* stream);
int cb_adamw_step(float* master, float* grad, float* exp_avg, float* exp_avg_sq, void* packed_bf16, const int64_t* chunks,
                  int nchunks, const float* hyper, const float* scales, const float* grad_sumsq, float max_norm, int zero_grad,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPBERT_B200_H_ */
