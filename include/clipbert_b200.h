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
 * mode CB_GEMM_WGRAD : out[m, t*N + n] += rowscale[m] * sum_p A[p, m] * B[p + shift_t, n]
 *     A: bf16 [P, M] (dY), B: bf16 [P, N] (X); both operands are read "MN-major" straight from
 *     the activation layout; fp32 red.global.add accumulation (split over P across grid.z).
 *
 * Epilogue (TN), applied in this order on the fp32 accumulator v of element (m, n):
 *     v = v * scale[n] + shift[n]        (FrozenBN affine / bias; either may be NULL)
 *     v = dropout(v)                     (if dropout_p > 0; counter RNG keyed by seed + out index)
 *     v += residual[m, n]                (bf16, optional)
 *     if out2: out2[row, n] = v          (bf16 pre-activation stash, optional)
 *     v = act(v)                         (none / relu / gelu(erf) / tanh)
 *     v *= auxfn(aux[m, n])              (backward masks: relu' , gelu', tanh' ; optional)
 *     out[row(m), n] = v                 (bf16 or fp32)
 * row(m) re-maps between compact NHWC pixel rows and zero-bordered rows (cb_rowmap).
 * ------------------------------------------------------------------------------------------ */
enum { CB_GEMM_TN = 0, CB_GEMM_WGRAD = 1 };
enum { CB_ACT_NONE = 0, CB_ACT_RELU = 1, CB_ACT_GELU = 2, CB_ACT_TANH = 3 };
enum {
  CB_AUX_NONE = 0,
  CB_AUX_RELU_MASK = 1, /* v *= (aux > 0)              aux = forward output of the ReLU       */
  CB_AUX_GELU_GRAD = 2, /* v *= gelu'(aux)             aux = forward pre-activation            */
  CB_AUX_TANH_GRAD = 3  /* v *= 1 - aux^2              aux = forward tanh output               */
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
  int32_t split_k;                /* WGRAD only; >= 1 */
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
  int32_t reserved;
} cb_gemm_desc;

int cb_gemm(const cb_gemm_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPBERT_B200_H_ */
