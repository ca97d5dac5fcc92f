/* eva_sm100.h -- C ABI of libeva_sm100.so: the B200 (sm_100a) EVA / LARA / causal-EVA attention
 * core (forward; backward for the EVA family).
 *
 * The reference (HKUNLP/efficient-attention) is pure Python/PyTorch and has no native boundary of
 * its own; the seam is cut just below `module.forward`: everything between `proj_and_split_heads`
 * and the output projection.  Each entry point names the reference code it replaces:
 *
 *   eva_chunk_stats        eva.py:155-196          causal_eva.py:676-719   (chunk pooling, adaptive
 *                                                  Linear+LayerNorm, phi-projection, beta)
 *   eva_window_attention   eva.py:200-227          causal_eva.py:722-783   (local + chunk logits, one
 *                          local_attention.py:134-182   abstract_attention.py:115-133   joint softmax, PV)
 *   eva_forward            eva.py:151-227 as one call (selects the fused sm_100a tcgen05/TMA kernel
 *                          when the geometry allows, else the two generic stages)
 *   eva_backward           what autograd derives from eva.py:151-227 / causal_eva.py:676-783 / local_attention.py:134-182
 *                          (vit/engine.py:47-62 trains through it): gradients of eva_forward / eva_window_attention
 *   rfa_forward / scatterbrain_forward / ra_forward   kernelized_attention.py, scatterbrain_attention.py, randomized_attention.py
 *   lara_forward           lara.py:84-175          (landmark pooling, Linear+LN, mixing, proposal stats) and
 *                          lara.py:201-246         (phi-projections, kv statistics, MIS weights, SNIS) in one call
 *
 * Conventions
 *   - every function returns 0 on success or a negative errno-style code; it never throws, never
 *     calls exit(), never allocates or frees device memory and never synchronises the device;
 *   - all buffers are caller-owned device memory (the Python host side hands out PyTorch tensors);
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and the call returns;
 *   - the library is re-entrant across streams and devices; the only state it keeps is per-device
 *     "attribute already set" flags (idempotent), a thread-local cache of encoded tensor maps keyed by
 *     the pointers / strides / sizes of a call, and a thread-local last-error string;
 *   - q, k, v are described as strided views [batch, tokens, heads, head_dim] with head_dim
 *     contiguous, so both the packed `qkv` Linear output ([B,N,3,h,d], abstract_attention.py:72-78)
 *     and separate q/k/v projections (causal_eva.py:511-536) are consumed without a permute copy;
 *   - the output is written as [batch, tokens, heads*head_dim] contiguous, which is what the
 *     output projection consumes (eva.py:228, causal_eva.py:784).
 */
#ifndef EVA_SM100_H_
#define EVA_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVA_SM100_ABI_VERSION 4

enum EvaDtype { EVA_F32 = 0, EVA_F16 = 1, EVA_BF16 = 2 };

enum EvaStatus {
  EVA_OK = 0,
  EVA_ERR_INVALID = -22,      /* EINVAL: malformed geometry / null pointer / misaligned view   */
  EVA_ERR_UNSUPPORTED = -95,  /* EOPNOTSUPP: legal in the reference but not built here         */
  EVA_ERR_CUDA = -5           /* EIO: a CUDA runtime/driver call failed (see eva_last_error)   */
};

/* One of q / k / v: element (b, n, h, e) lives at ptr[b*stride_b + n*stride_n + h*stride_h + e].
 * Strides are in elements. ptr must be 16-byte aligned and strides multiples of 8 elements. */
typedef struct EvaHeadsView {
  const void* ptr;
  int64_t stride_b, stride_n, stride_h;
} EvaHeadsView;

/* Geometry of one EVA-style forward (eva.py:119-165, causal_eva.py:353-365,676-690). */
typedef struct EvaGeometry {
  int32_t batch, heads, tokens, head_dim; /* tokens = padded sequence length N               */
  int32_t dims;                           /* 1 (sequence) or 2 (token grid)                   */
  int32_t grid_h, grid_w;                 /* dims==2: grid_h*grid_w == tokens                 */
  int32_t window;                         /* local window edge w (tokens % w == 0 / grid % w) */
  int32_t ext;                            /* halo ("ext_size") of the local windows           */
  int32_t halo_left_only;                 /* 1: causal_window_1d_partition (causal_eva.py:102)*/
  int32_t chunk;                          /* chunk edge (rf_win_size); 0 = no chunk keys      */
  int32_t chunk_ext;                      /* halo of the chunks (== ext in EVA, 0 in causal)  */
  int32_t causal;                         /* triu masks of causal_eva.py:725-739,765-771      */
  int32_t mask_queries;                   /* local logits of padded queries masked too        */
  int32_t mask_is_neg_inf;                /* 1: padded keys get -inf (softmax baseline),
                                             0: -5e4 (eva.py:139)                             */
  int32_t io_dtype;                       /* EvaDtype of q, k, v and out                      */
  int32_t bias_toeplitz;                  /* (ABI v2) 1: the caller guarantees bias[i][j] depends on i - j only
                                             (T5 bucketed bias, eva.py:31-65, causal_eva.py:62-97); a hint that lets a
                                             kernel read one column of the table instead of all of it; 0 is always valid */
  int32_t keep_stats;                     /* (ABI v3) 1: eva_forward leaves k_bar | beta (float32 [batch, heads, C_n, head_dim] each, the
                                             second one 256-byte aligned after the first) at the start of its workspace, whatever
                                             kernel runs -- eva_backward takes them back instead of recomputing them; 0: the
                                             workspace contents are unspecified on return */
} EvaGeometry;

/* adaptive_mu_q / adaptive_mu_k = Linear(d,d) [+ LayerNorm(d)] shared by all heads
 * (eva.py:78-98, causal_eva.py:381-396). float32, row-major [out,in].
 * w_q == NULL: adaptive_proj == 'none' (mu = 0).  ln_gain_* == NULL: no LayerNorm ('no-ln'). */
typedef struct EvaAdaptive {
  const float *w_q, *b_q, *ln_gain_q, *ln_bias_q;
  const float *w_k, *b_k, *ln_gain_k, *ln_bias_k;
  float mu_coeff; /* 0.5 in EVA (eva.py:182), 1.0 in causal EVA (causal_eva.py:708) */
  float ln_eps;   /* 1e-5 */
} EvaAdaptive;

int eva_sm100_abi_version(void);
/* Thread-local description of the last failure on this thread ("" if none). */
const char* eva_last_error(void);
/* Number of chunks the geometry produces (C_n), or a negative status. */
int eva_num_chunks(const EvaGeometry* g);

/* k_bar, beta: float32 [batch, heads, C_n, head_dim].  pad_mask: [batch, tokens] bytes, nonzero =
 * padding, may be NULL.  noise: float32 [batch, heads, C_n, head_dim] or NULL (eval mode). */
int eva_chunk_stats(const EvaGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                    const uint8_t* pad_mask, const EvaAdaptive* ada, const float* noise,
                    float* k_bar, float* beta, void* stream);

/* bias: float32 [bias_heads, L, J] added to the local logits (already scaled), NULL for none;
 * bias_stride_h = L*J, or 0 when one table is shared by all heads (causal T5 bias).
 * k_bar / beta may be NULL iff g->chunk == 0 (pure local / dense softmax attention).
 * out: io_dtype [batch, tokens, heads*head_dim]. */
int eva_window_attention(const EvaGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                         const uint8_t* pad_mask, const float* k_bar, const float* beta,
                         const float* bias, int64_t bias_stride_h, void* out, void* stream);

/* eva_window_attention that also keeps, for the backward, the log-sum-exp of every query row (base-2 logarithm, logits multiplied by
 * log2 e): lse float32 [batch, heads, tokens].  *lse_written = 1 when the kernel that ran wrote it (the tcgen05 kernels do; the
 * CUDA-core kernel does not: eva_backward then recomputes it), else 0.  lse / lse_written may be NULL. */
int eva_window_attention_lse(const EvaGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                             const uint8_t* pad_mask, const float* k_bar, const float* beta, const float* bias, int64_t bias_stride_h,
                             void* out, float* lse, int32_t* lse_written, void* stream);

/* Bytes of scratch eva_forward needs (k_bar + beta + fast-path staging); 256-byte aligned. */
int eva_forward_workspace_bytes(const EvaGeometry* g, size_t* bytes);

/* Both stages in one call.  *path_taken (optional) receives 1 when the fused sm_100a kernel ran,
 * 2 when the generic statistics kernel + the tcgen05 causal window kernel ran, 0 when the generic two-stage path ran; with
 * g->keep_stats, bit 0x100 is set in addition when the log-sum-exp of every query row (float32 [batch, heads, tokens], base 2) was
 * left in the LAST align256(batch * heads * tokens * 4) bytes of the workspace. */
int eva_forward(const EvaGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                const uint8_t* pad_mask, const EvaAdaptive* ada, const float* noise,
                const float* bias, int64_t bias_stride_h, void* out, void* workspace, size_t workspace_bytes,
                int32_t* path_taken, void* stream);

/* Backward of eva_forward / eva_window_attention (the reference differentiates eva.py:151-227 / causal_eva.py:676-783 /
 * local_attention.py:134-182 with autograd).  float32 CUDA-core kernels for every geometry the forward accepts; the
 * probabilities are recomputed from q, k, v (nothing but the forward OUTPUT is kept between the two calls).
 *   out, grad_out  io_dtype [batch, tokens, heads*head_dim] contiguous: the forward result and the gradient arriving at it
 *   grad_qkv       float32 [3, batch, tokens, heads, head_dim] = dq | dk | dv (need not be initialised: the call zeroes what it
 *                  accumulates into)
 *   grad_qkv_io    optional: io_dtype [batch, tokens, 3, heads, head_dim] -- the layout of the packed qkv projection
 *                  (abstract_attention.py:72-78).  When given, the final dq | dk | dv are written there rounded to io_dtype and the
 *                  contents of grad_qkv are unspecified on return (scratch)
 *   grad_bias      float32, the shape of `bias`; NULL: not wanted
 *   chunk_rows     float32 [12, batch, heads, C_n, head_dim]; NULL iff g->chunk == 0.  Slots on return:
 *                  0 k_bar | 1 beta (recomputed) | 2 d k_bar | 3 d beta | 4 dy_k | 5 dy_q (gradients at the adaptive Linear outputs) |
 *                  6 mean_k | 7 mean_q (the Linear inputs) | 8 n_k | 9 n_q (LayerNorm-normalised rows) | 10 dout_k | 11 dout_q
 *                  (gradients at the LayerNorm outputs).  The PARAMETER gradients are plain reductions over the chunk rows, left to
 *                  the caller's library: dW = dy^T mean, db = sum dy, d gain = sum dout * n, d ln_bias = sum dout.
 *   ada            may be NULL iff g->chunk == 0.
 *   k_bar, beta    the forward's chunk statistics (eva_forward with g->keep_stats, or eva_chunk_stats), or both NULL: recomputed
 *                  into slots 0 / 1.
 *   lse            the forward's row log-sum-exp (eva_forward: path_taken & 0x100; eva_window_attention_lse), or NULL: recomputed. */
int eva_backward(const EvaGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                 const uint8_t* pad_mask, const EvaAdaptive* ada, const float* noise, const float* bias, int64_t bias_stride_h,
                 const void* out, const void* grad_out, const float* k_bar, const float* beta, const float* lse, float* grad_qkv,
                 void* grad_qkv_io, float* grad_bias, float* chunk_rows, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LARA (lara.py).  Landmarks are the pooled q/k summaries; samples S = landmarks C, or 2C with
 * antithetic / multi-sample proposals at training time (lara.py:188-198). */
enum LaraMis { LARA_MIS_OPT = 0, LARA_MIS_BH = 1, LARA_MIS_BIASED = 2 };
enum LaraSample { LARA_SAMPLE_SINGLE = 0, LARA_SAMPLE_ANTITHETIC = 1, LARA_SAMPLE_MULTI = 2 };

typedef struct LaraGeometry {
  int32_t batch, heads, tokens, head_dim;
  int32_t dims;             /* 2: AdaptiveAvgPool2d landmarks (lara.py:129-175); 1: segment means (lara.py:84-127) */
  int32_t grid_h, grid_w;
  int32_t landmarks;        /* C: dims==2 -> int(sqrt(num_landmarks))^2 ; dims==1 -> min(num_landmarks, tokens) */
  int32_t per_token_proj;   /* 1: 'adaptive-1d' (Linear+LN on every token before pooling)   */
  int32_t mixed;            /* 0 none, 1 '-mixed', 2 '-vmixed' (lara.py:157-174)            */
  int32_t mis_type;         /* LaraMis                                                      */
  int32_t sample_mode;      /* LaraSample (only meaningful when noise != NULL)              */
  int32_t zero_padded;      /* 1: q,k,v of padded tokens are treated as 0 (1-D path, lara.py:87-91) */
  int32_t io_dtype;
  float alpha_coeff;
} LaraGeometry;

/* Scratch for lara_forward. */
int lara_forward_workspace_bytes(const LaraGeometry* g, size_t* bytes);

/* proj: the q_bar_gen / k_bar_gen Linear(+LayerNorm) parameters in an EvaAdaptive (w_q == NULL for
 * 'no-param-pool'; mu_coeff ignored).  noise: float32 [batch, heads, S or C (antithetic), head_dim]
 * or NULL.  out: io_dtype [batch, tokens, heads*head_dim]. */
int lara_forward(const LaraGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                 const uint8_t* pad_mask, const EvaAdaptive* proj, const float* noise,
                 void* out, void* workspace, size_t workspace_bytes, void* stream);

/* Same as lara_forward with the landmarks computed by the caller: pool_module_type == 'dense' (lara.py:36-39, 131-139) runs its
 * Linear / LayerNorm over ALL channels, i.e. across heads, which is a plain library GEMM on a [batch, C, dim] tensor.
 * landmarks: float32 [batch, heads, 3, C, head_dim] = q_bar | k_bar (before the '-mixed' step) | v_bar (read for '-vmixed' only). */
int lara_forward_given_landmarks(const LaraGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                                 const uint8_t* pad_mask, const float* landmarks, const float* noise,
                                 void* out, void* workspace, size_t workspace_bytes, void* stream);

/* LARA backward (what autograd derives from lara.py:201-246, mis-opt estimator, one sample per landmark, no padding mask): the three
 * fused steps BETWEEN the batched GEMMs of the explicit gradient.  Matrices are row-major [items][rows][tokens] in io_dtype
 * (items = batch x heads); vectors float32.
 *   which = 0  rows:  X [items, 2C, N] rows [C, 2C) = q_bar q^T in -> t = softmax_n(scale .) out;  Y [items, C, N] = omega k^T in ->
 *              Pk = softmax_m(scale . - v0) out, v0 [items, N] = scale |k|^2 / 2;  o0 = lse_B, o1 = lse_T [items, C]
 *   which = 1  columns:  X rows [0, C) = omega q^T in -> W = softmax_c(log w) out, rows [C, 2C) = t;  dW [items, C, N] = kv dO^T;
 *              M2 [items, 2C, N] out: rows [0, C) = d log w, rows [C, 2C) = dt;  v0 [items, N] = scale |q|^2 / 2, v1 = bh, v2 = lp,
 *              v3 = lse_B [items, C];  o0 += sum_n d log w, o1 += sum_n d alpha, o2 += sum_n t dt  [items, C], zeroed by the caller
 *   which = 2  rows:  X <- Y o (X - v0 + v1), rows of X / Y at x_item_stride / y_item_stride elements per item; v1 may be NULL */
int lara_backward_step(int32_t which, int32_t io_dtype, void* X, void* Y, const void* dW, void* M2, const float* v0, const float* v1,
                       const float* v2, const float* v3, float* o0, float* o1, float* o2, int64_t x_item_stride, int64_t y_item_stride,
                       int32_t items, int32_t landmarks, int32_t tokens, float scale, float alpha_coeff, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (ABI v4) Random-feature modules of the registry (__init__.py:53-62): 'performer' (kernelized_attention.py), 'ra'
 * (randomized_attention.py), 'scatterbrain' (scatterbrain_attention.py).  float32 math, q / k / v / out in io_dtype. */
enum RfaMethod {
  RFA_FAVORP = 0,        /* favorp_projection,       kernelized_attention.py:21-55  (positive random features, eps 1e-4)   */
  RFA_RELU = 1,          /* generalized_projection + relu, :92-113 (eps 1e-3)                                              */
  RFA_FOURIER = 2,       /* fourier_projection,      :57-87  (2 * proj_dim features)                                        */
  RFA_DPFP = 3,          /* dpfp_projection,         :12-19  (2 * head_dim * nu features, no projection matrix)            */
  RFA_RELU_ONLY = 4,     /* nonlinear_map(relu),     :89-90  (head_dim features, eps 0.1)                                   */
  RFA_SIGMOID_ONLY = 5,  /* nonlinear_map(sigmoid)                                                                          */
  RFA_GIVEN = 6          /* the caller computed phi(q), phi(k) itself ('mlp-fourier': a learnable Linear over all features,
                            :155-178, is a library GEMM on the host side)                                                   */
};

typedef struct RfaGeometry {
  int32_t batch, heads, tokens, head_dim;
  int32_t method;        /* RfaMethod */
  int32_t proj_dim;      /* rows of the projection matrix (approx_attn_dim) for FAVORP / RELU / FOURIER */
  int32_t nu;            /* DPFP */
  int32_t feat_dim;      /* GIVEN: width of q_feat / k_feat */
  int32_t cos_weighting; /* cosFormer re-weighting, kernelized_attention.py:122-153 */
  int32_t io_dtype;
} RfaGeometry;

/* Width of phi(.) after the optional cosFormer doubling, or a negative status. */
int rfa_feature_dim(const RfaGeometry* g);
int rfa_forward_workspace_bytes(const RfaGeometry* g, size_t* bytes);
/* KernelizedAttention._apply_attention (kernelized_attention.py:301-320): out = phi(q) (phi(k)^T v) / max(phi(q) . sum phi(k), 1e-2).
 * proj: float32 [heads, proj_dim, head_dim] (NULL for the methods without one); q_feat / k_feat: float32 [batch, heads, tokens,
 * feat_dim], RFA_GIVEN only; pad_mask: [batch, tokens] bytes or NULL (padded keys' features are zeroed after the stabilisers, as
 * the reference does).  out: io_dtype [batch, tokens, heads*head_dim]. */
int rfa_forward(const RfaGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v, const uint8_t* pad_mask,
                const float* proj, const float* q_feat, const float* k_feat, void* out, void* workspace, size_t workspace_bytes,
                void* stream);

/* ScatterBrain.forward between the qkv and the output projections (scatterbrain_attention.py:95-160), proj_method 'favorp':
 * halo-free local windows (<= 64 tokens) + proj_dim random-feature keys per window that stand for everything outside it. */
typedef struct SbGeometry {
  int32_t batch, heads, tokens, head_dim;
  int32_t dims, grid_h, grid_w, window;
  int32_t proj_dim;
  int32_t io_dtype;
} SbGeometry;
int scatterbrain_forward_workspace_bytes(const SbGeometry* g, size_t* bytes);
/* proj: float32 [heads, proj_dim, head_dim]; bias: float32 [heads, L, L] added to the local logits, or NULL. */
int scatterbrain_forward(const SbGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v, const uint8_t* pad_mask,
                         const float* proj, const float* bias, void* out, void* workspace, size_t workspace_bytes, void* stream);

/* RandomizedAttention._apply_attention (randomized_attention.py:24-55): out_n = softmax_m(scale w_n . k_m - scale |k_m|^2 / 2) v_m,
 * w_n = q_n + extra_n (+ noise_n).  mode 0: extra = mean of k (num_samples == 0; workspace: batch*heads*head_dim floats);
 * mode 1: extra = caller-supplied rows, io_dtype [batch, tokens, heads*head_dim] (num_samples == -1: E_pi[k], i.e.
 * eva_window_attention with v := k); mode 2: extra = k[k_ind[n]], k_ind int64 [batch, heads, tokens] (the multinomial draw).
 * noise: float32 [batch, heads, tokens, head_dim] or NULL.  The reference ignores the padding mask here; so does this. */
typedef struct RaGeometry {
  int32_t batch, heads, tokens, head_dim;
  int32_t mode;
  int32_t io_dtype;
} RaGeometry;
int ra_forward_workspace_bytes(const RaGeometry* g, size_t* bytes);
/* The draw of mode 2 (randomized_attention.py:36-40: one key index per query from pi_n = softmax_m(scale q_n . k_m), torch.multinomial
 * in the reference) without the [tokens, tokens] probabilities: k_ind[b, h, n] = argmax_m (scale q_n . k_m + G_nm), G i.i.d. standard
 * Gumbel (the Gumbel-max trick draws exactly from pi_n).  G comes from a counter-based hash of (seed, b, h, n, m), or from `gumbel`,
 * float32 [batch, heads, tokens, tokens], when that is not NULL (tests).  head_dim 64 and 16-bit q / k only: EVA_ERR_UNSUPPORTED
 * otherwise (the caller then draws with library ops, as the reference does).  g->mode is ignored. */
int ra_sample(const RaGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, uint64_t seed, const float* gumbel, int64_t* k_ind,
              void* stream);
int ra_forward(const RaGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v, const void* extra,
               const int64_t* k_ind, const float* noise, void* out, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EVA_SM100_H_ */
