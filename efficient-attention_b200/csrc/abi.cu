// C ABI of libeva_sm100.so (see include/eva_sm100.h).  Validates arguments, builds the internal
// geometry and enqueues kernels on the caller's stream.  Never throws, never allocates device memory.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "launch.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(EVA_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

int ipow(int b, int e) { return e == 2 ? b * b : b; }

int make_geo(const EvaGeometry* in, eva::Geo* g) {
  if (!in) return fail(EVA_ERR_INVALID, "geometry is NULL");
  if (in->batch <= 0 || in->heads <= 0 || in->tokens <= 0) return fail(EVA_ERR_INVALID, "batch/heads/tokens must be positive");
  if (in->head_dim != 16 && in->head_dim != 32 && in->head_dim != 64 && in->head_dim != 128)
    return fail(EVA_ERR_UNSUPPORTED, "head_dim %d not built (16, 32, 64, 128)", in->head_dim);
  if (in->dims != 1 && in->dims != 2) return fail(EVA_ERR_INVALID, "dims must be 1 or 2");
  if (in->io_dtype < EVA_F32 || in->io_dtype > EVA_BF16) return fail(EVA_ERR_INVALID, "unknown io_dtype %d", in->io_dtype);
  if (in->window <= 0 || in->ext < 0 || in->chunk < 0 || in->chunk_ext < 0) return fail(EVA_ERR_INVALID, "window must be > 0; ext/chunk >= 0");
  g->B = in->batch; g->H = in->heads; g->N = in->tokens; g->D = in->head_dim;
  g->dims = in->dims;
  g->window = in->window; g->ext = in->ext; g->left_only = in->halo_left_only ? 1 : 0;
  g->chunk = in->chunk; g->chunk_ext = in->chunk_ext;
  g->causal = in->causal ? 1 : 0; g->mask_queries = in->mask_queries ? 1 : 0;
  g->mask_fill = in->mask_is_neg_inf ? -INFINITY : eva::kMaskVal;
  g->bias_toeplitz = in->bias_toeplitz ? 1 : 0;
  if (in->dims == 2) {
    if (in->grid_h <= 0 || in->grid_w <= 0 || in->grid_h * in->grid_w != in->tokens)
      return fail(EVA_ERR_INVALID, "grid %dx%d does not cover %d tokens", in->grid_h, in->grid_w, in->tokens);
    if (in->grid_h % in->window || in->grid_w % in->window)
      return fail(EVA_ERR_INVALID, "grid %dx%d not divisible by window %d (eva.py:124-126)", in->grid_h, in->grid_w, in->window);
    if (in->causal || in->halo_left_only) return fail(EVA_ERR_INVALID, "causal / left-only halos are 1-D only");
    g->gh = in->grid_h; g->gw = in->grid_w;
    g->n_windows = (g->gh / g->window) * (g->gw / g->window);
    g->L = g->window * g->window;
    g->J = ipow(g->window + 2 * g->ext, 2);
    if (g->chunk > 0) {
      if (g->chunk_ext == 0 && (g->gh % g->chunk || g->gw % g->chunk))
        return fail(EVA_ERR_INVALID, "grid not divisible by chunk %d", g->chunk);
      g->n_chunks = (g->gh / g->chunk) * (g->gw / g->chunk);
      g->Jc = ipow(g->chunk + 2 * g->chunk_ext, 2);
    }
  } else {
    if (in->tokens % in->window) return fail(EVA_ERR_INVALID, "tokens %d not a multiple of window %d (pad first)", in->tokens, in->window);
    g->gh = 1; g->gw = in->tokens;
    g->n_windows = g->N / g->window;
    g->L = g->window;
    g->J = g->window + g->ext + (g->left_only ? 0 : g->ext);
    if (g->chunk > 0) {
      if (g->chunk_ext == 0 && g->N % g->chunk) return fail(EVA_ERR_INVALID, "tokens %d not divisible by chunk %d", g->N, g->chunk);
      g->n_chunks = g->N / g->chunk;
      g->Jc = g->chunk + g->chunk_ext + (g->left_only ? 0 : g->chunk_ext);
    }
  }
  if (g->chunk == 0) { g->n_chunks = 0; g->Jc = 0; }
  if (g->chunk > 0 && g->n_chunks <= 0) return fail(EVA_ERR_INVALID, "geometry yields no chunks");
  return EVA_OK;
}

int make_view(const EvaHeadsView* in, const char* name, eva::View* v) {
  if (!in || !in->ptr) return fail(EVA_ERR_INVALID, "%s view is NULL", name);
  if ((reinterpret_cast<uintptr_t>(in->ptr) & 15u) != 0) return fail(EVA_ERR_INVALID, "%s pointer not 16-byte aligned", name);
  if (in->stride_b % 8 || in->stride_n % 8 || in->stride_h % 8)
    return fail(EVA_ERR_INVALID, "%s strides must be multiples of 8 elements", name);
  v->ptr = in->ptr; v->sb = in->stride_b; v->sn = in->stride_n; v->sh = in->stride_h;
  return EVA_OK;
}

int check_ada(const EvaAdaptive* a) {
  if (!a) return fail(EVA_ERR_INVALID, "adaptive parameters are NULL");
  if (!a->w_k || !a->b_k) return fail(EVA_ERR_INVALID, "adaptive_mu_k Linear is required (eva.py:79-98)");
  if (a->w_q && !a->b_q) return fail(EVA_ERR_INVALID, "adaptive_mu_q bias missing");
  if ((a->ln_gain_k && !a->ln_bias_k) || (a->ln_gain_q && !a->ln_bias_q)) return fail(EVA_ERR_INVALID, "LayerNorm gain without bias");
  return EVA_OK;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

namespace eva {
int abi_fail(int code, const char* msg) { return fail(code, "%s", msg); }
int abi_cuda_fail(cudaError_t e, const char* what) { return cuda_fail(e, what); }
int abi_view(const EvaHeadsView* in, const char* name, View* v) { return make_view(in, name, v); }
}  // namespace eva

extern "C" {

int eva_sm100_abi_version(void) { return EVA_SM100_ABI_VERSION; }

const char* eva_last_error(void) { return g_err; }

int eva_num_chunks(const EvaGeometry* gin) {
  eva::Geo g{};
  const int rc = make_geo(gin, &g);
  return rc != EVA_OK ? rc : g.n_chunks;
}

int eva_chunk_stats(const EvaGeometry* gin, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                    const uint8_t* pad_mask, const EvaAdaptive* ada, const float* noise,
                    float* k_bar, float* beta, void* stream) {
  eva::Geo g{};
  eva::View vq, vk, vv;
  int rc;
  if ((rc = make_geo(gin, &g)) || (rc = make_view(q, "q", &vq)) || (rc = make_view(k, "k", &vk)) ||
      (rc = make_view(v, "v", &vv)) || (rc = check_ada(ada)))
    return rc;
  if (g.n_chunks == 0) return fail(EVA_ERR_INVALID, "chunk == 0: nothing to compute");
  if (!k_bar || !beta) return fail(EVA_ERR_INVALID, "k_bar / beta output is NULL");
  const cudaError_t e = eva::launch_chunk_stats(g, gin->io_dtype, vq, vk, vv, pad_mask, *ada, noise, k_bar, beta,
                                                reinterpret_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? EVA_OK : cuda_fail(e, "eva_chunk_stats");
}

int eva_window_attention(const EvaGeometry* gin, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                         const uint8_t* pad_mask, const float* k_bar, const float* beta,
                         const float* bias, int64_t bias_stride_h, void* out, void* stream) {
  return eva_window_attention_lse(gin, q, k, v, pad_mask, k_bar, beta, bias, bias_stride_h, out, nullptr, nullptr, stream);
}

int eva_window_attention_lse(const EvaGeometry* gin, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                             const uint8_t* pad_mask, const float* k_bar, const float* beta, const float* bias, int64_t bias_stride_h,
                             void* out, float* lse, int32_t* lse_written, void* stream) {
  if (lse_written) *lse_written = 0;
  eva::Geo g{};
  eva::View vq, vk, vv;
  int rc;
  if ((rc = make_geo(gin, &g)) || (rc = make_view(q, "q", &vq)) || (rc = make_view(k, "k", &vk)) ||
      (rc = make_view(v, "v", &vv)))
    return rc;
  if (g.n_chunks > 0 && (!k_bar || !beta)) return fail(EVA_ERR_INVALID, "k_bar / beta required when chunk > 0");
  if (!out) return fail(EVA_ERR_INVALID, "out is NULL");
  if (bias && bias_stride_h != 0 && bias_stride_h != (int64_t)g.L * g.J)
    return fail(EVA_ERR_INVALID, "bias_stride_h must be 0 or L*J = %d", g.L * g.J);
  if (g.n_chunks > 0 && eva::causal_window_supported(g, gin->io_dtype, vq, vk, vv, pad_mask, bias, bias_stride_h)) {
    const char* msg = "";
    const cudaError_t ec = eva::launch_causal_window(g, gin->io_dtype, vq, vk, vv, k_bar, beta, bias, out,
                                                     reinterpret_cast<cudaStream_t>(stream), &msg, nullptr, nullptr, nullptr, lse);
    if (lse && lse_written) *lse_written = 1;
    return ec == cudaSuccess ? EVA_OK : fail(EVA_ERR_CUDA, "eva_window_attention(causal window): %s: %s", msg, cudaGetErrorString(ec));
  }
  if (lse && lse_written && eva::window_tc_supported(g, gin->io_dtype)) *lse_written = 1;
  const cudaError_t e = eva::window_tc_supported(g, gin->io_dtype)
                            ? eva::launch_window_tc(g, gin->io_dtype, vq, vk, vv, pad_mask, k_bar, beta, bias, bias_stride_h, out,
                                                    reinterpret_cast<cudaStream_t>(stream), lse)
                            : eva::launch_window_attn(g, gin->io_dtype, vq, vk, vv, pad_mask, k_bar, beta, bias, bias_stride_h, out,
                                                      reinterpret_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? EVA_OK : cuda_fail(e, "eva_window_attention");
}

int eva_forward_workspace_bytes(const EvaGeometry* gin, size_t* bytes) {
  eva::Geo g{};
  const int rc = make_geo(gin, &g);
  if (rc) return rc;
  if (!bytes) return fail(EVA_ERR_INVALID, "bytes is NULL");
  const size_t stats = align256((size_t)g.B * g.H * g.n_chunks * g.D * sizeof(float));
  // k_bar | beta | fused-kernel scratch | per-window flags of the one-pass causal kernel
  // k_bar | beta | fused-kernel scratch | per-window flags | (keep_stats) log-sum-exp of every query row
  *bytes = 2 * stats + align256(eva::fused_workspace_bytes(g)) + align256((size_t)g.B * g.H * (g.n_windows > 0 ? g.n_windows : 1) * sizeof(unsigned int)) +
           (gin->keep_stats ? align256((size_t)g.B * g.H * g.N * sizeof(float)) : 0);
  return EVA_OK;
}

// diagnostic (not part of the public ABI): eva_forward calls by the path they took (0 generic, 1 fused, 2 causal tcgen05,
// 3 cluster-resident fused) -- lets module- and model-level tests prove which kernels ran
static int g_path_count[4] = {0, 0, 0, 0};
int eva_debug_path_count(int path) { return path >= 0 && path < 4 ? g_path_count[path] : -1; }

int eva_forward(const EvaGeometry* gin, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                const uint8_t* pad_mask, const EvaAdaptive* ada, const float* noise,
                const float* bias, int64_t bias_stride_h, void* out, void* workspace, size_t workspace_bytes,
                int32_t* path_taken, void* stream) {
  eva::Geo g{};
  eva::View vq, vk, vv;
  int rc;
  if ((rc = make_geo(gin, &g)) || (rc = make_view(q, "q", &vq)) || (rc = make_view(k, "k", &vk)) ||
      (rc = make_view(v, "v", &vv)) || (rc = check_ada(ada)))
    return rc;
  if (g.n_chunks == 0) return fail(EVA_ERR_INVALID, "eva_forward needs chunk > 0 (use eva_window_attention)");
  if (!out || !workspace) return fail(EVA_ERR_INVALID, "out / workspace is NULL");
  if (bias && bias_stride_h != 0 && bias_stride_h != (int64_t)g.L * g.J)
    return fail(EVA_ERR_INVALID, "bias_stride_h must be 0 or L*J = %d", g.L * g.J);
  size_t need = 0;
  eva_forward_workspace_bytes(gin, &need);
  if (workspace_bytes < need) return fail(EVA_ERR_INVALID, "workspace too small: %zu < %zu", workspace_bytes, need);
  if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return fail(EVA_ERR_INVALID, "workspace must be 256-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t stats = align256((size_t)g.B * g.H * g.n_chunks * g.D * sizeof(float));
  char* ws = reinterpret_cast<char*>(workspace);
  float* k_bar = reinterpret_cast<float*>(ws);
  float* beta = reinterpret_cast<float*>(ws + stats);
  if (!gin->keep_stats && eva::cluster_supported(g, gin->io_dtype, vq, vk, vv, pad_mask)) {   // the cluster kernel keeps its statistics on chip
    const char* msg = "";
    const cudaError_t e = eva::launch_cluster(g, gin->io_dtype, vq, vk, vv, *ada, noise, bias, bias_stride_h, out, ws + 2 * stats, st, &msg);
    if (path_taken) *path_taken = 3;
    ++g_path_count[3];
    if (e != cudaSuccess) return fail(EVA_ERR_CUDA, "eva_forward(cluster): %s: %s", msg, cudaGetErrorString(e));
    return EVA_OK;
  }
  if (eva::fused_supported(g, gin->io_dtype, vq, vk, vv, pad_mask, *ada, bias, bias_stride_h)) {
    const char* msg = "";
    const cudaError_t e = eva::launch_fused(g, gin->io_dtype, vq, vk, vv, *ada, noise, bias, bias_stride_h, out,
                                            ws + 2 * stats, st, &msg, gin->keep_stats ? k_bar : nullptr, gin->keep_stats ? beta : nullptr);
    if (path_taken) *path_taken = 1;
    ++g_path_count[1];
    if (e != cudaSuccess) return fail(EVA_ERR_CUDA, "eva_forward(fused): %s: %s", msg, cudaGetErrorString(e));
    return EVA_OK;
  }
  // Two kernels read q, k, v once each (c5 at batch 16: 1.67x algorithmic DRAM traffic).  Optionally they run back to back on
  // SLICES of the batch small enough for the statistics pass to leave its slice in L2 for the window pass
  // (EVA_SM100_L2_SLICE_MB=<MB>).  Measured on c5 (profiles/r02/README.md): 0.160 ms unsliced, 0.224 / 0.239 / 0.264 ms with
  // 64 / 40 / 24 MB slices -- four to eight small launches cost more than the HBM re-read saves -- so the default is OFF.
  static const long long slice_mb = [] { const char* e_ = getenv("EVA_SM100_L2_SLICE_MB"); return e_ ? atoll(e_) : 0LL; }();
  const long long elem = gin->io_dtype == EVA_F32 ? 4 : 2;
  const long long per_b = 3LL * g.N * g.H * g.D * elem;
  int nb = g.B;
  if (slice_mb > 0 && per_b > 0) {
    const long long fit = (slice_mb << 20) / per_b;
    nb = (int)(fit < 1 ? 1 : (fit > g.B ? g.B : fit));
  }
  const bool causal_fast = eva::causal_window_supported(g, gin->io_dtype, vq, vk, vv, pad_mask, bias, bias_stride_h);
  // training: the tcgen05 window kernels also leave the log-sum-exp of every query row (path_taken | 0x100 says so)
  float* lse = gin->keep_stats ? reinterpret_cast<float*>(ws + 2 * stats + align256(eva::fused_workspace_bytes(g)) +
                                                          align256((size_t)g.B * g.H * (g.n_windows > 0 ? g.n_windows : 1) * sizeof(unsigned int)))
                               : nullptr;
  const bool lse_kept = lse && nb == g.B && (causal_fast || eva::window_tc_supported(g, gin->io_dtype));
  if (!lse_kept) lse = nullptr;
  if (path_taken) *path_taken = (causal_fast ? 2 : 0) | (lse_kept ? 0x100 : 0);
  ++g_path_count[causal_fast ? 2 : 0];
  if (causal_fast && eva::causal_one_pass_supported(g, *ada)) {
    // one pass over q, k, v: the window kernel computes the chunk statistics itself
    unsigned int* flags = reinterpret_cast<unsigned int*>(ws + 2 * stats + align256(eva::fused_workspace_bytes(g)));
    const char* msg = "";
    const cudaError_t e1 = eva::launch_causal_window(g, gin->io_dtype, vq, vk, vv, k_bar, beta, bias, out, st, &msg, ada, noise, flags, lse);
    if (e1 != cudaSuccess) return fail(EVA_ERR_CUDA, "eva_forward(causal one-pass): %s: %s", msg, cudaGetErrorString(e1));
    return EVA_OK;
  }
  for (int b0 = 0; b0 < g.B; b0 += nb) {
    eva::Geo gs = g;
    gs.B = (g.B - b0) < nb ? (g.B - b0) : nb;
    auto shift = [&](const eva::View& v_) { eva::View r = v_; r.ptr = static_cast<const char*>(v_.ptr) + (long long)b0 * v_.sb * elem; return r; };
    const eva::View sq = shift(vq), sk = shift(vk), sv = shift(vv);
    const long long stat_off = (long long)b0 * g.H * g.n_chunks * g.D;
    const uint8_t* smask = pad_mask ? pad_mask + (long long)b0 * g.N : nullptr;
    const float* snoise = noise ? noise + stat_off : nullptr;
    void* sout = static_cast<char*>(out) + (long long)b0 * g.N * g.H * g.D * elem;
    cudaError_t e = eva::launch_chunk_stats(gs, gin->io_dtype, sq, sk, sv, smask, *ada, snoise, k_bar + stat_off, beta + stat_off, st);
    if (e != cudaSuccess) return cuda_fail(e, "eva_forward(chunk_stats)");
    if (causal_fast) {
      const char* msg = "";
      e = eva::launch_causal_window(gs, gin->io_dtype, sq, sk, sv, k_bar + stat_off, beta + stat_off, bias, sout, st, &msg, nullptr, nullptr, nullptr, lse);
      if (e != cudaSuccess) return fail(EVA_ERR_CUDA, "eva_forward(causal window): %s: %s", msg, cudaGetErrorString(e));
    } else {
      e = eva::window_tc_supported(gs, gin->io_dtype)
              ? eva::launch_window_tc(gs, gin->io_dtype, sq, sk, sv, smask, k_bar + stat_off, beta + stat_off, bias, bias_stride_h, sout, st, lse)
              : eva::launch_window_attn(gs, gin->io_dtype, sq, sk, sv, smask, k_bar + stat_off, beta + stat_off, bias, bias_stride_h, sout, st);
      if (e != cudaSuccess) return cuda_fail(e, "eva_forward(window_attention)");
    }
  }
  return EVA_OK;
}

int eva_backward(const EvaGeometry* gin, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                 const uint8_t* pad_mask, const EvaAdaptive* ada, const float* noise, const float* bias, int64_t bias_stride_h,
                 const void* out, const void* grad_out, const float* k_bar_in, const float* beta_in, const float* lse, float* grad_qkv,
                 void* grad_qkv_io, float* grad_bias, float* chunk_rows, void* stream) {
  eva::Geo g{};
  eva::View vq, vk, vv;
  int rc;
  if ((rc = make_geo(gin, &g)) || (rc = make_view(q, "q", &vq)) || (rc = make_view(k, "k", &vk)) || (rc = make_view(v, "v", &vv)))
    return rc;
  if (g.n_chunks > 0 && (rc = check_ada(ada))) return rc;
  if (g.n_chunks > 0 && !chunk_rows) return fail(EVA_ERR_INVALID, "chunk_rows is NULL");
  if (!out || !grad_out || !grad_qkv) return fail(EVA_ERR_INVALID, "out / grad_out / grad_qkv is NULL");
  if ((reinterpret_cast<uintptr_t>(grad_qkv) | reinterpret_cast<uintptr_t>(grad_qkv_io)) & 15u)
    return fail(EVA_ERR_INVALID, "grad_qkv / grad_qkv_io must be 16-byte aligned");
  if (((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(chunk_rows)) & 15u) != 0)
    return fail(EVA_ERR_INVALID, "out / grad_out / chunk_rows must be 16-byte aligned");
  if (bias && bias_stride_h != 0 && bias_stride_h != (int64_t)g.L * g.J)
    return fail(EVA_ERR_INVALID, "bias_stride_h must be 0 or L*J = %d", g.L * g.J);
  if (grad_bias && !bias) return fail(EVA_ERR_INVALID, "grad_bias without bias");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long slot = (long long)g.B * g.H * g.n_chunks * g.D;
  const long long tens = (long long)g.B * g.N * g.H * g.D;
  if ((k_bar_in == nullptr) != (beta_in == nullptr)) return fail(EVA_ERR_INVALID, "k_bar and beta come together");
  if (((reinterpret_cast<uintptr_t>(k_bar_in) | reinterpret_cast<uintptr_t>(beta_in)) & 15u) != 0)
    return fail(EVA_ERR_INVALID, "k_bar / beta must be 16-byte aligned");
  const float* kbar = k_bar_in ? k_bar_in : chunk_rows;
  const float* beta = beta_in ? beta_in : chunk_rows + slot;
  // accumulation targets start from zero: d k_bar | d beta, the bias gradient, and -- unless the tcgen05 window kernel runs, which
  // writes every dq / dk / dv row exactly once before the chunk-statistics kernel adds to it -- dq | dk | dv
  cudaError_t ez = cudaSuccess;
  if (g.n_chunks > 0) ez = cudaMemsetAsync(chunk_rows + 2 * slot, 0, 2 * (size_t)slot * sizeof(float), st);
  if (ez == cudaSuccess && grad_bias)
    ez = cudaMemsetAsync(grad_bias, 0, (size_t)(bias_stride_h ? g.H : 1) * g.L * g.J * sizeof(float), st);
  if (ez == cudaSuccess && !eva::window_bwd_tc_supported(g, gin->io_dtype, pad_mask))
    ez = cudaMemsetAsync(grad_qkv, 0, 3 * (size_t)tens * sizeof(float), st);
  if (ez != cudaSuccess) return cuda_fail(ez, "eva_backward(memset)");
  if (g.n_chunks > 0 && !k_bar_in) {
    const cudaError_t e0 = eva::launch_chunk_stats(g, gin->io_dtype, vq, vk, vv, pad_mask, *ada, noise, chunk_rows, chunk_rows + slot, st);
    if (e0 != cudaSuccess) return cuda_fail(e0, "eva_backward(chunk_stats)");
  }
  const cudaError_t e = eva::launch_eva_backward(g, gin->io_dtype, vq, vk, vv, pad_mask, ada, noise, kbar, beta, bias, bias_stride_h, out,
                                                 grad_out, grad_qkv, grad_qkv + tens, grad_qkv + 2 * tens, chunk_rows + 2 * slot,
                                                 chunk_rows + 3 * slot, grad_bias, chunk_rows + 4 * slot, grad_qkv_io, st, lse);
  return e == cudaSuccess ? EVA_OK : cuda_fail(e, "eva_backward");
}

int lara_backward_step(int32_t which, int32_t io_dtype, void* X, void* Y, const void* dW, void* M2, const float* v0, const float* v1,
                       const float* v2, const float* v3, float* o0, float* o1, float* o2, int64_t x_item_stride, int64_t y_item_stride,
                       int32_t items, int32_t landmarks, int32_t tokens, float scale, float alpha_coeff, void* stream) {
  if (which < 0 || which > 2) return fail(EVA_ERR_INVALID, "which must be 0, 1 or 2");
  if (io_dtype < EVA_F32 || io_dtype > EVA_BF16) return fail(EVA_ERR_INVALID, "unknown io_dtype %d", io_dtype);
  if (items <= 0 || landmarks <= 0 || tokens <= 0) return fail(EVA_ERR_INVALID, "items / landmarks / tokens must be positive");
  if (!X || !v0 || (which != 1 && !Y) || (which == 1 && (!dW || !M2 || !v1 || !v2 || !v3 || !o0 || !o1 || !o2)) || (which == 0 && (!o0 || !o1)))
    return fail(EVA_ERR_INVALID, "lara_backward_step(%d): a required pointer is NULL", which);
  if (which == 1 && items > 65535) return fail(EVA_ERR_UNSUPPORTED, "more than 65535 (batch x head) items in one call");
  const cudaError_t e = eva::launch_lara_bwd(which, io_dtype, X, Y, dW, M2, v0, v1, v2, v3, o0, o1, o2, x_item_stride, y_item_stride, items,
                                             landmarks, tokens, scale, alpha_coeff, reinterpret_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? EVA_OK : cuda_fail(e, "lara_backward_step");
}

// ---------------------------------------------------------------------------------------------
static int make_lara_geo(const LaraGeometry* in, eva::LaraGeo* g) {
  if (!in) return fail(EVA_ERR_INVALID, "geometry is NULL");
  if (in->batch <= 0 || in->heads <= 0 || in->tokens <= 0 || in->landmarks <= 0) return fail(EVA_ERR_INVALID, "batch/heads/tokens/landmarks must be positive");
  if (in->head_dim != 16 && in->head_dim != 32 && in->head_dim != 64 && in->head_dim != 128)
    return fail(EVA_ERR_UNSUPPORTED, "head_dim %d not built (16, 32, 64, 128)", in->head_dim);
  if (in->io_dtype < EVA_F32 || in->io_dtype > EVA_BF16) return fail(EVA_ERR_INVALID, "unknown io_dtype %d", in->io_dtype);
  if (in->mis_type < LARA_MIS_OPT || in->mis_type > LARA_MIS_BIASED) return fail(EVA_ERR_INVALID, "unknown mis_type");
  if (in->sample_mode < LARA_SAMPLE_SINGLE || in->sample_mode > LARA_SAMPLE_MULTI) return fail(EVA_ERR_INVALID, "unknown sample_mode");
  g->B = in->batch; g->H = in->heads; g->N = in->tokens; g->D = in->head_dim;
  g->dims = in->dims; g->gh = in->grid_h; g->gw = in->grid_w;
  g->C = in->landmarks;
  g->side = 0;
  if (in->dims == 2) {
    if (in->grid_h * in->grid_w != in->tokens) return fail(EVA_ERR_INVALID, "grid does not cover tokens");
    int side = 1;
    while ((side + 1) * (side + 1) <= in->landmarks) ++side;
    if (side * side != in->landmarks) return fail(EVA_ERR_INVALID, "2-D landmarks must be a square number");
    if (side > in->grid_h || side > in->grid_w) return fail(EVA_ERR_INVALID, "more landmarks per side than grid cells");
    if (in->per_token_proj) return fail(EVA_ERR_INVALID, "'adaptive-1d' proposals are 1-D only");
    g->side = side;
  } else if (in->dims == 1) {
    if (in->landmarks > in->tokens) return fail(EVA_ERR_INVALID, "1-D landmarks must be <= tokens (pass min(num_landmarks, tokens))");
    if (in->mixed) return fail(EVA_ERR_INVALID, "landmark mixing is 2-D only (lara.py:157)");
  } else {
    return fail(EVA_ERR_INVALID, "dims must be 1 or 2");
  }
  g->S = in->sample_mode == LARA_SAMPLE_SINGLE ? g->C : 2 * g->C;
  g->per_token_proj = in->per_token_proj ? 1 : 0;
  g->mixed = in->mixed; g->mis_type = in->mis_type; g->sample_mode = in->sample_mode;
  g->zero_padded = in->zero_padded ? 1 : 0;
  g->alpha_coeff = in->alpha_coeff;
  return EVA_OK;
}

int lara_forward_workspace_bytes(const LaraGeometry* gin, size_t* bytes) {
  eva::LaraGeo g{};
  const int rc = make_lara_geo(gin, &g);
  if (rc) return rc;
  if (!bytes) return fail(EVA_ERR_INVALID, "bytes is NULL");
  *bytes = eva::lara_workspace_bytes(g);
  return EVA_OK;
}

static int lara_forward_impl(const LaraGeometry* gin, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                             const uint8_t* pad_mask, const EvaAdaptive* proj, const float* noise,
                             void* out, void* workspace, size_t workspace_bytes, void* stream, const float* given);

int lara_forward(const LaraGeometry* gin, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                 const uint8_t* pad_mask, const EvaAdaptive* proj, const float* noise,
                 void* out, void* workspace, size_t workspace_bytes, void* stream) {
  return lara_forward_impl(gin, q, k, v, pad_mask, proj, noise, out, workspace, workspace_bytes, stream, nullptr);
}

int lara_forward_given_landmarks(const LaraGeometry* gin, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                                 const uint8_t* pad_mask, const float* landmarks, const float* noise,
                                 void* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!landmarks) return fail(EVA_ERR_INVALID, "landmarks is NULL");
  if (gin && gin->per_token_proj) return fail(EVA_ERR_INVALID, "'adaptive-1d' proposals are computed per token, not given");
  EvaAdaptive none;
  memset(&none, 0, sizeof(none));
  none.ln_eps = 1e-5f;
  return lara_forward_impl(gin, q, k, v, pad_mask, &none, noise, out, workspace, workspace_bytes, stream, landmarks);
}

static int lara_forward_impl(const LaraGeometry* gin, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v,
                             const uint8_t* pad_mask, const EvaAdaptive* proj, const float* noise,
                             void* out, void* workspace, size_t workspace_bytes, void* stream, const float* given) {
  eva::LaraGeo g{};
  eva::View vq, vk, vv;
  int rc;
  if ((rc = make_lara_geo(gin, &g)) || (rc = make_view(q, "q", &vq)) || (rc = make_view(k, "k", &vk)) ||
      (rc = make_view(v, "v", &vv)))
    return rc;
  if (!proj) return fail(EVA_ERR_INVALID, "proj is NULL (pass a zeroed struct for 'no-param-pool')");
  if (proj->w_q && (!proj->w_k || !proj->b_q || !proj->b_k || !proj->ln_gain_q || !proj->ln_gain_k || !proj->ln_bias_q || !proj->ln_bias_k))
    return fail(EVA_ERR_INVALID, "q_bar_gen / k_bar_gen need Linear and LayerNorm parameters for both sides");
  if (g.per_token_proj && !proj->w_q) return fail(EVA_ERR_INVALID, "'adaptive-1d' needs projection parameters");
  if (!noise && g.sample_mode != LARA_SAMPLE_SINGLE) return fail(EVA_ERR_INVALID, "sample_mode without noise");
  if (!out || !workspace) return fail(EVA_ERR_INVALID, "out / workspace is NULL");
  if (workspace_bytes < eva::lara_workspace_bytes(g)) return fail(EVA_ERR_INVALID, "workspace too small");
  if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return fail(EVA_ERR_INVALID, "workspace must be 256-byte aligned");
  const cudaError_t e = eva::launch_lara(g, gin->io_dtype, vq, vk, vv, pad_mask, *proj, noise, out, workspace,
                                         reinterpret_cast<cudaStream_t>(stream), given);
  if (e == cudaErrorInvalidConfiguration)
    return fail(EVA_ERR_UNSUPPORTED, "landmarks x head_dim too large for one CTA's shared memory");
  return e == cudaSuccess ? EVA_OK : cuda_fail(e, "lara_forward");
}

}  // extern "C"
