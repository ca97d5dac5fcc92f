// Internal C++ launch interface between the C ABI (abi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eva_sm100.h"
#include "common.cuh"

namespace eva {

// error plumbing of the C ABI (abi.cu), for the translation units that export entry points of their own (rfa_kernels.cu)
int abi_fail(int code, const char* msg);
int abi_cuda_fail(cudaError_t e, const char* what);
int abi_view(const EvaHeadsView* in, const char* name, View* v);

cudaError_t launch_chunk_stats(const Geo& g, int io_dtype, const View& q, const View& k, const View& v,
                               const uint8_t* mask, const EvaAdaptive& ada, const float* noise,
                               float* kbar, float* beta, cudaStream_t st);

cudaError_t launch_window_attn(const Geo& g, int io_dtype, const View& q, const View& k, const View& v,
                               const uint8_t* mask, const float* kbar, const float* beta,
                               const float* bias, long long bias_sh, void* out, cudaStream_t st);

// Fused tcgen05/TMA path (eva_fused_sm100.cu).  `supported` says whether the geometry qualifies.
bool fused_supported(const Geo& g, int io_dtype, const View& q, const View& k, const View& v,
                     const uint8_t* mask, const EvaAdaptive& ada, const float* bias, long long bias_sh);
size_t fused_workspace_bytes(const Geo& g);
// returns cudaSuccess, or an error with *msg describing it
cudaError_t launch_fused(const Geo& g, int io_dtype, const View& q, const View& k, const View& v,
                         const EvaAdaptive& ada, const float* noise, const float* bias, long long bias_sh,
                         void* out, void* workspace, cudaStream_t st, const char** msg, float* kbar_out = nullptr,
                         float* beta_out = nullptr);   // kbar_out / beta_out: also leave the chunk statistics in global memory (training)

// Cluster-resident fused path (eva_cluster_sm100.cu): the c3 geometry (28 x 28 tokens, window 7, 4 x 4 chunks), one item per
// two-CTA cluster with k / v resident in shared memory; same workspace layout as the streamed fused kernel
bool cluster_supported(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask);
cudaError_t launch_cluster(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const EvaAdaptive& ada,
                           const float* noise, const float* bias, long long bias_sh, void* out, void* workspace, cudaStream_t st,
                           const char** msg);

// Causal window attention on tcgen05 (eva_causal_sm100.cu): stage B of the causal layer for window 256 / head_dim 64 / 16-bit I/O
bool causal_window_supported(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask,
                             const float* bias, long long bias_sh);
// ada + flags != NULL: one-pass mode -- the kernel computes the chunk statistics of every window itself (kbar / beta are then
// OUTPUT scratch, flags = B * H * (N / 256) words of scratch); see causal_one_pass_supported
bool causal_one_pass_supported(const Geo& g, const EvaAdaptive& ada);
cudaError_t launch_causal_window(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const float* kbar,
                                 const float* beta, const float* bias, void* out, cudaStream_t st, const char** msg,
                                 const EvaAdaptive* ada = nullptr, const float* noise = nullptr, unsigned int* flags = nullptr,
                                 float* lse_out = nullptr);   // lse_out: log2-domain log-sum-exp per query row [B, H, N] (training)

// Backward of the two generic stages (eva_backward.cu); kbar / beta are the forward statistics, dkbar / dbeta zeroed scratch,
// rows = 8 per-chunk row slots (see chunk_stats_bwd_kernel); ada / noise / dkbar / dbeta / rows unused when g.n_chunks == 0;
// grad_io (optional): the final dq | dk | dv once more, rounded to the I/O format, packed [B, N, 3, H, D] (dq / dk / dv float32 are
// then scratch: the kernel that finishes a token may skip writing them back)
cudaError_t launch_eva_backward(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask,
                                const EvaAdaptive* ada, const float* noise, const float* kbar, const float* beta, const float* bias,
                                long long bias_sh, const void* out, const void* dout, float* dq, float* dk, float* dv, float* dkbar,
                                float* dbeta, float* dbias, float* rows, void* grad_io, cudaStream_t st, const float* lse = nullptr);

// Window attention on tcgen05 for any geometry with head_dim 64 and 16-bit I/O (eva_window_tc_sm100.cu); arguments as launch_window_attn
bool window_tc_supported(const Geo& g, int io_dtype);
cudaError_t launch_window_tc(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask,
                             const float* kbar, const float* beta, const float* bias, long long bias_sh, void* out, cudaStream_t st,
                             float* lse_out = nullptr, const float* key_bias = nullptr);   // key_bias: float32 [B, H, N] per-key logit addend

// tcgen05 window-attention backward (eva_bwd_sm100.cu): head_dim 64, 16-bit I/O, halo-free windows of <= 64 tokens, <= 64 chunks
bool window_bwd_tc_supported(const Geo& g, int io_dtype, const uint8_t* mask);
cudaError_t launch_window_bwd_tc(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const float* kbar,
                                 const float* beta, const float* bias, long long bias_sh, const void* out, const void* dout,
                                 float* dq, float* dk, float* dv, float* dkbar, float* dbeta, float* dbias, cudaStream_t st);

// tcgen05 window-attention backward for every other geometry with head_dim 64 and 16-bit I/O (eva_window_bwd_gen_sm100.cu)
bool bwd_tc_enabled();
void note_bwd_tc_launch();
bool window_bwd_gen_supported(const Geo& g, int io_dtype);
cudaError_t launch_window_bwd_gen(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask,
                                  const float* kbar, const float* beta, const float* bias, long long bias_sh, const void* out,
                                  const void* dout, float* dq, float* dk, float* dv, float* dkbar, float* dbeta, float* dbias,
                                  cudaStream_t st, const float* lse = nullptr);

// Performer / FAVOR+ on tcgen05 (rfa_tc_sm100.cu): 'favorp', 64 features, head_dim 64, 16-bit I/O
bool rfa_tc_supported(int method, int D, int m, int cosw, int io_dtype, const View& q, const View& k, const View& v);
cudaError_t launch_rfa_tc(int B, int H, int N, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask,
                          const float* proj, void* out, cudaStream_t st, float* sb_stabv = nullptr, float* sb_part = nullptr);
// sb_part != NULL: the ScatterBrain key statistics instead (per-feature maxima -> sb_stabv [items][64], KV | ksum -> sb_part
// [items][64 * 64 + 64], float32); q / out unused

// ScatterBrain window stage on tcgen05 (sb_window_tc_sm100.cu): 64 features, head_dim 64, 16-bit I/O, windows of <= 64 tokens
bool sb_window_tc_supported(int D, int m, int L, int io_dtype);
cudaError_t launch_sb_window_tc(int B, int H, int N, int dims, int gh, int gw, int w, int L, int n_windows, int io_dtype, const View& q,
                                const View& k, const View& v, const uint8_t* mask, const float* proj, const float* bias,
                                const float* stabv, const float* part, void* out, cudaStream_t st);

// Key draw of randomized attention by Gumbel-max on tcgen05 (ra_sample_tc_sm100.cu): head_dim 64, 16-bit I/O
bool ra_sample_tc_supported(int D, int io_dtype, const View& q, const View& k);
cudaError_t launch_ra_sample_tc(int B, int H, int N, int io_dtype, const View& q, const View& k, unsigned long long seed,
                                const float* gumbel, long long* k_ind, cudaStream_t st);

// LARA (lara_generic.cu)
struct LaraGeo {
  int B, H, N, D;
  int dims, gh, gw;
  int C, S, side;          // landmarks, samples, landmarks per grid side (2-D)
  int per_token_proj, mixed, mis_type, sample_mode, zero_padded;
  float alpha_coeff;
};
size_t lara_workspace_bytes(const LaraGeo& g);
// tcgen05 core (lara_core_sm100.cu): replaces the stats + out kernels for the DeiT geometry (mis-opt, S == C <= 64, N <= 224)
bool lara_core_supported(const LaraGeo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask);
bool lara_core_fuses_landmarks(const LaraGeo& g, const EvaAdaptive& proj);
cudaError_t launch_lara_core(const LaraGeo& g, int io_dtype, const View& q, const View& k, const View& v, const float* ws, void* out,
                             const EvaAdaptive* proj, const float* noise, cudaStream_t st);
cudaError_t launch_lara(const LaraGeo& g, int io_dtype, const View& q, const View& k, const View& v,
                        const uint8_t* mask, const EvaAdaptive& proj, const float* noise, void* out,
                        void* workspace, cudaStream_t st, const float* given_landmarks = nullptr);

// LARA backward helpers (lara_backward.cu): which = 0 rows1 (X, Bm, v0 = k2s -> o0 = lse_B, o1 = lse_T), 1 cols (X, dW, M2, v0 = q2s, v1 = bh,
// v2 = lp, v3 = lse_B -> o0 = d lse_B, o1 = d bh, o2 = R; zeroed by the caller), 2 row_affine (X <- Bm o (X - v0 + v1), item strides xs / ys)
cudaError_t launch_lara_bwd(int which, int io_dtype, void* X, void* Bm, const void* dW, void* M2, const float* v0, const float* v1,
                            const float* v2, const float* v3, float* o0, float* o1, float* o2, long long xs, long long ys, int BH, int C,
                            int N, float s, float coeff, cudaStream_t st);

}  // namespace eva
