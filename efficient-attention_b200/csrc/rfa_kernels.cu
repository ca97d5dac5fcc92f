// Random-feature attention for sm_100a: Performer-style linear attention (`rfa_forward`), randomized attention
// (`ra_forward`) and ScatterBrain (`scatterbrain_forward`).  Reference: kernelized_attention.py:12-153,301-320,
// randomized_attention.py:24-55, scatterbrain_attention.py:10-164.  float32 math on CUDA cores, 16-bit or float32 I/O; every
// kernel stages rows with 16-byte loads, keeps the projection matrix and the [features x head_dim] statistics in shared memory and
// never materialises a [tokens x features] or [tokens x tokens] matrix in HBM.
#include <math.h>
#include <stdio.h>

#include "common.cuh"
#include "launch.h"

namespace eva {
namespace rfa {

constexpr int kThreads = 256;
constexpr int kTT = 32;            // tokens per tile
constexpr int kMaxAcc = 16;        // feature rows per thread in the KV accumulation (M_eff <= 16 * 256 / (D / 4))
constexpr int RFA_LOG_FAVORP = 100;   // internal: log-features of scatterbrain_attention.py:10-44, exp(. - per-feature max) on the key side

struct Params {
  int B, H, N, D;
  int method, m, nu, Mb, Meff, cosw;      // m: projection rows; Mb: features before the cosFormer doubling; Meff: after
  int S;                                  // token splits per (batch, head) item
  const float* proj;                      // [H, m, D]
  const float* qf;                        // RFA_GIVEN: [B, H, N, Mb]
  const float* kf;
  const uint8_t* mask;                    // [B, N] or NULL
  float* stab;                            // [B*H][2]
  float* stabv;                           // log-FAVOR+ (ScatterBrain): [B*H][m] per-feature max of the keys' log-features
  float* part;                            // [B*H][S][Meff * D + Meff]
};

__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kThreads / 32; ++w) r = fmaxf(r, red[w]);
  return r;
}

// rows n0 .. n0 + kTT of one (batch, head) into xs[kTT][XS] (float32, zero past the sequence)
template <typename T>
__device__ __forceinline__ void load_rows(const View& x, int b, int h, int n0, int N, int D, int XS, float* xs) {
  const int pieces = D >> 3;
  for (int idx = threadIdx.x; idx < kTT * pieces; idx += kThreads) {
    const int r = idx / pieces, pc = idx - r * pieces;
    float o[8];
    if (n0 + r < N) load8<T>(x.row<T>(b, n0 + r, h) + 8 * pc, o);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
    }
    float4* dst = reinterpret_cast<float4*>(xs + r * XS + 8 * pc);
    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
  }
}

// dd[t][j] = dn * x_t . W_j for the tile (thread = token t, features jg, jg + 8, ...), hs[t] = dn^2 |x_t|^2 / 2
__device__ __forceinline__ void project_tile(const float* xs, int XS, const float* Ws, int m, int D, float dn, float* Fs, int MS, float* hs) {
  const int t = threadIdx.x & 31, jg = threadIdx.x >> 5;
  const float* xr = xs + t * XS;
  for (int j0 = jg; j0 < m; j0 += 64) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    float sq = 0.f;
    for (int d = 0; d < D; d += 4) {
      const float4 x4 = *reinterpret_cast<const float4*>(xr + d);
      if (j0 == jg) sq = fmaf(x4.x, x4.x, fmaf(x4.y, x4.y, fmaf(x4.z, x4.z, fmaf(x4.w, x4.w, sq))));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = j0 + 8 * i;
        if (j < m) {
          const float4 w4 = *reinterpret_cast<const float4*>(Ws + j * XS + d);
          acc[i] = fmaf(x4.x, w4.x, fmaf(x4.y, w4.y, fmaf(x4.z, w4.z, fmaf(x4.w, w4.w, acc[i]))));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = j0 + 8 * i;
      if (j < m) Fs[t * MS + j] = dn * acc[i];
    }
    if (j0 == jg && jg == 0) hs[t] = 0.5f * dn * dn * sq;
  }
}

// Turns the tile's raw projections (or rows) into features Fs[t][0 .. Meff).  is_q: query side.  stab: the (batch, head) stabiliser
// (favorp keys: max of dd; fourier: max of hs over the tokens).  Key rows that are padded or past the sequence become zero.
__device__ __forceinline__ void finish_features(const Params& p, bool is_q, const float* xs, int XS, float* Fs, int MS, const float* hs,
                                                float stab, int b, int bh, int n0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float ratio = rsqrtf((float)p.m);
  for (int r = warp; r < kTT; r += kThreads / 32) {
    const int n = n0 + r;
    float* f = Fs + r * MS;
    const bool dead = n >= p.N || (!is_q && p.mask && p.mask[(long long)b * p.N + n]);
    if (p.method == RFA_FAVORP) {
      float mx = kNegInf;
      if (is_q) {
        for (int j = lane; j < p.m; j += 32) mx = fmaxf(mx, f[j]);
        mx = warp_max(mx);
      } else mx = stab;
      const float sub = hs[r] + mx;
      for (int j = lane; j < p.m; j += 32) f[j] = ratio * expf(f[j] - sub) + 1e-4f;
    } else if (p.method == RFA_LOG_FAVORP) {          // keys only: exp(log-feature - max over the tokens), per feature
      const float sub = hs[r] + 0.5f * logf((float)p.m);
      for (int j = lane; j < p.m; j += 32) f[j] = expf(f[j] - sub - p.stabv[(long long)bh * p.m + j]);
    } else if (p.method == RFA_RELU) {
      for (int j = lane; j < p.m; j += 32) f[j] = fmaxf(ratio * f[j], 0.f) + 1e-3f;
    } else if (p.method == RFA_FOURIER) {
      const float c = ratio * expf(hs[r] - stab);
      for (int j = lane; j < p.m; j += 32) {
        float s_, c_;
        sincosf(f[j], &s_, &c_);
        f[j] = c * s_;
        f[p.m + j] = c * c_;
      }
    } else if (p.method == RFA_DPFP) {
      const int D2 = 2 * p.D;
      const float* xr = xs + r * XS;
      for (int j = lane; j < p.Mb; j += 32) {
        const int sh = j / D2 + 1, i = j - (sh - 1) * D2;
        int i2 = i - sh;
        if (i2 < 0) i2 += D2;
        const float a = i < p.D ? fmaxf(xr[i], 0.f) : fmaxf(-xr[i - p.D], 0.f);
        const float c = i2 < p.D ? fmaxf(xr[i2], 0.f) : fmaxf(-xr[i2 - p.D], 0.f);
        f[j] = a * c;
      }
    } else if (p.method == RFA_RELU_ONLY) {
      for (int j = lane; j < p.D; j += 32) f[j] = fmaxf(xs[r * XS + j], 0.f) + 0.1f;
    } else if (p.method == RFA_SIGMOID_ONLY) {
      for (int j = lane; j < p.D; j += 32) f[j] = 1.0f / (1.0f + expf(-xs[r * XS + j])) + 0.1f;
    } else {   // RFA_GIVEN
      const float* src = (is_q ? p.qf : p.kf) + ((long long)bh * p.N + (n < p.N ? n : 0)) * p.Mb;
      for (int j = lane; j < p.Mb; j += 32) f[j] = __ldg(src + j);
    }
    __syncwarp();
    if (p.cosw) {                       // cosFormer re-weighting: [phi cos(pi n / 2N) ; phi sin(pi n / 2N)]
      float s_, c_;
      sincosf(1.5707963267948966f * (float)n / (float)p.N, &s_, &c_);
      for (int j = lane; j < p.Mb; j += 32) { const float v = f[j]; f[j] = v * c_; f[p.Mb + j] = v * s_; }
    }
    __syncwarp();
    if (dead)
      for (int j = lane; j < p.Meff; j += 32) f[j] = 0.f;
  }
}

__device__ __forceinline__ bool needs_proj(int method) { return method == RFA_FAVORP || method == RFA_RELU || method == RFA_FOURIER || method == RFA_LOG_FAVORP; }

__device__ __forceinline__ void load_proj(const Params& p, int h, int XS, float* Ws) {
  if (!needs_proj(p.method)) return;
  const float* W = p.proj + (long long)h * p.m * p.D;
  for (int idx = threadIdx.x; idx < p.m * p.D; idx += kThreads) Ws[(idx / p.D) * XS + idx % p.D] = __ldg(W + idx);
}

// ---- stabilisers: favorp -> stab[bh][1] = max_{n, j} dd_k; fourier -> stab[bh][0 / 1] = max_n hs of q / k -----------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) rfa_stab_kernel(const View q, const View k, const Params p) {
  extern __shared__ float smf[];
  const int XS = p.D + 4, MS = p.Meff + 1;
  float* Ws = smf;
  float* xs = Ws + (size_t)p.m * XS;
  float* Fs = xs + kTT * XS;
  float* hs = Fs + kTT * MS;
  float* red = hs + kTT;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const float dn = rsqrtf(sqrtf((float)p.D));
  load_proj(p, h, XS, Ws);
  float mx[2] = {kNegInf, kNegInf};
  for (int side = (p.method == RFA_FOURIER ? 0 : 1); side < 2; ++side) {
    const View& x = side ? k : q;
    for (int n0 = 0; n0 < p.N; n0 += kTT) {
      __syncthreads();
      load_rows<T>(x, b, h, n0, p.N, p.D, XS, xs);
      __syncthreads();
      if (p.method == RFA_LOG_FAVORP) {
        project_tile(xs, XS, Ws, p.m, p.D, dn, Fs, MS, hs);
        __syncthreads();
        if (threadIdx.x < p.m)
          for (int r = 0; r < kTT && n0 + r < p.N; ++r)
            if (!(p.mask && p.mask[(long long)b * p.N + n0 + r])) mx[0] = fmaxf(mx[0], Fs[r * MS + threadIdx.x] - hs[r]);
      } else if (p.method == RFA_FAVORP) {
        project_tile(xs, XS, Ws, p.m, p.D, dn, Fs, MS, hs);
        __syncthreads();
        for (int idx = threadIdx.x; idx < kTT * p.m; idx += kThreads) {
          const int r = idx / p.m;
          if (n0 + r < p.N) mx[side] = fmaxf(mx[side], Fs[r * MS + idx - r * p.m]);
        }
      } else {
        const int r = threadIdx.x;
        if (r < kTT && n0 + r < p.N) {
          float sq = 0.f;
          for (int d = 0; d < p.D; ++d) sq = fmaf(xs[r * XS + d], xs[r * XS + d], sq);
          mx[side] = fmaxf(mx[side], 0.5f * dn * dn * sq);
        }
      }
    }
  }
  if (p.method == RFA_LOG_FAVORP) {
    if (threadIdx.x < p.m) p.stabv[(long long)bh * p.m + threadIdx.x] = mx[0] - 0.5f * logf((float)p.m);
    return;
  }
  const float m0 = block_max(mx[0], red), m1 = block_max(mx[1], red);
  if (threadIdx.x == 0) { p.stab[2 * bh] = m0; p.stab[2 * bh + 1] = m1; }
}

// ---- keys: partial KV[j][d] = sum_n phi(k_n)_j v_n[d], ksum[j] = sum_n phi(k_n)_j over the tiles of one split --------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) rfa_kv_kernel(const View k, const View v, const Params p) {
  extern __shared__ float smf[];
  const int XS = p.D + 4, MS = p.Meff + 1;
  float* Ws = smf;
  float* xs = Ws + (size_t)p.m * XS;
  float* vs = xs + kTT * XS;
  float* Fs = vs + kTT * XS;
  float* hs = Fs + kTT * MS;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H, sp = blockIdx.y;
  const float dn = rsqrtf(sqrtf((float)p.D));
  load_proj(p, h, XS, Ws);
  const float stab = p.stab[2 * bh + 1];
  const int dgs = p.D >> 2, mgs = kThreads / dgs;              // thread = (4 consecutive d, feature rows mg, mg + mgs, ...)
  const int dg = threadIdx.x % dgs, mg = threadIdx.x / dgs;
  const int nI = (p.Meff + mgs - 1) / mgs;
  float acc[kMaxAcc][4], ks[kMaxAcc];
#pragma unroll
  for (int i = 0; i < kMaxAcc; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; ks[i] = 0.f; }
  for (int n0 = sp * kTT; n0 < p.N; n0 += p.S * kTT) {
    __syncthreads();
    load_rows<T>(k, b, h, n0, p.N, p.D, XS, xs);
    load_rows<T>(v, b, h, n0, p.N, p.D, XS, vs);
    __syncthreads();
    if (needs_proj(p.method)) {
      project_tile(xs, XS, Ws, p.m, p.D, dn, Fs, MS, hs);
      __syncthreads();
    }
    finish_features(p, false, xs, XS, Fs, MS, hs, stab, b, bh, n0);
    __syncthreads();
    for (int t = 0; t < kTT; ++t) {
      const float4 v4 = *reinterpret_cast<const float4*>(vs + t * XS + 4 * dg);
#pragma unroll
      for (int i = 0; i < kMaxAcc; ++i) {
        if (i < nI) {
          const int j = mg + i * mgs;
          const float f = j < p.Meff ? Fs[t * MS + j] : 0.f;
          acc[i][0] = fmaf(f, v4.x, acc[i][0]); acc[i][1] = fmaf(f, v4.y, acc[i][1]);
          acc[i][2] = fmaf(f, v4.z, acc[i][2]); acc[i][3] = fmaf(f, v4.w, acc[i][3]);
          ks[i] += f;
        }
      }
    }
  }
  float* out = p.part + ((long long)bh * p.S + sp) * ((long long)p.Meff * p.D + p.Meff);
#pragma unroll
  for (int i = 0; i < kMaxAcc; ++i) {
    const int j = mg + i * mgs;
    if (i < nI && j < p.Meff) {
      *reinterpret_cast<float4*>(out + (long long)j * p.D + 4 * dg) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      if (dg == 0) out[(long long)p.Meff * p.D + j] = ks[i];
    }
  }
}

// ---- queries: out_n = phi(q_n) KV / max(phi(q_n) . ksum, 1e-2) ---------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) rfa_out_kernel(const View q, T* __restrict__ out, const Params p) {
  extern __shared__ float smf[];
  const int XS = p.D + 4, MS = p.Meff + 1;
  float* Ws = smf;
  float* xs = Ws + (size_t)p.m * XS;
  float* Fs = xs + kTT * XS;
  float* hs = Fs + kTT * MS;
  float* KV = hs + kTT;                       // [Meff][D] then ksum [Meff]
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H, sp = blockIdx.y;
  const float dn = rsqrtf(sqrtf((float)p.D));
  load_proj(p, h, XS, Ws);
  const int kvn = p.Meff * p.D + p.Meff;
  for (int idx = threadIdx.x; idx < kvn; idx += kThreads) {
    float s = 0.f;
    for (int u = 0; u < p.S; ++u) s += p.part[((long long)bh * p.S + u) * kvn + idx];
    KV[idx] = s;
  }
  const float* ksum = KV + p.Meff * p.D;
  const float stab = p.stab[2 * bh];
  const int t = threadIdx.x >> 3, dgp = threadIdx.x & 7, DPT = p.D >> 3;     // thread = (token t, DPT consecutive features)
  for (int n0 = sp * kTT; n0 < p.N; n0 += p.S * kTT) {
    __syncthreads();
    load_rows<T>(q, b, h, n0, p.N, p.D, XS, xs);
    __syncthreads();
    if (needs_proj(p.method)) {
      project_tile(xs, XS, Ws, p.m, p.D, dn, Fs, MS, hs);
      __syncthreads();
    }
    finish_features(p, true, xs, XS, Fs, MS, hs, stab, b, bh, n0);
    __syncthreads();
    float o[16], den = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = 0.f;
    for (int j = 0; j < p.Meff; ++j) {
      const float f = Fs[t * MS + j];
      den = fmaf(f, ksum[j], den);
      const float* kvr = KV + j * p.D + dgp * DPT;
      if ((DPT & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          if (i < DPT) {
            const float4 kv4 = *reinterpret_cast<const float4*>(kvr + i);
            o[i] = fmaf(f, kv4.x, o[i]); o[i + 1] = fmaf(f, kv4.y, o[i + 1]); o[i + 2] = fmaf(f, kv4.z, o[i + 2]); o[i + 3] = fmaf(f, kv4.w, o[i + 3]);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (i < DPT) o[i] = fmaf(f, kvr[i], o[i]);
      }
    }
    const int n = n0 + t;
    if (n < p.N) {
      const float inv = 1.0f / fmaxf(den, 1e-2f);
      T* dst = out + ((long long)b * p.N + n) * ((long long)p.H * p.D) + (long long)h * p.D + dgp * DPT;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < DPT) dst[i] = from_f32<T>(o[i] * inv);
    }
  }
}

struct Plan { Params p; size_t smem_stab, smem_kv, smem_out, ws_bytes; };

static int make_plan(const RfaGeometry* g, Plan* pl, const char** why) {
  Params& p = pl->p;
  p = Params{};
  if (!g) { *why = "geometry is NULL"; return EVA_ERR_INVALID; }
  if (g->batch <= 0 || g->heads <= 0 || g->tokens <= 0) { *why = "batch / heads / tokens must be positive"; return EVA_ERR_INVALID; }
  if (g->head_dim % 8 || g->head_dim < 8 || g->head_dim > 128) { *why = "head_dim must be a multiple of 8, at most 128"; return EVA_ERR_UNSUPPORTED; }
  if (g->io_dtype < EVA_F32 || g->io_dtype > EVA_BF16) { *why = "unknown io_dtype"; return EVA_ERR_INVALID; }
  p.B = g->batch; p.H = g->heads; p.N = g->tokens; p.D = g->head_dim; p.method = g->method; p.m = 0; p.nu = g->nu; p.cosw = g->cos_weighting ? 1 : 0;
  switch (g->method) {
    case RFA_FAVORP: case RFA_RELU: p.m = g->proj_dim; p.Mb = g->proj_dim; break;
    case RFA_FOURIER: p.m = g->proj_dim; p.Mb = 2 * g->proj_dim; break;
    case RFA_DPFP: p.Mb = 2 * g->head_dim * g->nu; break;
    case RFA_RELU_ONLY: case RFA_SIGMOID_ONLY: p.Mb = g->head_dim; break;
    case RFA_GIVEN: p.Mb = g->feat_dim; break;
    default: *why = "unknown feature method"; return EVA_ERR_INVALID;
  }
  if (p.Mb <= 0) { *why = "feature dimension must be positive"; return EVA_ERR_INVALID; }
  p.Meff = p.cosw ? 2 * p.Mb : p.Mb;
  const int mgs = kThreads / (p.D / 4);
  if (p.Meff > kMaxAcc * mgs) { *why = "too many features for this head_dim (M_eff * head_dim <= 16384)"; return EVA_ERR_UNSUPPORTED; }
  const int tiles = (p.N + kTT - 1) / kTT;
  int S = (296 + p.B * p.H - 1) / (p.B * p.H);
  S = S < 1 ? 1 : (S > 8 ? 8 : S);
  p.S = S > tiles ? tiles : S;
  const size_t XS = p.D + 4, MS = p.Meff + 1;
  const size_t base = ((size_t)p.m * XS + kTT * XS + kTT * MS + kTT) * 4;
  pl->smem_stab = base + 64;
  pl->smem_kv = base + kTT * XS * 4;
  pl->smem_out = base + ((size_t)p.Meff * p.D + p.Meff) * 4;
  if (pl->smem_kv > 200 * 1024 || pl->smem_out > 200 * 1024) { *why = "feature / projection sizes exceed shared memory"; return EVA_ERR_UNSUPPORTED; }
  pl->ws_bytes = 256 + (size_t)p.B * p.H * 2 * 4 + (size_t)p.B * p.H * p.S * ((size_t)p.Meff * p.D + p.Meff) * 4;
  pl->ws_bytes = (pl->ws_bytes + 255) & ~(size_t)255;
  return EVA_OK;
}

template <typename T>
static cudaError_t run(const Plan& pl, const View& q, const View& k, const View& v, void* out, cudaStream_t st) {
  const Params& p = pl.p;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(rfa_stab_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_stab)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(rfa_kv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_kv)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(rfa_out_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_out)) != cudaSuccess) return e;
  if (p.method == RFA_FAVORP || p.method == RFA_FOURIER) rfa_stab_kernel<T><<<p.B * p.H, kThreads, pl.smem_stab, st>>>(q, k, p);
  rfa_kv_kernel<T><<<dim3(p.B * p.H, p.S), kThreads, pl.smem_kv, st>>>(k, v, p);
  rfa_out_kernel<T><<<dim3(p.B * p.H, p.S), kThreads, pl.smem_out, st>>>(q, reinterpret_cast<T*>(out), p);
  return cudaGetLastError();
}


// =====================================================================================================================================
// ScatterBrain (scatterbrain_attention.py:95-160): per window, local logits and m "random-feature keys" whose logit is
// log phi(q_i)_c + log(sum over the keys OUTSIDE the window of phi(k)_c) and whose value is the phi-weighted mean of v over those
// keys, in one joint softmax.  The global sums come from rfa_stab_kernel / rfa_kv_kernel (method RFA_LOG_FAVORP).
// =====================================================================================================================================
struct SbParams {
  Params r;
  int dims, gh, gw, w, L;                 // windows without halo: L = w (1-D) or w * w (2-D) tokens, queries == keys
  const float* bias;                      // [H, L, L] or NULL
};

__device__ __forceinline__ int window_token(const SbParams& p, int g, int l) {
  if (p.dims == 2) {
    const int ngx = p.gw / p.w;
    return ((g / ngx) * p.w + l / p.w) * p.gw + (g % ngx) * p.w + l % p.w;
  }
  return g * p.w + l;
}

template <typename T>
__device__ __forceinline__ void load_window_rows(const View& x, const SbParams& p, int b, int h, int g, int l0, int XS, float* xs) {
  const int pieces = p.r.D >> 3;
  for (int idx = threadIdx.x; idx < kTT * pieces; idx += kThreads) {
    const int r = idx / pieces, pc = idx - r * pieces;
    float o[8];
    if (l0 + r < p.L) load8<T>(x.row<T>(b, window_token(p, g, l0 + r), h) + 8 * pc, o);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
    }
    float4* dst = reinterpret_cast<float4*>(xs + r * XS + 8 * pc);
    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) sb_window_kernel(const View q, const View k, const View v, T* __restrict__ out, const SbParams p) {
  extern __shared__ float smf[];
  const Params& r = p.r;
  const int D = r.D, m = r.m, XS = D + 4, MS = m + 1, LP = 64;       // up to 64 tokens per window (two tiles of 32)
  float* Ws = smf;
  float* qs = Ws + (size_t)m * XS;
  float* ks = qs + LP * XS;
  float* vs = ks + LP * XS;
  float* LQ = vs + LP * XS;              // [LP][MS] log phi(q)
  float* PK = LQ + LP * MS;              // [LP][MS] exp(log phi(k) - max)
  float* KV = PK + LP * MS;              // [m][D] values of the random-feature keys
  float* nl = KV + (size_t)m * D;        // [m] log of the non-local feature mass
  float* hq = nl + m;                    // [LP]
  float* hk = hq + LP;                   // [LP]
  float* Ps = hk + LP;                   // [8 warps][LP + m] probabilities of the row in flight
  uint8_t* dead = reinterpret_cast<uint8_t*>(Ps + 8 * (LP + m));   // [LP] key is padding
  const int bh = blockIdx.x, b = bh / r.H, h = bh % r.H, g = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float dn = rsqrtf(sqrtf((float)D)), half_log_m = 0.5f * logf((float)m), scale = rsqrtf((float)D);
  load_proj(r, h, XS, Ws);
  for (int l0 = 0; l0 < p.L; l0 += kTT) {
    load_window_rows<T>(q, p, b, h, g, l0, XS, qs + l0 * XS);
    load_window_rows<T>(k, p, b, h, g, l0, XS, ks + l0 * XS);
    load_window_rows<T>(v, p, b, h, g, l0, XS, vs + l0 * XS);
  }
  for (int l = threadIdx.x; l < LP; l += kThreads)
    dead[l] = (l >= p.L || (r.mask && r.mask[(long long)b * r.N + window_token(p, g, l)])) ? 1 : 0;
  __syncthreads();
  for (int l0 = 0; l0 < p.L; l0 += kTT) {
    project_tile(qs + l0 * XS, XS, Ws, m, D, dn, LQ + l0 * MS, MS, hq + l0);
    project_tile(ks + l0 * XS, XS, Ws, m, D, dn, PK + l0 * MS, MS, hk + l0);
  }
  __syncthreads();
  const float* stabv = r.stabv + (long long)bh * m;
  for (int idx = threadIdx.x; idx < p.L * m; idx += kThreads) {
    const int l = idx / m, c = idx - l * m;
    LQ[l * MS + c] = LQ[l * MS + c] - hq[l] - half_log_m;
    PK[l * MS + c] = dead[l] ? 0.f : expf(PK[l * MS + c] - hk[l] - half_log_m - stabv[c]);
  }
  __syncthreads();
  {   // values and log-mass of the random-feature keys: global sums minus this window's
    const int dgs = D >> 2, mgs = kThreads / dgs, dg = threadIdx.x % dgs, mg = threadIdx.x / dgs;
    const int kvn = m * D + m;
    const float* part = r.part + (long long)bh * r.S * kvn;
    for (int c = mg; c < m; c += mgs) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, ls = 0.f;
      for (int l = 0; l < p.L; ++l) {
        const float f = PK[l * MS + c];
        const float4 v4 = *reinterpret_cast<const float4*>(vs + l * XS + 4 * dg);
        a0 = fmaf(f, v4.x, a0); a1 = fmaf(f, v4.y, a1); a2 = fmaf(f, v4.z, a2); a3 = fmaf(f, v4.w, a3);
        ls += f;
      }
      float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f, gs = 0.f;
      for (int u = 0; u < r.S; ++u) {
        const float4 g4 = *reinterpret_cast<const float4*>(part + (long long)u * kvn + c * D + 4 * dg);
        g0 += g4.x; g1 += g4.y; g2 += g4.z; g3 += g4.w;
        gs += part[(long long)u * kvn + m * D + c];
      }
      const float inv = 1.0f / fmaxf(gs - ls, 1e-3f);
      *reinterpret_cast<float4*>(KV + c * D + 4 * dg) = make_float4((g0 - a0) * inv, (g1 - a1) * inv, (g2 - a2) * inv, (g3 - a3) * inv);
      if (dg == 0) {   // log_add_exp(glse, llse, mask = (1, -1)) of attn_utils.py:44-51, both relative to the per-feature max
        const float glse = logf(gs), llse = logf(ls), a = fmaxf(glse, llse);
        nl[c] = stabv[c] + a + logf(expf(glse - a) - expf(llse - a) + 1e-5f);
      }
    }
  }
  __syncthreads();
  float* Pw = Ps + warp * (LP + m);
  const int DPL = (D + 31) >> 5;
  for (int i = warp; i < p.L; i += kThreads / 32) {
    // logits of row i: lane = key j (two per lane), then the m feature keys
    float sl[2], mx = kNegInf;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = lane + 32 * u;
      float s = kNegInf;
      if (j < p.L && !dead[j]) {
        float acc = 0.f;
        for (int d = 0; d < D; d += 4) {
          const float4 a = *reinterpret_cast<const float4*>(qs + i * XS + d), c4 = *reinterpret_cast<const float4*>(ks + j * XS + d);
          acc = fmaf(a.x, c4.x, fmaf(a.y, c4.y, fmaf(a.z, c4.z, fmaf(a.w, c4.w, acc))));
        }
        s = scale * acc + (p.bias ? __ldg(p.bias + ((long long)h * p.L + i) * p.L + j) : 0.f);
      }
      sl[u] = s;
      mx = fmaxf(mx, s);
    }
    for (int c = lane; c < m; c += 32) mx = fmaxf(mx, LQ[i * MS + c] + nl[c]);
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float e = expf(sl[u] - mx);
      Pw[lane + 32 * u] = e;
      sum += e;
    }
    for (int c = lane; c < m; c += 32) {
      const float e = expf(LQ[i * MS + c] + nl[c] - mx);
      Pw[LP + c] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < p.L; ++j) {
      const float pj = Pw[j];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < DPL && lane + 32 * u < D) o[u] = fmaf(pj, vs[j * XS + lane + 32 * u], o[u]);
    }
    for (int c = 0; c < m; ++c) {
      const float pc = Pw[LP + c];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < DPL && lane + 32 * u < D) o[u] = fmaf(pc, KV[c * D + lane + 32 * u], o[u]);
    }
    const float inv = 1.0f / sum;
    T* dst = out + ((long long)b * r.N + window_token(p, g, i)) * ((long long)r.H * D) + (long long)h * D;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < DPL && lane + 32 * u < D) dst[lane + 32 * u] = from_f32<T>(o[u] * inv);
    __syncwarp();
  }
}

template <typename T>
static cudaError_t run_sb(const Plan& pl, const SbParams& sp, size_t smem_win, int windows, const View& q, const View& k, const View& v,
                          void* out, cudaStream_t st, bool sb_global_tc, int io_dtype) {
  const Params& p = pl.p;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(rfa_stab_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_stab)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(rfa_kv_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_kv)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(sb_window_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_win)) != cudaSuccess) return e;
  if (sb_global_tc) {          // global key statistics on tcgen05 (rfa_tc_sm100.cu, kLogF); p.S == 1 there
    if ((e = launch_rfa_tc(p.B, p.H, p.N, io_dtype, k, k, v, p.mask, p.proj, nullptr, st, p.stabv, p.part)) != cudaSuccess) return e;
  } else {
    rfa_stab_kernel<T><<<p.B * p.H, kThreads, pl.smem_stab, st>>>(k, k, p);
    rfa_kv_kernel<T><<<dim3(p.B * p.H, p.S), kThreads, pl.smem_kv, st>>>(k, v, p);
  }
  if (sb_global_tc && sb_window_tc_supported(p.D, p.m, sp.L, io_dtype))
    return launch_sb_window_tc(p.B, p.H, p.N, sp.dims, sp.gh, sp.gw, sp.w, sp.L, windows, io_dtype, q, k, v, p.mask, p.proj, sp.bias, p.stabv,
                               p.part, out, st);
  sb_window_kernel<T><<<dim3(p.B * p.H, windows), kThreads, smem_win, st>>>(q, k, v, reinterpret_cast<T*>(out), sp);
  return cudaGetLastError();
}

// =====================================================================================================================================
// Randomized attention (randomized_attention.py:24-55): out_n = softmax_m(scale w_n . k_m - scale |k_m|^2 / 2) v_m with
// w_n = q_n + extra_n (+ noise_n); extra = mean of k | a caller-supplied row (E_pi[k]) | k[k_ind[n]].
// =====================================================================================================================================
struct RaParams {
  int B, H, N, D, mode;                   // mode 0: mean of k, 1: extra rows given, 2: gather by k_ind
  const void* extra;                      // mode 1: io_dtype [B, N, H * D]
  const long long* k_ind;                 // mode 2: int64 [B, H, N]
  const float* noise;                     // [B, H, N, D] or NULL
  float* kmean;                           // mode 0: [B * H][D] scratch
};

template <typename T>
__global__ void __launch_bounds__(kThreads) ra_kmean_kernel(const View k, const RaParams p) {
  __shared__ float red[kThreads / 32][128];
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n = warp; n < p.N; n += kThreads / 32) {
    const T* row = k.row<T>(b, n, h);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (lane + 32 * u < p.D) acc[u] += to_f32<T>(row[lane + 32 * u]);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) red[warp][lane + 32 * u] = acc[u];
  __syncthreads();
  if (threadIdx.x < p.D) {
    float s = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) s += red[w][threadIdx.x];
    p.kmean[(long long)bh * p.D + threadIdx.x] = s / (float)p.N;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) ra_attn_kernel(const View q, const View k, const View v, T* __restrict__ out, const RaParams p) {
  extern __shared__ float smf[];
  const int D = p.D, XS = D + 4;
  float* wsm = smf;                      // [32][XS] scale * (q + extra + noise)
  float* ks = wsm + kTT * XS;
  float* vs = ks + kTT * XS;
  float* kb = vs + kTT * XS;             // [32] -scale |k|^2 / 2 (-inf past the sequence)
  float* Ps = kb + kTT;                  // [8 warps][4 rows][32]
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H, n0 = blockIdx.y * kTT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale = rsqrtf((float)D);
  load_rows<T>(q, b, h, n0, p.N, D, XS, wsm);
  __syncthreads();
  for (int idx = threadIdx.x; idx < kTT * D; idx += kThreads) {
    const int r = idx / D, d = idx - r * D, n = n0 + r;
    float w = wsm[r * XS + d];
    if (n < p.N) {
      if (p.mode == 0) w += p.kmean[(long long)bh * D + d];
      else if (p.mode == 1) w += to_f32<T>(reinterpret_cast<const T*>(p.extra)[((long long)b * p.N + n) * ((long long)p.H * D) + (long long)h * D + d]);
      else w += to_f32<T>(k.row<T>(b, (int)p.k_ind[(long long)bh * p.N + n], h)[d]);
      if (p.noise) w += __ldg(p.noise + ((long long)bh * p.N + n) * D + d);
    }
    wsm[r * XS + d] = scale * w;
  }
  const int DPL = (D + 31) >> 5;
  float mrun[4], lrun[4], o[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a) { mrun[a] = kNegInf; lrun[a] = 0.f; o[a][0] = o[a][1] = o[a][2] = o[a][3] = 0.f; }
  for (int m0 = 0; m0 < p.N; m0 += kTT) {
    __syncthreads();
    load_rows<T>(k, b, h, m0, p.N, D, XS, ks);
    load_rows<T>(v, b, h, m0, p.N, D, XS, vs);
    __syncthreads();
    if (threadIdx.x < kTT) {
      float sq = 0.f;
      for (int d = 0; d < D; ++d) sq = fmaf(ks[threadIdx.x * XS + d], ks[threadIdx.x * XS + d], sq);
      kb[threadIdx.x] = m0 + threadIdx.x < p.N ? -0.5f * scale * sq : kNegInf;
    }
    __syncthreads();
    // the warp's four query rows against key `lane`
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int d = 0; d < D; d += 4) {
      const float4 k4 = *reinterpret_cast<const float4*>(ks + lane * XS + d);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float4 w4 = *reinterpret_cast<const float4*>(wsm + (4 * warp + a) * XS + d);
        s[a] = fmaf(w4.x, k4.x, fmaf(w4.y, k4.y, fmaf(w4.z, k4.z, fmaf(w4.w, k4.w, s[a]))));
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float lg = s[a] + kb[lane];
      const float mn = fmaxf(mrun[a], warp_max(lg));
      const float corr = expf(mrun[a] - mn), e = expf(lg - mn);
      mrun[a] = mn;
      lrun[a] = lrun[a] * corr + warp_sum(e);
#pragma unroll
      for (int u = 0; u < 4; ++u) o[a][u] *= corr;
      Ps[(warp * 4 + a) * 32 + lane] = e;
    }
    __syncwarp();
    for (int j = 0; j < kTT; ++j) {
      float vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) vv[u] = (u < DPL && lane + 32 * u < D) ? vs[j * XS + lane + 32 * u] : 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float pj = Ps[(warp * 4 + a) * 32 + j];
#pragma unroll
        for (int u = 0; u < 4; ++u) o[a][u] = fmaf(pj, vv[u], o[a][u]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int n = n0 + 4 * warp + a;
    if (n < p.N) {
      const float inv = 1.0f / lrun[a];
      T* dst = out + ((long long)b * p.N + n) * ((long long)p.H * D) + (long long)h * D;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < DPL && lane + 32 * u < D) dst[lane + 32 * u] = from_f32<T>(o[a][u] * inv);
    }
  }
}


// w = q + extra (+ noise) as a contiguous [B, N, H, D] tensor in the I/O format and kb[b, h, n] = -scale |k_n|^2 / 2: the inputs of the
// tensor-core route (the dense tcgen05 window kernel with a per-key logit addend)
template <typename T>
__global__ void __launch_bounds__(kThreads) ra_prepare_kernel(const View q, const View k, const RaParams p, T* __restrict__ w, float* __restrict__ kb) {
  const long long rows = (long long)p.B * p.N * p.H;
  const float scale = rsqrtf((float)p.D);
  const int pieces = p.D >> 3;
  for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < rows * pieces; idx += (long long)gridDim.x * kThreads) {
    const int pc = (int)(idx % pieces);
    const long long row = idx / pieces;
    const int h = (int)(row % p.H), n = (int)((row / p.H) % p.N), b = (int)(row / ((long long)p.H * p.N));
    const int bh = b * p.H + h;
    float x[8], e[8], kk[8];
    load8<T>(q.row<T>(b, n, h) + 8 * pc, x);
    load8<T>(k.row<T>(b, n, h) + 8 * pc, kk);
    if (p.mode == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) e[i] = p.kmean[(long long)bh * p.D + 8 * pc + i];
    } else if (p.mode == 1) {
      load8<T>(reinterpret_cast<const T*>(p.extra) + ((long long)b * p.N + n) * ((long long)p.H * p.D) + (long long)h * p.D + 8 * pc, e);
    } else {
      load8<T>(k.row<T>(b, (int)p.k_ind[(long long)bh * p.N + n], h) + 8 * pc, e);
    }
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      x[i] += e[i];
      if (p.noise) x[i] += __ldg(p.noise + ((long long)bh * p.N + n) * p.D + 8 * pc + i);
      sq = fmaf(kk[i], kk[i], sq);
    }
    T* dst = w + row * p.D + 8 * pc;
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = from_f32<T>(x[i]);
    // the `pieces` lanes of a row are consecutive lanes of one warp (pieces divides 32 for D = 64)
    const unsigned act = __activemask();              // groups of `pieces` lanes are active or inactive together
    for (int o = pieces >> 1; o > 0; o >>= 1) sq += __shfl_xor_sync(act, sq, o);
    if (pc == 0) kb[(long long)bh * p.N + n] = -0.5f * scale * sq;
  }
}

template <typename T>
static cudaError_t run_ra_tc(const RaParams& p, int io_dtype, const View& q, const View& k, const View& v, void* out, void* w16, float* kb,
                             cudaStream_t st) {
  if (p.mode == 0) ra_kmean_kernel<T><<<p.B * p.H, kThreads, 0, st>>>(k, p);
  const long long work = (long long)p.B * p.N * p.H * (p.D >> 3);
  const int grid = (int)((work + kThreads - 1) / kThreads < 148 * 8 ? (work + kThreads - 1) / kThreads : 148 * 8);
  ra_prepare_kernel<T><<<grid, kThreads, 0, st>>>(q, k, p, reinterpret_cast<T*>(w16), kb);
  Geo g{};
  g.B = p.B; g.H = p.H; g.N = p.N; g.D = p.D; g.dims = 1; g.gh = 1; g.gw = p.N; g.window = p.N; g.n_windows = 1; g.L = p.N; g.J = p.N;
  g.mask_fill = kNegInf;
  View vw{w16, (long long)p.N * p.H * p.D, (long long)p.H * p.D, (long long)p.D};
  return launch_window_tc(g, io_dtype, vw, k, v, nullptr, nullptr, nullptr, nullptr, 0, out, st, nullptr, kb);
}

template <typename T>
static cudaError_t run_ra(const RaParams& p, const View& q, const View& k, const View& v, void* out, cudaStream_t st) {
  if (p.mode == 0) ra_kmean_kernel<T><<<p.B * p.H, kThreads, 0, st>>>(k, p);
  const size_t smem = ((size_t)3 * kTT * (p.D + 4) + kTT + 8 * 4 * 32) * 4;
  ra_attn_kernel<T><<<dim3(p.B * p.H, (p.N + kTT - 1) / kTT), kThreads, smem, st>>>(q, k, v, reinterpret_cast<T*>(out), p);
  return cudaGetLastError();
}

}  // namespace rfa
}  // namespace eva

extern "C" {

int rfa_feature_dim(const RfaGeometry* g) {
  eva::rfa::Plan pl;
  const char* why = "";
  const int rc = eva::rfa::make_plan(g, &pl, &why);
  if (rc != EVA_OK) return eva::abi_fail(rc, why);
  return pl.p.Meff;
}

int rfa_forward_workspace_bytes(const RfaGeometry* g, size_t* bytes) {
  eva::rfa::Plan pl;
  const char* why = "";
  const int rc = eva::rfa::make_plan(g, &pl, &why);
  if (rc != EVA_OK) return eva::abi_fail(rc, why);
  if (!bytes) return eva::abi_fail(EVA_ERR_INVALID, "bytes is NULL");
  *bytes = pl.ws_bytes;
  return EVA_OK;
}

int rfa_forward(const RfaGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v, const uint8_t* pad_mask,
                const float* proj, const float* q_feat, const float* k_feat, void* out, void* workspace, size_t workspace_bytes,
                void* stream) {
  eva::rfa::Plan pl;
  const char* why = "";
  int rc = eva::rfa::make_plan(g, &pl, &why);
  if (rc != EVA_OK) return eva::abi_fail(rc, why);
  eva::View vq, vk, vv;
  if ((rc = eva::abi_view(q, "q", &vq)) || (rc = eva::abi_view(k, "k", &vk)) || (rc = eva::abi_view(v, "v", &vv))) return rc;
  eva::rfa::Params& p = pl.p;
  if (p.m > 0 && !proj) return eva::abi_fail(EVA_ERR_INVALID, "this feature method needs the projection matrix");
  if (p.method == RFA_GIVEN && (!q_feat || !k_feat)) return eva::abi_fail(EVA_ERR_INVALID, "RFA_GIVEN needs q_feat and k_feat");
  if (!out || !workspace) return eva::abi_fail(EVA_ERR_INVALID, "out / workspace is NULL");
  if (workspace_bytes < pl.ws_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return eva::abi_fail(EVA_ERR_INVALID, "workspace too small or not 256-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (eva::rfa_tc_supported(p.method, p.D, p.m, p.cosw, g->io_dtype, vq, vk, vv)) {      // tcgen05 path (rfa_tc_sm100.cu)
    const cudaError_t e2 = eva::launch_rfa_tc(p.B, p.H, p.N, g->io_dtype, vq, vk, vv, pad_mask, proj, out, st);
    return e2 == cudaSuccess ? EVA_OK : eva::abi_cuda_fail(e2, "rfa_forward (tcgen05)");
  }
  p.proj = proj; p.qf = q_feat; p.kf = k_feat; p.mask = pad_mask;
  p.stab = reinterpret_cast<float*>(workspace);
  p.part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + ((((size_t)p.B * p.H * 8) + 255) & ~(size_t)255));
  cudaError_t e;
  if (g->io_dtype == EVA_F32) e = eva::rfa::run<float>(pl, vq, vk, vv, out, st);
  else if (g->io_dtype == EVA_F16) e = eva::rfa::run<__half>(pl, vq, vk, vv, out, st);
  else e = eva::rfa::run<__nv_bfloat16>(pl, vq, vk, vv, out, st);
  if (e != cudaSuccess) return eva::abi_cuda_fail(e, "rfa_forward");
  return EVA_OK;
}

/* ---- ScatterBrain ---- */
static int sb_plan(const SbGeometry* g, eva::rfa::Plan* pl, eva::rfa::SbParams* sp, size_t* smem_win, int* windows) {
  if (!g) return eva::abi_fail(EVA_ERR_INVALID, "geometry is NULL");
  RfaGeometry rg{};
  rg.batch = g->batch; rg.heads = g->heads; rg.tokens = g->tokens; rg.head_dim = g->head_dim; rg.method = RFA_FAVORP;
  rg.proj_dim = g->proj_dim; rg.io_dtype = g->io_dtype;
  const char* why = "";
  const int rc = eva::rfa::make_plan(&rg, pl, &why);
  if (rc != EVA_OK) return eva::abi_fail(rc, why);
  pl->p.method = eva::rfa::RFA_LOG_FAVORP;
  if (g->proj_dim > 256) return eva::abi_fail(EVA_ERR_UNSUPPORTED, "at most 256 random features");
  if (g->dims != 1 && g->dims != 2) return eva::abi_fail(EVA_ERR_INVALID, "dims must be 1 or 2");
  if (g->window <= 0) return eva::abi_fail(EVA_ERR_INVALID, "window must be positive");
  sp->dims = g->dims; sp->w = g->window;
  if (g->dims == 2) {
    if (g->grid_h * g->grid_w != g->tokens || g->grid_h % g->window || g->grid_w % g->window)
      return eva::abi_fail(EVA_ERR_INVALID, "grid does not cover the tokens or is not divisible by the window");
    sp->gh = g->grid_h; sp->gw = g->grid_w; sp->L = g->window * g->window;
    *windows = (g->grid_h / g->window) * (g->grid_w / g->window);
  } else {
    if (g->tokens % g->window) return eva::abi_fail(EVA_ERR_INVALID, "tokens not a multiple of the window (pad first)");
    sp->gh = 1; sp->gw = g->tokens; sp->L = g->window;
    *windows = g->tokens / g->window;
  }
  if (sp->L > 64) return eva::abi_fail(EVA_ERR_UNSUPPORTED, "windows of more than 64 tokens are not built");
  if (*windows > 65535) return eva::abi_fail(EVA_ERR_UNSUPPORTED, "more than 65535 windows");
  const size_t D = g->head_dim, m = g->proj_dim, XS = D + 4, MS = m + 1, LP = 64;
  *smem_win = (m * XS + 3 * LP * XS + 2 * LP * MS + m * D + m + 2 * LP + 8 * (LP + m)) * 4 + LP + 16;
  if (*smem_win > 220 * 1024) return eva::abi_fail(EVA_ERR_UNSUPPORTED, "feature / head sizes exceed shared memory");
  pl->ws_bytes = 256 + (((size_t)g->batch * g->heads * m * 4 + 255) & ~(size_t)255) +
                 (size_t)g->batch * g->heads * pl->p.S * (m * D + m) * 4;
  pl->ws_bytes = (pl->ws_bytes + 255) & ~(size_t)255;
  return EVA_OK;
}

int scatterbrain_forward_workspace_bytes(const SbGeometry* g, size_t* bytes) {
  eva::rfa::Plan pl;
  eva::rfa::SbParams sp{};
  size_t smem = 0;
  int windows = 0;
  const int rc = sb_plan(g, &pl, &sp, &smem, &windows);
  if (rc != EVA_OK) return rc;
  if (!bytes) return eva::abi_fail(EVA_ERR_INVALID, "bytes is NULL");
  *bytes = pl.ws_bytes;
  return EVA_OK;
}

int scatterbrain_forward(const SbGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v, const uint8_t* pad_mask,
                         const float* proj, const float* bias, void* out, void* workspace, size_t workspace_bytes, void* stream) {
  eva::rfa::Plan pl;
  eva::rfa::SbParams sp{};
  size_t smem = 0;
  int windows = 0;
  int rc = sb_plan(g, &pl, &sp, &smem, &windows);
  if (rc != EVA_OK) return rc;
  eva::View vq, vk, vv;
  if ((rc = eva::abi_view(q, "q", &vq)) || (rc = eva::abi_view(k, "k", &vk)) || (rc = eva::abi_view(v, "v", &vv))) return rc;
  if (!proj || !out || !workspace) return eva::abi_fail(EVA_ERR_INVALID, "proj / out / workspace is NULL");
  if (workspace_bytes < pl.ws_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return eva::abi_fail(EVA_ERR_INVALID, "workspace too small or not 256-byte aligned");
  eva::rfa::Params& p = pl.p;
  p.proj = proj; p.mask = pad_mask;
  p.stabv = reinterpret_cast<float*>(workspace);
  p.stab = p.stabv;
  p.part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + (((size_t)p.B * p.H * p.m * 4 + 255) & ~(size_t)255));
  const bool tc = eva::rfa_tc_supported(RFA_FAVORP, p.D, p.m, 0, g->io_dtype, vk, vk, vv);
  if (tc) p.S = 1;                      // the tcgen05 statistics kernel writes one partial per item
  sp.r = p; sp.bias = bias;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (g->io_dtype == EVA_F32) e = eva::rfa::run_sb<float>(pl, sp, smem, windows, vq, vk, vv, out, st, false, g->io_dtype);
  else if (g->io_dtype == EVA_F16) e = eva::rfa::run_sb<__half>(pl, sp, smem, windows, vq, vk, vv, out, st, tc, g->io_dtype);
  else e = eva::rfa::run_sb<__nv_bfloat16>(pl, sp, smem, windows, vq, vk, vv, out, st, tc, g->io_dtype);
  if (e != cudaSuccess) return eva::abi_cuda_fail(e, "scatterbrain_forward");
  return EVA_OK;
}

/* ---- randomized attention ---- */
static bool ra_uses_tensor_cores(const RaGeometry* g) {
  if (g->head_dim != 64 || (g->io_dtype != EVA_F16 && g->io_dtype != EVA_BF16)) return false;
  eva::Geo geo{};
  geo.D = 64; geo.J = g->tokens;
  return eva::window_tc_supported(geo, g->io_dtype);
}

int ra_sample(const RaGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, uint64_t seed, const float* gumbel, int64_t* k_ind,
              void* stream) {
  if (!g) return eva::abi_fail(EVA_ERR_INVALID, "geometry is NULL");
  if (g->batch <= 0 || g->heads <= 0 || g->tokens <= 0) return eva::abi_fail(EVA_ERR_INVALID, "batch / heads / tokens must be positive");
  eva::View vq, vk;
  int rc;
  if ((rc = eva::abi_view(q, "q", &vq)) || (rc = eva::abi_view(k, "k", &vk))) return rc;
  if (!k_ind) return eva::abi_fail(EVA_ERR_INVALID, "k_ind is NULL");
  if (!eva::ra_sample_tc_supported(g->head_dim, g->io_dtype, vq, vk))
    return eva::abi_fail(EVA_ERR_UNSUPPORTED, "ra_sample: head_dim 64 with 16-bit q / k only (draw with library ops otherwise)");
  const cudaError_t e = eva::launch_ra_sample_tc(g->batch, g->heads, g->tokens, g->io_dtype, vq, vk, seed, gumbel,
                                                 reinterpret_cast<long long*>(k_ind), reinterpret_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? EVA_OK : eva::abi_cuda_fail(e, "ra_sample");
}

int ra_forward_workspace_bytes(const RaGeometry* g, size_t* bytes) {
  if (!g || !bytes) return eva::abi_fail(EVA_ERR_INVALID, "geometry / bytes is NULL");
  size_t n = ((size_t)g->batch * g->heads * g->head_dim * 4 + 255) & ~(size_t)255;                 /* mean of k */
  if (ra_uses_tensor_cores(g))
    n += (((size_t)g->batch * g->tokens * g->heads * g->head_dim * 2 + 255) & ~(size_t)255) +       /* w, 16-bit */
         (((size_t)g->batch * g->heads * g->tokens * 4 + 255) & ~(size_t)255);                      /* per-key addend */
  *bytes = n;
  return EVA_OK;
}

int ra_forward(const RaGeometry* g, const EvaHeadsView* q, const EvaHeadsView* k, const EvaHeadsView* v, const void* extra,
               const int64_t* k_ind, const float* noise, void* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!g) return eva::abi_fail(EVA_ERR_INVALID, "geometry is NULL");
  if (g->batch <= 0 || g->heads <= 0 || g->tokens <= 0) return eva::abi_fail(EVA_ERR_INVALID, "batch / heads / tokens must be positive");
  if (g->head_dim % 8 || g->head_dim < 8 || g->head_dim > 128) return eva::abi_fail(EVA_ERR_UNSUPPORTED, "head_dim must be a multiple of 8, at most 128");
  if (g->io_dtype < EVA_F32 || g->io_dtype > EVA_BF16) return eva::abi_fail(EVA_ERR_INVALID, "unknown io_dtype");
  if (g->mode < 0 || g->mode > 2) return eva::abi_fail(EVA_ERR_INVALID, "mode must be 0 (mean), 1 (given rows) or 2 (gather)");
  if ((long long)((g->tokens + 31) / 32) > 65535) return eva::abi_fail(EVA_ERR_UNSUPPORTED, "sequence too long");
  eva::View vq, vk, vv;
  int rc;
  if ((rc = eva::abi_view(q, "q", &vq)) || (rc = eva::abi_view(k, "k", &vk)) || (rc = eva::abi_view(v, "v", &vv))) return rc;
  if (!out) return eva::abi_fail(EVA_ERR_INVALID, "out is NULL");
  if (g->mode == 1 && !extra) return eva::abi_fail(EVA_ERR_INVALID, "mode 1 needs the extra rows");
  if (g->mode == 2 && !k_ind) return eva::abi_fail(EVA_ERR_INVALID, "mode 2 needs k_ind");
  size_t need = 0;
  ra_forward_workspace_bytes(g, &need);
  if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255u))
    return eva::abi_fail(EVA_ERR_INVALID, "workspace too small (ra_forward_workspace_bytes) or not 256-byte aligned");
  eva::rfa::RaParams p{};
  p.B = g->batch; p.H = g->heads; p.N = g->tokens; p.D = g->head_dim; p.mode = g->mode;
  p.extra = extra; p.k_ind = reinterpret_cast<const long long*>(k_ind); p.noise = noise; p.kmean = reinterpret_cast<float*>(workspace);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (ra_uses_tensor_cores(g)) {      // prepare w, kb -> dense tcgen05 window kernel with a per-key addend (eva_window_tc_sm100.cu)
    uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
    const size_t off_w = ((size_t)g->batch * g->heads * g->head_dim * 4 + 255) & ~(size_t)255;
    const size_t off_kb = off_w + (((size_t)g->batch * g->tokens * g->heads * g->head_dim * 2 + 255) & ~(size_t)255);
    if (g->io_dtype == EVA_F16) e = eva::rfa::run_ra_tc<__half>(p, g->io_dtype, vq, vk, vv, out, base + off_w, reinterpret_cast<float*>(base + off_kb), st);
    else e = eva::rfa::run_ra_tc<__nv_bfloat16>(p, g->io_dtype, vq, vk, vv, out, base + off_w, reinterpret_cast<float*>(base + off_kb), st);
    if (e != cudaSuccess) return eva::abi_cuda_fail(e, "ra_forward (tcgen05)");
    return EVA_OK;
  }
  if (g->io_dtype == EVA_F32) e = eva::rfa::run_ra<float>(p, vq, vk, vv, out, st);
  else if (g->io_dtype == EVA_F16) e = eva::rfa::run_ra<__half>(p, vq, vk, vv, out, st);
  else e = eva::rfa::run_ra<__nv_bfloat16>(p, vq, vk, vv, out, st);
  if (e != cudaSuccess) return eva::abi_cuda_fail(e, "ra_forward");
  return EVA_OK;
}

}  // extern "C"
