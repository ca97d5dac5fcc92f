// Shared device helpers for libeva_sm100 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eva_sm100.h"

namespace eva {

constexpr float kMaskVal = -5.0e4f;  // eva.py:139
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kNegInf = -INFINITY;

// ---- io-type helpers: everything is computed in fp32, T is only the HBM format ----------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive elements (16-byte aligned for 2-byte types, 32-byte span for float)
template <typename T> __device__ __forceinline__ void load8(const T* p, float* o);
template <> __device__ __forceinline__ void load8<float>(const float* p, float* o) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__half>(const __half* p, float* o) {
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* o) {
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// exp(x) for x <= 0 (softmax terms); exp(-inf) = 0
__device__ __forceinline__ float exp_nonpos(float x) { return exp2f(x * kLog2e); }

// ---- warp-per-row helpers: lane l owns features l, l+32, ... of a D-vector (D = 16, 32, 64, 128) ----
template <int D> struct Feat {
  static constexpr int kPerLane = (D + 31) / 32;
  static __device__ __forceinline__ bool has(int lane, int i) { return (D % 32 == 0) || (lane + 32 * i < D); }
};

// y = W x + b with W^T staged in shared memory as Wt[in][out]
template <int D>
__device__ __forceinline__ void warp_linear(const float* __restrict__ Wt, const float* __restrict__ bias,
                                            const float (&x)[Feat<D>::kPerLane], float (&y)[Feat<D>::kPerLane], int lane) {
  constexpr int DPL = Feat<D>::kPerLane;
#pragma unroll
  for (int i = 0; i < DPL; ++i) y[i] = (bias && Feat<D>::has(lane, i)) ? __ldg(bias + lane + 32 * i) : 0.f;
#pragma unroll
  for (int ii = 0; ii < DPL; ++ii) {
#pragma unroll 8
    for (int jj = 0; jj < (D < 32 ? D : 32); ++jj) {
      const float m = __shfl_sync(0xffffffffu, x[ii], jj);
      const float* wrow = Wt + (jj + 32 * ii) * D + lane;
#pragma unroll
      for (int i = 0; i < DPL; ++i)
        if (Feat<D>::has(lane, i)) y[i] = fmaf(wrow[32 * i], m, y[i]);
    }
  }
}

template <int D>
__device__ __forceinline__ void warp_layer_norm(float (&y)[Feat<D>::kPerLane], const float* __restrict__ gain,
                                                const float* __restrict__ bias, float eps, int lane) {
  constexpr int DPL = Feat<D>::kPerLane;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < DPL; ++i) s += Feat<D>::has(lane, i) ? y[i] : 0.f;
  const float mean = warp_sum(s) * (1.0f / D);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < DPL; ++i) { const float c = Feat<D>::has(lane, i) ? y[i] - mean : 0.f; v = fmaf(c, c, v); }
  const float inv = 1.0f / sqrtf(warp_sum(v) * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < DPL; ++i)
    if (Feat<D>::has(lane, i)) y[i] = (y[i] - mean) * inv * __ldg(gain + lane + 32 * i) + __ldg(bias + lane + 32 * i);
}

// ---- strided q/k/v view ------------------------------------------------------------------------
struct View {
  const void* ptr;
  long long sb, sn, sh;
  template <typename T> __device__ __forceinline__ const T* row(int b, int n, int h) const {
    return reinterpret_cast<const T*>(ptr) + (long long)b * sb + (long long)n * sn + (long long)h * sh;
  }
};

// ---- window / chunk geometry -------------------------------------------------------------------
struct Geo {
  int B, H, N, D;
  int dims, gh, gw;
  int window, ext, left_only;
  int chunk, chunk_ext;
  int causal, mask_queries;
  int n_windows, L, J;    // queries per window, local keys per window
  int n_chunks, Jc;       // chunk keys, tokens per chunk (incl. halo)
  float mask_fill;        // -5e4, or -inf for the dense softmax baseline
  int bias_toeplitz;      // caller's hint: bias[i][j] = f(i - j)
};

// token id of slot `slot` of group `grp` (edge `size`, halo `ext`), -1 when off the sequence.
// 2-D: groups row-major over the grid, slots row-major inside the (size+2ext)^2 box
// (attn_utils.py:172-210).  1-D: slot 0 is `ext` tokens left of the group start (attn_utils.py:155-166).
__device__ __forceinline__ int group_token(const Geo& g, int grp, int slot, int size, int ext) {
  if (g.dims == 2) {
    const int t = size + 2 * ext;
    const int ngx = g.gw / size;
    const int y = (grp / ngx) * size - ext + slot / t;
    const int x = (grp % ngx) * size - ext + slot % t;
    return (y >= 0 && y < g.gh && x >= 0 && x < g.gw) ? y * g.gw + x : -1;
  }
  const int p = grp * size - ext + slot;
  return (p >= 0 && p < g.N) ? p : -1;
}

// ---- head_dim 64 vectors held as feature pairs (2 lane, 2 lane + 1): 4-byte loads of 16-bit rows ----------------------
template <typename T> struct Pair16;
template <> struct Pair16<__half> {
  static __device__ __forceinline__ float2 up(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
  static __device__ __forceinline__ uint32_t pk(float a, float b) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
};
template <> struct Pair16<__nv_bfloat16> {
  static __device__ __forceinline__ float2 up(uint32_t v) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v)); }
  static __device__ __forceinline__ uint32_t pk(float a, float b) { const __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
};

// y = W x + b, x / y distributed as (2 lane, 2 lane + 1); Wt[in][out] in shared memory
__device__ __forceinline__ float2 pair_linear(const float* __restrict__ Wt, const float* __restrict__ bias, float2 x, int lane) {
  float2 y = bias ? __ldg(reinterpret_cast<const float2*>(bias) + lane) : make_float2(0.f, 0.f);
#pragma unroll 8
  for (int j = 0; j < 32; ++j) {
    const float x0 = __shfl_sync(0xffffffffu, x.x, j), x1 = __shfl_sync(0xffffffffu, x.y, j);
    const float2 w0 = *reinterpret_cast<const float2*>(Wt + (2 * j) * 64 + 2 * lane);
    const float2 w1 = *reinterpret_cast<const float2*>(Wt + (2 * j + 1) * 64 + 2 * lane);
    y.x = fmaf(w0.x, x0, fmaf(w1.x, x1, y.x));
    y.y = fmaf(w0.y, x0, fmaf(w1.y, x1, y.y));
  }
  return y;
}
// dx = W^T dy, W row-major [out][in] in global memory (L1)
__device__ __forceinline__ float2 pair_linear_bwd(const float* __restrict__ W, float2 dy, int lane) {
  float2 dx = make_float2(0.f, 0.f);
#pragma unroll 8
  for (int j = 0; j < 32; ++j) {
    const float d0 = __shfl_sync(0xffffffffu, dy.x, j), d1 = __shfl_sync(0xffffffffu, dy.y, j);
    const float2 w0 = __ldg(reinterpret_cast<const float2*>(W + (2 * j) * 64) + lane);
    const float2 w1 = __ldg(reinterpret_cast<const float2*>(W + (2 * j + 1) * 64) + lane);
    dx.x = fmaf(w0.x, d0, fmaf(w1.x, d1, dx.x));
    dx.y = fmaf(w0.y, d0, fmaf(w1.y, d1, dx.y));
  }
  return dx;
}
// LayerNorm over 64 features held as pairs: normalised row and 1 / sigma
__device__ __forceinline__ float2 pair_ln(float2 y, float eps, float& inv) {
  const float mean = warp_sum(y.x + y.y) * (1.0f / 64);
  const float cx = y.x - mean, cy = y.y - mean;
  inv = 1.0f / sqrtf(warp_sum(cx * cx + cy * cy) * (1.0f / 64) + eps);
  return make_float2(cx * inv, cy * inv);
}
__device__ __forceinline__ float2 pair_ln_bwd(float2 dout, float2 n, float inv, const float* __restrict__ gain, int lane) {
  if (!gain) return dout;
  const float2 gg = __ldg(reinterpret_cast<const float2*>(gain) + lane);
  const float dx = dout.x * gg.x, dy = dout.y * gg.y;
  const float s1 = warp_sum(dx + dy) * (1.0f / 64);
  const float s2 = warp_sum(dx * n.x + dy * n.y) * (1.0f / 64);
  return make_float2(inv * (dx - s1 - n.x * s2), inv * (dy - s1 - n.y * s2));
}

// 8 consecutive features per lane (piece p8 = lane & 7 of a 128-byte row), partial over the tokens of sub-index ts = lane >> 3:
// sum over the four ts groups, then hand lane l its feature pair (2l, 2l + 1)
__device__ __forceinline__ float2 pieces_to_pair(float (&a)[8], int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] += __shfl_xor_sync(0xffffffffu, a[i], 8);
    a[i] += __shfl_xor_sync(0xffffffffu, a[i], 16);
  }
  float2 r = make_float2(0.f, 0.f);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float x = __shfl_sync(0xffffffffu, a[2 * u], lane >> 2), y = __shfl_sync(0xffffffffu, a[2 * u + 1], lane >> 2);
    if ((lane & 3) == u) r = make_float2(x, y);
  }
  return r;
}
template <typename T>
__device__ __forceinline__ void add8(const uint4& raw, float w, float (&a)[8]) {
  const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float2 f = Pair16<T>::up(w4[u]);
    a[2 * u] = fmaf(w, f.x, a[2 * u]); a[2 * u + 1] = fmaf(w, f.y, a[2 * u + 1]);
  }
}

// logit of (query row li / token tq, key gj) after bias and masks, exactly as window_attn_kernel (eva_generic.cu); returns whether it still depends
// on q . k (false: the forward overwrote it with a constant, so no gradient flows to q, k or the bias)
__device__ __forceinline__ bool finish_logit(const Geo& g, float& sv, int flag, int gj, int li, int tq, int qpad,
                                             const float* __restrict__ bias, long long bias_off) {
  if (flag == 2) { sv = kNegInf; return false; }
  if (gj < g.J) {
    bool live = true;
    if (bias) sv += __ldg(bias + bias_off + (long long)li * g.J + gj);
    if (flag == 1 || (g.mask_queries && qpad)) { sv = g.mask_fill; live = false; }
    if (g.causal && gj > li + g.ext) { sv = kMaskVal; live = false; }
    return live;
  }
  if (g.causal && (gj - g.J) >= tq / g.chunk) { sv = kMaskVal; return false; }
  return true;
}

}  // namespace eva
