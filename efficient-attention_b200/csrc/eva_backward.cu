// Backward of the EVA attention core for every geometry the forward kernels accept: 1-D / 2-D, halos, padding
// masks, causal, chunk-less local / dense attention, head_dim 16 .. 128.  float32 maths on CUDA cores (gradients are accumulated
// in float32 whatever the I/O format); T is only the HBM format of q, k, v, out and grad_out.
//
//   window_attn_bwd_kernel    gradient of eva.py:200-227 / causal_eva.py:722-783 / local_attention.py:134-182: the joint softmax over
//                             [local keys | chunk keys] is recomputed tile by tile (no probabilities are stored by the forward),
//                             dS = P o (dP - rowsum(dO o O)); writes dq, accumulates dk / dv / d k_bar / d beta / d bias
//   chunk_stats_bwd_kernel    gradient of eva.py:155-196 / causal_eva.py:676-719: chunk softmax -> omega -> LayerNorm -> Linear ->
//                             chunk means, back to dq / dk / dv; leaves per-chunk rows from which the caller forms the PARAMETER
//                             gradients with library reductions (dW = dy^T mean, ...)
#include <stdlib.h>

#include "common.cuh"
#include "launch.h"

namespace eva {

constexpr int kBR = 64;    // query rows per CTA, 8 per warp
constexpr int kBK = 64;    // keys per tile: 2 per lane in the logit phase, 8 per warp in the dK / dV phase
constexpr int kBRw = 8;
constexpr int kStS = kBR + 4;  // row stride of the transposed dS tile (16-byte aligned rows)

template <int D> struct BwdSmem {
  static constexpr int DP = D + 1;
  static constexpr int kQ = 0;
  static constexpr int kG = kQ + kBR * D;                      // grad_out rows
  static constexpr int kK = kG + kBR * D;
  static constexpr int kV = kK + kBK * DP;                     // v tile, later the transposed dS tile
  static constexpr int kVsz = kBK * DP > kBK * kStS ? kBK * DP : kBK * kStS;
  static constexpr int kP = kV + kVsz;
  static constexpr int kS = kP + kBR * kBK;
  static constexpr int kFloats = kS + kBR * kBK;
  static constexpr size_t kBytes = (size_t)kFloats * sizeof(float) + (size_t)(kBK + 2 * kBR) * sizeof(int);
};

template <typename T, int D>
__global__ void __launch_bounds__(256, (D <= 64 ? 2 : 1))
window_attn_bwd_kernel(const Geo g, const View q, const View k, const View v, const uint8_t* __restrict__ mask,
                       const float* __restrict__ kbar, const float* __restrict__ beta, const float* __restrict__ bias,
                       const long long bias_sh, const T* __restrict__ out, const T* __restrict__ dout, float* __restrict__ dq,
                       float* __restrict__ dk, float* __restrict__ dv, float* __restrict__ dkbar, float* __restrict__ dbeta,
                       float* __restrict__ dbias) {
  using L = BwdSmem<D>;
  constexpr int DPL = Feat<D>::kPerLane;
  constexpr int DP = L::DP;
  extern __shared__ float sm[];
  float* Qs = sm + L::kQ;       // [kBR][D] pre-scaled by d^-1/2
  float* Gs = sm + L::kG;       // [kBR][D] grad_out
  float* Ks = sm + L::kK;       // [kBK][DP]
  float* Vs = sm + L::kV;       // [kBK][DP]
  float* St = sm + L::kV;       // [kBK][kStS] dS, key-major (overwrites the v tile once dP is done)
  float* Ps = sm + L::kP;       // [kBR][kBK]
  float* Ss = sm + L::kS;       // [kBR][kBK] dS, row-major
  int* kflag = reinterpret_cast<int*>(sm + L::kFloats);   // [kBK] 0 live, 1 masked, 2 absent
  int* qtok = kflag + kBK;                                // [kBR]
  int* qpad = qtok + kBR;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_rb = (g.L + kBR - 1) / kBR;
  const long long cta = blockIdx.x;
  const int rb = (int)(cta % n_rb), win = (int)((cta / n_rb) % g.n_windows);
  const int bh = (int)(cta / ((long long)n_rb * g.n_windows));
  const int b = bh / g.H, h = bh % g.H;
  const float scale = rsqrtf((float)D);
  const int n_keys = g.J + g.n_chunks;
  const long long HD = (long long)g.H * D;
  const long long bias_off = (long long)h * bias_sh;

  if (tid < kBR) {
    const int li = rb * kBR + tid;
    const int tok = li < g.L ? group_token(g, win, li, g.window, 0) : -1;
    qtok[tid] = tok;
    qpad[tid] = (tok >= 0 && mask) ? (int)mask[(long long)b * g.N + tok] : 0;
  }
  for (int idx = tid; idx < kBR * (D / 8); idx += blockDim.x) {
    const int r = idx / (D / 8), part = idx % (D / 8);
    const int li = rb * kBR + r;
    const int tok = li < g.L ? group_token(g, win, li, g.window, 0) : -1;
    float f[8], gg[8];
    if (tok >= 0) {
      load8<T>(q.row<T>(b, tok, h) + part * 8, f);
      load8<T>(dout + ((long long)b * g.N + tok) * HD + (long long)h * D + part * 8, gg);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      Qs[r * D + part * 8 + i] = tok >= 0 ? f[i] * scale : 0.f;
      Gs[r * D + part * 8 + i] = tok >= 0 ? gg[i] : 0.f;
    }
  }
  __syncthreads();

  // delta_r = <grad_out_r, out_r>: the softmax Jacobian's rank-one term
  float delta[kBRw];
#pragma unroll
  for (int r = 0; r < kBRw; ++r) {
    const int row = warp * kBRw + r;
    const int tq = qtok[row];
    float part = 0.f;
    if (tq >= 0) {
      const T* orow = out + ((long long)b * g.N + tq) * HD + (long long)h * D;
#pragma unroll
      for (int i = 0; i < DPL; ++i)
        if (Feat<D>::has(lane, i)) part = fmaf(to_f32(orow[lane + 32 * i]), Gs[row * D + lane + 32 * i], part);
    }
    delta[r] = warp_sum(part);
  }

  auto load_tile = [&](int kt0, bool with_v) {
    for (int idx = tid; idx < kBK * (D / 8); idx += blockDim.x) {
      const int j = idx / (D / 8), part = idx % (D / 8);
      const int gj = kt0 + j;
      float fk[8], fv[8];
      int flag = 0;
      bool have = false;
      if (gj < g.J) {
        const int tok = group_token(g, win, gj, g.window, g.ext);
        if (tok >= 0) {
          load8<T>(k.row<T>(b, tok, h) + part * 8, fk);
          if (with_v) load8<T>(v.row<T>(b, tok, h) + part * 8, fv);
          have = true;
          flag = (mask && mask[(long long)b * g.N + tok]) ? 1 : 0;
        } else {
          flag = 1;
        }
      } else if (gj < n_keys) {
        const long long base = (((long long)b * g.H + h) * g.n_chunks + (gj - g.J)) * D + part * 8;
        load8<float>(kbar + base, fk);
        if (with_v) load8<float>(beta + base, fv);
        have = true;
      } else {
        flag = 2;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        Ks[j * DP + part * 8 + i] = have ? fk[i] : 0.f;
        if (with_v) Vs[j * DP + part * 8 + i] = have ? fv[i] : 0.f;
      }
      if (part == 0) kflag[j] = flag;
    }
  };

  // ---- pass 0: row maxima and normalisers of the joint softmax ----
  float m[kBRw], linv[kBRw];
#pragma unroll
  for (int r = 0; r < kBRw; ++r) { m[r] = kNegInf; linv[r] = 0.f; }
  // causal: a tile of local keys that lies entirely above the diagonal of this row block holds only masked logits (P = dS = 0)
  const int last_visible = rb * kBR + kBR - 1 + g.ext;
  for (int kt0 = 0; kt0 < n_keys; kt0 += kBK) {
    if (g.causal && kt0 > last_visible && kt0 + kBK <= g.J) continue;
    __syncthreads();
    load_tile(kt0, false);
    __syncthreads();
    float s[kBRw][2];
#pragma unroll
    for (int r = 0; r < kBRw; ++r) s[r][0] = s[r][1] = 0.f;
    const float* k0 = Ks + lane * DP;
    const float* k1 = Ks + (lane + 32) * DP;
#pragma unroll 2
    for (int e = 0; e < D; e += 4) {
      const float a0 = k0[e], a1 = k0[e + 1], a2 = k0[e + 2], a3 = k0[e + 3];
      const float c0 = k1[e], c1 = k1[e + 1], c2 = k1[e + 2], c3 = k1[e + 3];
#pragma unroll
      for (int r = 0; r < kBRw; ++r) {
        const float4 qv = *reinterpret_cast<const float4*>(Qs + (warp * kBRw + r) * D + e);
        s[r][0] = fmaf(qv.x, a0, fmaf(qv.y, a1, fmaf(qv.z, a2, fmaf(qv.w, a3, s[r][0]))));
        s[r][1] = fmaf(qv.x, c0, fmaf(qv.y, c1, fmaf(qv.z, c2, fmaf(qv.w, c3, s[r][1]))));
      }
    }
    const int f0 = kflag[lane], f1 = kflag[lane + 32];
#pragma unroll
    for (int r = 0; r < kBRw; ++r) {
      const int row = warp * kBRw + r;
      const int tq = qtok[row];
      if (tq < 0) continue;
      const int li = rb * kBR + row;
      float s0 = s[r][0], s1 = s[r][1];
      finish_logit(g, s0, f0, kt0 + lane, li, tq, qpad[row], bias, bias_off);
      finish_logit(g, s1, f1, kt0 + lane + 32, li, tq, qpad[row], bias, bias_off);
      const float mn = fmaxf(m[r], warp_max(fmaxf(s0, s1)));
      float corr = 1.f, p = 0.f;
      if (mn != kNegInf) { corr = exp_nonpos(m[r] - mn); p = exp_nonpos(s0 - mn) + exp_nonpos(s1 - mn); }
      linv[r] = fmaf(linv[r], corr, warp_sum(p));
      m[r] = mn;
    }
  }
#pragma unroll
  for (int r = 0; r < kBRw; ++r) linv[r] = linv[r] > 0.f ? 1.0f / linv[r] : 0.f;

  // ---- pass 1: per key tile dS, then dQ (registers), dK / dV (atomics) ----
  float dqa[kBRw][DPL];
#pragma unroll
  for (int r = 0; r < kBRw; ++r)
#pragma unroll
    for (int i = 0; i < DPL; ++i) dqa[r][i] = 0.f;

  for (int kt0 = 0; kt0 < n_keys; kt0 += kBK) {
    if (g.causal && kt0 > last_visible && kt0 + kBK <= g.J) continue;
    __syncthreads();
    load_tile(kt0, true);
    __syncthreads();
    float s[kBRw][2], dp[kBRw][2];
#pragma unroll
    for (int r = 0; r < kBRw; ++r) s[r][0] = s[r][1] = dp[r][0] = dp[r][1] = 0.f;
    {
      const float* k0 = Ks + lane * DP;
      const float* k1 = Ks + (lane + 32) * DP;
      const float* v0 = Vs + lane * DP;
      const float* v1 = Vs + (lane + 32) * DP;
#pragma unroll 1
      for (int e = 0; e < D; e += 4) {
        const float a0 = k0[e], a1 = k0[e + 1], a2 = k0[e + 2], a3 = k0[e + 3];
        const float c0 = k1[e], c1 = k1[e + 1], c2 = k1[e + 2], c3 = k1[e + 3];
        const float x0 = v0[e], x1 = v0[e + 1], x2 = v0[e + 2], x3 = v0[e + 3];
        const float y0 = v1[e], y1 = v1[e + 1], y2 = v1[e + 2], y3 = v1[e + 3];
#pragma unroll
        for (int r = 0; r < kBRw; ++r) {
          const float4 qv = *reinterpret_cast<const float4*>(Qs + (warp * kBRw + r) * D + e);
          const float4 gv = *reinterpret_cast<const float4*>(Gs + (warp * kBRw + r) * D + e);
          s[r][0] = fmaf(qv.x, a0, fmaf(qv.y, a1, fmaf(qv.z, a2, fmaf(qv.w, a3, s[r][0]))));
          s[r][1] = fmaf(qv.x, c0, fmaf(qv.y, c1, fmaf(qv.z, c2, fmaf(qv.w, c3, s[r][1]))));
          dp[r][0] = fmaf(gv.x, x0, fmaf(gv.y, x1, fmaf(gv.z, x2, fmaf(gv.w, x3, dp[r][0]))));
          dp[r][1] = fmaf(gv.x, y0, fmaf(gv.y, y1, fmaf(gv.z, y2, fmaf(gv.w, y3, dp[r][1]))));
        }
      }
    }
    const int f0 = kflag[lane], f1 = kflag[lane + 32];
#pragma unroll
    for (int r = 0; r < kBRw; ++r) {
      const int row = warp * kBRw + r;
      const int tq = qtok[row];
      const int li = rb * kBR + row;
      float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
      if (tq >= 0 && m[r] != kNegInf) {
        float s0 = s[r][0], s1 = s[r][1];
        const bool l0 = finish_logit(g, s0, f0, kt0 + lane, li, tq, qpad[row], bias, bias_off);
        const bool l1 = finish_logit(g, s1, f1, kt0 + lane + 32, li, tq, qpad[row], bias, bias_off);
        p0 = exp_nonpos(s0 - m[r]) * linv[r];
        p1 = exp_nonpos(s1 - m[r]) * linv[r];
        d0 = l0 ? p0 * (dp[r][0] - delta[r]) : 0.f;
        d1 = l1 ? p1 * (dp[r][1] - delta[r]) : 0.f;
        if (dbias) {
          if (l0 && kt0 + lane < g.J) atomicAdd(dbias + bias_off + (long long)li * g.J + kt0 + lane, d0);
          if (l1 && kt0 + lane + 32 < g.J) atomicAdd(dbias + bias_off + (long long)li * g.J + kt0 + lane + 32, d1);
        }
      }
      s[r][0] = p0; s[r][1] = p1; dp[r][0] = d0; dp[r][1] = d1;
    }
    __syncthreads();   // every warp is done with the v tile: the transposed dS tile may take its place
#pragma unroll
    for (int r = 0; r < kBRw; ++r) {
      const int row = warp * kBRw + r;
      Ps[row * kBK + lane] = s[r][0];           Ps[row * kBK + lane + 32] = s[r][1];
      Ss[row * kBK + lane] = dp[r][0];          Ss[row * kBK + lane + 32] = dp[r][1];
      St[lane * kStS + row] = dp[r][0];         St[(lane + 32) * kStS + row] = dp[r][1];
    }
    __syncthreads();
    // dQ rows of this warp: sum_j dS[r][j] K[j][:]
#pragma unroll 2
    for (int j = 0; j < kBK; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(St + j * kStS + warp * kBRw);
      const float4 c = *reinterpret_cast<const float4*>(St + j * kStS + warp * kBRw + 4);
      const float ds[kBRw] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
      float kk[DPL];
#pragma unroll
      for (int i = 0; i < DPL; ++i) kk[i] = Feat<D>::has(lane, i) ? Ks[j * DP + lane + 32 * i] : 0.f;
#pragma unroll
      for (int r = 0; r < kBRw; ++r)
#pragma unroll
        for (int i = 0; i < DPL; ++i) dqa[r][i] = fmaf(ds[r], kk[i], dqa[r][i]);
    }
    // dK / dV of keys 8 warp .. 8 warp + 7 of the tile: sum over the CTA's rows
    float dka[8][DPL], dva[8][DPL];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
#pragma unroll
      for (int i = 0; i < DPL; ++i) dka[jj][i] = dva[jj][i] = 0.f;
#pragma unroll 2
    for (int r = 0; r < kBR; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(Ss + r * kBK + warp * 8);
      const float4 c = *reinterpret_cast<const float4*>(Ss + r * kBK + warp * 8 + 4);
      const float4 pa = *reinterpret_cast<const float4*>(Ps + r * kBK + warp * 8);
      const float4 pc = *reinterpret_cast<const float4*>(Ps + r * kBK + warp * 8 + 4);
      const float ds[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
      const float pp[8] = {pa.x, pa.y, pa.z, pa.w, pc.x, pc.y, pc.z, pc.w};
      float qf[DPL], gf[DPL];
#pragma unroll
      for (int i = 0; i < DPL; ++i) {
        qf[i] = Feat<D>::has(lane, i) ? Qs[r * D + lane + 32 * i] : 0.f;
        gf[i] = Feat<D>::has(lane, i) ? Gs[r * D + lane + 32 * i] : 0.f;
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
#pragma unroll
        for (int i = 0; i < DPL; ++i) {
          dka[jj][i] = fmaf(ds[jj], qf[i], dka[jj][i]);
          dva[jj][i] = fmaf(pp[jj], gf[i], dva[jj][i]);
        }
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = warp * 8 + jj, gj = kt0 + j;
      float* pk = nullptr;
      float* pv = nullptr;
      if (gj < g.J) {
        const int tok = group_token(g, win, gj, g.window, g.ext);
        if (tok >= 0) {
          const long long base = (((long long)b * g.N + tok) * g.H + h) * D;
          pk = dk + base; pv = dv + base;
        }
      } else if (gj < n_keys) {
        const long long base = (((long long)b * g.H + h) * g.n_chunks + (gj - g.J)) * D;
        pk = dkbar + base; pv = dbeta + base;
      }
      if (!pk) continue;
#pragma unroll
      for (int i = 0; i < DPL; ++i)
        if (Feat<D>::has(lane, i)) {
          atomicAdd(pk + lane + 32 * i, dka[jj][i]);     // Qs carries d^-1/2 already
          atomicAdd(pv + lane + 32 * i, dva[jj][i]);
        }
    }
  }
#pragma unroll
  for (int r = 0; r < kBRw; ++r) {
    const int tq = qtok[warp * kBRw + r];
    if (tq < 0) continue;
    float* dst = dq + (((long long)b * g.N + tq) * g.H + h) * D;
#pragma unroll
    for (int i = 0; i < DPL; ++i)
      if (Feat<D>::has(lane, i)) atomicAdd(dst + lane + 32 * i, dqa[r][i] * scale);
  }
}

// ------------------------------------------------------------------------------------------------
// one warp per (batch, head, chunk); lane l owns features l, l + 32, ... (as chunk_stats_kernel)
//   rows[slot][chunk row][D]: 0 dy_k  1 dy_q (at the Linear outputs)   2 mean_k  3 mean_q (Linear inputs)
//                             4 n_k   5 n_q  (LayerNorm-normalised)    6 dout_k  7 dout_q (at the LayerNorm outputs)
// ------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void warp_ln_stats(const float (&y)[Feat<D>::kPerLane], float eps, int lane, float (&n)[Feat<D>::kPerLane],
                                              float& inv) {
  constexpr int DPL = Feat<D>::kPerLane;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < DPL; ++i) s += Feat<D>::has(lane, i) ? y[i] : 0.f;
  const float mean = warp_sum(s) * (1.0f / D);
  float vv = 0.f;
#pragma unroll
  for (int i = 0; i < DPL; ++i) { const float c = Feat<D>::has(lane, i) ? y[i] - mean : 0.f; vv = fmaf(c, c, vv); }
  inv = 1.0f / sqrtf(warp_sum(vv) * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < DPL; ++i) n[i] = Feat<D>::has(lane, i) ? (y[i] - mean) * inv : 0.f;
}

// gradient at the Linear output from the gradient at the (optional) LayerNorm output
template <int D>
__device__ __forceinline__ void warp_ln_bwd(const float (&dout)[Feat<D>::kPerLane], const float (&n)[Feat<D>::kPerLane], float inv,
                                            const float* __restrict__ gain, int lane, float (&dy)[Feat<D>::kPerLane]) {
  constexpr int DPL = Feat<D>::kPerLane;
  if (!gain) {
#pragma unroll
    for (int i = 0; i < DPL; ++i) dy[i] = dout[i];
    return;
  }
  float dn[DPL], s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < DPL; ++i) {
    dn[i] = Feat<D>::has(lane, i) ? dout[i] * __ldg(gain + lane + 32 * i) : 0.f;
    s1 += dn[i];
    s2 = fmaf(dn[i], n[i], s2);
  }
  s1 = warp_sum(s1) * (1.0f / D);
  s2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
  for (int i = 0; i < DPL; ++i) dy[i] = Feat<D>::has(lane, i) ? inv * (dn[i] - s1 - n[i] * s2) : 0.f;
}

// dx[i] = sum_e W[e][i] dy[e]   (W row-major [out][in], read through L1)
template <int D>
__device__ __forceinline__ void warp_linear_bwd(const float* __restrict__ W, const float (&dy)[Feat<D>::kPerLane],
                                                float (&dx)[Feat<D>::kPerLane], int lane) {
  constexpr int DPL = Feat<D>::kPerLane;
#pragma unroll
  for (int i = 0; i < DPL; ++i) dx[i] = 0.f;
#pragma unroll
  for (int ee = 0; ee < DPL; ++ee) {
#pragma unroll 8
    for (int jj = 0; jj < (D < 32 ? D : 32); ++jj) {
      const float d = __shfl_sync(0xffffffffu, dy[ee], jj);
      const float* wrow = W + (long long)(jj + 32 * ee) * D + lane;
#pragma unroll
      for (int i = 0; i < DPL; ++i)
        if (Feat<D>::has(lane, i)) dx[i] = fmaf(__ldg(wrow + 32 * i), d, dx[i]);
    }
  }
}

template <typename T, int D>
__global__ void __launch_bounds__(256)
chunk_stats_bwd_kernel(const Geo g, const View q, const View k, const View v, const uint8_t* __restrict__ mask,
                       const EvaAdaptive ada, const float* __restrict__ noise, const float* __restrict__ dkbar,
                       const float* __restrict__ dbeta, float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv,
                       float* __restrict__ rows) {
  constexpr int DPL = Feat<D>::kPerLane;
  extern __shared__ float sm[];
  float* WtK = sm;
  float* WtQ = sm + D * D;
  for (int idx = threadIdx.x; idx < D * D; idx += blockDim.x) {
    const int e = idx / D, i = idx % D;
    WtK[i * D + e] = __ldg(ada.w_k + idx);
    if (ada.w_q) WtQ[i * D + e] = __ldg(ada.w_q + idx);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const float scale = rsqrtf((float)D);
  const float inv_cnt = 1.0f / (float)g.Jc;
  const long long total = (long long)g.B * g.H * g.n_chunks;
  const long long slot = total * D;
  for (long long wg = (long long)blockIdx.x * wpb + warp; wg < total; wg += (long long)gridDim.x * wpb) {
    const int c = (int)(wg % g.n_chunks);
    const int h = (int)((wg / g.n_chunks) % g.H);
    const int b = (int)(wg / ((long long)g.n_chunks * g.H));
    const long long obase = wg * D;
    // ---- forward, recomputed (chunk_stats_kernel) ----
    float sq[DPL], sk[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) sq[i] = sk[i] = 0.f;
#pragma unroll 4
    for (int s = 0; s < g.Jc; ++s) {
      const int tok = group_token(g, c, s, g.chunk, g.chunk_ext);
      if (tok < 0 || (mask && mask[(long long)b * g.N + tok])) continue;
      const T* qr = q.row<T>(b, tok, h);
      const T* kr = k.row<T>(b, tok, h);
#pragma unroll
      for (int i = 0; i < DPL; ++i)
        if (Feat<D>::has(lane, i)) { sq[i] += to_f32(qr[lane + 32 * i]); sk[i] += to_f32(kr[lane + 32 * i]); }
    }
#pragma unroll
    for (int i = 0; i < DPL; ++i) { sq[i] *= inv_cnt; sk[i] *= inv_cnt; }
    float yk[DPL], nk[DPL], kb[DPL], om[DPL], nq[DPL];
    float inv_k = 1.f, inv_q = 1.f;
    warp_linear<D>(WtK, ada.b_k, sk, yk, lane);
    if (ada.ln_gain_k) {
      warp_ln_stats<D>(yk, ada.ln_eps, lane, nk, inv_k);
#pragma unroll
      for (int i = 0; i < DPL; ++i)
        kb[i] = Feat<D>::has(lane, i) ? nk[i] * __ldg(ada.ln_gain_k + lane + 32 * i) + __ldg(ada.ln_bias_k + lane + 32 * i) : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < DPL; ++i) { kb[i] = yk[i]; nk[i] = 0.f; }
    }
#pragma unroll
    for (int i = 0; i < DPL; ++i) { om[i] = 0.f; nq[i] = 0.f; }
    if (ada.w_q) {
      float yq[DPL], qb[DPL];
      warp_linear<D>(WtQ, ada.b_q, sq, yq, lane);
      if (ada.ln_gain_q) {
        warp_ln_stats<D>(yq, ada.ln_eps, lane, nq, inv_q);
#pragma unroll
        for (int i = 0; i < DPL; ++i)
          qb[i] = Feat<D>::has(lane, i) ? nq[i] * __ldg(ada.ln_gain_q + lane + 32 * i) + __ldg(ada.ln_bias_q + lane + 32 * i) : 0.f;
      } else {
#pragma unroll
        for (int i = 0; i < DPL; ++i) qb[i] = yq[i];
      }
#pragma unroll
      for (int i = 0; i < DPL; ++i) om[i] = ada.mu_coeff * (qb[i] + kb[i]);
    }
    float db[DPL], dkb_in[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      db[i] = dkb_in[i] = 0.f;
      if (!Feat<D>::has(lane, i)) continue;
      if (noise) om[i] += __ldg(noise + obase + lane + 32 * i);
      db[i] = dbeta[obase + lane + 32 * i];
      dkb_in[i] = dkbar[obase + lane + 32 * i];
    }
    float m = kNegInf, l = 0.f, acc[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) acc[i] = 0.f;
#pragma unroll 4
    for (int s = 0; s < g.Jc; ++s) {
      const int tok = group_token(g, c, s, g.chunk, g.chunk_ext);
      const bool dead = tok < 0 || (mask && mask[(long long)b * g.N + tok]);
      float lg = kMaskVal, vv[DPL];
#pragma unroll
      for (int i = 0; i < DPL; ++i) vv[i] = 0.f;
      if (!dead) {
        const T* kr = k.row<T>(b, tok, h);
        const T* vr = v.row<T>(b, tok, h);
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < DPL; ++i) {
          if (!Feat<D>::has(lane, i)) continue;
          const float kk = to_f32(kr[lane + 32 * i]);
          part = fmaf(kk, om[i] - 0.5f * kk, part);
          vv[i] = to_f32(vr[lane + 32 * i]);
        }
        lg = scale * warp_sum(part);
      }
      const float mn = fmaxf(m, lg);
      const float corr = exp_nonpos(m - mn), p = exp_nonpos(lg - mn);
      l = fmaf(l, corr, p);
#pragma unroll
      for (int i = 0; i < DPL; ++i) acc[i] = fmaf(acc[i], corr, p * vv[i]);
      m = mn;
    }
    const float inv_l = 1.0f / l;
    float dsum = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) dsum = fmaf(db[i], acc[i] * inv_l, dsum);
    dsum = warp_sum(dsum);            // <d beta, beta>
    // ---- chunk softmax backward: dv, dk (logit path), d omega ----
    float dom[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) dom[i] = 0.f;
#pragma unroll 4
    for (int s = 0; s < g.Jc; ++s) {
      const int tok = group_token(g, c, s, g.chunk, g.chunk_ext);
      if (tok < 0 || (mask && mask[(long long)b * g.N + tok])) continue;     // constant logit, zero value: no gradient
      const T* kr = k.row<T>(b, tok, h);
      const T* vr = v.row<T>(b, tok, h);
      float kk[DPL], p1 = 0.f, p2 = 0.f;
#pragma unroll
      for (int i = 0; i < DPL; ++i) {
        kk[i] = 0.f;
        if (!Feat<D>::has(lane, i)) continue;
        kk[i] = to_f32(kr[lane + 32 * i]);
        p1 = fmaf(kk[i], om[i] - 0.5f * kk[i], p1);
        p2 = fmaf(db[i], to_f32(vr[lane + 32 * i]), p2);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { p1 += __shfl_xor_sync(0xffffffffu, p1, o); p2 += __shfl_xor_sync(0xffffffffu, p2, o); }
      const float a = exp_nonpos(scale * p1 - m) * inv_l;
      const float dlg = scale * a * (p2 - dsum);
      const long long base = (((long long)b * g.N + tok) * g.H + h) * D;
#pragma unroll
      for (int i = 0; i < DPL; ++i) {
        if (!Feat<D>::has(lane, i)) continue;
        atomicAdd(dv + base + lane + 32 * i, a * db[i]);
        atomicAdd(dk + base + lane + 32 * i, dlg * (om[i] - kk[i]));
        dom[i] = fmaf(dlg, kk[i], dom[i]);
      }
    }
    // ---- omega -> LayerNorm -> Linear -> means ----
    float dok[DPL], doq[DPL], dyk[DPL], dyq[DPL], dmk[DPL], dmq[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      dok[i] = dkb_in[i] + (ada.w_q ? ada.mu_coeff * dom[i] : 0.f);
      doq[i] = ada.w_q ? ada.mu_coeff * dom[i] : 0.f;
      dyq[i] = dmq[i] = 0.f;
    }
    warp_ln_bwd<D>(dok, nk, inv_k, ada.ln_gain_k, lane, dyk);
    warp_linear_bwd<D>(ada.w_k, dyk, dmk, lane);
    if (ada.w_q) {
      warp_ln_bwd<D>(doq, nq, inv_q, ada.ln_gain_q, lane, dyq);
      warp_linear_bwd<D>(ada.w_q, dyq, dmq, lane);
    }
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      if (!Feat<D>::has(lane, i)) continue;
      const long long o = obase + lane + 32 * i;
      rows[0 * slot + o] = dyk[i];  rows[1 * slot + o] = dyq[i];
      rows[2 * slot + o] = sk[i];   rows[3 * slot + o] = sq[i];
      rows[4 * slot + o] = nk[i];   rows[5 * slot + o] = nq[i];
      rows[6 * slot + o] = dok[i];  rows[7 * slot + o] = doq[i];
    }
#pragma unroll 4
    for (int s = 0; s < g.Jc; ++s) {
      const int tok = group_token(g, c, s, g.chunk, g.chunk_ext);
      if (tok < 0 || (mask && mask[(long long)b * g.N + tok])) continue;
      const long long base = (((long long)b * g.N + tok) * g.H + h) * D;
#pragma unroll
      for (int i = 0; i < DPL; ++i) {
        if (!Feat<D>::has(lane, i)) continue;
        atomicAdd(dk + base + lane + 32 * i, dmk[i] * inv_cnt);
        if (ada.w_q) atomicAdd(dq + base + lane + 32 * i, dmq[i] * inv_cnt);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Chunk-statistics gradient, fast case: head_dim 64, 16-bit I/O, chunks without halo of at most 32 tokens, no padding mask (the
// DeiT geometries).  Every token then belongs to exactly one chunk, so the warp that owns a chunk also FINISHES the gradient of the
// chunk's tokens: it adds the chunk-path terms to the window kernel's float32 dq / dk / dv rows and writes the result once -- no
// atomics -- either back in place or, when `gio` is given, rounded to the I/O format in the packed [B, N, 3, H, 64] layout of the
// qkv projection (the caller's convert-and-interleave pass disappears).  One warp per chunk; lane l owns features 2l, 2l + 1 of
// every vector (4-byte loads of 16-bit rows); the per-token dot products run with lane = token on k / v rows staged in shared
// memory, so the 2 x Jc warp reductions of chunk_stats_bwd_kernel become two.
// ------------------------------------------------------------------------------------------------
constexpr int kCgMaxJc = 32;
constexpr int kCgRow = 132;          // bytes per staged 16-bit row: 33 words, conflict-free for lane = token and for lane = feature pair

template <typename T>
__global__ void __launch_bounds__(256)
chunk_grad_fast_kernel(const Geo g, const View q, const View k, const View v, const EvaAdaptive ada, const float* __restrict__ noise,
                       const float* __restrict__ beta_fw, const float* __restrict__ dkbar, const float* __restrict__ dbeta,
                       float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv, float* __restrict__ rows,
                       T* __restrict__ gio) {
  extern __shared__ float sm[];
  float* WtK = sm;
  float* WtQ = sm + 64 * 64;
  for (int idx = threadIdx.x; idx < 64 * 64; idx += blockDim.x) {
    const int e = idx >> 6, i = idx & 63;
    WtK[i * 64 + e] = __ldg(ada.w_k + idx);
    if (ada.w_q) WtQ[i * 64 + e] = __ldg(ada.w_q + idx);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int Jc = g.Jc;
  uint8_t* tiles = reinterpret_cast<uint8_t*>(sm + 2 * 64 * 64) + (size_t)warp * (2 * Jc * kCgRow + 512);
  uint8_t* Kt = tiles;                                              // [Jc][kCgRow] k rows
  uint8_t* Vt = tiles + Jc * kCgRow;                                // [Jc][kCgRow] v rows
  float* vec = reinterpret_cast<float*>(tiles + 2 * Jc * kCgRow);   // omega [64] | d beta [64]
  const float scale = 0.125f;
  const float inv_cnt = 1.0f / (float)Jc;
  const long long total = (long long)g.B * g.H * g.n_chunks;
  const long long slot = total * 64;
  for (long long wg = (long long)blockIdx.x * wpb + warp; wg < total; wg += (long long)gridDim.x * wpb) {
    const int c = (int)(wg % g.n_chunks);
    const int h = (int)((wg / g.n_chunks) % g.H);
    const int b = (int)(wg / ((long long)g.n_chunks * g.H));
    const long long obase = wg * 64;
    const int my_tok = lane < Jc ? group_token(g, c, lane, g.chunk, 0) : 0;
    // ---- stage the chunk's k / v rows; column sums of q and k ----
    __syncwarp();
    for (int idx = lane; idx < Jc * 8; idx += 32) {
      const int t = idx >> 3, piece = idx & 7;
      const int tok = __shfl_sync(0xffffffffu, my_tok, t);
      const uint4 rk = __ldg(reinterpret_cast<const uint4*>(k.row<T>(b, tok, h)) + piece);
      const uint4 rv = __ldg(reinterpret_cast<const uint4*>(v.row<T>(b, tok, h)) + piece);
      uint32_t* dk_ = reinterpret_cast<uint32_t*>(Kt + t * kCgRow + piece * 16);
      uint32_t* dv_ = reinterpret_cast<uint32_t*>(Vt + t * kCgRow + piece * 16);
      dk_[0] = rk.x; dk_[1] = rk.y; dk_[2] = rk.z; dk_[3] = rk.w;
      dv_[0] = rv.x; dv_[1] = rv.y; dv_[2] = rv.z; dv_[3] = rv.w;
    }
    float2 sq = make_float2(0.f, 0.f), sk = sq;
    for (int t = 0; t < Jc; ++t) {
      const int tok = __shfl_sync(0xffffffffu, my_tok, t);
      const float2 a = Pair16<T>::up(__ldg(reinterpret_cast<const uint32_t*>(q.row<T>(b, tok, h)) + lane));
      const float2 c2 = Pair16<T>::up(__ldg(reinterpret_cast<const uint32_t*>(k.row<T>(b, tok, h)) + lane));
      sq.x += a.x; sq.y += a.y; sk.x += c2.x; sk.y += c2.y;
    }
    sq.x *= inv_cnt; sq.y *= inv_cnt; sk.x *= inv_cnt; sk.y *= inv_cnt;
    // ---- forward statistics: k_bar, omega ----
    float inv_k = 1.f, inv_q = 1.f;
    float2 nk = make_float2(0.f, 0.f), nq = nk, om = nk;
    float2 kb = pair_linear(WtK, ada.b_k, sk, lane);
    if (ada.ln_gain_k) {
      nk = pair_ln(kb, ada.ln_eps, inv_k);
      const float2 gg = __ldg(reinterpret_cast<const float2*>(ada.ln_gain_k) + lane), bb = __ldg(reinterpret_cast<const float2*>(ada.ln_bias_k) + lane);
      kb = make_float2(fmaf(nk.x, gg.x, bb.x), fmaf(nk.y, gg.y, bb.y));
    }
    if (ada.w_q) {
      float2 qb = pair_linear(WtQ, ada.b_q, sq, lane);
      if (ada.ln_gain_q) {
        nq = pair_ln(qb, ada.ln_eps, inv_q);
        const float2 gg = __ldg(reinterpret_cast<const float2*>(ada.ln_gain_q) + lane), bb = __ldg(reinterpret_cast<const float2*>(ada.ln_bias_q) + lane);
        qb = make_float2(fmaf(nq.x, gg.x, bb.x), fmaf(nq.y, gg.y, bb.y));
      }
      om = make_float2(ada.mu_coeff * (qb.x + kb.x), ada.mu_coeff * (qb.y + kb.y));
    }
    if (noise) { const float2 z = __ldg(reinterpret_cast<const float2*>(noise + obase) + lane); om.x += z.x; om.y += z.y; }
    const float2 db = *(reinterpret_cast<const float2*>(dbeta + obase) + lane);
    const float2 dkb_in = *(reinterpret_cast<const float2*>(dkbar + obase) + lane);
    const float2 bf = __ldg(reinterpret_cast<const float2*>(beta_fw + obase) + lane);
    const float dsum = warp_sum(db.x * bf.x + db.y * bf.y);                  // <d beta, beta>
    reinterpret_cast<float2*>(vec)[lane] = om;
    reinterpret_cast<float2*>(vec + 64)[lane] = db;
    __syncwarp();
    // ---- lane = token: phi-logit and <d beta, v_t> ----
    float p1 = 0.f, p2 = 0.f;
    if (lane < Jc) {
      const uint32_t* kr = reinterpret_cast<const uint32_t*>(Kt + lane * kCgRow);
      const uint32_t* vr = reinterpret_cast<const uint32_t*>(Vt + lane * kCgRow);
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        const float2 kk = Pair16<T>::up(kr[i]), vv = Pair16<T>::up(vr[i]);
        const float2 oo = reinterpret_cast<const float2*>(vec)[i], dd = reinterpret_cast<const float2*>(vec + 64)[i];
        p1 = fmaf(kk.x, oo.x - 0.5f * kk.x, fmaf(kk.y, oo.y - 0.5f * kk.y, p1));
        p2 = fmaf(dd.x, vv.x, fmaf(dd.y, vv.y, p2));
      }
    }
    const float lg = lane < Jc ? scale * p1 : kNegInf;
    const float mx = warp_max(lg);
    const float e = exp_nonpos(lg - mx);
    const float a = e / warp_sum(e);                                           // softmax over the chunk's tokens (0 on idle lanes)
    const float dlg = scale * a * (p2 - dsum);
    // ---- d omega = sum_t dlg_t k_t ----
    float2 dom = make_float2(0.f, 0.f);
    for (int t = 0; t < Jc; ++t) {
      const float d = __shfl_sync(0xffffffffu, dlg, t);
      const float2 kk = Pair16<T>::up(reinterpret_cast<const uint32_t*>(Kt + t * kCgRow)[lane]);
      dom.x = fmaf(d, kk.x, dom.x); dom.y = fmaf(d, kk.y, dom.y);
    }
    // ---- omega -> LayerNorm -> Linear -> means ----
    const float cq = ada.w_q ? ada.mu_coeff : 0.f;
    const float2 dok = make_float2(dkb_in.x + cq * dom.x, dkb_in.y + cq * dom.y);
    const float2 doq = make_float2(cq * dom.x, cq * dom.y);
    const float2 dyk = pair_ln_bwd(dok, nk, inv_k, ada.ln_gain_k, lane);
    const float2 dmk = pair_linear_bwd(ada.w_k, dyk, lane);
    float2 dyq = make_float2(0.f, 0.f), dmq = dyq;
    if (ada.w_q) {
      dyq = pair_ln_bwd(doq, nq, inv_q, ada.ln_gain_q, lane);
      dmq = pair_linear_bwd(ada.w_q, dyq, lane);
    }
    {
      float2* r2 = reinterpret_cast<float2*>(rows + obase) + lane;
      const long long s2 = slot / 2;
      r2[0] = dyk; r2[s2] = dyq; r2[2 * s2] = sk; r2[3 * s2] = sq; r2[4 * s2] = nk; r2[5 * s2] = nq; r2[6 * s2] = dok; r2[7 * s2] = doq;
    }
    // ---- finish the chunk's tokens: window-path rows + chunk-path terms, written once ----
    for (int t = 0; t < Jc; ++t) {
      const int tok = __shfl_sync(0xffffffffu, my_tok, t);
      const float at = __shfl_sync(0xffffffffu, a, t), dt = __shfl_sync(0xffffffffu, dlg, t);
      const float2 kk = Pair16<T>::up(reinterpret_cast<const uint32_t*>(Kt + t * kCgRow)[lane]);
      const long long base = (((long long)b * g.N + tok) * g.H + h) * 64;
      float2* pq = reinterpret_cast<float2*>(dq + base) + lane;
      float2* pk = reinterpret_cast<float2*>(dk + base) + lane;
      float2* pv = reinterpret_cast<float2*>(dv + base) + lane;
      float2 gq = *pq, gk = *pk, gv = *pv;
      gq.x = fmaf(dmq.x, inv_cnt, gq.x); gq.y = fmaf(dmq.y, inv_cnt, gq.y);
      gk.x += fmaf(dt, om.x - kk.x, dmk.x * inv_cnt); gk.y += fmaf(dt, om.y - kk.y, dmk.y * inv_cnt);
      gv.x = fmaf(at, db.x, gv.x); gv.y = fmaf(at, db.y, gv.y);
      if (gio) {
        uint32_t* o = reinterpret_cast<uint32_t*>(gio + ((((long long)b * g.N + tok) * 3) * g.H + h) * 64) + lane;
        const long long plane = (long long)g.H * 32;       // one of q | k | v, in 32-bit words
        o[0] = Pair16<T>::pk(gq.x, gq.y); o[plane] = Pair16<T>::pk(gk.x, gk.y); o[2 * plane] = Pair16<T>::pk(gv.x, gv.y);
      } else {
        *pq = gq; *pk = gk; *pv = gv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Chunk-statistics gradient for head_dim 64 with 16-bit I/O and chunks of at most 256 slots, ANY geometry (halo, padding, 1-D /
// 2-D): the backward twin of chunk_stats_fast_kernel.  Tokens may belong to several chunks here, so the token gradients are
// accumulated with vector reductions (red.global.add.v4.f32) instead of being finished in place; everything else follows the two
// fast kernels: one token-index computation per slot, four tokens per 16-byte load instruction in the feature passes, lane = token
// for the two per-token dot products (phi-logit, <d beta, v_t>).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
chunk_stats_bwd_fast_kernel(const Geo g, const View q, const View k, const View v, const uint8_t* __restrict__ mask,
                            const EvaAdaptive ada, const float* __restrict__ noise, const float* __restrict__ beta_fw,
                            const float* __restrict__ dkbar, const float* __restrict__ dbeta, float* __restrict__ dq,
                            float* __restrict__ dk, float* __restrict__ dv, float* __restrict__ rows) {
  extern __shared__ float sm[];
  float* WtK = sm;
  float* WtQ = sm + 64 * 64;
  for (int idx = threadIdx.x; idx < 64 * 64; idx += blockDim.x) {
    const int e = idx >> 6, i = idx & 63;
    WtK[i * 64 + e] = __ldg(ada.w_k + idx);
    if (ada.w_q) WtQ[i * 64 + e] = __ldg(ada.w_q + idx);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* vec = sm + 2 * 64 * 64 + warp * 320;         // omega | d beta | dm_k / Jc | dm_q / Jc | (spare)   64 floats each
  const float scale = 0.125f;
  const float inv_cnt = 1.0f / (float)g.Jc;
  const int n_rounds = (g.Jc + 31) >> 5;
  const int ts = lane >> 3, p8 = lane & 7;
  const long long total = (long long)g.B * g.H * g.n_chunks;
  const long long slot = total * 64;
  for (long long wg = (long long)blockIdx.x * wpb + warp; wg < total; wg += (long long)gridDim.x * wpb) {
    const int c = (int)(wg % g.n_chunks);
    const int h = (int)((wg / g.n_chunks) % g.H);
    const int b = (int)(wg / ((long long)g.n_chunks * g.H));
    const long long obase = wg * 64;
    int tokr[8];
#pragma unroll
    for (int rd = 0; rd < 8; ++rd) {
      const int s = 32 * rd + lane;
      int t = (rd < n_rounds && s < g.Jc) ? group_token(g, c, s, g.chunk, g.chunk_ext) : -1;
      if (t >= 0 && mask && mask[(long long)b * g.N + t]) t = -1;
      tokr[rd] = t;
    }
    // ---- forward statistics, recomputed ----
    float aq[8], ak[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) aq[i] = ak[i] = 0.f;
#pragma unroll
    for (int rd = 0; rd < 8; ++rd) {
      if (rd >= n_rounds) break;
#pragma unroll 4
      for (int i = 0; i < 8; ++i) {
        const int t = __shfl_sync(0xffffffffu, tokr[rd], 4 * i + ts);
        if (t < 0) continue;
        add8<T>(__ldg(reinterpret_cast<const uint4*>(q.row<T>(b, t, h)) + p8), 1.f, aq);
        add8<T>(__ldg(reinterpret_cast<const uint4*>(k.row<T>(b, t, h)) + p8), 1.f, ak);
      }
    }
    float2 sq = pieces_to_pair(aq, lane), sk = pieces_to_pair(ak, lane);
    sq.x *= inv_cnt; sq.y *= inv_cnt; sk.x *= inv_cnt; sk.y *= inv_cnt;
    float inv_k = 1.f, inv_q = 1.f;
    float2 nk = make_float2(0.f, 0.f), nq = nk, om = nk;
    float2 kb = pair_linear(WtK, ada.b_k, sk, lane);
    if (ada.ln_gain_k) {
      nk = pair_ln(kb, ada.ln_eps, inv_k);
      const float2 gg = __ldg(reinterpret_cast<const float2*>(ada.ln_gain_k) + lane), bb = __ldg(reinterpret_cast<const float2*>(ada.ln_bias_k) + lane);
      kb = make_float2(fmaf(nk.x, gg.x, bb.x), fmaf(nk.y, gg.y, bb.y));
    }
    if (ada.w_q) {
      float2 qb = pair_linear(WtQ, ada.b_q, sq, lane);
      if (ada.ln_gain_q) {
        nq = pair_ln(qb, ada.ln_eps, inv_q);
        const float2 gg = __ldg(reinterpret_cast<const float2*>(ada.ln_gain_q) + lane), bb = __ldg(reinterpret_cast<const float2*>(ada.ln_bias_q) + lane);
        qb = make_float2(fmaf(nq.x, gg.x, bb.x), fmaf(nq.y, gg.y, bb.y));
      }
      om = make_float2(ada.mu_coeff * (qb.x + kb.x), ada.mu_coeff * (qb.y + kb.y));
    }
    if (noise) { const float2 z = __ldg(reinterpret_cast<const float2*>(noise + obase) + lane); om.x += z.x; om.y += z.y; }
    const float2 db = *(reinterpret_cast<const float2*>(dbeta + obase) + lane);
    const float2 dkb_in = *(reinterpret_cast<const float2*>(dkbar + obase) + lane);
    const float2 bf = __ldg(reinterpret_cast<const float2*>(beta_fw + obase) + lane);
    const float dsum = warp_sum(db.x * bf.x + db.y * bf.y);                  // <d beta, beta>
    __syncwarp();
    reinterpret_cast<float2*>(vec)[lane] = om;
    reinterpret_cast<float2*>(vec + 64)[lane] = db;
    __syncwarp();
    // ---- lane = token: phi-logit and <d beta, v_t> ----
    float lg[8], p2[8];
    float mx = kNegInf;
#pragma unroll
    for (int rd = 0; rd < 8; ++rd) {
      lg[rd] = kNegInf; p2[rd] = 0.f;
      if (rd >= n_rounds) continue;
      if (32 * rd + lane < g.Jc) {
        lg[rd] = kMaskVal;
        const int t = tokr[rd];
        if (t >= 0) {
          const uint4* kr = reinterpret_cast<const uint4*>(k.row<T>(b, t, h));
          const uint4* vr = reinterpret_cast<const uint4*>(v.row<T>(b, t, h));
          float part = 0.f, pv = 0.f;
#pragma unroll
          for (int pc = 0; pc < 8; ++pc) {
            const uint4 rk = __ldg(kr + pc), rv = __ldg(vr + pc);
            const uint32_t wk4[4] = {rk.x, rk.y, rk.z, rk.w}, wv4[4] = {rv.x, rv.y, rv.z, rv.w};
            const float4 o0 = *reinterpret_cast<const float4*>(vec + 8 * pc), o1 = *reinterpret_cast<const float4*>(vec + 8 * pc + 4);
            const float4 d0 = *reinterpret_cast<const float4*>(vec + 64 + 8 * pc), d1 = *reinterpret_cast<const float4*>(vec + 64 + 8 * pc + 4);
            const float oo[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
            const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 kk = Pair16<T>::up(wk4[u]), vv = Pair16<T>::up(wv4[u]);
              part = fmaf(kk.x, oo[2 * u] - 0.5f * kk.x, fmaf(kk.y, oo[2 * u + 1] - 0.5f * kk.y, part));
              pv = fmaf(dd[2 * u], vv.x, fmaf(dd[2 * u + 1], vv.y, pv));
            }
          }
          lg[rd] = scale * part;
          p2[rd] = pv;
        }
      }
      mx = fmaxf(mx, lg[rd]);
    }
    mx = warp_max(mx);
    float den = 0.f;
#pragma unroll
    for (int rd = 0; rd < 8; ++rd) { lg[rd] = exp_nonpos(lg[rd] - mx); den += lg[rd]; }
    const float inv_l = 1.0f / warp_sum(den);
    float dlg[8];
#pragma unroll
    for (int rd = 0; rd < 8; ++rd) {
      lg[rd] *= inv_l;                                                     // a_t
      dlg[rd] = tokr[rd] >= 0 ? scale * lg[rd] * (p2[rd] - dsum) : 0.f;    // masked slots: constant logit, no gradient
    }
    // ---- d omega = sum_t dlg_t k_t ----
    float ad[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ad[i] = 0.f;
#pragma unroll
    for (int rd = 0; rd < 8; ++rd) {
      if (rd >= n_rounds) break;
#pragma unroll 4
      for (int i = 0; i < 8; ++i) {
        const int t = __shfl_sync(0xffffffffu, tokr[rd], 4 * i + ts);
        const float d = __shfl_sync(0xffffffffu, dlg[rd], 4 * i + ts);
        if (t < 0) continue;
        add8<T>(__ldg(reinterpret_cast<const uint4*>(k.row<T>(b, t, h)) + p8), d, ad);
      }
    }
    const float2 dom = pieces_to_pair(ad, lane);
    // ---- omega -> LayerNorm -> Linear -> means ----
    const float cq = ada.w_q ? ada.mu_coeff : 0.f;
    const float2 dok = make_float2(dkb_in.x + cq * dom.x, dkb_in.y + cq * dom.y);
    const float2 doq = make_float2(cq * dom.x, cq * dom.y);
    const float2 dyk = pair_ln_bwd(dok, nk, inv_k, ada.ln_gain_k, lane);
    const float2 dmk = pair_linear_bwd(ada.w_k, dyk, lane);
    float2 dyq = make_float2(0.f, 0.f), dmq = dyq;
    if (ada.w_q) {
      dyq = pair_ln_bwd(doq, nq, inv_q, ada.ln_gain_q, lane);
      dmq = pair_linear_bwd(ada.w_q, dyq, lane);
    }
    {
      float2* r2 = reinterpret_cast<float2*>(rows + obase) + lane;
      const long long s2 = slot / 2;
      r2[0] = dyk; r2[s2] = dyq; r2[2 * s2] = sk; r2[3 * s2] = sq; r2[4 * s2] = nk; r2[5 * s2] = nq; r2[6 * s2] = dok; r2[7 * s2] = doq;
    }
    reinterpret_cast<float2*>(vec + 128)[lane] = make_float2(dmk.x * inv_cnt, dmk.y * inv_cnt);
    reinterpret_cast<float2*>(vec + 192)[lane] = make_float2(dmq.x * inv_cnt, dmq.y * inv_cnt);
    __syncwarp();
    // ---- token gradients: four tokens per instruction, 8 features per lane, vector reductions ----
    float o8[8], d8[8], mk8[8], mq8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o8[i] = vec[8 * p8 + i]; d8[i] = vec[64 + 8 * p8 + i]; mk8[i] = vec[128 + 8 * p8 + i]; mq8[i] = vec[192 + 8 * p8 + i]; }
#pragma unroll
    for (int rd = 0; rd < 8; ++rd) {
      if (rd >= n_rounds) break;
#pragma unroll 2
      for (int i = 0; i < 8; ++i) {
        const int t = __shfl_sync(0xffffffffu, tokr[rd], 4 * i + ts);
        const float at = __shfl_sync(0xffffffffu, lg[rd], 4 * i + ts), dt = __shfl_sync(0xffffffffu, dlg[rd], 4 * i + ts);
        if (t < 0) continue;
        const uint4 rk = __ldg(reinterpret_cast<const uint4*>(k.row<T>(b, t, h)) + p8);
        const uint32_t w4[4] = {rk.x, rk.y, rk.z, rk.w};
        float gk[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 kk = Pair16<T>::up(w4[u]);
          gk[2 * u] = fmaf(dt, o8[2 * u] - kk.x, mk8[2 * u]);
          gk[2 * u + 1] = fmaf(dt, o8[2 * u + 1] - kk.y, mk8[2 * u + 1]);
        }
        const long long base = (((long long)b * g.N + t) * g.H + h) * 64 + 8 * p8;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dk + base), "f"(gk[0]), "f"(gk[1]), "f"(gk[2]), "f"(gk[3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dk + base + 4), "f"(gk[4]), "f"(gk[5]), "f"(gk[6]), "f"(gk[7]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dv + base), "f"(at * d8[0]), "f"(at * d8[1]), "f"(at * d8[2]), "f"(at * d8[3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dv + base + 4), "f"(at * d8[4]), "f"(at * d8[5]), "f"(at * d8[6]), "f"(at * d8[7]) : "memory");
        if (ada.w_q) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dq + base), "f"(mq8[0]), "f"(mq8[1]), "f"(mq8[2]), "f"(mq8[3]) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dq + base + 4), "f"(mq8[4]), "f"(mq8[5]), "f"(mq8[6]), "f"(mq8[7]) : "memory");
        }
      }
    }
    __syncwarp();
  }
}

// float32 [3, B, N, H, D] -> io format, packed [B, N, 3, H, D] (the cases the fast kernel above does not finish itself)
template <typename T>
__global__ void pack_grad_kernel(const float* __restrict__ g3, T* __restrict__ gio, long long tens, int HD, int N) {
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;      // 4 elements of one (which, b, n) row
  if (i4 >= 3 * tens) return;
  const long long which = i4 / tens, rem = i4 % tens;
  const long long bn = rem / HD, e = rem % HD;
  const float4 x = *reinterpret_cast<const float4*>(g3 + i4);
  T* o = gio + (bn * 3 + which) * HD + e;
  o[0] = from_f32<T>(x.x); o[1] = from_f32<T>(x.y); o[2] = from_f32<T>(x.z); o[3] = from_f32<T>(x.w);
}

// ------------------------------------------------------------------------------------------------
template <typename T, int D>
static cudaError_t launch_bwd_t(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask,
                                const EvaAdaptive* ada, const float* noise, const float* kbar, const float* beta, const float* bias,
                                long long bias_sh, const void* out, const void* dout, float* dq, float* dk, float* dv, float* dkbar,
                                float* dbeta, float* dbias, float* rows, void* gio, cudaStream_t st, const float* lse) {
  cudaError_t e;
  const long long tens = (long long)g.B * g.N * g.H * g.D;
  auto pack = [&]() -> cudaError_t {          // generic finish: convert and interleave
    if (!gio) return cudaSuccess;
    const long long n4 = (3 * tens + 3) / 4;
    pack_grad_kernel<T><<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(dq, reinterpret_cast<T*>(gio), tens, g.H * g.D, g.N);
    return cudaGetLastError();
  };
  if (window_bwd_tc_supported(g, io_dtype, mask)) {
    e = launch_window_bwd_tc(g, io_dtype, q, k, v, kbar, beta, bias, bias_sh, out, dout, dq, dk, dv, dkbar, dbeta, dbias, st);
  } else if (window_bwd_gen_supported(g, io_dtype)) {
    e = launch_window_bwd_gen(g, io_dtype, q, k, v, mask, kbar, beta, bias, bias_sh, out, dout, dq, dk, dv, dkbar, dbeta, dbias, st, lse);
  } else {
    auto kern = window_attn_bwd_kernel<T, D>;
    const size_t smem = BwdSmem<D>::kBytes;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long ctas = (long long)((g.L + kBR - 1) / kBR) * g.n_windows * g.B * g.H;
    if (ctas > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    kern<<<(unsigned)ctas, 256, smem, st>>>(g, q, k, v, mask, kbar, beta, bias, bias_sh, reinterpret_cast<const T*>(out),
                                            reinterpret_cast<const T*>(dout), dq, dk, dv, dkbar, dbeta, dbias);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return e;
  if (g.n_chunks == 0) return pack();
  if constexpr (D == 64 && sizeof(T) == 2) {
    static const bool slow_only = [] { const char* e_ = getenv("EVA_SM100_CHUNK_BWD_GENERIC"); return e_ && e_[0] == '1'; }();
    if (!slow_only && !mask && g.chunk_ext == 0 && g.Jc <= kCgMaxJc) {
      auto kf = chunk_grad_fast_kernel<T>;
      const size_t smf = 2 * 64 * 64 * sizeof(float) + 8 * ((size_t)2 * g.Jc * kCgRow + 512);
      e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smf);
      if (e != cudaSuccess) return e;
      const long long total_f = (long long)g.B * g.H * g.n_chunks;
      long long blocks_f = (total_f + 7) / 8;
      if (blocks_f > 148LL * 12) blocks_f = 148LL * 12;
      kf<<<(unsigned)blocks_f, 256, smf, st>>>(g, q, k, v, *ada, noise, beta, dkbar, dbeta, dq, dk, dv, rows, reinterpret_cast<T*>(gio));
      return cudaGetLastError();
    }
  }
  if constexpr (D == 64 && sizeof(T) == 2) {
    static const bool slow_only2 = [] { const char* e_ = getenv("EVA_SM100_CHUNK_BWD_GENERIC"); return e_ && e_[0] == '1'; }();
    if (!slow_only2 && g.Jc <= 256) {
      auto kf = chunk_stats_bwd_fast_kernel<T>;
      const size_t smf = (2 * 64 * 64 + 8 * 320) * sizeof(float);
      e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smf);
      if (e != cudaSuccess) return e;
      const long long total_f = (long long)g.B * g.H * g.n_chunks;
      long long blocks_f = (total_f + 7) / 8;
      if (blocks_f > 148LL * 16) blocks_f = 148LL * 16;
      kf<<<(unsigned)blocks_f, 256, smf, st>>>(g, q, k, v, mask, *ada, noise, beta, dkbar, dbeta, dq, dk, dv, rows);
      e = cudaGetLastError();
      return e != cudaSuccess ? e : pack();
    }
  }
  auto kern2 = chunk_stats_bwd_kernel<T, D>;
  const size_t smem2 = 2 * (size_t)D * D * sizeof(float);
  e = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  if (e != cudaSuccess) return e;
  const long long total = (long long)g.B * g.H * g.n_chunks;
  long long blocks = (total + 7) / 8;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  kern2<<<(unsigned)blocks, 256, smem2, st>>>(g, q, k, v, mask, *ada, noise, dkbar, dbeta, dq, dk, dv, rows);
  e = cudaGetLastError();
  return e != cudaSuccess ? e : pack();
}

cudaError_t launch_eva_backward(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask,
                                const EvaAdaptive* ada, const float* noise, const float* kbar, const float* beta, const float* bias,
                                long long bias_sh, const void* out, const void* dout, float* dq, float* dk, float* dv, float* dkbar,
                                float* dbeta, float* dbias, float* rows, void* gio, cudaStream_t st, const float* lse) {
#define EVA_BWD_CASE(DT, TY, DD) \
  case DT * 256 + DD: return launch_bwd_t<TY, DD>(g, io_dtype, q, k, v, mask, ada, noise, kbar, beta, bias, bias_sh, out, dout, dq, dk, dv, dkbar, dbeta, dbias, rows, gio, st, lse);
  switch (io_dtype * 256 + g.D) {
    EVA_BWD_CASE(EVA_F32, float, 16) EVA_BWD_CASE(EVA_F32, float, 32) EVA_BWD_CASE(EVA_F32, float, 64) EVA_BWD_CASE(EVA_F32, float, 128)
    EVA_BWD_CASE(EVA_F16, __half, 16) EVA_BWD_CASE(EVA_F16, __half, 32) EVA_BWD_CASE(EVA_F16, __half, 64) EVA_BWD_CASE(EVA_F16, __half, 128)
    EVA_BWD_CASE(EVA_BF16, __nv_bfloat16, 16) EVA_BWD_CASE(EVA_BF16, __nv_bfloat16, 32) EVA_BWD_CASE(EVA_BF16, __nv_bfloat16, 64)
    EVA_BWD_CASE(EVA_BF16, __nv_bfloat16, 128)
    default: return cudaErrorInvalidValue;
  }
#undef EVA_BWD_CASE
}

}  // namespace eva
