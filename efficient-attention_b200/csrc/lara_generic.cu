// LARA (LinearRA) forward core, generic CUDA-core implementation: fp32 math, T = HBM format.
//
//   lara_landmark_kernel   lara.py:84-175 + 182-198 + the [S,C] proposal statistics of :221-236
//   lara_stats_kernel      lara.py:202-211 (kv statistics, lse of phi(k)) and :222-223 (lse of t_nc)
//   lara_out_kernel        lara.py:201, 214-246 (phi(q), MIS weights, self-normalised combine)
//
// Workspace per (batch, head), float32:
//   qbar[C,D] mu[C,D] omega[S,D] kv[S,D] lp[S] bh[S] lse_k[S] lse_t[C]
#include <math.h>

#include <type_traits>

#include "common.cuh"
#include "launch.h"

#include "lara_ws.cuh"

namespace eva {

__device__ __forceinline__ void bin_of(int i, int n_in, int n_out, int& lo, int& hi) {
  // AdaptiveAvgPool bins: [floor(i*n_in/n_out), ceil((i+1)*n_in/n_out))
  lo = (i * n_in) / n_out;
  hi = ((i + 1) * n_in + n_out - 1) / n_out;
}

template <int D>
__device__ __forceinline__ void warp_linear_ln(const float* Wt, const EvaAdaptive& p, bool q_side,
                                               const float (&x)[Feat<D>::kPerLane], float (&y)[Feat<D>::kPerLane], int lane) {
  warp_linear<D>(Wt, q_side ? p.b_q : p.b_k, x, y, lane);
  const float* gain = q_side ? p.ln_gain_q : p.ln_gain_k;
  if (gain) warp_layer_norm<D>(y, gain, q_side ? p.ln_bias_q : p.ln_bias_k, p.ln_eps, lane);
}

// Ms[r][c] = <X_r, Y_c> for r < R, c < C.  X: rows with stride D+1 in shared memory; Yal: 16-byte aligned [C][D] copy of Y.
// One work item = (row, block of 10 columns): the row lives in registers, every Y row is a broadcast read -- ~13x fewer
// shared-memory instructions than one 64-step dot product per lane.
template <int D>
__device__ __forceinline__ void dots_rows_cols(const float* __restrict__ X, int R, const float* __restrict__ Yal, int C,
                                               float* __restrict__ Ms, int tid, int nthreads) {
  constexpr int DP = D + 1, CB = 10;
  const int n_cb = (C + CB - 1) / CB;
  for (int w = tid; w < R * n_cb; w += nthreads) {
    const int r = w % R, c0 = (w / R) * CB;
    float x[D];
#pragma unroll
    for (int e = 0; e < D; ++e) x[e] = X[r * DP + e];
    const int c1 = c0 + CB < C ? c0 + CB : C;
    for (int c = c0; c < c1; ++c) {
      const float4* y = reinterpret_cast<const float4*>(Yal + c * D);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int i4 = 0; i4 < D / 4; ++i4) {
        const float4 y4 = y[i4];
        a0 = fmaf(x[4 * i4], y4.x, a0); a1 = fmaf(x[4 * i4 + 1], y4.y, a1);
        a2 = fmaf(x[4 * i4 + 2], y4.z, a2); a3 = fmaf(x[4 * i4 + 3], y4.w, a3);
      }
      Ms[r * C + c] = (a0 + a1) + (a2 + a3);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// L1: landmarks + proposal statistics.  One CTA per (batch, head).
// ------------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(256)
lara_landmark_kernel(const LaraGeo g, const View q, const View k, const View v, const uint8_t* __restrict__ mask,
                     const EvaAdaptive proj, const float* __restrict__ noise, float* __restrict__ ws_base,
                     const float* __restrict__ given) {
  constexpr int DPL = Feat<D>::kPerLane, DP = D + 1;
  extern __shared__ float sm[];
  const int C = g.C, S = g.S;
  float* qb = sm;                  // [C][DP]  q landmarks, later mu
  float* kb = qb + C * DP;         // [C][DP]
  float* vb = kb + C * DP;         // [C][DP]  (vmixed)
  float* kb2 = vb + C * DP;        // [C][DP]  mixed k landmarks
  float* om = kb2 + C * DP;        // [S][DP]
  float* Wt = om + S * DP;         // [D][D]
  float* rowbuf = Wt + D * D;      // [8][C]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / g.H, h = blockIdx.x % g.H;
  const float scale = rsqrtf((float)D);
  const LaraWs ws = lara_ws_at(ws_base, blockIdx.x, C, S, D);
  const bool has_proj = proj.w_q != nullptr;

  bool coop = false;
  if constexpr (D == 64) coop = g.dims == 2 && !g.per_token_proj;
  // `given`: landmarks computed by the caller (pool_module_type == 'dense', lara.py:36-39,131-139: Linear / LayerNorm over ALL
  // channels, i.e. across heads), float32 [batch * heads][3][C][D] = q_bar | k_bar (before mixing) | v_bar ('-vmixed' only)
  const bool gen = given == nullptr;
  if (!gen) {
    const float* src = given + (long long)blockIdx.x * 3 * C * D;
    for (int idx = tid; idx < (g.mixed == 2 ? 3 : 2) * C * D; idx += blockDim.x) {
      const int part = idx / (C * D), r = idx % (C * D);
      float* dst = part == 0 ? qb : (part == 1 ? kb : vb);
      dst[(r / D) * DP + r % D] = __ldg(src + idx);
    }
  }
  if constexpr (D == 64) if (coop && gen) {
    // Cooperative path for the pooled 2-D proposals (DeiT): the warp-per-landmark loop below serialises global-load latency
    // (4 dependent-looking token loads per landmark) and a 64-shuffle Linear per landmark.
    //   1. pooling: one work item per (side, landmark, 8 features), all threads, 16-byte loads
    //   2. Linear:  thread = (output feature, row group); its weight row lives in registers, the means are broadcast reads
    //   3. LayerNorm: one warp per row, in place
    // [3][C][D] means (16-byte aligned for float4 reads); kb2 | om | Wt | rowbuf are dead until the mixing step
    float* xm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(kb2) + 15) & ~(uintptr_t)15);
    const int n_sides = g.mixed == 2 ? 3 : 2;
    for (int idx = tid; idx < n_sides * C * (D / 8); idx += blockDim.x) {
      const int part = idx % (D / 8), c = (idx / (D / 8)) % C, side = idx / (C * (D / 8));
      const View& src = side == 0 ? q : (side == 1 ? k : v);
      int y0, y1, x0, x1;
      bin_of(c / g.side, g.gh, g.side, y0, y1);
      bin_of(c % g.side, g.gw, g.side, x0, x1);
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
          const int tok = y * g.gw + x;
          if (g.zero_padded && mask && mask[(long long)b * g.N + tok]) continue;
          float f[8];
          load8<T>(src.row<T>(b, tok, h) + part * 8, f);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += f[i];
        }
      const float inv = 1.0f / (float)((y1 - y0) * (x1 - x0));
#pragma unroll
      for (int i = 0; i < 8; ++i) xm[(side * C + c) * D + part * 8 + i] = acc[i] * inv;
    }
    __syncthreads();
    if (g.mixed == 2)
      for (int idx = tid; idx < C * D; idx += blockDim.x) vb[(idx / D) * DP + idx % D] = xm[2 * C * D + idx];
    for (int side = 0; side < 2; ++side) {
      float* dst = side == 0 ? qb : kb;
      const float* xs = xm + side * C * D;
      if (has_proj) {
        const int o = tid & (D - 1), grp = tid / D, n_grp = blockDim.x / D;
        const float* W = (side == 0 ? proj.w_q : proj.w_k) + o * D;
        const float* bias = side == 0 ? proj.b_q : proj.b_k;
        float wrow[D];
#pragma unroll
        for (int i4 = 0; i4 < D / 4; ++i4) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(W) + i4);
          wrow[4 * i4] = w4.x; wrow[4 * i4 + 1] = w4.y; wrow[4 * i4 + 2] = w4.z; wrow[4 * i4 + 3] = w4.w;
        }
        const float b0 = bias ? __ldg(bias + o) : 0.f;
        for (int c = grp; c < C; c += n_grp) {
          const float4* xr = reinterpret_cast<const float4*>(xs + c * D);
          float a0 = b0, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int i4 = 0; i4 < D / 4; ++i4) {
            const float4 x4 = xr[i4];
            a0 = fmaf(wrow[4 * i4], x4.x, a0); a1 = fmaf(wrow[4 * i4 + 1], x4.y, a1);
            a2 = fmaf(wrow[4 * i4 + 2], x4.z, a2); a3 = fmaf(wrow[4 * i4 + 3], x4.w, a3);
          }
          dst[c * DP + o] = (a0 + a1) + (a2 + a3);
        }
      } else {
        for (int idx = tid; idx < C * D; idx += blockDim.x) dst[(idx / D) * DP + idx % D] = xs[idx];
      }
    }
    __syncthreads();
    if (has_proj) {
      for (int r = warp; r < 2 * C; r += 8) {
        const int side = r / C, c = r % C;
        const float* gain = side == 0 ? proj.ln_gain_q : proj.ln_gain_k;
        if (!gain) continue;
        const float* lb = side == 0 ? proj.ln_bias_q : proj.ln_bias_k;
        float* row = (side == 0 ? qb : kb) + c * DP;
        const float y0v = row[lane], y1v = row[lane + 32];
        const float mean = warp_sum(y0v + y1v) * (1.0f / D);
        const float d0 = y0v - mean, d1 = y1v - mean;
        const float inv = 1.0f / sqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.0f / D) + proj.ln_eps);
        row[lane] = d0 * inv * __ldg(gain + lane) + __ldg(lb + lane);
        row[lane + 32] = d1 * inv * __ldg(gain + lane + 32) + __ldg(lb + lane + 32);
      }
    }
  }
  if (!coop && gen)
  for (int side = 0; side < 3; ++side) {  // 0: q, 1: k, 2: v (only for '-vmixed')
    if (side == 2 && g.mixed != 2) break;
    const bool per_tok = g.per_token_proj && side < 2;
    if (has_proj && side < 2) {
      __syncthreads();
      const float* W = side == 0 ? proj.w_q : proj.w_k;
      for (int idx = tid; idx < D * D; idx += blockDim.x) Wt[(idx % D) * D + idx / D] = __ldg(W + idx);
      __syncthreads();
    }
    const View& src = side == 0 ? q : (side == 1 ? k : v);
    float* dst = side == 0 ? qb : (side == 1 ? kb : vb);
    for (int c = warp; c < C; c += 8) {
      float acc[DPL];
#pragma unroll
      for (int i = 0; i < DPL; ++i) acc[i] = 0.f;
      int y0 = 0, y1 = 1, x0, x1;
      if (g.dims == 2) {
        bin_of(c / g.side, g.gh, g.side, y0, y1);
        bin_of(c % g.side, g.gw, g.side, x0, x1);
      } else if (g.N <= C) {
        x0 = c; x1 = c + 1;
      } else if (g.N % C == 0) {
        x0 = c * (g.N / C); x1 = x0 + g.N / C;
      } else {
        const int seg = g.N / C, n_short = (seg + 1) * C - g.N;  // lara.py:111-124
        if (c < n_short) { x0 = c * seg; x1 = x0 + seg; }
        else { x0 = n_short * seg + (c - n_short) * (seg + 1); x1 = x0 + seg + 1; }
      }
      const int width = g.dims == 2 ? g.gw : 0;
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
          const int tok = y * width + x;
          float xv[DPL];
          const bool zero = g.zero_padded && mask && mask[(long long)b * g.N + tok];
          const T* r = src.row<T>(b, tok, h);
#pragma unroll
          for (int i = 0; i < DPL; ++i) xv[i] = (zero || !Feat<D>::has(lane, i)) ? 0.f : to_f32(r[lane + 32 * i]);
          if (per_tok) {
            float yv[DPL];
            warp_linear_ln<D>(Wt, proj, side == 0, xv, yv, lane);
#pragma unroll
            for (int i = 0; i < DPL; ++i) acc[i] += yv[i];
          } else {
#pragma unroll
            for (int i = 0; i < DPL; ++i) acc[i] += xv[i];
          }
        }
      const float inv = 1.0f / (float)((y1 - y0) * (x1 - x0));
      float mean[DPL];
#pragma unroll
      for (int i = 0; i < DPL; ++i) mean[i] = acc[i] * inv;
      if (has_proj && !g.per_token_proj && side < 2 && g.dims == 2) {
        float yv[DPL];
        warp_linear_ln<D>(Wt, proj, side == 0, mean, yv, lane);
#pragma unroll
        for (int i = 0; i < DPL; ++i) mean[i] = yv[i];
      }
#pragma unroll
      for (int i = 0; i < DPL; ++i)
        if (Feat<D>::has(lane, i)) dst[c * DP + lane + 32 * i] = mean[i];
    }
  }
  __syncthreads();

  // landmark mixing: k_bar <- softmax(scale k_bar k_bar^T [+ log|v_bar|]) k_bar  (lara.py:157-174)
  float* kfin = kb;
  bool tiled = false;                                  // register-tiled mixing / proposal statistics (cooperative path only)
  if constexpr (D == 64) tiled = coop && g.mixed != 2 && S == C;
  float* const Yal = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(Wt) + 15) & ~(uintptr_t)15);   // [C][D], Wt is unused on this path
  float* const Ms = vb;                                // [C][C] (vb only holds data for '-vmixed')
  if (g.mixed && tiled) {
    if constexpr (D == 64) {
      for (int idx = tid; idx < C * D; idx += blockDim.x) Yal[idx] = kb[(idx / D) * DP + idx % D];
      __syncthreads();
      dots_rows_cols<D>(kb, C, Yal, C, Ms, tid, blockDim.x);
      __syncthreads();
      for (int pr_ = warp; pr_ < C; pr_ += 8) {          // row softmax, in place
        float mx = kNegInf;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, Ms[pr_ * C + c] * scale);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int c = lane; c < C; c += 32) { const float e_ = exp_nonpos(Ms[pr_ * C + c] * scale - mx); Ms[pr_ * C + c] = e_; sum += e_; }
        const float inv = 1.0f / warp_sum(sum);
        for (int c = lane; c < C; c += 32) Ms[pr_ * C + c] *= inv;
      }
      __syncthreads();
      for (int w = tid; w < C * (D / 16); w += blockDim.x) {   // k_bar2[p][16 f .. 16 f + 15] = sum_c P[p][c] k_bar[c][..]
        const int pp = w % C, fb = w / C;
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        for (int c = 0; c < C; ++c) {
          const float pw = Ms[pp * C + c];
          const float4* y = reinterpret_cast<const float4*>(Yal + c * D + 16 * fb);
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const float4 y4 = y[i4];
            acc[4 * i4] = fmaf(pw, y4.x, acc[4 * i4]); acc[4 * i4 + 1] = fmaf(pw, y4.y, acc[4 * i4 + 1]);
            acc[4 * i4 + 2] = fmaf(pw, y4.z, acc[4 * i4 + 2]); acc[4 * i4 + 3] = fmaf(pw, y4.w, acc[4 * i4 + 3]);
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) kb2[pp * DP + 16 * fb + i] = acc[i];
      }
      kfin = kb2;
      __syncthreads();
    }
  } else if (g.mixed) {
    float* pr = rowbuf + warp * C;
    for (int p = warp; p < C; p += 8) {
      float mx = kNegInf;
      for (int c = lane; c < C; c += 32) {
        float s = 0.f;
        for (int e = 0; e < D; ++e) s = fmaf(kb[p * DP + e], kb[c * DP + e], s);
        s *= scale;
        if (g.mixed == 2) {
          float n2 = 0.f;
          for (int e = 0; e < D; ++e) n2 = fmaf(vb[c * DP + e], vb[c * DP + e], n2);
          s += logf(sqrtf(n2) + 1e-4f);
        }
        pr[c] = s;
        mx = fmaxf(mx, s);
      }
      mx = warp_max(mx);
      float sum = 0.f;
      for (int c = lane; c < C; c += 32) { const float e_ = exp_nonpos(pr[c] - mx); pr[c] = e_; sum += e_; }
      sum = warp_sum(sum);
      __syncwarp();
      const float inv = 1.0f / sum;
#pragma unroll
      for (int i = 0; i < DPL; ++i) {
        if (!Feat<D>::has(lane, i)) continue;
        float a = 0.f;
        for (int c = 0; c < C; ++c) a = fmaf(pr[c], kb[c * DP + lane + 32 * i], a);
        kb2[p * DP + lane + 32 * i] = a * inv;
      }
      __syncwarp();
    }
    kfin = kb2;
    __syncthreads();
  }

  // q_bar -> workspace; mu = q_bar + k_bar (lara.py:182); omega = proposal samples (lara.py:188-198)
  for (int idx = tid; idx < C * D; idx += blockDim.x) {
    const int c = idx / D, e = idx % D;
    const float qv = qb[c * DP + e];
    const float m = qv + kfin[c * DP + e];
    ws.qbar[idx] = qv;
    ws.mu[idx] = m;
    qb[c * DP + e] = m;  // qb now holds mu
  }
  __syncthreads();
  const float* nz = noise ? noise + (long long)blockIdx.x * (g.sample_mode == LARA_SAMPLE_ANTITHETIC ? C : S) * D : nullptr;
  for (int idx = tid; idx < S * D; idx += blockDim.x) {
    const int s = idx / D, e = idx % D;
    float o = qb[(s % C) * DP + e];
    if (nz) {
      if (g.sample_mode == LARA_SAMPLE_ANTITHETIC) o += (s < C ? 1.f : -1.f) * __ldg(nz + (s % C) * D + e);
      else o += __ldg(nz + idx);
    }
    om[s * DP + e] = o;
    ws.omega[idx] = o;
  }
  __syncthreads();

  // proposal statistics over Lm[s][c] = prm(mu_c, omega_s)   (lara.py:215,228-236)
  const float log_rep = logf((float)(S / C));
  if (tiled) {
    if constexpr (D == 64) {
      for (int idx = tid; idx < C * D; idx += blockDim.x) Yal[idx] = qb[(idx / D) * DP + idx % D];      // mu
      __syncthreads();
      for (int c = tid; c < C; c += blockDim.x) {
        float n2 = 0.f;
        for (int e = 0; e < D; ++e) n2 = fmaf(Yal[c * D + e], Yal[c * D + e], n2);
        rowbuf[c] = n2;
      }
      dots_rows_cols<D>(om, S, Yal, C, Ms, tid, blockDim.x);
      __syncthreads();
      for (int s_ = warp; s_ < S; s_ += 8) {
        float mx = kNegInf;
        for (int c = lane; c < C; c += 32) {
          const float lv = scale * (Ms[s_ * C + c] - 0.5f * rowbuf[c]);
          Ms[s_ * C + c] = lv;
          mx = fmaxf(mx, lv);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int c = lane; c < C; c += 32) sum += exp_nonpos(Ms[s_ * C + c] - mx);
        const float lse = mx + logf(warp_sum(sum));
        __syncwarp();
        if (lane == 0) {
          if (g.mis_type == LARA_MIS_OPT) {
            const float lp = Ms[s_ * C + s_ % C];
            ws.lp[s_] = lp;
            ws.bh[s_] = expf(lp - (lse + log_rep));
          } else {
            ws.lp[s_] = lse;
            ws.bh[s_] = 0.f;
          }
        }
      }
    }
    return;
  }
  float* pr = rowbuf + warp * C;
  for (int s = warp; s < S; s += 8) {
    float mx = kNegInf;
    for (int c = lane; c < C; c += 32) {
      float dot = 0.f, n2 = 0.f;
      for (int e = 0; e < D; ++e) {
        const float m = qb[c * DP + e];
        dot = fmaf(om[s * DP + e], m, dot);
        n2 = fmaf(m, m, n2);
      }
      const float lv = scale * (dot - 0.5f * n2);
      pr[c] = lv;
      mx = fmaxf(mx, lv);
    }
    mx = warp_max(mx);
    __syncwarp();
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += exp_nonpos(pr[c] - mx);
    const float lse = mx + logf(warp_sum(sum));
    if (lane == 0) {
      if (g.mis_type == LARA_MIS_OPT) {
        const float lp = pr[s % C];
        ws.lp[s] = lp;
        ws.bh[s] = expf(lp - (lse + log_rep));
      } else {
        ws.lp[s] = lse;
        ws.bh[s] = 0.f;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// L2: rows x all tokens with online softmax.  blockIdx.x < kv_blocks: rows = omega samples, keys = k,
// values = v  -> kv[s], lse_k[s].  Otherwise (mis-opt): rows = q_bar landmarks, keys = q -> lse_t[c].
// ------------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(128)
lara_stats_kernel(const LaraGeo g, const View q, const View k, const View v, const uint8_t* __restrict__ mask,
                  float* __restrict__ ws_base, const int kv_blocks) {
  constexpr int DPL = Feat<D>::kPerLane, DP = D + 1;
  constexpr int ROWS = 16, RPW = 4, KT = 32;
  extern __shared__ float sm[];
  float* Rs = sm;                 // [ROWS][D]
  float* Ks = Rs + ROWS * D;      // [KT][DP]
  float* Vs = Ks + KT * DP;       // [KT][DP]
  float* Ps = Vs + KT * DP;       // [4][RPW][KT]
  int* kflag = reinterpret_cast<int*>(Ps + 4 * RPW * KT);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // grid = (batch * head, row blocks): gridDim.x has no 65535 cap
  const int b = blockIdx.x / g.H, h = blockIdx.x % g.H;
  const LaraWs ws = lara_ws_at(ws_base, blockIdx.x, g.C, g.S, D);
  const bool kv_mode = (int)blockIdx.y < kv_blocks;
  const int rb = kv_mode ? blockIdx.y : blockIdx.y - kv_blocks;
  const int n_rows = kv_mode ? g.S : g.C;
  const float* rows = kv_mode ? ws.omega : ws.qbar;
  const View& keys = kv_mode ? k : q;
  const float scale = rsqrtf((float)D);

  for (int idx = tid; idx < ROWS * D; idx += blockDim.x) {
    const int r = idx / D, row = rb * ROWS + r;
    Rs[idx] = row < n_rows ? rows[(long long)row * D + idx % D] : 0.f;
  }
  float m[RPW], l[RPW], o[RPW][DPL];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    m[r] = kNegInf; l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) o[r][i] = 0.f;
  }
  for (int kt0 = 0; kt0 < g.N; kt0 += KT) {
    __syncthreads();
    for (int idx = tid; idx < KT * (D / 8); idx += blockDim.x) {
      const int j = idx / (D / 8), part = idx % (D / 8);
      const int tok = kt0 + j;
      float fk[8], fv[8];
      int flag = 2;
      bool have = false;
      if (tok < g.N) {
        const bool padded = mask && mask[(long long)b * g.N + tok];
        flag = (padded && kv_mode) ? 1 : 0;
        if (!(padded && g.zero_padded)) {
          load8<T>(keys.row<T>(b, tok, h) + part * 8, fk);
          if (kv_mode) load8<T>(v.row<T>(b, tok, h) + part * 8, fv);
          have = true;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        Ks[j * DP + part * 8 + i] = have ? fk[i] : 0.f;
        Vs[j * DP + part * 8 + i] = (have && kv_mode) ? fv[i] : 0.f;
      }
      if (part == 0) kflag[j] = flag;
    }
    __syncthreads();
    float s[RPW], n2 = 0.f;
#pragma unroll
    for (int r = 0; r < RPW; ++r) s[r] = 0.f;
    const float* krow = Ks + lane * DP;
    const float* rrow = Rs + (warp * RPW) * D;
#pragma unroll 8
    for (int e = 0; e < D; ++e) {
      const float kk = krow[e];
      n2 = fmaf(kk, kk, n2);
#pragma unroll
      for (int r = 0; r < RPW; ++r) s[r] = fmaf(rrow[r * D + e], kk, s[r]);
    }
    const int flag = kflag[lane];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      float sv = kv_mode ? scale * (s[r] - 0.5f * n2) : scale * s[r];
      if (flag != 0) sv = kNegInf;  // absent, or padded key of phi(k) (lara.py:204-208)
      const float mt = warp_max(sv);
      const float mn = fmaxf(m[r], mt);
      float corr = 1.f, p = 0.f;
      if (mn != kNegInf) { corr = exp_nonpos(m[r] - mn); p = exp_nonpos(sv - mn); }
      l[r] = fmaf(l[r], corr, warp_sum(p));
      m[r] = mn;
#pragma unroll
      for (int i = 0; i < DPL; ++i) o[r][i] *= corr;
      Ps[(warp * RPW + r) * KT + lane] = p;
    }
    __syncwarp();
    if (kv_mode) {
      const float* prow = Ps + (warp * RPW) * KT;
#pragma unroll 4
      for (int j = 0; j < KT; ++j) {
        float vv[DPL];
#pragma unroll
        for (int i = 0; i < DPL; ++i) vv[i] = Feat<D>::has(lane, i) ? Vs[j * DP + lane + 32 * i] : 0.f;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          const float p = prow[r * KT + j];
#pragma unroll
          for (int i = 0; i < DPL; ++i) o[r][i] = fmaf(p, vv[i], o[r][i]);
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int row = rb * ROWS + warp * RPW + r;
    if (row >= n_rows) continue;
    const float lse = m[r] + logf(l[r]);
    if (kv_mode) {
      const float inv = 1.0f / l[r];
#pragma unroll
      for (int i = 0; i < DPL; ++i)
        if (Feat<D>::has(lane, i)) ws.kv[(long long)row * D + lane + 32 * i] = o[r][i] * inv;
      if (lane == 0) ws.lse_k[row] = lse;
    } else if (lane == 0) {
      ws.lse_t[row] = lse;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// L3: per token.  CTA stages omega, q_bar|mu, kv and the per-sample constants; one warp per token.
// ------------------------------------------------------------------------------------------------
constexpr int kLaraTokPerCta = 64;

template <typename T, int D>
__global__ void __launch_bounds__(256)
lara_out_kernel(const LaraGeo g, const View q, const uint8_t* __restrict__ mask, const float* __restrict__ ws_base,
                T* __restrict__ out) {
  constexpr int DPL = Feat<D>::kPerLane, DP = D + 1;
  extern __shared__ float sm[];
  const int C = g.C, S = g.S;
  float* om = sm;               // [S][DP]
  float* aux = om + S * DP;     // [C][DP]  q_bar (mis-opt) or mu (mis-biased)
  float* kvs = aux + C * DP;    // [S][DP]
  float* cst = kvs + S * DP;    // [S]  lse_k - lp
  float* bhs = cst + S;         // [S]
  float* lset = bhs + S;        // [C]
  float* qv = lset + C;         // [8][D]
  float* wv = qv + 8 * D;       // [8][S]
  float* tb = wv + 8 * S;       // [8][C]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / g.H, h = blockIdx.x % g.H;      // grid = (batch * head, token blocks)
  const LaraWs ws = lara_ws_at(const_cast<float*>(ws_base), blockIdx.x, C, S, D);
  const float scale = rsqrtf((float)D);
  const float* auxsrc = g.mis_type == LARA_MIS_OPT ? ws.qbar : ws.mu;
  for (int idx = tid; idx < S * D; idx += blockDim.x) {
    om[(idx / D) * DP + idx % D] = ws.omega[idx];
    kvs[(idx / D) * DP + idx % D] = ws.kv[idx];
  }
  for (int idx = tid; idx < C * D; idx += blockDim.x) aux[(idx / D) * DP + idx % D] = auxsrc[idx];
  for (int s = tid; s < S; s += blockDim.x) { cst[s] = ws.lse_k[s] - ws.lp[s]; bhs[s] = ws.bh[s]; }
  for (int c = tid; c < C; c += blockDim.x) lset[c] = ws.lse_t[c];
  __syncthreads();

  float* myq = qv + warp * D;
  float* myw = wv + warp * S;
  float* myt = tb + warp * C;
  const int t_end = min(g.N, (int)(blockIdx.y + 1) * kLaraTokPerCta);
  for (int tok = blockIdx.y * kLaraTokPerCta + warp; tok < t_end; tok += 8) {
    const bool zero = g.zero_padded && mask && mask[(long long)b * g.N + tok];
    const T* qr = q.row<T>(b, tok, h);
    float n2 = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      if (!Feat<D>::has(lane, i)) continue;
      const float x = zero ? 0.f : to_f32(qr[lane + 32 * i]);
      myq[lane + 32 * i] = x;
      n2 = fmaf(x, x, n2);
    }
    n2 = warp_sum(n2);
    __syncwarp();
    float mean_t = 0.f;
    if (g.mis_type == LARA_MIS_OPT) {
      float acc = 0.f;
      for (int c = lane; c < C; c += 32) {
        float d_ = 0.f;
        for (int e = 0; e < D; ++e) d_ = fmaf(aux[c * DP + e], myq[e], d_);
        const float t = expf(scale * d_ - lset[c]);
        myt[c] = t;
        acc += t;
      }
      mean_t = warp_sum(acc) / (float)C;
      __syncwarp();
    }
    float mx = kNegInf;
    for (int s = lane; s < S; s += 32) {
      float d_ = 0.f;
      for (int e = 0; e < D; ++e) d_ = fmaf(om[s * DP + e], myq[e], d_);
      const float A = scale * (d_ - 0.5f * n2);
      float log_alpha = 0.f;
      if (g.mis_type == LARA_MIS_OPT) {
        const float alpha = bhs[s] + g.alpha_coeff * (myt[s % C] - mean_t);
        log_alpha = logf(fmaxf(alpha, 1e-8f));
      } else if (g.mis_type == LARA_MIS_BIASED) {
        float dm = 0.f;
        for (int e = 0; e < D; ++e) dm = fmaf(aux[(s % C) * DP + e], myq[e], dm);
        log_alpha = scale * dm;
      }
      const float lw = log_alpha + A + cst[s];
      myw[s] = lw;
      mx = fmaxf(mx, lw);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s = lane; s < S; s += 32) { const float e_ = exp_nonpos(myw[s] - mx); myw[s] = e_; sum += e_; }
    const float inv = 1.0f / warp_sum(sum);
    __syncwarp();
    float o[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) o[i] = 0.f;
    for (int s = 0; s < S; ++s) {
      const float w = myw[s];
#pragma unroll
      for (int i = 0; i < DPL; ++i)
        if (Feat<D>::has(lane, i)) o[i] = fmaf(w, kvs[s * DP + lane + 32 * i], o[i]);
    }
    T* orow = out + ((long long)b * g.N + tok) * ((long long)g.H * D) + (long long)h * D;
#pragma unroll
    for (int i = 0; i < DPL; ++i)
      if (Feat<D>::has(lane, i)) orow[lane + 32 * i] = from_f32<T>(o[i] * inv);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
static size_t landmark_smem(const LaraGeo& g) {
  const int DP = g.D + 1;
  return ((size_t)(4 * g.C + g.S) * DP + (size_t)g.D * g.D + 8 * (size_t)g.C) * sizeof(float);
}
static size_t out_smem(const LaraGeo& g) {
  const int DP = g.D + 1;
  return ((size_t)(2 * g.S + g.C) * DP + 2 * (size_t)g.S + g.C + 8 * (size_t)(g.D + g.S + g.C)) * sizeof(float);
}

size_t lara_workspace_bytes(const LaraGeo& g) {
  const size_t per = lara_ws_floats_per_bh(g.C, g.S, g.D);
  return (((size_t)g.B * g.H * per * sizeof(float)) + 255) & ~(size_t)255;
}

template <typename T, int D>
static cudaError_t launch_lara_t(const LaraGeo& g, const View& q, const View& k, const View& v, const uint8_t* mask,
                                 const EvaAdaptive& proj, const float* noise, void* out, void* workspace,
                                 cudaStream_t st, const float* given) {
  float* ws = reinterpret_cast<float*>(workspace);
  const size_t sm1 = landmark_smem(g), sm3 = out_smem(g);
  if (sm1 > 227 * 1024 || sm3 > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e;
  if constexpr (D == 64 && !std::is_same<T, float>::value) {
    constexpr int io = std::is_same<T, __half>::value ? EVA_F16 : EVA_BF16;
    if (!given && lara_core_supported(g, io, q, k, v, mask) && lara_core_fuses_landmarks(g, proj))
      return launch_lara_core(g, io, q, k, v, ws, out, &proj, noise, st);      // landmarks, statistics and output in one kernel
  }
  {
    auto kern = lara_landmark_kernel<T, D>;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1)) != cudaSuccess) return e;
    kern<<<g.B * g.H, 256, sm1, st>>>(g, q, k, v, mask, proj, noise, ws, given);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  if constexpr (D == 64 && !std::is_same<T, float>::value) {
    constexpr int io = std::is_same<T, __half>::value ? EVA_F16 : EVA_BF16;
    if (lara_core_supported(g, io, q, k, v, mask)) return launch_lara_core(g, io, q, k, v, ws, out, nullptr, nullptr, st);
  }
  {
    auto kern = lara_stats_kernel<T, D>;
    const size_t sm2 = (size_t)(16 * D + 2 * 32 * (D + 1) + 4 * 4 * 32) * sizeof(float) + 32 * sizeof(int);
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2)) != cudaSuccess) return e;
    const int kv_blocks = (g.S + 15) / 16;
    const int t_blocks = g.mis_type == LARA_MIS_OPT ? (g.C + 15) / 16 : 0;
    dim3 grid(g.B * g.H, kv_blocks + t_blocks);
    kern<<<grid, 128, sm2, st>>>(g, q, k, v, mask, ws, kv_blocks);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  {
    auto kern = lara_out_kernel<T, D>;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3)) != cudaSuccess) return e;
    dim3 grid(g.B * g.H, (g.N + kLaraTokPerCta - 1) / kLaraTokPerCta);
    kern<<<grid, 256, sm3, st>>>(g, q, mask, ws, reinterpret_cast<T*>(out));
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_lara(const LaraGeo& g, int io_dtype, const View& q, const View& k, const View& v,
                        const uint8_t* mask, const EvaAdaptive& proj, const float* noise, void* out,
                        void* workspace, cudaStream_t st, const float* given) {
  switch (io_dtype * 256 + g.D) {
    case EVA_F32 * 256 + 16: return launch_lara_t<float, 16>(g, q, k, v, mask, proj, noise, out, workspace, st, given);
    case EVA_F32 * 256 + 32: return launch_lara_t<float, 32>(g, q, k, v, mask, proj, noise, out, workspace, st, given);
    case EVA_F32 * 256 + 64: return launch_lara_t<float, 64>(g, q, k, v, mask, proj, noise, out, workspace, st, given);
    case EVA_F32 * 256 + 128: return launch_lara_t<float, 128>(g, q, k, v, mask, proj, noise, out, workspace, st, given);
    case EVA_F16 * 256 + 16: return launch_lara_t<__half, 16>(g, q, k, v, mask, proj, noise, out, workspace, st, given);
    case EVA_F16 * 256 + 32: return launch_lara_t<__half, 32>(g, q, k, v, mask, proj, noise, out, workspace, st, given);
    case EVA_F16 * 256 + 64: return launch_lara_t<__half, 64>(g, q, k, v, mask, proj, noise, out, workspace, st, given);
    case EVA_F16 * 256 + 128: return launch_lara_t<__half, 128>(g, q, k, v, mask, proj, noise, out, workspace, st, given);
    case EVA_BF16 * 256 + 16: return launch_lara_t<__nv_bfloat16, 16>(g, q, k, v, mask, proj, noise, out, workspace, st, given);
    case EVA_BF16 * 256 + 32: return launch_lara_t<__nv_bfloat16, 32>(g, q, k, v, mask, proj, noise, out, workspace, st, given);
    case EVA_BF16 * 256 + 64: return launch_lara_t<__nv_bfloat16, 64>(g, q, k, v, mask, proj, noise, out, workspace, st, given);
    case EVA_BF16 * 256 + 128: return launch_lara_t<__nv_bfloat16, 128>(g, q, k, v, mask, proj, noise, out, workspace, st, given);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace eva
