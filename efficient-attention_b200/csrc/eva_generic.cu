// Generic (any geometry) EVA kernels: fp32 math on CUDA cores, T only the HBM format.
// These cover every configuration the reference accepts (1-D / 2-D, halos, padding masks, causal,
// chunk-less local / dense attention, head_dim 32/64/128); the fused tcgen05/TMA kernel in
// eva_fused_sm100.cu takes over for the geometries it is specialised for.
//
//   chunk_stats_kernel    eva.py:155-196, causal_eva.py:676-719
//   window_attn_kernel    eva.py:200-227, causal_eva.py:722-783, local_attention.py:134-182,
//                         abstract_attention.py:115-133
#include <stdlib.h>

#include "common.cuh"
#include "launch.h"

namespace eva {

// ------------------------------------------------------------------------------------------------
// Stage A: one warp per (batch, head, chunk).  Lane l owns features l, l+32, ...
// ------------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(256)
chunk_stats_kernel(const Geo g, const View q, const View k, const View v, const uint8_t* __restrict__ mask,
                   const EvaAdaptive ada, const float* __restrict__ noise, float* __restrict__ kbar_out,
                   float* __restrict__ beta_out) {
  constexpr int DPL = Feat<D>::kPerLane;
  extern __shared__ float sm[];
  float* WtK = sm;
  float* WtQ = sm + D * D;
  for (int idx = threadIdx.x; idx < D * D; idx += blockDim.x) {
    const int e = idx / D, i = idx % D;  // W[e][i] -> Wt[i][e]
    WtK[i * D + e] = __ldg(ada.w_k + idx);
    if (ada.w_q) WtQ[i * D + e] = __ldg(ada.w_q + idx);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const float scale = rsqrtf((float)D);
  const float inv_cnt = 1.0f / (float)g.Jc;
  const long long total = (long long)g.B * g.H * g.n_chunks;
  for (long long wg = (long long)blockIdx.x * wpb + warp; wg < total; wg += (long long)gridDim.x * wpb) {
    const int c = (int)(wg % g.n_chunks);
    const int h = (int)((wg / g.n_chunks) % g.H);
    const int b = (int)(wg / ((long long)g.n_chunks * g.H));
    // pass 1: chunk means; padded / off-sequence slots count as zeros in the denominator (eva.py:174-180)
    float sq[DPL], sk[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) sq[i] = sk[i] = 0.f;
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
    float kb[DPL], om[DPL];
    warp_linear<D>(WtK, ada.b_k, sk, kb, lane);
    if (ada.ln_gain_k) warp_layer_norm<D>(kb, ada.ln_gain_k, ada.ln_bias_k, ada.ln_eps, lane);
    if (ada.w_q) {
      float qb[DPL];
      warp_linear<D>(WtQ, ada.b_q, sq, qb, lane);
      if (ada.ln_gain_q) warp_layer_norm<D>(qb, ada.ln_gain_q, ada.ln_bias_q, ada.ln_eps, lane);
#pragma unroll
      for (int i = 0; i < DPL; ++i) om[i] = ada.mu_coeff * (qb[i] + kb[i]);
    } else {
#pragma unroll
      for (int i = 0; i < DPL; ++i) om[i] = 0.f;
    }
    const long long obase = wg * D;
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      if (!Feat<D>::has(lane, i)) continue;
      if (noise) om[i] += __ldg(noise + obase + lane + 32 * i);
      kbar_out[obase + lane + 32 * i] = kb[i];
    }
    // pass 2: beta = softmax_j(prm(k_j, omega)) . v_j, online over the chunk's slots (eva.py:192-196)
    float m = kNegInf, l = 0.f, acc[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) acc[i] = 0.f;
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
#pragma unroll
    for (int i = 0; i < DPL; ++i)
      if (Feat<D>::has(lane, i)) beta_out[obase + lane + 32 * i] = acc[i] * inv_l;
  }
}

// ------------------------------------------------------------------------------------------------
// Stage A for long 1-D chunks (causal LM: 256 tokens per chunk): one CTA of 8 warps per (batch, head, chunk) instead
// of one warp -- the warp kernel above walks the chunk token by token with a shuffle reduction per logit.
//   phase 1  means: warp w sums tokens w, w+8, ...; lane = feature pair (coalesced 128-byte rows)
//   phase 2  Linear + LayerNorm: thread t < 64 -> k side feature t, 64 <= t < 128 -> q side; omega -> shared memory
//   phase 3  logits: thread = token (its 128-byte k row from global / L2), block softmax
//   phase 4  beta: warp w accumulates p_s v_s over its tokens, lane = feature pair; cross-warp sum
// head_dim 64, no halo, no padding mask (the warp kernel keeps those cases).  Semantics as chunk_stats_kernel.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
chunk_stats_cta_kernel(const Geo g, const View q, const View k, const View v, const EvaAdaptive ada,
                       const float* __restrict__ noise, float* __restrict__ kbar_out, float* __restrict__ beta_out) {
  constexpr int D = 64;
  extern __shared__ float sm[];
  float* part = sm;                 // [8][128]  per-warp partial sums (q | k), later beta partials [8][64]
  float* mean = part + 8 * 128;     // [128]     q means | k means
  float* yv = mean + 128;           // [128]     Linear outputs, q side | k side
  float* om = yv + 128;             // [64]      omega
  float* red = om + 64;             // [16]      block reductions
  float* pj = red + 16;             // [Jc]      softmax weights
  // the chunk's k rows staged by phase 1 for the per-token logits of phase 3 (row stride 144 B: conflict-free 16-byte reads)
  constexpr int kKs = 144;
  uint8_t* kst = reinterpret_cast<uint8_t*>(pj + ((g.Jc + 3) & ~3));
  const bool stage_k = sizeof(T) == 2 && g.Jc <= 512;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long wg = blockIdx.x;
  const int c = (int)(wg % g.n_chunks);
  const int h = (int)((wg / g.n_chunks) % g.H);
  const int b = (int)(wg / ((long long)g.n_chunks * g.H));
  const int t0 = c * g.chunk;
  const float scale = 0.125f;
  // ---- phase 1 ----
  if constexpr (sizeof(T) == 2) {
    // 16-byte loads: 8 lanes cover one 128-byte row, a warp covers 4 tokens per instruction (4x the bytes in flight of the
    // 4-byte-per-lane version); lane = (token sub-index ts, 16-byte piece p8)
    const int ts = lane >> 3, p8 = lane & 7;
    float sq[8], sk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sq[i] = 0.f; sk[i] = 0.f; }
#pragma unroll 2
    for (int s = 4 * warp + ts; s < g.Jc; s += 32) {
      const uint4 rq = __ldg(reinterpret_cast<const uint4*>(q.row<T>(b, t0 + s, h)) + p8);
      const uint4 rk = __ldg(reinterpret_cast<const uint4*>(k.row<T>(b, t0 + s, h)) + p8);
      const T* eq = reinterpret_cast<const T*>(&rq);
      const T* ek = reinterpret_cast<const T*>(&rk);
#pragma unroll
      for (int i = 0; i < 8; ++i) { sq[i] += to_f32(eq[i]); sk[i] += to_f32(ek[i]); }
      if (stage_k) *reinterpret_cast<uint4*>(kst + s * kKs + p8 * 16) = rk;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], 8);  sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], 16);
      sk[i] += __shfl_xor_sync(0xffffffffu, sk[i], 8);  sk[i] += __shfl_xor_sync(0xffffffffu, sk[i], 16);
    }
    if (ts == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { part[warp * 128 + 8 * p8 + i] = sq[i]; part[warp * 128 + 64 + 8 * p8 + i] = sk[i]; }
    }
  } else {
    float sq0 = 0.f, sq1 = 0.f, sk0 = 0.f, sk1 = 0.f;
#pragma unroll 4
    for (int s = warp; s < g.Jc; s += 8) {
      const T* qr = q.row<T>(b, t0 + s, h);
      const T* kr = k.row<T>(b, t0 + s, h);
      sq0 += to_f32(qr[2 * lane]); sq1 += to_f32(qr[2 * lane + 1]);
      sk0 += to_f32(kr[2 * lane]); sk1 += to_f32(kr[2 * lane + 1]);
    }
    part[warp * 128 + 2 * lane] = sq0; part[warp * 128 + 2 * lane + 1] = sq1;
    part[warp * 128 + 64 + 2 * lane] = sk0; part[warp * 128 + 64 + 2 * lane + 1] = sk1;
  }
  __syncthreads();
  if (tid < 128) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += part[w * 128 + tid];
    mean[tid] = a / (float)g.Jc;
  }
  __syncthreads();
  // ---- phase 2 ----
  if (tid < 128) {
    const bool kside = tid < 64;
    const int f = tid & 63;
    const float* W = kside ? ada.w_k : ada.w_q;
    const float* bias = kside ? ada.b_k : ada.b_q;
    const float* mv = mean + (kside ? 64 : 0);
    float y = 0.f;
    if (W) {
      y = bias ? __ldg(bias + f) : 0.f;
      const float4* wr = reinterpret_cast<const float4*>(W + f * D);
#pragma unroll 4
      for (int i4 = 0; i4 < D / 4; ++i4) {
        const float4 w4 = __ldg(wr + i4);
        y = fmaf(w4.x, mv[4 * i4], y); y = fmaf(w4.y, mv[4 * i4 + 1], y);
        y = fmaf(w4.z, mv[4 * i4 + 2], y); y = fmaf(w4.w, mv[4 * i4 + 3], y);
      }
    }
    yv[kside ? 64 + f : f] = y;
  }
  __syncthreads();
  if (tid < 128) {
    const bool kside = tid < 64;
    const int f = tid & 63;
    const float* yy = yv + (kside ? 64 : 0);
    const float* gain = kside ? ada.ln_gain_k : ada.ln_gain_q;
    const float* lb = kside ? ada.ln_bias_k : ada.ln_bias_q;
    float y = yy[f];
    if (gain) {
      float s1 = 0.f;
#pragma unroll 8
      for (int i = 0; i < D; ++i) s1 += yy[i];
      const float mu = s1 * (1.0f / D);
      float s2 = 0.f;
#pragma unroll 8
      for (int i = 0; i < D; ++i) { const float d = yy[i] - mu; s2 = fmaf(d, d, s2); }
      y = (y - mu) * (1.0f / sqrtf(s2 * (1.0f / D) + ada.ln_eps)) * __ldg(gain + f) + __ldg(lb + f);
    }
    if (kside) kbar_out[wg * D + f] = y;
    mean[kside ? 64 + f : f] = y;          // normalised q_bar | k_bar (the means are dead)
  }
  __syncthreads();
  if (tid < 64) {
    float o = ada.w_q ? ada.mu_coeff * (mean[tid] + mean[64 + tid]) : 0.f;
    if (noise) o += __ldg(noise + wg * D + tid);
    om[tid] = o;
  }
  __syncthreads();
  // ---- phase 3 ----
  float mloc = kNegInf;
  for (int s = tid; s < g.Jc; s += 256) {
    const T* kr = stage_k ? reinterpret_cast<const T*>(kst + s * kKs) : k.row<T>(b, t0 + s, h);
    float acc = 0.f;
#pragma unroll
    for (int p8 = 0; p8 < 8; ++p8) {
      float f[8];
      if (stage_k) {                 // staged rows live in shared memory: plain 16-byte read (load8 is ld.global.nc)
        const uint4 raw = *reinterpret_cast<const uint4*>(kr + 8 * p8);
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = to_f32(e[i]);
      } else {
        load8<T>(kr + 8 * p8, f);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(f[i], om[8 * p8 + i] - 0.5f * f[i], acc);
    }
    const float lg = scale * acc;
    pj[s] = lg;
    mloc = fmaxf(mloc, lg);
  }
  mloc = warp_max(mloc);
  if (lane == 0) red[warp] = mloc;
  __syncthreads();
  float mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  float lsum = 0.f;
  for (int s = tid; s < g.Jc; s += 256) {
    const float pe = exp_nonpos(pj[s] - mx);
    pj[s] = pe;
    lsum += pe;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red[8 + warp] = lsum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[8 + w];
  // ---- phase 4 ----
  if constexpr (sizeof(T) == 2) {
    const int ts = lane >> 3, p8 = lane & 7;
    float bacc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bacc[i] = 0.f;
#pragma unroll 2
    for (int s = 4 * warp + ts; s < g.Jc; s += 32) {
      const uint4 rv = __ldg(reinterpret_cast<const uint4*>(v.row<T>(b, t0 + s, h)) + p8);
      const T* ev = reinterpret_cast<const T*>(&rv);
      const float pe = pj[s];
#pragma unroll
      for (int i = 0; i < 8; ++i) bacc[i] = fmaf(pe, to_f32(ev[i]), bacc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      bacc[i] += __shfl_xor_sync(0xffffffffu, bacc[i], 8);
      bacc[i] += __shfl_xor_sync(0xffffffffu, bacc[i], 16);
    }
    if (ts == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) part[warp * 64 + 8 * p8 + i] = bacc[i];
    }
  } else {
    float b0 = 0.f, b1 = 0.f;
    for (int s = warp; s < g.Jc; s += 8) {
      const T* vr = v.row<T>(b, t0 + s, h);
      const float pe = pj[s];
      b0 = fmaf(pe, to_f32(vr[2 * lane]), b0);
      b1 = fmaf(pe, to_f32(vr[2 * lane + 1]), b1);
    }
    part[warp * 64 + 2 * lane] = b0;
    part[warp * 64 + 2 * lane + 1] = b1;
  }
  __syncthreads();
  if (tid < 64) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += part[w * 64 + tid];
    beta_out[wg * D + tid] = a / tot;
  }
}

// ------------------------------------------------------------------------------------------------
// Stage A for head_dim 64 with 16-bit I/O and chunks of at most 128 slots (any geometry: halo, padding, 1-D / 2-D).  Same
// semantics as chunk_stats_kernel with an order of magnitude fewer instructions per chunk: the token index of a slot is computed
// by ONE lane per slot (not by every lane for every slot and pass), vectors are held as feature pairs (4-byte loads of the
// 16-bit rows), and the per-token phi-logits run with lane = token (a 64-term dot product from the token's own 128-byte row
// against omega in shared memory) -- two warp reductions per chunk instead of one per token.
// ------------------------------------------------------------------------------------------------
constexpr int kFastMaxJc = 128;

template <typename T>
__global__ void __launch_bounds__(256)
chunk_stats_fast_kernel(const Geo g, const View q, const View k, const View v, const uint8_t* __restrict__ mask,
                        const EvaAdaptive ada, const float* __restrict__ noise, float* __restrict__ kbar_out,
                        float* __restrict__ beta_out) {
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
  float* omv = sm + 2 * 64 * 64 + warp * 64;          // omega of the warp's chunk
  const float scale = 0.125f;
  const float inv_cnt = 1.0f / (float)g.Jc;
  const int n_rounds = (g.Jc + 31) >> 5;
  const long long total = (long long)g.B * g.H * g.n_chunks;
  for (long long wg = (long long)blockIdx.x * wpb + warp; wg < total; wg += (long long)gridDim.x * wpb) {
    const int c = (int)(wg % g.n_chunks);
    const int h = (int)((wg / g.n_chunks) % g.H);
    const int b = (int)(wg / ((long long)g.n_chunks * g.H));
    const long long obase = wg * 64;
    // token of slot 32 rd + lane: -1 when off the sequence or padded (such slots count as zeros)
    int tokr[4];
#pragma unroll
    for (int rd = 0; rd < 4; ++rd) {
      const int s = 32 * rd + lane;
      int t = (rd < n_rounds && s < g.Jc) ? group_token(g, c, s, g.chunk, g.chunk_ext) : -1;
      if (t >= 0 && mask && mask[(long long)b * g.N + t]) t = -1;
      tokr[rd] = t;
    }
    // ---- chunk means: lane = (token sub-index ts, 16-byte piece p8), four tokens per load instruction ----
    const int ts = lane >> 3, p8 = lane & 7;
    float aq[8], ak[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) aq[i] = ak[i] = 0.f;
#pragma unroll
    for (int rd = 0; rd < 4; ++rd) {
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
    // ---- k_bar, omega ----
    float inv_unused;
    float2 kb = pair_linear(WtK, ada.b_k, sk, lane);
    if (ada.ln_gain_k) {
      const float2 n = pair_ln(kb, ada.ln_eps, inv_unused);
      const float2 gg = __ldg(reinterpret_cast<const float2*>(ada.ln_gain_k) + lane), bb = __ldg(reinterpret_cast<const float2*>(ada.ln_bias_k) + lane);
      kb = make_float2(fmaf(n.x, gg.x, bb.x), fmaf(n.y, gg.y, bb.y));
    }
    float2 om = make_float2(0.f, 0.f);
    if (ada.w_q) {
      float2 qb = pair_linear(WtQ, ada.b_q, sq, lane);
      if (ada.ln_gain_q) {
        const float2 n = pair_ln(qb, ada.ln_eps, inv_unused);
        const float2 gg = __ldg(reinterpret_cast<const float2*>(ada.ln_gain_q) + lane), bb = __ldg(reinterpret_cast<const float2*>(ada.ln_bias_q) + lane);
        qb = make_float2(fmaf(n.x, gg.x, bb.x), fmaf(n.y, gg.y, bb.y));
      }
      om = make_float2(ada.mu_coeff * (qb.x + kb.x), ada.mu_coeff * (qb.y + kb.y));
    }
    if (noise) { const float2 z = __ldg(reinterpret_cast<const float2*>(noise + obase) + lane); om.x += z.x; om.y += z.y; }
    reinterpret_cast<float2*>(kbar_out + obase)[lane] = kb;
    __syncwarp();
    reinterpret_cast<float2*>(omv)[lane] = om;
    __syncwarp();
    // ---- phi-logits (lane = token): scale (omega . k - |k|^2 / 2); padded / off-sequence slots: -5e4 ----
    float lg[4];
    float mx = kNegInf;
#pragma unroll
    for (int rd = 0; rd < 4; ++rd) {
      lg[rd] = kNegInf;
      if (rd >= n_rounds) continue;
      if (32 * rd + lane < g.Jc) {
        lg[rd] = kMaskVal;
        const int t = tokr[rd];
        if (t >= 0) {
          const uint4* kr = reinterpret_cast<const uint4*>(k.row<T>(b, t, h));
          float part = 0.f;
#pragma unroll
          for (int pc = 0; pc < 8; ++pc) {
            const uint4 raw = __ldg(kr + pc);
            const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
            const float4 o0 = *reinterpret_cast<const float4*>(omv + 8 * pc), o1 = *reinterpret_cast<const float4*>(omv + 8 * pc + 4);
            const float oo[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 kk = Pair16<T>::up(w4[u]);
              part = fmaf(kk.x, oo[2 * u] - 0.5f * kk.x, fmaf(kk.y, oo[2 * u + 1] - 0.5f * kk.y, part));
            }
          }
          lg[rd] = scale * part;
        }
      }
      mx = fmaxf(mx, lg[rd]);
    }
    mx = warp_max(mx);
    float den = 0.f;
#pragma unroll
    for (int rd = 0; rd < 4; ++rd) { lg[rd] = exp_nonpos(lg[rd] - mx); den += lg[rd]; }     // exp(-inf) = 0 on slots that do not exist
    const float inv_l = 1.0f / warp_sum(den);
    // ---- beta = sum_t softmax_t v_t: same four-tokens-per-instruction layout ----
    float av[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) av[i] = 0.f;
#pragma unroll
    for (int rd = 0; rd < 4; ++rd) {
      if (rd >= n_rounds) break;
#pragma unroll 4
      for (int i = 0; i < 8; ++i) {
        const int t = __shfl_sync(0xffffffffu, tokr[rd], 4 * i + ts);
        const float a = __shfl_sync(0xffffffffu, lg[rd], 4 * i + ts);
        if (t < 0) continue;
        add8<T>(__ldg(reinterpret_cast<const uint4*>(v.row<T>(b, t, h)) + p8), a, av);
      }
    }
    const float2 acc = pieces_to_pair(av, lane);
    reinterpret_cast<float2*>(beta_out + obase)[lane] = make_float2(acc.x * inv_l, acc.y * inv_l);
  }
}

// ------------------------------------------------------------------------------------------------
// Stage B: CTA = (batch*head, window, block of 16 query rows); keys = local window slots followed by
// the chunk keys (k_bar / beta); key tiles of 32 staged in smem, online softmax per query row.
// ------------------------------------------------------------------------------------------------
constexpr int kRows = 16;   // query rows per CTA (4 per warp)
constexpr int kRpw = 4;
constexpr int kKt = 32;     // keys per tile

template <typename T, int D>
__global__ void __launch_bounds__(128)
window_attn_kernel(const Geo g, const View q, const View k, const View v, const uint8_t* __restrict__ mask,
                   const float* __restrict__ kbar, const float* __restrict__ beta,
                   const float* __restrict__ bias, const long long bias_sh, T* __restrict__ out) {
  constexpr int DPL = Feat<D>::kPerLane;
  constexpr int DP = D + 1;
  extern __shared__ float sm[];
  float* Qs = sm;                       // [kRows][D], pre-scaled
  float* Ks = Qs + kRows * D;           // [kKt][DP]
  float* Vs = Ks + kKt * DP;            // [kKt][DP]
  float* Ps = Vs + kKt * DP;            // [4][kRpw][kKt]
  int* kflag = reinterpret_cast<int*>(Ps + 4 * kRpw * kKt);  // [kKt] 0 live, 1 masked, 2 absent
  int* qtok = kflag + kKt;              // [kRows] token id or -1
  int* qpad = qtok + kRows;             // [kRows]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // (row block, window, batch * head) folded into gridDim.x: gridDim.y / .z stop at 65535, the reference has no such limit
  const int n_rb = (g.L + kRows - 1) / kRows;
  const long long cta = blockIdx.x;
  const int rb = (int)(cta % n_rb), win = (int)((cta / n_rb) % g.n_windows);
  const int bh = (int)(cta / ((long long)n_rb * g.n_windows));
  const int b = bh / g.H, h = bh % g.H;
  const float scale = rsqrtf((float)D);
  const int n_keys = g.J + g.n_chunks;

  if (tid < kRows) {
    const int li = rb * kRows + tid;
    const int tok = li < g.L ? group_token(g, win, li, g.window, 0) : -1;
    qtok[tid] = tok;
    qpad[tid] = (tok >= 0 && mask) ? (int)mask[(long long)b * g.N + tok] : 0;
  }
  for (int idx = tid; idx < kRows * (D / 8); idx += blockDim.x) {
    const int r = idx / (D / 8), part = idx % (D / 8);
    const int li = rb * kRows + r;
    const int tok = li < g.L ? group_token(g, win, li, g.window, 0) : -1;
    float f[8];
    if (tok >= 0) load8<T>(q.row<T>(b, tok, h) + part * 8, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) Qs[r * D + part * 8 + i] = tok >= 0 ? f[i] * scale : 0.f;
  }

  float m[kRpw], l[kRpw], o[kRpw][DPL];
#pragma unroll
  for (int r = 0; r < kRpw; ++r) {
    m[r] = kNegInf; l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) o[r][i] = 0.f;
  }

  for (int kt0 = 0; kt0 < n_keys; kt0 += kKt) {
    __syncthreads();  // previous tile fully consumed (and Qs/qtok visible on the first pass)
    for (int idx = tid; idx < kKt * (D / 8); idx += blockDim.x) {
      const int j = idx / (D / 8), part = idx % (D / 8);
      const int gj = kt0 + j;
      float fk[8], fv[8];
      int flag = 0;
      bool have = false;
      if (gj < g.J) {
        const int tok = group_token(g, win, gj, g.window, g.ext);
        if (tok >= 0) {
          load8<T>(k.row<T>(b, tok, h) + part * 8, fk);
          load8<T>(v.row<T>(b, tok, h) + part * 8, fv);
          have = true;
          flag = (mask && mask[(long long)b * g.N + tok]) ? 1 : 0;
        } else {
          flag = 1;  // halo outside the sequence: zero key/value, masked logit (pad_val=1)
        }
      } else if (gj < n_keys) {
        const long long base = (((long long)b * g.H + h) * g.n_chunks + (gj - g.J)) * D + part * 8;
        load8<float>(kbar + base, fk);
        load8<float>(beta + base, fv);
        have = true;
      } else {
        flag = 2;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        Ks[j * DP + part * 8 + i] = have ? fk[i] : 0.f;
        Vs[j * DP + part * 8 + i] = have ? fv[i] : 0.f;
      }
      if (part == 0) kflag[j] = flag;
    }
    __syncthreads();

    // logits: lane = key, 4 rows of this warp share each K read
    float s[kRpw];
#pragma unroll
    for (int r = 0; r < kRpw; ++r) s[r] = 0.f;
    const float* krow = Ks + lane * DP;
    const float* qrow = Qs + (warp * kRpw) * D;
#pragma unroll 8
    for (int e = 0; e < D; ++e) {
      const float kk = krow[e];
#pragma unroll
      for (int r = 0; r < kRpw; ++r) s[r] = fmaf(qrow[r * D + e], kk, s[r]);
    }
    const int gj = kt0 + lane;
    const int flag = kflag[lane];
#pragma unroll
    for (int r = 0; r < kRpw; ++r) {
      const int row = warp * kRpw + r;
      const int li = rb * kRows + row;
      const int tq = qtok[row];
      if (tq < 0) continue;  // warp-uniform: row does not exist
      float sv = s[r];
      if (flag == 2) {
        sv = kNegInf;
      } else if (gj < g.J) {
        if (bias) sv += __ldg(bias + (long long)h * bias_sh + (long long)li * g.J + gj);
        const bool dead = flag == 1 || (g.mask_queries && qpad[row]);
        if (dead) sv = g.mask_fill;
        if (g.causal && gj > li + g.ext) sv = kMaskVal;
      } else {
        if (g.causal && (gj - g.J) >= tq / g.chunk) sv = kMaskVal;
      }
      const float mt = warp_max(sv);
      const float mn = fmaxf(m[r], mt);
      float corr = 1.f, p = 0.f;
      if (mn != kNegInf) { corr = exp_nonpos(m[r] - mn); p = exp_nonpos(sv - mn); }
      l[r] = fmaf(l[r], corr, warp_sum(p));
      m[r] = mn;
#pragma unroll
      for (int i = 0; i < DPL; ++i) o[r][i] *= corr;
      Ps[(warp * kRpw + r) * kKt + lane] = p;
    }
    __syncwarp();
    // PV: lane owns features lane + 32 i
    const float* prow = Ps + (warp * kRpw) * kKt;
#pragma unroll 4
    for (int j = 0; j < kKt; ++j) {
      float vv[DPL];
#pragma unroll
      for (int i = 0; i < DPL; ++i) vv[i] = Feat<D>::has(lane, i) ? Vs[j * DP + lane + 32 * i] : 0.f;
#pragma unroll
      for (int r = 0; r < kRpw; ++r) {
        const float p = prow[r * kKt + j];
#pragma unroll
        for (int i = 0; i < DPL; ++i) o[r][i] = fmaf(p, vv[i], o[r][i]);
      }
    }
    __syncwarp();
  }

#pragma unroll
  for (int r = 0; r < kRpw; ++r) {
    const int tq = qtok[warp * kRpw + r];
    if (tq < 0) continue;
    const float inv = 1.0f / l[r];
    T* orow = out + ((long long)b * g.N + tq) * ((long long)g.H * D) + (long long)h * D;
#pragma unroll
    for (int i = 0; i < DPL; ++i)
      if (Feat<D>::has(lane, i)) orow[lane + 32 * i] = from_f32<T>(o[r][i] * inv);
  }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
template <typename T, int D>
static cudaError_t launch_chunk_stats_t(const Geo& g, const View& q, const View& k, const View& v,
                                        const uint8_t* mask, const EvaAdaptive& ada, const float* noise,
                                        float* kbar, float* beta, cudaStream_t st) {
  constexpr int DPL = Feat<D>::kPerLane;
  if constexpr (D == 64) {
    // long 1-D chunks (causal LM): one CTA per chunk
    static const bool generic_only = [] { const char* e = getenv("EVA_SM100_DISABLE_FUSED"); return e && e[0] == '1'; }();
    if (!generic_only && g.dims == 1 && g.chunk_ext == 0 && !mask && g.Jc >= 64 && g.Jc <= 8192) {
      const size_t smem_cta = (size_t)(8 * 128 + 128 + 128 + 64 + 16 + ((g.Jc + 3) & ~3)) * sizeof(float) + (sizeof(T) == 2 && g.Jc <= 512 ? (size_t)g.Jc * 144 : 0);
      if (smem_cta > 48 * 1024) {
        const cudaError_t ea = cudaFuncSetAttribute(chunk_stats_cta_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta);
        if (ea != cudaSuccess) return ea;
      }
      const long long total_cta = (long long)g.B * g.H * g.n_chunks;
      chunk_stats_cta_kernel<T><<<(unsigned)total_cta, 256, smem_cta, st>>>(g, q, k, v, ada, noise, kbar, beta);
      return cudaGetLastError();
    }
  }
  if constexpr (D == 64 && sizeof(T) == 2) {
    static const bool slow_only = [] { const char* e = getenv("EVA_SM100_STATS_GENERIC"); return e && e[0] == '1'; }();
    if (!slow_only && g.Jc <= kFastMaxJc) {
      auto kf = chunk_stats_fast_kernel<T>;
      const size_t smf = (2 * 64 * 64 + 8 * 64) * sizeof(float);
      cudaError_t ef = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smf);
      if (ef != cudaSuccess) return ef;
      const long long total_f = (long long)g.B * g.H * g.n_chunks;
      long long blocks_f = (total_f + 7) / 8;
      if (blocks_f > 148LL * 16) blocks_f = 148LL * 16;
      kf<<<(unsigned)blocks_f, 256, smf, st>>>(g, q, k, v, mask, ada, noise, kbar, beta);
      return cudaGetLastError();
    }
  }
  const size_t smem = 2 * (size_t)D * D * sizeof(float);
  auto kern = chunk_stats_kernel<T, D>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long total = (long long)g.B * g.H * g.n_chunks;
  const int wpb = 8;
  long long blocks = (total + wpb - 1) / wpb;
  if (blocks > 148LL * 16) blocks = 148LL * 16;  // grid-stride beyond 16 CTAs per SM
  kern<<<(unsigned)blocks, wpb * 32, smem, st>>>(g, q, k, v, mask, ada, noise, kbar, beta);
  return cudaGetLastError();
}

template <typename T, int D>
static cudaError_t launch_window_attn_t(const Geo& g, const View& q, const View& k, const View& v,
                                        const uint8_t* mask, const float* kbar, const float* beta,
                                        const float* bias, long long bias_sh, void* out, cudaStream_t st) {
  constexpr int DPL = Feat<D>::kPerLane;
  const size_t smem = (size_t)(kRows * D + 2 * kKt * (D + 1) + 4 * kRpw * kKt) * sizeof(float) +
                      (size_t)(kKt + 2 * kRows) * sizeof(int);
  auto kern = window_attn_kernel<T, D>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long ctas = (long long)((g.L + kRows - 1) / kRows) * g.n_windows * g.B * g.H;
  if (ctas > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  kern<<<(unsigned)ctas, 128, smem, st>>>(g, q, k, v, mask, kbar, beta, bias, bias_sh, reinterpret_cast<T*>(out));
  return cudaGetLastError();
}

#define EVA_DISPATCH(FN, ...)                                                             \
  switch (io_dtype * 256 + g.D) {                                                         \
    case EVA_F32 * 256 + 16: return FN<float, 16>(__VA_ARGS__);                           \
    case EVA_F32 * 256 + 32: return FN<float, 32>(__VA_ARGS__);                           \
    case EVA_F32 * 256 + 64: return FN<float, 64>(__VA_ARGS__);                           \
    case EVA_F32 * 256 + 128: return FN<float, 128>(__VA_ARGS__);                         \
    case EVA_F16 * 256 + 16: return FN<__half, 16>(__VA_ARGS__);                          \
    case EVA_F16 * 256 + 32: return FN<__half, 32>(__VA_ARGS__);                          \
    case EVA_F16 * 256 + 64: return FN<__half, 64>(__VA_ARGS__);                          \
    case EVA_F16 * 256 + 128: return FN<__half, 128>(__VA_ARGS__);                        \
    case EVA_BF16 * 256 + 16: return FN<__nv_bfloat16, 16>(__VA_ARGS__);                  \
    case EVA_BF16 * 256 + 32: return FN<__nv_bfloat16, 32>(__VA_ARGS__);                  \
    case EVA_BF16 * 256 + 64: return FN<__nv_bfloat16, 64>(__VA_ARGS__);                  \
    case EVA_BF16 * 256 + 128: return FN<__nv_bfloat16, 128>(__VA_ARGS__);                \
    default: return cudaErrorInvalidValue;                                                \
  }

cudaError_t launch_chunk_stats(const Geo& g, int io_dtype, const View& q, const View& k, const View& v,
                               const uint8_t* mask, const EvaAdaptive& ada, const float* noise,
                               float* kbar, float* beta, cudaStream_t st) {
  EVA_DISPATCH(launch_chunk_stats_t, g, q, k, v, mask, ada, noise, kbar, beta, st)
}

cudaError_t launch_window_attn(const Geo& g, int io_dtype, const View& q, const View& k, const View& v,
                               const uint8_t* mask, const float* kbar, const float* beta,
                               const float* bias, long long bias_sh, void* out, cudaStream_t st) {
  EVA_DISPATCH(launch_window_attn_t, g, q, k, v, mask, kbar, beta, bias, bias_sh, out, st)
}

}  // namespace eva
