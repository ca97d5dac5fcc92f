// LARA forward core on tcgen05 / TMEM / TMA for sm_100a: replaces lara_stats_kernel + lara_out_kernel (lara.py:201-246) for
// the DeiT geometry (BASELINE config c4): mis-opt weights, one proposal sample per landmark (S == C <= 64), N <= 224 tokens,
// head_dim 64, 16-bit I/O, no padding mask.  lara_landmark_kernel still produces q_bar, omega, lp, bh per (batch, head).
//
// Per (batch, head) item, everything from one TMA load of q, k, v (N x 128 B each):
//   phase S (TMEM lane = landmark): two 128-row A tiles, T1 = [omega ; 0] and T2 = [0 ; q_bar], ONE accumulator
//       D = T1 K^T + T2 Q^T   lanes 0-63:   phi-logits of the keys  -> softmax over the tokens (thread-local) -> P, lse_k
//                             lanes 64-127: q_bar q^T               -> lse_t (the normaliser of t_nc, lara.py:222-223)
//                             (both row softmaxes run side by side on the four compute warps)
//       kv = P V      (A operand from TMEM)                -> 16-bit kv tile in shared memory
//   phase O (TMEM lane = token, two blocks of 128 tokens):
//       D3 = Q T1^T + Q T2^T   columns 0-63: q . omega_c (phi(q)), columns 64-127: q . q_bar_c (t_nc)
//       per token: t, alpha = bh + coeff (t - mean_c t), log w = log alpha + phi(q) + lse_k - lp, softmax over c (thread-local)
//       O = W kv      (A operand from TMEM) -> normalise -> staged in the (dead) q tile -> TMA store
// Two persistent CTAs per SM (<= 113 KB shared memory, 256 TMEM columns) hide each other's serial per-item chain: warps 0-3
// compute, warp 4 producer (TMA + the fp32 -> 16-bit tiles), warp 5 MMA issuer, warps 6-7 only complete the second warpgroup:
// the kernel launches with 128 registers per thread and re-balances with setmaxnreg (compute 208, the others 48).  One q tile
// and one k/v tile per CTA: v is loaded into the k tile once the D MMAs have read it.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>

#include <mutex>

#define EVA_MBAR_WAIT_NO_CALL   // setmaxnreg below: no out-of-line calls in this translation unit's kernels
#include "common.cuh"
#include "lara_ws.cuh"
#include "launch.h"
#include "sm100_ptx.cuh"

namespace eva {
namespace laracore {

constexpr int kThreads = 256;         // two warpgroups: compute (warps 0-3) | producer, MMA issuer and two idle warps (setmaxnreg works on whole warpgroups)
constexpr int kRegsCompute = 208, kRegsOther = 48;   // after re-balancing; launch allocation: 128 x 256 with two CTAs per SM
static_assert(128 * (kRegsCompute + kRegsOther) <= 128 * kThreads, "setmaxnreg budget");
enum Bar { kFullQK0, kFullQK1, kFullV0, kFullV1, kFullAW0, kFullAW1, kFree0, kFree1,
           kSFull, kPFull, kKvFull, kKvTile, kD3Full, kP2Full0, kP2Full1, kOFull0, kOFull1, kEpiDone,
           kFullW0, kFullW1, kWFree, kToMma, kToCompute, kKFree, kS2Full, kLtDone, kNumBars };

struct Params {
  int B, H, N, NP, C, items;
  float alpha_coeff;
  const float* ws;               // lara workspace (float32) written by lara_landmark_kernel
  int sl;                        // bytes of one q / k / v tile = NP * 128
  // fused landmark phase (phase L): pooled 2-D proposals, Linear + LayerNorm, optional landmark mixing, proposal statistics
  int fuse, side, gh, gw, mixed, has_proj;
  const float *b_q, *g_q, *beta_q, *b_k, *g_k, *beta_k;
  float ln_eps;
  const float* noise;            // [B, H, C, 64] or NULL
};

template <typename T> struct Fmt;
template <> struct Fmt<__half> {
  static constexpr uint32_t kUmma = ptx::kFmtF16;
  static __device__ __forceinline__ uint32_t pack2(float a, float b) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
  static __device__ __forceinline__ float2 unpack2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
};
template <> struct Fmt<__nv_bfloat16> {
  static constexpr uint32_t kUmma = ptx::kFmtBF16;
  static __device__ __forceinline__ uint32_t pack2(float a, float b) { const __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
  static __device__ __forceinline__ float2 unpack2(uint32_t v) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v)); }
};
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// TMEM columns
constexpr uint32_t cD1 = 0, cD2 = 0, cKv = 128;            // phase S: D1 [128 x NP], then (after the kv read-back) D2 in the same columns; kv [128 x 64] in the
                                                           // columns of D1 that are dead once P (NP / 2 <= 112 columns) has been written
constexpr uint32_t cD3 = 0, cO = 64;                       // phase O, token block rb at 128 rb: D3 [128 x 128]; O over the dead t half (+64)

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
lara_core_kernel(const __grid_constant__ CUtensorMap t_q, const __grid_constant__ CUtensorMap t_k,
                 const __grid_constant__ CUtensorMap t_v, const __grid_constant__ CUtensorMap t_o,
                 const __grid_constant__ CUtensorMap t_w, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* const sm = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int SL = p.sl;
  uint8_t* const aw0 = sm + 2 * SL;                  // two [128][128 B] tiles: T1 = [omega ; 0], T2 = [0 ; q_bar] (T1 first holds the means tile)
  uint8_t* const kvt = aw0 + 2 * 16384;              // [128][128 B]: kv (rows = landmarks; phase L: k_bar, then mu); rows 64-127 stay zero (M = 128 A operand of the mixing MMA)
  float* const n2k = reinterpret_cast<float*>(kvt + 16384);    // [256] |k_n|^2 scale log2(e) / 2 (+inf for n >= N)
  uint32_t* const bins = reinterpret_cast<uint32_t*>(n2k + 256);   // [64] pooling bin of landmark c: y0 | y1 << 8 | x0 << 16 | x1 << 24 (|q_n|^2 is not needed: it cancels in the softmax over c)
  float* const lpS = n2k + 512;                                 // [2][64] lp  (per stage, from the workspace)
  float* const bhS = lpS + 128;                                 // [2][64] bh
  float* const cst2 = bhS + 128;                                // [64] (lse_k - lp) log2(e)
  float* const lse2t = cst2 + 64;                               // [64] lse of the t logits, log2 units
  float* const mu2 = lse2t + 64;                                // [64] |mu_c|^2 (phase L)
  float* const lnp = mu2 + 64;                                  // [6][64] Linear bias, LayerNorm gain / bias: q side | k side
  const uint32_t bars = ptx::smem_u32(reinterpret_cast<uint8_t*>(lnp + 384));
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(reinterpret_cast<uint8_t*>(lnp + 384) + kNumBars * 8);
  auto bar = [&](int i) { return bars + 8u * i; };
  auto tile = [&](int, int which) { return sm + (which == 0 ? 0 : SL); };      // which: 0 q, 1 k, 2 v (v takes over the k tile)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, NP = p.NP, C = p.C;

  for (int i = tid; i < (3 * 16384) / 16; i += kThreads) reinterpret_cast<uint4*>(aw0)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 4 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(bar(kFullQK0 + s), 1);
      ptx::mbar_init(bar(kFullV0 + s), 1);
      ptx::mbar_init(bar(kFullAW0 + s), 1);
      ptx::mbar_init(bar(kFree0 + s), 3);            // MMA commit + one output-store drain per token block
      ptx::mbar_init(bar(kP2Full0 + s), 128);
      ptx::mbar_init(bar(kOFull0 + s), 1);
    }
    ptx::mbar_init(bar(kSFull), 1);
    ptx::mbar_init(bar(kPFull), 128);
    ptx::mbar_init(bar(kKvFull), 1);
    ptx::mbar_init(bar(kKvTile), 128);
    ptx::mbar_init(bar(kD3Full), 1);
    ptx::mbar_init(bar(kEpiDone), 128);
    ptx::mbar_init(bar(kFullW0), 1);
    ptx::mbar_init(bar(kFullW1), 1);
    ptx::mbar_init(bar(kWFree), 1);
    ptx::mbar_init(bar(kToMma), 128);
    ptx::mbar_init(bar(kToCompute), 1);
    ptx::mbar_init(bar(kKFree), 1);
    ptx::mbar_init(bar(kS2Full), 1);
    ptx::mbar_init(bar(kLtDone), 128);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&t_q); ptx::prefetch_tmap(&t_k); ptx::prefetch_tmap(&t_v); ptx::prefetch_tmap(&t_o);
  }
  if (p.fuse) {
    if (tid < 64) {                                   // AdaptiveAvgPool2d bins, once per CTA (no integer divisions in the item loop)
      const int cc = tid < C ? tid : 0, by = cc / p.side, bx = cc % p.side;
      const int y0 = (by * p.gh) / p.side, y1 = ((by + 1) * p.gh + p.side - 1) / p.side;
      const int x0 = (bx * p.gw) / p.side, x1 = ((bx + 1) * p.gw + p.side - 1) / p.side;
      bins[tid] = (uint32_t)y0 | ((uint32_t)y1 << 8) | ((uint32_t)x0 << 16) | ((uint32_t)x1 << 24);
    }
    const float* src[6] = {p.b_q, p.g_q, p.beta_q, p.b_k, p.g_k, p.beta_k};
    for (int idx = tid; idx < 384; idx += kThreads) lnp[idx] = src[idx >> 6] ? __ldg(src[idx >> 6] + (idx & 63)) : 0.f;
  }
  if (warp == 5) ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr)), 256);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const float scale_log2 = 0.125f * kLog2e;

  // register re-balancing: each role changes its budget INSIDE its own branch (ptxas gives code after a join of branches with
  // different budgets the smallest one)
  if (warp >= 6) {
    ptx::setmaxnreg_dec<kRegsOther>();              // only there to complete the second warpgroup
  } else if (warp == 4) {
    ptx::setmaxnreg_dec<kRegsOther>();
    // =================================== producer ==============================================
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int s = 0;
      const int h = item % p.H, b = item / p.H;
      if (it >= 1) ptx::mbar_wait(bar(kFree0 + s), (it - 1) & 1);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(bar(kFullQK0 + s), 2 * SL);
        ptx::tma_load_4d(ptx::smem_u32(tile(s, 0)), &t_q, bar(kFullQK0 + s), 0, h, 0, b);
        ptx::tma_load_4d(ptx::smem_u32(tile(s, 1)), &t_k, bar(kFullQK0 + s), 0, h, 0, b);
        if (p.fuse && p.has_proj) {                  // [W_q ; W_k] borrows the kv tile until the Linear MMA has read it
          ptx::mbar_arrive_expect_tx(bar(kFullW0 + s), 16384);
          ptx::tma_load_2d(ptx::smem_u32(kvt), &t_w, bar(kFullW0 + s), 0, 0);
        }
      }
      auto load_v = [&]() {                          // v takes over the k tile once D1 has read it
        ptx::mbar_wait(bar(kKFree), it & 1);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(bar(kFullV0 + s), SL);
          ptx::tma_load_4d(ptx::smem_u32(tile(s, 2)), &t_v, bar(kFullV0 + s), 0, h, 0, b);
          // one q tile and one k/v tile per CTA: the next item's loads cannot start before this item is finished, so at least
          // make them L2 hits
          const int nxt = item + (int)gridDim.x;
          if (nxt < p.items) {
            const int hn = nxt % p.H, bn = nxt / p.H;
            ptx::tma_prefetch_4d(&t_q, 0, hn, 0, bn);
            ptx::tma_prefetch_4d(&t_k, 0, hn, 0, bn);
            ptx::tma_prefetch_4d(&t_v, 0, hn, 0, bn);
          }
        }
      };
      if (p.fuse) {
        load_v();
        continue;                                    // the AW tile, lp and bh are produced on chip (phase L)
      }
      const LaraWs w = lara_ws_at(const_cast<float*>(p.ws), item, C, C, 64);
      uint8_t* aw = aw0 + s * 16384;
      for (int idx = lane; idx < C * 8; idx += 32) {
        const int row = idx >> 3, ch = idx & 7;
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(w.omega + row * 64 + ch * 8));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(w.omega + row * 64 + ch * 8) + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(w.qbar + row * 64 + ch * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(w.qbar + row * 64 + ch * 8) + 1);
        const int off = row * 128 + ((ch ^ (row & 7)) << 4);       // row and row + 64 share (row & 7)
        *reinterpret_cast<uint4*>(aw + off) = make_uint4(Fmt<T>::pack2(a0.x, a0.y), Fmt<T>::pack2(a0.z, a0.w), Fmt<T>::pack2(a1.x, a1.y), Fmt<T>::pack2(a1.z, a1.w));
        *reinterpret_cast<uint4*>(aw + 16384 + 8192 + off) = make_uint4(Fmt<T>::pack2(b0.x, b0.y), Fmt<T>::pack2(b0.z, b0.w), Fmt<T>::pack2(b1.x, b1.y), Fmt<T>::pack2(b1.z, b1.w));
      }
      for (int c = lane; c < C; c += 32) { lpS[s * 64 + c] = __ldg(w.lp + c); bhS[s * 64 + c] = __ldg(w.bh + c); }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (ptx::elect_one()) ptx::mbar_arrive(bar(kFullAW0 + s));
      load_v();
    }
  } else if (warp == 5) {
    ptx::setmaxnreg_dec<kRegsOther>();
    // =================================== MMA issuer ============================================
    constexpr uint32_t fmt = Fmt<T>::kUmma;
    const uint32_t id_s = ptx::umma_idesc(fmt, fmt, 0, 0, 128, (uint32_t)NP);
    constexpr uint32_t id_d3 = ptx::umma_idesc(fmt, fmt, 0, 0, 128, 128);
    constexpr uint32_t id_pv = ptx::umma_idesc(fmt, fmt, 0, 1, 128, 64);
    constexpr uint32_t id_m64 = ptx::umma_idesc(fmt, fmt, 0, 0, 128, 64);
    const uint64_t dKV = ptx::umma_desc_sw128(ptx::smem_u32(kvt));
    uint32_t it = 0, hand = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int s = 0;
      const uint32_t ph = it & 1, pi = it & 1;
      const uint64_t dQ = ptx::umma_desc_sw128(ptx::smem_u32(tile(s, 0))), dK = ptx::umma_desc_sw128(ptx::smem_u32(tile(s, 1)));
      const uint64_t dV = ptx::umma_desc_sw128(ptx::smem_u32(tile(s, 2))), dAW = ptx::umma_desc_sw128(ptx::smem_u32(aw0 + s * 16384));
      const uint64_t dT2 = dAW + (uint64_t)(16384 >> 4);            // [0 ; q_bar]
      ptx::mbar_wait(bar(kFullQK0 + s), ph);
      if (!p.fuse) {
        ptx::mbar_wait(bar(kFullAW0 + s), ph);
        if (it > 0) ptx::mbar_wait(bar(kEpiDone), (it - 1) & 1);    // the previous item's TMEM has been read
        ptx::mbar_wait(bar(kToMma), hand & 1);                       // |k|^2 taken: D1 may release the k tile to the v load
        ++hand;
      } else {
        // phase L: a strict ping-pong with the compute warps (kToMma: 128 arrivals, kToCompute: one commit per hand-off)
        const uint64_t dKBm = dKV;                                   // k_bar / mu tile lives in the kv tile until the kv read-back
        auto wait_compute = [&]() { ptx::mbar_wait(bar(kToMma), hand & 1); ++hand; ptx::tc_fence_after(); };
        if (p.has_proj) {
          wait_compute();                                            // means tile written (in the AW tile of this stage)
          ptx::mbar_wait(bar(kFullW0 + s), ph);
          if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + 0, dAW + 2 * ks, dKV + 2 * ks, id_d3, ks > 0);
            ptx::umma_commit(bar(kToCompute));
          }
        }
        if (p.mixed) {
          wait_compute();                                            // k_bar tile written
          if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + 128, dKBm + 2 * ks, dKBm + 2 * ks, id_m64, ks > 0);
            ptx::umma_commit(bar(kToCompute));
          }
          wait_compute();                                            // mixing weights P written over the logits
          if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) ptx::umma_ts(tmem + 192, tmem + 128 + 8 * ks, dKBm + 128 * ks, id_pv, ks > 0);
            ptx::umma_commit(bar(kToCompute));
          }
        }
        wait_compute();                                              // AW (omega | q_bar) and mu tiles written
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + 0, dAW + 2 * ks, dKBm + 2 * ks, id_m64, ks > 0);   // the Linear result is dead
          ptx::umma_commit(bar(kToCompute));
        }
        wait_compute();                                              // lp / bh done, phase-L TMEM reads finished
      }
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cD1, dAW + 2 * ks, dK + 2 * ks, id_s, ks > 0);
        // [omega ; 0] K^T + [0 ; q_bar] Q^T in ONE accumulator: lanes 0-63 hold the phi-logits of the keys, lanes 64-127 the
        // t-logits, so that the two row softmaxes run side by side on all four compute warps
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cD1, dT2 + 2 * ks, dQ + 2 * ks, id_s, 1);
        ptx::umma_commit(bar(kSFull));
        ptx::umma_commit(bar(kKFree));
      }
      ptx::mbar_wait(bar(kPFull), pi);
      ptx::mbar_wait(bar(kFullV0 + s), ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll 1
        for (int ks = 0; ks < NP / 16; ++ks) ptx::umma_ts(tmem + cKv, tmem + cD1 + 8 * ks, dV + 128 * ks, id_pv, ks > 0);
        ptx::umma_commit(bar(kKvFull));
      }
      ptx::mbar_wait(bar(kKvTile), pi);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll 1
        for (int rb = 0; rb < 2; ++rb) {                             // D3 = Q [omega ; 0]^T + Q [0 ; q_bar]^T
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_ss(tmem + cD3 + 128 * rb, dQ + (uint64_t)(rb * (16384 >> 4)) + 2 * ks, dAW + 2 * ks, id_d3, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_ss(tmem + cD3 + 128 * rb, dQ + (uint64_t)(rb * (16384 >> 4)) + 2 * ks, dT2 + 2 * ks, id_d3, 1);
        }
        ptx::umma_commit(bar(kD3Full));
      }
#pragma unroll 1
      for (int rb = 0; rb < 2; ++rb) {
        ptx::mbar_wait(bar(kP2Full0 + rb), pi);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_ts(tmem + cO + 128 * rb, tmem + cD3 + 128 * rb + 8 * ks, dKV + 128 * ks, id_pv, ks > 0);
          ptx::umma_commit(bar(kOFull0 + rb));
          if (rb == 1) ptx::umma_commit(bar(kFree0 + s));
        }
      }
    }
  } else {
    ptx::setmaxnreg_inc<kRegsCompute>();
    // =================================== compute warps =========================================
    const uint32_t trow = tmem + ((uint32_t)(32 * warp) << 16);
    const bool d1_side = tid < 64;                     // warps 0-1: omega rows (D1); warps 2-3: q_bar rows (D2)
    const int c_row = tid & 63;
    uint32_t it = 0, hc = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int s = 0;
      const uint32_t ph = it & 1, pi = it & 1;
      const int h = item % p.H, b = item / p.H;
      // |k_n|^2 from the tile: two threads per token row (64 B each), packed fp32x2 FMAs
      ptx::mbar_wait(bar(kFullQK0 + s), ph);
      for (int idx = tid; idx < 2 * NP; idx += 128) {
        const int n = idx >> 1, half = idx & 1;
        uint64_t acc2 = 0ull;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int ch = 4 * half + c4;
          const uint4 rk = *reinterpret_cast<const uint4*>(tile(s, 1) + n * 128 + ((ch ^ (n & 7)) << 4));
          const uint32_t wk[4] = {rk.x, rk.y, rk.z, rk.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 fk = Fmt<T>::unpack2(wk[j]);
            const uint64_t f2 = ptx::pk2(fk.x, fk.y);
            acc2 = ptx::fma2(f2, f2, acc2);
          }
        }
        float a0, a1;
        ptx::upk2(acc2, a0, a1);
        float ak = a0 + a1;
        ak += __shfl_xor_sync(0xffffffffu, ak, 1);
        if (!half) n2k[n] = n < N ? 0.5f * scale_log2 * ak : __int_as_float(0x7f800000);   // pre-scaled; +inf masks the padding columns of D1
      }
      ptx::named_bar_sync(1, 128);
      if (!p.fuse) ptx::mbar_arrive(bar(kToMma));
      if (p.fuse) {
        // ---- phase L: landmarks on chip (lara.py:129-175, 182-198, 221-232) ----
        uint8_t* const awt = aw0 + s * 16384;           // first the means tile, finally [omega ; q_bar]
        const int ws_ = tid >> 6, c = tid & 63;         // lanes 0-63: q side, 64-127: k side; c = landmark
        auto wait_mma = [&]() { ptx::mbar_wait(bar(kToCompute), hc & 1); ++hc; ptx::tc_fence_after(); };
        auto to_mma = [&]() { ptx::fence_proxy_async_smem(); ptx::tc_fence_before(); ptx::mbar_arrive(bar(kToMma)); };
        // L1: adaptive average pooling of q and k (AdaptiveAvgPool2d bins), 16-byte pieces -> means tile
        for (int idx = tid; idx < 2 * C * 8; idx += 128) {
          const int part = idx & 7, u = idx >> 3, sd = u >= C ? 1 : 0, cc = u - (sd ? C : 0);
          const uint32_t bin = bins[cc];
          const int y0 = bin & 255, y1 = (bin >> 8) & 255, x0 = (bin >> 16) & 255, x1 = bin >> 24;
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = 0.f;
          for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {
              const int n = y * p.gw + x;
              const uint4 raw = *reinterpret_cast<const uint4*>(tile(s, sd) + n * 128 + ((part ^ (n & 7)) << 4));
              const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) { const float2 f = Fmt<T>::unpack2(w4[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
            }
          const float inv = __fdividef(1.0f, (float)((y1 - y0) * (x1 - x0)));
          const int r = 64 * sd + cc;
          *reinterpret_cast<uint4*>(awt + r * 128 + ((part ^ (r & 7)) << 4)) =
              make_uint4(Fmt<T>::pack2(acc[0] * inv, acc[1] * inv), Fmt<T>::pack2(acc[2] * inv, acc[3] * inv),
                         Fmt<T>::pack2(acc[4] * inv, acc[5] * inv), Fmt<T>::pack2(acc[6] * inv, acc[7] * inv));
        }
        to_mma();
        // L2: Linear (MMA) -> bias, LayerNorm
        wait_mma();
        float y[64];
#pragma unroll
        for (int g = 0; g < 4; ++g) ptx::tmem_ld16(trow + (ws_ ? 64u : 0u) + 16 * g, reinterpret_cast<uint32_t*>(y) + 16 * g);
        ptx::tmem_ld_wait();
        {
          const float* lp_ = lnp + (ws_ ? 192 : 0);
#pragma unroll
          for (int e = 0; e < 64; ++e) y[e] += lp_[e];
          if (ws_ ? (p.g_k != nullptr) : (p.g_q != nullptr)) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int e = 0; e < 64; e += 2) { s0 += y[e]; s1 += y[e + 1]; }
            const float mean = (s0 + s1) * (1.0f / 64);
            float v0 = 0.f, v1 = 0.f;
#pragma unroll
            for (int e = 0; e < 64; e += 2) { const float d0 = y[e] - mean, d1 = y[e + 1] - mean; v0 = fmaf(d0, d0, v0); v1 = fmaf(d1, d1, v1); }
            const float rs = rsqrtf((v0 + v1) * (1.0f / 64) + p.ln_eps);
#pragma unroll
            for (int e = 0; e < 64; ++e) y[e] = (y[e] - mean) * rs * lp_[64 + e] + lp_[128 + e];
          }
        }
        if (ws_ == 1) {                                  // k_bar rows -> tile (rows >= C zero)
          uint8_t* row = kvt + c * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<uint4*>(row + ((ch ^ (c & 7)) << 4)) = c < C ?
                make_uint4(Fmt<T>::pack2(y[8 * ch], y[8 * ch + 1]), Fmt<T>::pack2(y[8 * ch + 2], y[8 * ch + 3]),
                           Fmt<T>::pack2(y[8 * ch + 4], y[8 * ch + 5]), Fmt<T>::pack2(y[8 * ch + 6], y[8 * ch + 7])) : make_uint4(0, 0, 0, 0);
        }
        float k2[64], lp_exact = 0.f;
        if (p.mixed) {
          // L3: k_bar <- softmax(scale k_bar k_bar^T) k_bar (lara.py:157-174); row p on lane p
          to_mma();
          wait_mma();
          if (ws_ == 0) {
            float v[64];
#pragma unroll
            for (int g = 0; g < 4; ++g) ptx::tmem_ld16(trow + 128 + 16 * g, reinterpret_cast<uint32_t*>(v) + 16 * g);
            ptx::tmem_ld_wait();
            float mx = kNegInf;
#pragma unroll
            for (int e = 0; e < 64; ++e) mx = fmaxf(mx, e < C ? v[e] : kNegInf);
            float sum = 0.f;
#pragma unroll
            for (int e = 0; e < 64; ++e) { v[e] = e < C ? ex2((v[e] - mx) * scale_log2) : 0.f; sum += v[e]; }
            const float inv = __fdividef(1.0f, sum);
            uint32_t pk[32];
#pragma unroll
            for (int e = 0; e < 64; e += 2) pk[e >> 1] = Fmt<T>::pack2(v[e] * inv, v[e + 1] * inv);
            ptx::tmem_st16(trow + 128, pk);
            ptx::tmem_st16(trow + 144, pk + 16);
            ptx::tmem_st_wait();
          }
          to_mma();
          wait_mma();
          if (ws_ == 0) {
#pragma unroll
            for (int g = 0; g < 4; ++g) ptx::tmem_ld16(trow + 192 + 16 * g, reinterpret_cast<uint32_t*>(k2) + 16 * g);
            ptx::tmem_ld_wait();
          }
        } else {
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(1, 128);                   // k_bar rows visible to the q-side lanes
          if (ws_ == 0) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const uint4 raw = *reinterpret_cast<const uint4*>(kvt + c * 128 + ((ch ^ (c & 7)) << 4));
              const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) { const float2 f = Fmt<T>::unpack2(w4[j]); k2[8 * ch + 2 * j] = f.x; k2[8 * ch + 2 * j + 1] = f.y; }
            }
          }
        }
        // L4: mu = q_bar + k_bar (lara.py:182), omega = mu (+ noise); tiles for the proposal statistics and the later phases
        if (ws_ == 1) {                                  // the k-side means of T1 are dead: T1 becomes [omega ; 0]
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(awt + (64 + c) * 128 + (ch << 4)) = make_uint4(0, 0, 0, 0);
        }
        if (ws_ == 0) {
          const bool live = c < C;
          const float* nz = (p.noise && live) ? p.noise + ((long long)item * C + c) * 64 : nullptr;
          float m2 = 0.f, dot = 0.f;
          uint8_t* const r_om = awt + c * 128;
          uint8_t* const r_qb = awt + 16384 + (64 + c) * 128;   // T2
          uint8_t* const r_mu = kvt + c * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            float mu[8], om[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              mu[j] = live ? y[8 * ch + j] + k2[8 * ch + j] : 0.f;
              om[j] = mu[j] + (nz ? __ldg(nz + 8 * ch + j) : 0.f);
              m2 = fmaf(mu[j], mu[j], m2);
              dot = fmaf(om[j], mu[j], dot);
            }
            const int sw = (ch ^ (c & 7)) << 4;
            *reinterpret_cast<uint4*>(r_om + sw) = make_uint4(Fmt<T>::pack2(om[0], om[1]), Fmt<T>::pack2(om[2], om[3]), Fmt<T>::pack2(om[4], om[5]), Fmt<T>::pack2(om[6], om[7]));
            *reinterpret_cast<uint4*>(r_mu + sw) = make_uint4(Fmt<T>::pack2(mu[0], mu[1]), Fmt<T>::pack2(mu[2], mu[3]), Fmt<T>::pack2(mu[4], mu[5]), Fmt<T>::pack2(mu[6], mu[7]));
            *reinterpret_cast<uint4*>(r_qb + sw) = live ?
                make_uint4(Fmt<T>::pack2(y[8 * ch], y[8 * ch + 1]), Fmt<T>::pack2(y[8 * ch + 2], y[8 * ch + 3]),
                           Fmt<T>::pack2(y[8 * ch + 4], y[8 * ch + 5]), Fmt<T>::pack2(y[8 * ch + 6], y[8 * ch + 7])) : make_uint4(0, 0, 0, 0);
          }
          mu2[c] = m2;
          lp_exact = scale_log2 * (dot - 0.5f * m2);   // L[s][s] from the fp32 rows (the MMA below sees 16-bit copies), log2 units
        }
        to_mma();
        // L5: proposal statistics L[s][c] = prm(mu_c, omega_s): lp = L[s][s], bh = exp(lp - lse_c L[s][c]) (mis-opt)
        wait_mma();
        ptx::named_bar_sync(1, 128);                     // mu2 of every landmark
        if (ws_ == 0) {
          float v[64];
#pragma unroll
          for (int g = 0; g < 4; ++g) ptx::tmem_ld16(trow + 0 + 16 * g, reinterpret_cast<uint32_t*>(v) + 16 * g);
          ptx::tmem_ld_wait();
          float mx = kNegInf;
          const float lp = lp_exact;
#pragma unroll
          for (int e = 0; e < 64; ++e) {
            v[e] = scale_log2 * (v[e] - 0.5f * mu2[e]);  // log2 units
            v[e] = e == c ? lp : v[e];                   // keep the diagonal consistent with lp
            mx = fmaxf(mx, e < C ? v[e] : kNegInf);
          }
          float sum = 0.f;
#pragma unroll
          for (int e = 0; e < 64; ++e) sum += e < C ? ex2(v[e] - mx) : 0.f;
          const float lse2 = mx + lg2(sum);
          lpS[s * 64 + c] = lp * (1.0f / kLog2e);        // natural-log units, as the workspace of the landmark kernel
          bhS[s * 64 + c] = ex2(lp - lse2);
        }
        to_mma();
      }
      // ---- phase S: softmax over the tokens, thread = landmark row ----
      // one two-pass softmax routine, used on D1 by the omega rows (lanes 0-63) and later on D2 by the q_bar rows (64-127)
      float sum = 0.f, lse2 = 0.f;
      auto row_softmax = [&](const bool with_k2, const bool write_p) {
        float m0 = kNegInf;
#pragma unroll 1
        for (int g = 0; g < NP / 16; ++g) {
          float v[16];
          ptx::tmem_ld16(trow + 16 * g, reinterpret_cast<uint32_t*>(v));
          ptx::tmem_ld_wait();
          float hk[16];
          if (with_k2) {
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 x = *reinterpret_cast<const float4*>(n2k + 16 * g + 4 * e4);
              hk[4 * e4] = x.x; hk[4 * e4 + 1] = x.y; hk[4 * e4 + 2] = x.z; hk[4 * e4 + 3] = x.w;
            }
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int n = 16 * g + e;
            if (with_k2) m0 = fmaxf(m0, fmaf(scale_log2, v[e], -hk[e]));
            else m0 = fmaxf(m0, n < N ? scale_log2 * v[e] : kNegInf);
          }
        }
        sum = 0.f;
#pragma unroll 1
        for (int g = 0; g < NP / 16; ++g) {
          float v[16];
          uint32_t pk[8];
          ptx::tmem_ld16(trow + 16 * g, reinterpret_cast<uint32_t*>(v));
          ptx::tmem_ld_wait();
          float hk[16];
          if (with_k2) {
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 x = *reinterpret_cast<const float4*>(n2k + 16 * g + 4 * e4);
              hk[4 * e4] = x.x + m0; hk[4 * e4 + 1] = x.y + m0; hk[4 * e4 + 2] = x.z + m0; hk[4 * e4 + 3] = x.w + m0;
            }
          }
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const int n = 16 * g + e;
            float a, c2;
            if (with_k2) {                                             // padding columns: hk = +inf -> ex2(-inf) = 0
              a = ex2(fmaf(scale_log2, v[e], -hk[e]));
              c2 = ex2(fmaf(scale_log2, v[e + 1], -hk[e + 1]));
            } else {
              a = n < N ? ex2(fmaf(scale_log2, v[e], -m0)) : 0.f;
              c2 = n + 1 < N ? ex2(fmaf(scale_log2, v[e + 1], -m0)) : 0.f;
            }
            sum += a + c2;
            pk[e >> 1] = Fmt<T>::pack2(a, c2);
          }
          if (write_p) ptx::tmem_st8(trow + 8 * g, pk);               // P over the first half of the columns already read
        }
        lse2 = m0 + lg2(sum);                                          // log2 units
      };
      ptx::mbar_wait(bar(kSFull), pi);
      if (!p.fuse) ptx::mbar_wait(bar(kFullAW0 + s), ph);          // lp / bh of this stage (written by the producer warp) are visible
      ptx::tc_fence_after();
      if (d1_side) {
        row_softmax(true, true);
        cst2[c_row] = lse2 - lpS[s * 64 + c_row] * kLog2e;
        ptx::tmem_st_wait();
      } else {                                                        // the q_bar rows: lse_t, side by side with the omega rows
        row_softmax(false, false);
        lse2t[c_row] = lse2;
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(kPFull));
      // ---- kv rows -> 16-bit tile ----
      ptx::mbar_wait(bar(kKvFull), pi);
      ptx::tc_fence_after();
      if (d1_side) {
        float o[64];
#pragma unroll
        for (int g = 0; g < 4; ++g) ptx::tmem_ld16(trow + cKv + 16 * g, reinterpret_cast<uint32_t*>(o) + 16 * g);
        ptx::tmem_ld_wait();
        if (c_row < C) {
          const float inv = __fdividef(1.0f, sum);
          uint8_t* row = kvt + c_row * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<uint4*>(row + ((ch ^ (c_row & 7)) << 4)) =
                make_uint4(Fmt<T>::pack2(o[8 * ch] * inv, o[8 * ch + 1] * inv), Fmt<T>::pack2(o[8 * ch + 2] * inv, o[8 * ch + 3] * inv),
                           Fmt<T>::pack2(o[8 * ch + 4] * inv, o[8 * ch + 5] * inv), Fmt<T>::pack2(o[8 * ch + 6] * inv, o[8 * ch + 7] * inv));
        } else {
          uint8_t* row = kvt + c_row * 128;              // rows >= C must be zero (K dimension of the output MMA)
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(row + (ch << 4)) = make_uint4(0, 0, 0, 0);
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(kKvTile));
      // ---- phase O: thread = token ----
      ptx::mbar_wait(bar(kD3Full), pi);
      ptx::tc_fence_after();
      ptx::named_bar_sync(1, 128);                                    // cst2 / lse2t written by all rows
#pragma unroll 1
      for (int rb = 0; rb < 2; ++rb) {
        const int n = 128 * rb + tid;
        const uint32_t cB = cD3 + 128 * rb;
        float a[64], t[64];
#pragma unroll
        for (int g = 0; g < 4; ++g) ptx::tmem_ld16(trow + cB + 16 * g, reinterpret_cast<uint32_t*>(a) + 16 * g);
#pragma unroll
        for (int g = 0; g < 4; ++g) ptx::tmem_ld16(trow + cB + 64 + 16 * g, reinterpret_cast<uint32_t*>(t) + 16 * g);
        ptx::tmem_ld_wait();
        // landmark columns in groups of 8; groups past C are skipped (uniform), per-landmark constants come as 16-byte broadcasts
        const int G8 = (C + 7) >> 3;
        auto ld8 = [](const float* src, float* dst) {
          const float4 x = *reinterpret_cast<const float4*>(src), y = *reinterpret_cast<const float4*>(src + 4);
          dst[0] = x.x; dst[1] = x.y; dst[2] = x.z; dst[3] = x.w; dst[4] = y.x; dst[5] = y.y; dst[6] = y.z; dst[7] = y.w;
        };
        float tsum = 0.f;
#pragma unroll
        for (int g8 = 0; g8 < 8; ++g8) {
          if (g8 < G8) {
            float ls[8];
            ld8(lse2t + 8 * g8, ls);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = 8 * g8 + j;
              t[c] = ex2(fmaf(t[c], scale_log2, -ls[j]));             // t_nc = softmax_n(scale q_bar_c . q_n)
              tsum += c < C ? t[c] : 0.f;
            }
          }
        }
        const float mean_t = __fdividef(tsum, (float)C);
        float mw = kNegInf;
#pragma unroll
        for (int g8 = 0; g8 < 8; ++g8) {
          if (g8 < G8) {
            float bh8[8], cs8[8];
            ld8(bhS + s * 64 + 8 * g8, bh8);
            ld8(cst2 + 8 * g8, cs8);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = 8 * g8 + j;
              const float alpha = bh8[j] + p.alpha_coeff * (t[c] - mean_t);
              // log2 of the importance weight, up to the per-token constant -|q_n|^2/2 that the softmax over c removes
              a[c] = lg2(fmaxf(alpha, 1e-8f)) + fmaf(scale_log2, a[c], cs8[j]);
              mw = fmaxf(mw, c < C ? a[c] : kNegInf);
            }
          }
        }
        float wsum = 0.f;
        uint32_t pk[32];
#pragma unroll
        for (int g8 = 0; g8 < 8; ++g8) {
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const int c = 8 * g8 + j;
            float w0 = 0.f, w1 = 0.f;
            if (g8 < G8) {
              w0 = c < C ? ex2(a[c] - mw) : 0.f;
              w1 = c + 1 < C ? ex2(a[c + 1] - mw) : 0.f;
            }
            wsum += w0 + w1;
            pk[c >> 1] = Fmt<T>::pack2(w0, w1);
          }
        }
        ptx::tmem_st16(trow + cB, pk);
        ptx::tmem_st16(trow + cB + 16, pk + 16);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar(kP2Full0 + rb));
        ptx::mbar_wait(bar(kOFull0 + rb), pi);
        ptx::tc_fence_after();
        float o[64];
#pragma unroll
        for (int g = 0; g < 4; ++g) ptx::tmem_ld16(trow + cO + 128 * rb + 16 * g, reinterpret_cast<uint32_t*>(o) + 16 * g);
        ptx::tmem_ld_wait();
        const float inv = __fdividef(1.0f, wsum);
        uint8_t* row = tile(s, 0) + n * 128;                           // the q tile is dead: every D3 MMA has completed
        if (n < NP) {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<uint4*>(row + ((ch ^ (n & 7)) << 4)) =
                make_uint4(Fmt<T>::pack2(o[8 * ch] * inv, o[8 * ch + 1] * inv), Fmt<T>::pack2(o[8 * ch + 2] * inv, o[8 * ch + 3] * inv),
                           Fmt<T>::pack2(o[8 * ch + 4] * inv, o[8 * ch + 5] * inv), Fmt<T>::pack2(o[8 * ch + 6] * inv, o[8 * ch + 7] * inv));
        }
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(1, 128);
        if (warp == 0 && ptx::elect_one()) {
          if (128 * rb < N) {                                          // rows >= N are clipped by the tensor map
            ptx::tma_store_4d(&t_o, ptx::smem_u32(tile(s, 0)) + rb * 16384, 0, h, 128 * rb, b);
            ptx::bulk_commit_group();
            ptx::bulk_wait_read0();
          }
          ptx::mbar_arrive(bar(kFree0 + s));
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(kEpiDone));
    }
    if (warp == 0 && ptx::elect_one()) ptx::bulk_wait_all();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 5) ptx::tmem_dealloc(tmem, 256);
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

static bool make_seq_map(CUtensorMap* tm, const void* ptr, long long sb, long long sn, long long sh, const LaraGeo& g, int io_dtype, int rows) {
  auto enc = get_encode();
  if (!enc) return false;
  const cuuint64_t dims[4] = {64, (cuuint64_t)g.H, (cuuint64_t)g.N, (cuuint64_t)g.B};
  const cuuint64_t strides[3] = {(cuuint64_t)sh * 2, (cuuint64_t)sn * 2, (cuuint64_t)sb * 2};
  const cuuint32_t box[4] = {64, 1, (cuuint32_t)rows, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tm, io_dtype == EVA_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims,
             strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [W_q ; W_k] (fp32 [64][64] each, row-major [out][in]) -> 16-bit [128][64] for the Linear MMA of phase L
template <typename T>
__global__ void pack_proj_weights(const float* __restrict__ wq, const float* __restrict__ wk, T* __restrict__ w16) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < 128 * 64) w16[idx] = from_f32<T>((idx < 64 * 64 ? wq : wk)[idx & 4095]);
}

template <typename T>
static cudaError_t launch_t(const LaraGeo& g, int io_dtype, const View& q, const View& k, const View& v, const float* ws, void* out,
                            const EvaAdaptive* proj, const float* noise, cudaStream_t st) {
  const int NP = (g.N + 15) & ~15;
  CUtensorMap tq, tk, tv, to;
  if (!make_seq_map(&tq, q.ptr, q.sb, q.sn, q.sh, g, io_dtype, NP) || !make_seq_map(&tk, k.ptr, k.sb, k.sn, k.sh, g, io_dtype, NP) ||
      !make_seq_map(&tv, v.ptr, v.sb, v.sn, v.sh, g, io_dtype, NP) ||
      !make_seq_map(&to, out, (long long)g.N * g.H * 64, (long long)g.H * 64, 64, g, io_dtype, 128))
    return cudaErrorInvalidValue;
  Params p{};
  p.B = g.B; p.H = g.H; p.N = g.N; p.NP = NP; p.C = g.C; p.items = g.B * g.H;
  p.alpha_coeff = g.alpha_coeff; p.ws = ws; p.sl = NP * 128;
  CUtensorMap tw = tq;                       // placeholder when phase L is not fused
  if (proj) {                                // fused landmark phase: the workspace only holds the packed projection weights
    p.fuse = 1; p.side = g.side; p.gh = g.gh; p.gw = g.gw; p.mixed = g.mixed; p.has_proj = 1;
    p.b_q = proj->b_q; p.g_q = proj->ln_gain_q; p.beta_q = proj->ln_bias_q;
    p.b_k = proj->b_k; p.g_k = proj->ln_gain_k; p.beta_k = proj->ln_bias_k;
    p.ln_eps = proj->ln_eps; p.noise = noise;
    T* w16 = reinterpret_cast<T*>(const_cast<float*>(ws));
    pack_proj_weights<T><<<32, 256, 0, st>>>(proj->w_q, proj->w_k, w16);
    auto enc = get_encode();
    const cuuint64_t dims[2] = {64, 128};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, 128};
    const cuuint32_t estr[2] = {1, 1};
    if (!enc || enc(&tw, io_dtype == EVA_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w16, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  const int dyn = 2 * p.sl + 3 * 16384 + (2 * 256 + 2 * 128 + 3 * 64 + 384) * 4 + kNumBars * 8 + 16 + 1024;
  auto kern = lara_core_kernel<T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.items < 2 * sms ? p.items : 2 * sms;
  kern<<<grid, kThreads, dyn, st>>>(tq, tk, tv, to, tw, p);
  return cudaGetLastError();
}

}  // namespace laracore

bool lara_core_supported(const LaraGeo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask) {
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("EVA_SM100_DISABLE_FUSED"); disabled = (e && e[0] == '1') ? 1 : 0; }
  if (disabled) return false;
  if (g.D != 64 || g.mis_type != LARA_MIS_OPT || g.sample_mode != LARA_SAMPLE_SINGLE || g.S != g.C || g.C > 64 || mask) return false;
  if (g.N > 224 || g.N < 16) return false;
  if (io_dtype != EVA_F16 && io_dtype != EVA_BF16) return false;
  for (const View* x : {&q, &k, &v}) {
    if (x->sh * 2 % 16 || x->sn * 2 % 16 || x->sb * 2 % 16) return false;
    if (x->sh <= 0 || x->sn <= 0 || x->sb <= 0) return false;
    if (reinterpret_cast<uintptr_t>(x->ptr) % 16) return false;
  }
  return laracore::get_encode() != nullptr;
}

static int g_core_launches = 0;
// diagnostic (not part of the public ABI): how many times the tcgen05 core has been launched by this process
extern "C" int eva_debug_lara_core_launches(void) { return g_core_launches; }

// proj != NULL: the landmark phase is fused too (lara_landmark_kernel is not needed); see lara_core_fuses_landmarks
// pooled 2-D proposals with Linear + LayerNorm ('pool', 'pool-mixed'), at most 208 tokens: phase L runs inside the core kernel
bool lara_core_fuses_landmarks(const LaraGeo& g, const EvaAdaptive& proj) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("EVA_SM100_LARA_NO_FUSED_LANDMARKS"); off = (e && e[0] == '1') ? 1 : 0; }
  return !off && g.dims == 2 && !g.per_token_proj && g.mixed <= 1 && proj.w_q && proj.w_k && ((g.N + 15) & ~15) <= 208;
}

cudaError_t launch_lara_core(const LaraGeo& g, int io_dtype, const View& q, const View& k, const View& v, const float* ws, void* out,
                             const EvaAdaptive* proj, const float* noise, cudaStream_t st) {
  ++g_core_launches;
  if (io_dtype == EVA_F16) return laracore::launch_t<__half>(g, io_dtype, q, k, v, ws, out, proj, noise, st);
  return laracore::launch_t<__nv_bfloat16>(g, io_dtype, q, k, v, ws, out, proj, noise, st);
}

}  // namespace eva
