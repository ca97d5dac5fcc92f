// Window-attention gradient on tcgen05 for EVERY geometry with head_dim 64 and 16-bit I/O (the backward twin of
// eva_window_tc_sm100.cu): halos, 1-D, padding masks, causal, chunk keys or none, any window length.  The DeiT geometries keep
// eva_bwd_sm100.cu (one window per CTA, TMA boxes); float32 I/O and the other head dims keep window_attn_bwd_kernel (CUDA cores).
//
// CTA iteration = (batch x head, window, block of 128 query rows), 8 warps, one CTA per SM (all 512 tensor-memory columns):
//   rows: Q, dO -> 128-byte-swizzled tiles; delta_r = <dO_r, O_r>
//   pass 0 over the key tiles: S = Q K^T, finished logits, online (max, sum) -> lse_r          (nothing is kept by the forward)
//   pass 1 over the key tiles:
//     S = Q K^T, dP = dO V^T                         (K = [local k rows | k_bar rows], V = [local v rows | beta rows])
//     P = exp2(logit - lse), dS = P o (dP - delta) where the logit still depends on q . k (not overwritten by a mask rule)
//        -> 16-bit row-major tiles [128 rows][2 x 64 keys] in shared memory
//     dQ [128 x 64] += dS K            (A K-major, B = the k tile MN-major; accumulates over the key tiles in tensor memory)
//     dK [128 keys x 64] = dS^T Q      (A = the same dS tile read MN-major: two 64-key atoms 16 KB apart; B = the q tile MN-major)
//     dV [128 keys x 64] = P^T dO
//     dK / dV rows -> red.global.add.v4.f32 (token rows of dk / dv, or d k_bar / d beta for chunk keys)
//   dQ rows -> float4 stores (each query row belongs to one block)
#include <stdlib.h>

#include "common.cuh"
#include "fused_common.cuh"
#include "launch.h"
#include "sm100_ptx.cuh"

namespace eva {
namespace bwdgen {

using fused::IoFmt;
using fused::tile_off;
using fused::tmem_ld_cols;
using fused::ex2;

constexpr int kThreads = 256;
constexpr int kQ = 0, kG = 16384, kK = 32768, kV = 49152, kP = 65536, kdS = 98304, kMisc = 131072;
constexpr int kFac = kMisc, kKtok = kFac + 2048, kQtok = kKtok + 512, kQpad = kQtok + 512, kDelta = kQpad + 512, kPm = kDelta + 512,
              kPl = kPm + 1024, kBiasT = kPl + 1024, kBar = kBiasT + 8 * 32 * 17 * 4, kSlot = kBar + 16, kSmemBytes = kSlot + 16 + 1024;
constexpr uint32_t kTmemCols = 512, cS = 0, cDP = 128, cDQ = 256, cDK = 320, cDV = 384;

struct Params {
  Geo g;
  View q, k, v;
  const uint8_t* mask;
  const float* kbar; const float* beta; const float* bias;
  long long bias_sh;
  const void* out; const void* dout;
  const float* lse;              // log2-domain log-sum-exp per query row [B, H, N] kept by the forward, or NULL (recomputed: pass 0)
  float* dq; float* dk; float* dv; float* dkbar; float* dbeta; float* dbias;
  long long total;
  int trace;
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
eva_window_bwd_gen_kernel(const Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (ptx::smem_u32(raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Geo& g = p.g;
  // [128] per key: logit = s * mul + add with (1, 0) live | (0, mask_fill) masked | (0, -inf) absent, and the causal rule as two
  // thresholds: overwritten by -5e4 when row index < thr_row (local keys: key slot - halo) or row chunk < thr_chunk (chunk keys: c + 1)
  float4* kfac = reinterpret_cast<float4*>(sm + kFac);
  int* ktok = reinterpret_cast<int*>(sm + kKtok);        // [128] destination of a key's gradient: token >= 0 | -1 none | -2 - c chunk c
  int* qtok = reinterpret_cast<int*>(sm + kQtok);
  int* qpad = reinterpret_cast<int*>(sm + kQpad);
  float* delta = reinterpret_cast<float*>(sm + kDelta);
  float* pm = reinterpret_cast<float*>(sm + kPm);        // [2][128]
  float* pl = reinterpret_cast<float*>(sm + kPl);        // [2][128]
  float* biasT = reinterpret_cast<float*>(sm + kBiasT);  // [8 warps][32][17] bias transposition
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + kSlot);
  const uint32_t bar = ptx::smem_u32(sm + kBar);
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(slot), kTmemCols);
  if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  constexpr uint32_t fmt = IoFmt<T>::kUmma;
  constexpr uint32_t id_s = ptx::umma_idesc(fmt, fmt, 0, 0, 128, 128);
  constexpr uint32_t id_dq = ptx::umma_idesc(fmt, fmt, 0, 1, 128, 64);
  constexpr uint32_t id_dk = ptx::umma_idesc(fmt, fmt, 1, 1, 128, 64);
  const uint64_t dQd = ptx::umma_desc_sw128(ptx::smem_u32(sm + kQ)), dGd = ptx::umma_desc_sw128(ptx::smem_u32(sm + kG));
  const uint64_t dKd = ptx::umma_desc_sw128(ptx::smem_u32(sm + kK)), dVd = ptx::umma_desc_sw128(ptx::smem_u32(sm + kV));
  const uint64_t dSd = ptx::umma_desc_sw128(ptx::smem_u32(sm + kdS));
  const uint64_t dStd = ptx::umma_desc_sw128_mn(ptx::smem_u32(sm + kdS), 16384), dPtd = ptx::umma_desc_sw128_mn(ptx::smem_u32(sm + kP), 16384);
  const int qr = warp & 3, hf = warp >> 2;
  const int r = 32 * qr + lane;                          // query row of the block = TMEM lane (also: key row of the dK / dV tiles)
  const uint32_t trow = tmem + ((uint32_t)(32 * qr) << 16);
  const float scale = 0.125f;
  const int n_rb = (g.L + 127) / 128;
  const int n_keys = g.J + g.n_chunks;
  const long long HD = (long long)g.H * 64;
  uint32_t ph = 0;

  for (long long item = blockIdx.x; item < p.total; item += gridDim.x) {
    const int rb = (int)(item % n_rb), win = (int)((item / n_rb) % g.n_windows);
    const int bh = (int)(item / ((long long)n_rb * g.n_windows));
    const int b = bh / g.H, h = bh % g.H;
    const long long bias_off = (long long)h * p.bias_sh;
    __syncthreads();
    if (tid < 128) {
      const int li = rb * 128 + tid;
      const int tok = li < g.L ? group_token(g, win, li, g.window, 0) : -1;
      qtok[tid] = tok;
      qpad[tid] = (tok >= 0 && p.mask) ? (int)p.mask[(long long)b * g.N + tok] : 0;
    }
    for (int idx = tid; idx < 128 * 8; idx += kThreads) {
      const int row = idx >> 3, piece = idx & 7;
      const int li = rb * 128 + row;
      const int tok = li < g.L ? group_token(g, win, li, g.window, 0) : -1;
      uint4 zq = make_uint4(0, 0, 0, 0), zg = zq;
      float part = 0.f;
      if (tok >= 0) {
        zq = __ldg(reinterpret_cast<const uint4*>(p.q.row<T>(b, tok, h)) + piece);
        const long long o = ((long long)b * g.N + tok) * HD + (long long)h * 64;
        zg = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.dout) + o) + piece);
        const uint4 zo = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.out) + o) + piece);
        const uint32_t* a = reinterpret_cast<const uint32_t*>(&zg);
        const uint32_t* c = reinterpret_cast<const uint32_t*>(&zo);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 x = IoFmt<T>::unpack2(a[i]), y = IoFmt<T>::unpack2(c[i]);
          part = fmaf(x.x, y.x, fmaf(x.y, y.y, part));
        }
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      if (piece == 0) delta[row] = part;
      const int off = tile_off(row, 8 * piece);
      *reinterpret_cast<uint4*>(sm + kQ + off) = zq;
      *reinterpret_cast<uint4*>(sm + kG + off) = zg;
    }
    const int last_visible = rb * 128 + 127 + g.ext;
    auto skip_tile = [&](int kt0) { return g.causal && kt0 > last_visible && kt0 + 128 <= g.J; };
    auto load_tile = [&](int kt0, bool with_v) -> int {
      int flags = 0;
      for (int idx = tid; idx < 128 * 8; idx += kThreads) {
        const int j = idx >> 3, piece = idx & 7;
        const int gj = kt0 + j;
        uint4 zk = make_uint4(0, 0, 0, 0), zv = zk;
        int flag = 0, dest = -1;
        if (gj < g.J) {
          const int tok = group_token(g, win, gj, g.window, g.ext);
          if (tok >= 0) {
            zk = __ldg(reinterpret_cast<const uint4*>(p.k.row<T>(b, tok, h)) + piece);
            if (with_v) zv = __ldg(reinterpret_cast<const uint4*>(p.v.row<T>(b, tok, h)) + piece);
            flag = (p.mask && p.mask[(long long)b * g.N + tok]) ? 1 : 0;
            dest = tok;
          } else {
            flag = 1;
          }
        } else if (gj < n_keys) {
          const long long base = ((long long)bh * g.n_chunks + (gj - g.J)) * 64 + 8 * piece;
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.kbar + base)), a1 = __ldg(reinterpret_cast<const float4*>(p.kbar + base) + 1);
          zk = make_uint4(IoFmt<T>::pack2(a0.x, a0.y), IoFmt<T>::pack2(a0.z, a0.w), IoFmt<T>::pack2(a1.x, a1.y), IoFmt<T>::pack2(a1.z, a1.w));
          if (with_v) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + base)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + base) + 1);
            zv = make_uint4(IoFmt<T>::pack2(b0.x, b0.y), IoFmt<T>::pack2(b0.z, b0.w), IoFmt<T>::pack2(b1.x, b1.y), IoFmt<T>::pack2(b1.z, b1.w));
          }
          dest = -2 - (gj - g.J);
        } else {
          flag = 2;
        }
        const int off = tile_off(j, 8 * piece);
        *reinterpret_cast<uint4*>(sm + kK + off) = zk;
        if (with_v) *reinterpret_cast<uint4*>(sm + kV + off) = zv;
        if (piece == 0) {
          ktok[j] = dest;
          const int never = -2147483647 - 1;
          const int thr_row = (g.causal && gj < g.J) ? gj - g.ext : never;
          const int thr_chunk = (g.causal && gj >= g.J && gj < n_keys) ? gj - g.J + 1 : never;
          kfac[j] = make_float4(flag == 0 ? 1.f : 0.f, flag == 0 ? 0.f : (flag == 1 ? g.mask_fill : kNegInf), __int_as_float(thr_row),
                                __int_as_float(thr_chunk));
        }
        flags |= flag;
      }
      return flags;
    };
    // S (and dP) of the staged tiles; returns the block-wide OR of the key flags
    auto mma_s = [&](int my_flags, bool with_dp) -> bool {
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      const int any = __syncthreads_or(my_flags);
      if (tid == 0) {
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cS, dQd + 2 * ks, dKd + 2 * ks, id_s, ks > 0);
        if (with_dp) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cDP, dGd + 2 * ks, dVd + 2 * ks, id_s, ks > 0);
        }
        ptx::umma_commit(bar);
      }
      ptx::mbar_wait(bar, ph & 1);
      ++ph;
      ptx::tc_fence_after();
      return any != 0;
    };
    __syncthreads();
    const int tq_row = qtok[r], li_row = rb * 128 + r, qp_row = qpad[r];
    const int tq_chunk = (g.causal && g.chunk > 0 && tq_row >= 0) ? tq_row / g.chunk : 0;
    const float* brow = (p.bias && tq_row >= 0) ? p.bias + bias_off + (long long)li_row * g.J : nullptr;
    const bool rules = p.bias != nullptr || g.causal || g.mask_queries;
    const bool row_masked = g.mask_queries && qp_row;
    // finished logits of my 64 columns (log2 domain); live: bit j set when the logit still depends on q . k
    auto logits = [&](int kt0, bool tile_flags, float (&x)[64], uint32_t (&live)[2]) {
      tmem_ld_cols<64>(trow + cS + 64 * hf, reinterpret_cast<uint32_t*>(x));
      ptx::tmem_ld_wait();
      live[0] = live[1] = 0xffffffffu;
      if (!rules && !tile_flags) {
#pragma unroll
        for (int j = 0; j < 64; ++j) x[j] *= scale * kLog2e;
        return;
      }
      const int c0 = kt0 + 64 * hf;
      const float4* kf = kfac + 64 * hf;
      const bool any_row_masked = __any_sync(0xffffffffu, row_masked);
      live[0] = live[1] = 0u;
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        float bb[16];
        if (p.bias) {                                  // coalesced: two rows x 16 columns per load instruction, transposed through
          float* tb = biasT + warp * (32 * 17);        // shared memory (a per-thread row read would touch 32 lines per instruction)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int row = 2 * i + (lane >> 4), col = lane & 15;
            const int li2 = rb * 128 + 32 * qr + row, gj = c0 + 16 * blk + col;
            tb[row * 17 + col] = (li2 < g.L && gj < g.J) ? __ldg(p.bias + bias_off + (long long)li2 * g.J + gj) : 0.f;
          }
          __syncwarp();
#pragma unroll
          for (int e = 0; e < 16; ++e) bb[e] = tb[lane * 17 + e];
          __syncwarp();
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) bb[e] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int j = 16 * blk + e;
          const float4 f = kf[j];
          float sv = fmaf(fmaf(x[j], scale, bb[e]), f.x, f.y);
          const bool cm = (li_row < __float_as_int(f.z)) | (tq_chunk < __float_as_int(f.w));
          bool lv = f.x != 0.f;
          if (any_row_masked) {                        // padded query rows of the causal layer (warp-uniform branch)
            const bool rm = row_masked && (c0 + j < g.J);
            sv = rm ? g.mask_fill : sv;
            lv = lv && !rm;
          }
          sv = cm ? kMaskVal : sv;
          lv = lv && !cm;
          x[j] = sv * kLog2e;
          live[j >> 5] |= (lv ? 1u : 0u) << (j & 31);
        }
      }
    };
    long long tk[8];
    const bool tr_on = p.trace && item == blockIdx.x + gridDim.x && blockIdx.x == 0;
    if (tr_on) tk[0] = clock64();
    // ---- pass 0: lse of every row (skipped when the forward kept it) ----
    float mrow = kNegInf, lrow = 0.f;
    for (int kt0 = 0; kt0 < n_keys && !p.lse; kt0 += 128) {
      if (skip_tile(kt0)) continue;
      __syncthreads();
      const bool tf = mma_s(load_tile(kt0, false), false);
      float x[64];
      uint32_t live[2];
      logits(kt0, tf, x, live);
      float mloc = kNegInf;
#pragma unroll
      for (int j = 0; j < 64; ++j) mloc = fmaxf(mloc, x[j]);
      pm[hf * 128 + r] = mloc;
      __syncthreads();
      const float mnew = fmaxf(mrow, fmaxf(pm[r], pm[128 + r]));
      const float alpha = (mnew == kNegInf || mrow == mnew) ? 1.f : ex2(mrow - mnew);
      mrow = mnew;
      lrow *= alpha;
      if (mrow != kNegInf) {
#pragma unroll
        for (int j = 0; j < 64; ++j) lrow += ex2(x[j] - mrow);
      }
    }
    pl[hf * 128 + r] = lrow;
    __syncthreads();
    const float ltot = pl[r] + pl[128 + r];
    bool row_ok = tq_row >= 0 && mrow != kNegInf && ltot > 0.f;
    float lse = row_ok ? mrow + log2f(ltot) : 0.f;
    if (p.lse) {
      lse = tq_row >= 0 ? __ldg(p.lse + (long long)bh * g.N + tq_row) : 0.f;
      row_ok = tq_row >= 0 && lse > -1e30f && lse < 1e30f;
    }
    const float dl = delta[r];
    float* dbrow = (p.dbias && tq_row >= 0) ? p.dbias + bias_off + (long long)li_row * g.J : nullptr;
    // ---- pass 1 ----
    bool first = true;
    if (tr_on) tk[1] = clock64();
    for (int kt0 = 0; kt0 < n_keys; kt0 += 128) {
      if (skip_tile(kt0)) continue;
      __syncthreads();
      if (tr_on) tk[2] = clock64();
      const int lf = load_tile(kt0, true);
      if (tr_on) tk[3] = clock64();
      const bool tf = mma_s(lf, true);
      if (tr_on) tk[4] = clock64();
      float x[64];
      uint32_t live[2];
      logits(kt0, tf, x, live);
      if (tr_on) tk[5] = clock64();
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        float dp[16];
        ptx::tmem_ld16(trow + cDP + 64 * hf + 16 * blk, reinterpret_cast<uint32_t*>(dp));
        ptx::tmem_ld_wait();
        uint32_t pk[8], pp[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const int j = 16 * blk + e;
          const float p0 = row_ok ? ex2(x[j] - lse) : 0.f, p1 = row_ok ? ex2(x[j + 1] - lse) : 0.f;
          const float d0 = ((live[j >> 5] >> (j & 31)) & 1u) ? p0 * (dp[e] - dl) : 0.f;
          const float d1 = ((live[(j + 1) >> 5] >> ((j + 1) & 31)) & 1u) ? p1 * (dp[e + 1] - dl) : 0.f;
          if (dbrow) {
            const int gj = kt0 + 64 * hf + j;
            if (gj < g.J && d0 != 0.f) atomicAdd(dbrow + gj, d0);
            if (gj + 1 < g.J && d1 != 0.f) atomicAdd(dbrow + gj + 1, d1);
          }
          pk[e >> 1] = IoFmt<T>::pack2(d0, d1);
          pp[e >> 1] = IoFmt<T>::pack2(p0, p1);
        }
        const int o0 = hf * 16384 + tile_off(r, 16 * blk), o1 = hf * 16384 + tile_off(r, 16 * blk + 8);
        *reinterpret_cast<uint4*>(sm + kdS + o0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(sm + kdS + o1) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        *reinterpret_cast<uint4*>(sm + kP + o0) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
        *reinterpret_cast<uint4*>(sm + kP + o1) = make_uint4(pp[4], pp[5], pp[6], pp[7]);
      }
      if (tr_on) tk[6] = clock64();
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          ptx::umma_ss(tmem + cDQ, dSd + (uint64_t)((ks >> 2) * (16384 >> 4) + 2 * (ks & 3)), dKd + 128 * ks, id_dq, (first ? 0u : 1u) | (ks > 0 ? 1u : 0u));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) ptx::umma_ss(tmem + cDK, dStd + 128 * ks, dQd + 128 * ks, id_dk, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) ptx::umma_ss(tmem + cDV, dPtd + 128 * ks, dGd + 128 * ks, id_dk, ks > 0);
        ptx::umma_commit(bar);
      }
      first = false;
      ptx::mbar_wait(bar, ph & 1);
      ++ph;
      ptx::tc_fence_after();
      if (tr_on) tk[7] = clock64();
      // dK / dV rows of the tile: lane = key row
      {
        float y[32], z[32];
        tmem_ld_cols<32>(trow + cDK + 32 * hf, reinterpret_cast<uint32_t*>(y));
        tmem_ld_cols<32>(trow + cDV + 32 * hf, reinterpret_cast<uint32_t*>(z));
        ptx::tmem_ld_wait();
        const int dest = ktok[r];
        float* pk_ = nullptr;
        float* pv_ = nullptr;
        if (dest >= 0) {
          const long long base = (((long long)b * g.N + dest) * g.H + h) * 64 + 32 * hf;
          pk_ = p.dk + base; pv_ = p.dv + base;
        } else if (dest <= -2) {
          const long long base = ((long long)bh * g.n_chunks + (-2 - dest)) * 64 + 32 * hf;
          pk_ = p.dkbar + base; pv_ = p.dbeta + base;
        }
        if (pk_) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(pk_ + i), "f"(y[i] * scale), "f"(y[i + 1] * scale),
                         "f"(y[i + 2] * scale), "f"(y[i + 3] * scale) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(pv_ + i), "f"(z[i]), "f"(z[i + 1]), "f"(z[i + 2]), "f"(z[i + 3]) : "memory");
          }
        }
      }
      ptx::tc_fence_before();
      if (tr_on && (tid == 0 || tid == 200) && kt0 == 0)
        printf("bwd gen trace tid %d: pass0 %lld | load %lld | S,dP mma %lld | logits %lld | P,dS %lld | mma2 %lld | dK,dV out %lld\n", tid,
               tk[1] - tk[0], tk[3] - tk[2], tk[4] - tk[3], tk[5] - tk[4], tk[6] - tk[5], tk[7] - tk[6], clock64() - tk[7]);
    }
    // ---- dQ rows ----
    {
      float y[32];
      tmem_ld_cols<32>(trow + cDQ + 32 * hf, reinterpret_cast<uint32_t*>(y));
      ptx::tmem_ld_wait();
      if (tq_row >= 0) {
        float* dst = p.dq + (((long long)b * g.N + tq_row) * g.H + h) * 64 + 32 * hf;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(first ? 0.f : y[i] * scale), "f"(first ? 0.f : y[i + 1] * scale),
                       "f"(first ? 0.f : y[i + 2] * scale), "f"(first ? 0.f : y[i + 3] * scale) : "memory");
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, kTmemCols);
}

}  // namespace bwdgen

bool window_bwd_gen_supported(const Geo& g, int io_dtype) {
  return bwd_tc_enabled() && g.D == 64 && (io_dtype == EVA_F16 || io_dtype == EVA_BF16) && g.J + g.n_chunks >= 1;
}

cudaError_t launch_window_bwd_gen(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask,
                                  const float* kbar, const float* beta, const float* bias, long long bias_sh, const void* out,
                                  const void* dout, float* dq, float* dk, float* dv, float* dkbar, float* dbeta, float* dbias,
                                  cudaStream_t st, const float* lse) {
  bwdgen::Params p;
  p.g = g; p.q = q; p.k = k; p.v = v; p.mask = mask;
  p.kbar = kbar; p.beta = beta; p.bias = bias; p.bias_sh = bias_sh;
  p.out = out; p.dout = dout; p.lse = lse;
  p.dq = dq; p.dk = dk; p.dv = dv; p.dkbar = dkbar; p.dbeta = dbeta; p.dbias = dbias;
  p.total = (long long)((g.L + 127) / 128) * g.n_windows * g.B * g.H;
  static const int trace = [] { const char* e = getenv("EVA_SM100_TRACE"); return (e && e[0] == '1') ? 1 : 0; }();
  p.trace = trace;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long grid = p.total < sms ? p.total : sms;
  note_bwd_tc_launch();
  cudaError_t e;
  if (io_dtype == EVA_F16) {
    e = cudaFuncSetAttribute(bwdgen::eva_window_bwd_gen_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwdgen::kSmemBytes);
    if (e != cudaSuccess) return e;
    bwdgen::eva_window_bwd_gen_kernel<__half><<<(unsigned)grid, bwdgen::kThreads, bwdgen::kSmemBytes, st>>>(p);
  } else {
    e = cudaFuncSetAttribute(bwdgen::eva_window_bwd_gen_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwdgen::kSmemBytes);
    if (e != cudaSuccess) return e;
    bwdgen::eva_window_bwd_gen_kernel<__nv_bfloat16><<<(unsigned)grid, bwdgen::kThreads, bwdgen::kSmemBytes, st>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace eva
