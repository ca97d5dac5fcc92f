// ScatterBrain window stage on tcgen05 / TMEM for sm_100a (scatterbrain_attention.py:95-160): 'favorp' log-features, 64 random
// features, head_dim 64, 16-bit I/O, halo-free windows of at most 64 tokens.  The SIMT kernel of rfa_kernels.cu keeps everything else.
//
// One PAIR of windows of one (batch, head) item per iteration, 256 threads (TMEM lane = row of the pair's tile: rows 0-63 window a,
// 64-127 window b; warps w and w + 4 share a lane quarter and split the columns of every epilogue), with W' = d^-1/4 W in 16 bits
// and the item's global statistics G = sum_n phi(k_n) v_n, gs = sum_n phi(k_n), mx (per
// feature; rfa_favorp_tc_kernel<kLogF>) in global memory:
//   M1  DDk = [Ka ; Kb] W'^T                      E1  PK = exp(DDk - |k|^2 term - log(m)/2 - mx)  (0 for padding)      -> 16-bit tile
//   M2  Lc_w = PK_w^T V_w, ls_w = PK_w^T 1 (w = a, b; M = 64 features), R = [Qa ; Qb] W'^T
//                                                  E2  thread = feature: values (G - Lc_w) / max(gs - ls_w, 1e-3) -> 16-bit tile
//                                                      KVS [a features ; b features][d], log-mass nl_w = log(sum outside the window)
//   M3  S = [Qa ; Qb] [Ka ; Kb]^T (128 x 128; the diagonal blocks are used)
//                                                  E3  thread = query row: local logits (scale, bias, -inf for padding) and feature
//                                                      logits R - |q|^2 term - log(m)/2 + nl_w, ONE softmax over both, P (16-bit,
//                                                      block-diagonal [local a, local b | features a, features b]) -> tensor memory
//   M4  O = P [Va ; Vb ; KVS]  (A from tensor memory, K = 256)
//                                                  E4  O / row sum -> 128-byte row stores
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "fused_common.cuh"
#include "launch.h"
#include "sm100_ptx.cuh"

namespace eva {
namespace sbtc {

using fused::tmem_ld_cols;
using fused::ex2;

constexpr int kThreads = 256;       // warps w and w + 4 share a TMEM lane quarter and split the columns of every epilogue
constexpr uint32_t cDD = 0, cLa = 64, cLsa = 128, cLb = 144, cLsb = 208, cS = 64, cP = 64, cO = 192;
constexpr int kQ = 0, kK = 16384, kV = 32768, kPK = 49152, kKVS = 65536, kW = 81920, kOnes = 90112, kNl = 98304, kMxs = kNl + 512,
              kDead = kMxs + 256, kTok = kDead + 256, kDmask = kTok + 1024, kPm = kDmask + 16, kPd = kPm + 1024, kBar = kPd + 1024, kTmemPtr = kBar + 16, kBias = kTmemPtr + 16;     // bias: L * L floats at the end

struct Params {
  int B, H, N, items;
  int dims, gh, gw, w, L, n_windows;
  const float* proj;       // [H, 64, 64]
  const uint8_t* mask;     // [B, N] or NULL
  const float* bias;       // [H, L, L] or NULL
  const float* stabv;      // [items][64]
  const float* part;       // [items][64 * 64 + 64]
  int trace;               // EVA_SM100_TRACE=1: phase clocks of one pair (block 0, thread 0)
};

template <typename T> struct Fmt;
template <> struct Fmt<__half> { static constexpr uint32_t kUmma = ptx::kFmtF16; };
template <> struct Fmt<__nv_bfloat16> { static constexpr uint32_t kUmma = ptx::kFmtBF16; };

__device__ __forceinline__ int window_token(const Params& p, int g, int l) {
  if (p.dims == 2) {
    const int ngx = p.gw / p.w;
    return ((g / ngx) * p.w + l / p.w) * p.gw + (g % ngx) * p.w + l % p.w;
  }
  return g * p.w + l;
}

template <typename T>
__device__ __forceinline__ float row_sq(const uint8_t* tile, int row) {
  float s = 0.f;
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    const uint4 raw = *reinterpret_cast<const uint4*>(tile + row * 128 + ((ch ^ (row & 7)) << 4));
    const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) { const float2 f = Pair16<T>::up(w4[u]); s = fmaf(f.x, f.x, fmaf(f.y, f.y, s)); }
  }
  return s;
}

template <typename T>
__device__ __forceinline__ void store_row16(uint8_t* tile, int row, const float (&f)[64]) {
#pragma unroll
  for (int ch = 0; ch < 8; ++ch)
    *reinterpret_cast<uint4*>(tile + row * 128 + ((ch ^ (row & 7)) << 4)) =
        make_uint4(Pair16<T>::pk(f[8 * ch], f[8 * ch + 1]), Pair16<T>::pk(f[8 * ch + 2], f[8 * ch + 3]),
                   Pair16<T>::pk(f[8 * ch + 4], f[8 * ch + 5]), Pair16<T>::pk(f[8 * ch + 6], f[8 * ch + 7]));
}

template <typename T>
__device__ __forceinline__ void store_half16(uint8_t* tile, int row, int hf, const float (&f)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(tile + row * 128 + (((4 * hf + c) ^ (row & 7)) << 4)) =
        make_uint4(Pair16<T>::pk(f[8 * c], f[8 * c + 1]), Pair16<T>::pk(f[8 * c + 2], f[8 * c + 3]),
                   Pair16<T>::pk(f[8 * c + 4], f[8 * c + 5]), Pair16<T>::pk(f[8 * c + 6], f[8 * c + 7]));
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2) sb_window_tc_kernel(const View q, const View k, const View v, T* __restrict__ out, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* const sm = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qr = warp & 3, hf = warp >> 2, r = 32 * qr + lane;     // my TMEM lane quarter, column half, row of the pair's tile
  float* const pm = reinterpret_cast<float*>(sm + kPm);          // [2][128] partial row maxima of the two column halves
  float* const pd = reinterpret_cast<float*>(sm + kPd);          // [2][128] partial row sums
  const uint32_t bar = ptx::smem_u32(sm + kBar);
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + kTmemPtr);
  float* const nl = reinterpret_cast<float*>(sm + kNl);          // [2][64]
  float* const mxs = reinterpret_cast<float*>(sm + kMxs);        // [64]
  uint8_t* const dead_all = sm + kDead;                          // [2][128] key row is padding / absent (per pair parity)
  int* const tokS = reinterpret_cast<int*>(sm + kTok);           // [2][128] token of every row of the pair (-1: no row), per pair parity
  uint32_t* const dmask = reinterpret_cast<uint32_t*>(sm + kDmask);   // [4] the dead flags of the pair as one ballot per warp
  float* const biasS = reinterpret_cast<float*>(sm + kBias);     // [L][L]
  constexpr uint32_t fmt = Fmt<T>::kUmma;
  constexpr uint32_t id_dd = ptx::umma_idesc(fmt, fmt, 0, 0, 128, 64);
  constexpr uint32_t id_kv = ptx::umma_idesc(fmt, fmt, 1, 1, 64, 64);
  constexpr uint32_t id_ks = ptx::umma_idesc(fmt, fmt, 1, 1, 64, 16);
  constexpr uint32_t id_s = ptx::umma_idesc(fmt, fmt, 0, 0, 128, 128);
  constexpr uint32_t id_o = ptx::umma_idesc(fmt, fmt, 0, 1, 128, 64);
  {
    const uint32_t one2 = Pair16<T>::pk(1.0f, 1.0f);
    for (int i = tid; i < 8192 / 16; i += kThreads) reinterpret_cast<uint4*>(sm + kOnes)[i] = make_uint4(one2, one2, one2, one2);
  }
  if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr)), 256);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t trow = tmem + ((uint32_t)(32 * qr) << 16);
  const uint64_t dQ = ptx::umma_desc_sw128(ptx::smem_u32(sm + kQ)), dK = ptx::umma_desc_sw128(ptx::smem_u32(sm + kK));
  const uint64_t dV = ptx::umma_desc_sw128(ptx::smem_u32(sm + kV)), dPK = ptx::umma_desc_sw128(ptx::smem_u32(sm + kPK));
  const uint64_t dKVS = ptx::umma_desc_sw128(ptx::smem_u32(sm + kKVS)), dW = ptx::umma_desc_sw128(ptx::smem_u32(sm + kW));
  const uint64_t dOnes = ptx::umma_desc_sw128(ptx::smem_u32(sm + kOnes));
  const float dn = 0.35355339059327373f, half_dn2 = 0.5f * dn * dn, hlm = 0.5f * 4.1588830833596715f, scale = 0.125f;
  uint32_t phase = 0;
  auto mma_wait = [&]() { ptx::mbar_wait(bar, phase & 1); ++phase; ptx::tc_fence_after(); };
  auto hand_over = [&]() { ptx::fence_proxy_async_smem(); ptx::tc_fence_before(); __syncthreads(); };
  const int L = p.L, pairs = (p.n_windows + 1) >> 1;
  const int w2 = r >> 6, li = r & 63;                      // my row: window a / b of the pair, slot inside it
  // rows of the two windows of pair `pr_` of item `item_` -> tile `which` (0 q, 1 k, 2 v), cp.async with zero fill where there is no
  // row; the key flags of that pair go to the parity buffer of pair counter `np_`
  auto issue_rows = [&](int item_, int pr_, int which, uint32_t np_) {
    const int b_ = item_ / p.H, h_ = item_ % p.H;
    const View& x = which == 0 ? q : (which == 1 ? k : v);
    uint8_t* const dst = sm + (which == 0 ? kQ : (which == 1 ? kK : kV));
#pragma unroll
    for (int it = 0; it < 1024 / kThreads; ++it) {
      const int idx = it * kThreads + tid, row = idx >> 3, ch = idx & 7;
      const int tk_ = tokS[128 * (np_ & 1) + row];                 // (index arithmetic with divisions: once per row and pair)
      const bool ok = tk_ >= 0;
      const uint4* src = reinterpret_cast<const uint4*>(x.row<T>(b_, ok ? tk_ : 0, h_)) + ch;
      const uint32_t d = ptx::smem_u32(dst + row * 128 + ((ch ^ (row & 7)) << 4));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    if (which == 1 && tid < 128) {
      const int tk_ = tokS[128 * (np_ & 1) + tid];
      dead_all[128 * (np_ & 1) + tid] = (tk_ < 0 || (p.mask && p.mask[(long long)b_ * p.N + tk_])) ? 1 : 0;
    }
  };
  uint32_t np = 0;                                  // pairs processed by this CTA
  long long tk[12];
#define SB_MARK(i) if (p.trace && np == 5 && blockIdx.x == 0) tk[i] = clock64();
  int h_loaded = -1;
  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    const int b = item / p.H, h = item % p.H;
    if (h != h_loaded) {
      const float* W = p.proj + (long long)h * 4096;
      for (int idx = tid; idx < 512; idx += kThreads) {
        const int row = idx >> 3, ch = idx & 7;
        const float4 a = __ldg(reinterpret_cast<const float4*>(W + row * 64 + 8 * ch)), c = __ldg(reinterpret_cast<const float4*>(W + row * 64 + 8 * ch) + 1);
        *reinterpret_cast<uint4*>(sm + kW + row * 128 + ((ch ^ (row & 7)) << 4)) =
            make_uint4(Pair16<T>::pk(dn * a.x, dn * a.y), Pair16<T>::pk(dn * a.z, dn * a.w), Pair16<T>::pk(dn * c.x, dn * c.y), Pair16<T>::pk(dn * c.z, dn * c.w));
      }
      for (int idx = tid; idx < L * L; idx += kThreads) biasS[idx] = p.bias ? __ldg(p.bias + (long long)h * L * L + idx) : 0.f;
      h_loaded = h;
    }
    if (tid < 64) mxs[tid] = __ldg(p.stabv + (long long)item * 64 + tid);
    const float* part = p.part + (long long)item * (64 * 64 + 64);
    for (int pr = 0; pr < pairs; ++pr, ++np) {
      auto pair_token = [&](int pr_) { const int win_ = 2 * pr_ + w2; return (win_ < p.n_windows && li < L) ? window_token(p, win_, li) : -1; };
      if (np == 0) { if (hf == 0) tokS[r] = pair_token(pr); __syncthreads(); }
      // successor pair (possibly of the CTA's next item): its row tokens now, its rows under this pair's softmax / output MMA
      int nitem = item, npr = pr + 1;
      if (npr == pairs) { nitem = item + (int)gridDim.x; npr = 0; }
      const bool more = nitem < p.items;
      if (hf == 0) tokS[128 * ((np + 1) & 1) + r] = more ? pair_token(npr) : -1;
      const int tok = tokS[128 * (np & 1) + r];
      const bool have_row = tok >= 0;
      uint8_t* const dead = dead_all + 128 * (np & 1);
      SB_MARK(0)
      // ---- tiles of this pair: requested during the previous pair (q, k after its S MMA, v after its output MMA) ----
      if (np == 0) { issue_rows(item, pr, 0, np); issue_rows(item, pr, 1, np); issue_rows(item, pr, 2, np); }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      hand_over();
      SB_MARK(1)
      if (warp == 0 && ptx::elect_one()) {
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cDD, dK + 2 * ks, dW + 2 * ks, id_dd, ks > 0);
        ptx::umma_commit(bar);
      }
      mma_wait();
      SB_MARK(2)
      const float qsub = half_dn2 * row_sq<T>(sm + kQ, r) + hlm;         // (the q tile is handed to the next pair after the S MMA)
      {   // E1: exp-features of my key row, my half of the 64 features
        float f[32];
        tmem_ld_cols<32>(trow + cDD + 32 * hf, reinterpret_cast<uint32_t*>(f));
        ptx::tmem_ld_wait();
        const float sub = half_dn2 * row_sq<T>(sm + kK, r) + hlm;
        const bool dd_ = dead[r] != 0;
        if (hf == 0) {                              // (warp-uniform)
          const unsigned bal = __ballot_sync(0xffffffffu, dd_);
          if (lane == 0) dmask[qr] = bal;
        }
        const float live = dd_ ? 0.f : 1.f;         // branch-free: a select around every exponential compiled to a divergence region each
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = live * __expf(fminf(f[j] - sub - mxs[32 * hf + j], 0.f));      // (live rows: <= 0 by the definition of mxs)
        store_half16<T>(sm + kPK, r, hf, f);
      }
      SB_MARK(3)
      hand_over();
      if (warp == 0 && ptx::elect_one()) {
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cLa, dPK + 128 * ks, dV + 128 * ks, id_kv, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cLsa, dPK + 128 * ks, dOnes + 128 * ks, id_ks, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cLb, dPK + 512 + 128 * ks, dV + 512 + 128 * ks, id_kv, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cLsb, dPK + 512 + 128 * ks, dOnes + 128 * ks, id_ks, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cDD, dQ + 2 * ks, dW + 2 * ks, id_dd, ks > 0);      // R (DDk is dead)
        ptx::umma_commit(bar);
      }
      mma_wait();
      SB_MARK(4)
      {   // E2: thread = feature (lanes < 16 of each warp hold the M = 64 rows; the loads are warp-collective); warps 0-3 take
        const int c = 16 * qr + (lane & 15);        // window a, warps 4-7 window b
        const int ww = hf;
        const float gsum = __ldg(part + 64 * 64 + c);
        float lc[64];
        uint32_t ls0;
        tmem_ld_cols<64>(trow + (ww ? cLb : cLa), reinterpret_cast<uint32_t*>(lc));
        ptx::tmem_ld1(trow + (ww ? cLsb : cLsa), ls0);
        ptx::tmem_ld_wait();
        if (lane < 16) {
          const float ls = __uint_as_float(ls0);
          const float inv = 1.0f / fmaxf(gsum - ls, 1e-3f);
#pragma unroll
          for (int d4 = 0; d4 < 16; ++d4) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(part + c * 64) + d4);
            lc[4 * d4] = (g4.x - lc[4 * d4]) * inv; lc[4 * d4 + 1] = (g4.y - lc[4 * d4 + 1]) * inv;
            lc[4 * d4 + 2] = (g4.z - lc[4 * d4 + 2]) * inv; lc[4 * d4 + 3] = (g4.w - lc[4 * d4 + 3]) * inv;
          }
          store_row16<T>(sm + kKVS, 64 * ww + c, lc);
          // log_add_exp(glse, llse, mask = (1, -1)) of attn_utils.py:44-51; both log-sums are relative to the per-feature maximum
          const float glse = __logf(gsum), llse = __logf(ls), a = fmaxf(glse, llse);
          nl[64 * ww + c] = mxs[c] + a + __logf(__expf(glse - a) - __expf(llse - a) + 1e-5f);
        }
      }
      SB_MARK(5)
      hand_over();
      if (warp == 0 && ptx::elect_one()) {
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cS, dQ + 2 * ks, dK + 2 * ks, id_s, ks > 0);
        ptx::umma_commit(bar);
      }
      mma_wait();
      if (more) {       // the successor's q and k rows travel under the softmax and the output MMA
        issue_rows(nitem, npr, 0, np + 1); issue_rows(nitem, npr, 1, np + 1); }
      SB_MARK(6)
      {   // E3: joint softmax of my query row over [local keys of my window | the 64 feature keys]; my half of either
        float s[32], rr[32];
        tmem_ld_cols<32>(trow + cS + 64 * w2 + 32 * hf, reinterpret_cast<uint32_t*>(s));
        tmem_ld_cols<32>(trow + cDD + 32 * hf, reinterpret_cast<uint32_t*>(rr));
        ptx::tmem_ld_wait();
        const float* brow = biasS + (li < L ? li : 0) * L;
        const unsigned dm = dmask[2 * w2 + hf];     // the dead flags of my 32 local keys
        float mx = kNegInf;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int jj = 32 * hf + j;
          s[j] = (jj < L && !((dm >> j) & 1u)) ? fmaf(scale, s[j], brow[jj < L ? jj : 0]) : kNegInf;
          rr[j] = rr[j] - qsub + nl[64 * w2 + jj];
          mx = fmaxf(mx, fmaxf(s[j], rr[j]));
        }
        pm[128 * hf + r] = mx;
        __syncthreads();                            // the two column halves of a row meet
        mx = fmaxf(pm[r], pm[128 + r]);
        const float mx2 = -mx * kLog2e;             // (every row has finite feature logits: mx is finite)
        float rsum = 0.f;
        uint32_t pl[16], pr_[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float e0 = ex2(fmaf(s[j], kLog2e, mx2)), e1 = ex2(fmaf(s[j + 1], kLog2e, mx2)), f0 = ex2(fmaf(rr[j], kLog2e, mx2)), f1 = ex2(fmaf(rr[j + 1], kLog2e, mx2));
          rsum += (e0 + e1) + (f0 + f1);
          pl[j >> 1] = Pair16<T>::pk(e0, e1);
          pr_[j >> 1] = Pair16<T>::pk(f0, f1);
        }
        pd[128 * hf + r] = rsum;
        uint32_t zero[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) zero[j] = 0u;
        // P row: 256 16-bit values = 128 columns: [local a | local b | features a | features b], mine in my window's blocks, my half
        fused::tmem_st_cols<16>(trow + cP + 32 * w2 + 16 * hf, pl);
        fused::tmem_st_cols<16>(trow + cP + 32 * (1 - w2) + 16 * hf, zero);
        fused::tmem_st_cols<16>(trow + cP + 64 + 32 * w2 + 16 * hf, pr_);
        fused::tmem_st_cols<16>(trow + cP + 64 + 32 * (1 - w2) + 16 * hf, zero);
        ptx::tmem_st_wait();
      }
      SB_MARK(7)
      ptx::tc_fence_before();
      __syncthreads();
      if (warp == 0 && ptx::elect_one()) {
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) ptx::umma_ts(tmem + cO, tmem + cP + 8 * ks, dV + 128 * ks, id_o, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) ptx::umma_ts(tmem + cO, tmem + cP + 64 + 8 * ks, dKVS + 128 * ks, id_o, 1);
        ptx::umma_commit(bar);
      }
      mma_wait();
      SB_MARK(8)
      if (more) issue_rows(nitem, npr, 2, np + 1);
      {   // E4
        float o[32];
        tmem_ld_cols<32>(trow + cO + 32 * hf, reinterpret_cast<uint32_t*>(o));
        ptx::tmem_ld_wait();
        if (have_row) {
          const float inv = 1.0f / (pd[r] + pd[128 + r]);
          uint4* dst = reinterpret_cast<uint4*>(out + ((long long)b * p.N + tok) * ((long long)p.H * 64) + (long long)h * 64) + 4 * hf;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            dst[ch] = make_uint4(Pair16<T>::pk(o[8 * ch] * inv, o[8 * ch + 1] * inv), Pair16<T>::pk(o[8 * ch + 2] * inv, o[8 * ch + 3] * inv),
                                 Pair16<T>::pk(o[8 * ch + 4] * inv, o[8 * ch + 5] * inv), Pair16<T>::pk(o[8 * ch + 6] * inv, o[8 * ch + 7] * inv));
        }
      }
      SB_MARK(9)
      ptx::tc_fence_before();
      __syncthreads();
      if (p.trace && np == 5 && blockIdx.x == 0 && tid == 0)
        printf("sb pair trace (cycles): loads+wait %lld | sync %lld | M1 %lld | E1 %lld | sync+M2 %lld | E2 %lld | sync+M3 %lld | E3 %lld | sync+M4 %lld | E4 %lld\n",
               tk[1] - tk[0], 0LL, tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3], tk[5] - tk[4], tk[6] - tk[5], tk[7] - tk[6], tk[8] - tk[7], tk[9] - tk[8]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

static int g_launches = 0;

}  // namespace sbtc

extern "C" int eva_debug_sb_tc_launches(void) { return sbtc::g_launches; }

bool sb_window_tc_supported(int D, int m, int L, int io_dtype) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("EVA_SM100_DISABLE_FUSED"); off = (e && e[0] == '1') ? 1 : 0; }
  return !off && D == 64 && m == 64 && L <= 64 && (io_dtype == EVA_F16 || io_dtype == EVA_BF16);
}

cudaError_t launch_sb_window_tc(int B, int H, int N, int dims, int gh, int gw, int w, int L, int n_windows, int io_dtype, const View& q,
                                const View& k, const View& v, const uint8_t* mask, const float* proj, const float* bias,
                                const float* stabv, const float* part, void* out, cudaStream_t st) {
  static const int trace = [] { const char* e = getenv("EVA_SM100_TRACE"); return (e && e[0] == '1') ? 1 : 0; }();
  sbtc::Params p{B, H, N, B * H, dims, gh, gw, w, L, n_windows, proj, mask, bias, stabv, part, trace};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int dyn = sbtc::kBias + L * L * 4 + 1024;
  const int grid = p.items < 2 * sms ? p.items : 2 * sms;
  ++sbtc::g_launches;
  auto go = [&](auto kern, auto* o) -> cudaError_t {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return e;
    kern<<<grid, sbtc::kThreads, dyn, st>>>(q, k, v, o, p);
    return cudaGetLastError();
  };
  if (io_dtype == EVA_F16) return go(sbtc::sb_window_tc_kernel<__half>, reinterpret_cast<__half*>(out));
  return go(sbtc::sb_window_tc_kernel<__nv_bfloat16>, reinterpret_cast<__nv_bfloat16*>(out));
}

}  // namespace eva
