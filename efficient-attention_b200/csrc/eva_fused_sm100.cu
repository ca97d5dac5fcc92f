// Fused EVA forward for sm_100a: one persistent kernel does, per (batch, head) work item,
//   phase A  chunk pooling -> adaptive Linear (tcgen05.mma) -> LayerNorm -> k_bar, omega -> beta     (eva.py:155-196)
//   phase B  per pair of windows: S = Q [K_w ; k_bar]^T (tcgen05.mma, operands landed by TMA),
//            joint row softmax in registers (TMEM -> RF), P back to TMEM, O = P [V_w ; beta] (tcgen05.mma,
//            A operand from TMEM), normalise, store                                                   (eva.py:200-227)
// q/k/v are read from HBM once (phase A); the window tiles of phase B are TMA loads that hit L2 because
// the same CTA just streamed that (batch, head).  Geometry: 2-D grid, no halo, head_dim 64, window w with
// L = w*w <= 64 queries, CN <= 64 chunks, fp16 / bf16 I/O, no padding mask.
//
// CTA = 6 warps: warps 0-3 compute (thread t <-> TMEM lane t <-> query row t of the pair tile),
// warp 4 = TMA producer, warp 5 = MMA issuer.  Two CTAs are resident per SM (256 TMEM columns and
// ~103 KB shared memory each) so that one CTA's softmax overlaps the other's loads and MMAs.
//
// Pair tile (M = 128): rows 0..L-1 = window a, rows 64..64+L-1 = window b.  K/V tile (112 rows for L=49):
// rows 0..L-1 = window a, rows LP8..LP8+L-1 = window b (LP8 = L rounded up to 8).  S columns:
// [0,2*LP8) local logits (the off-diagonal blocks are computed and ignored), [2*LP8, 2*LP8+64) chunk logits.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>

#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "launch.h"
#include "sm100_ptx.cuh"

namespace eva {
namespace fused {

constexpr int kD = 64;
constexpr int kThreads = 192;
constexpr int kComputeThreads = 128;
constexpr uint32_t kTmemCols = 256;

enum Bar { kQkFull = 0, kVFull, kQkFree, kVFree, kSFull, kPFull, kOFull, kOFree, kAFull, kLinFull, kStatsFull, kNumBars };

struct Params {
  int B, H, N, gh, gw;
  int nwx, n_windows, n_pairs;   // windows per grid row, total, pairs per item
  int chunk, ncx, Jc;            // chunk edge, chunks per grid row, tokens per chunk
  int items;
  View q, k, v;
  const float *w_q, *b_q, *g_q, *beta_q, *w_k, *b_k, *g_k, *beta_k;
  float mu_coeff, ln_eps;
  const float* noise;
  const float* bias;
  long long bias_sh;
  void* out;
};

template <int L, int CN> struct Layout {
  static constexpr int LP8 = (L + 7) & ~7;
  static constexpr int LS = L | 1;  // bias row stride (odd: conflict-free across rows)
  static constexpr int kQ = 0;                                  // [128][128 B]
  static constexpr int kK = kQ + 128 * 128;                     // [2*LP8][128 B]
  static constexpr int kV = kK + 2 * LP8 * 128;
  static constexpr int kKbar = kV + 2 * LP8 * 128;              // [64][128 B]
  static constexpr int kBeta = kKbar + 64 * 128;                // [64][128 B]
  static constexpr int kA = kBeta + 64 * 128;                   // [128][128 B] chunk means (fp16); later omega fp32 [CN][65]
  static constexpr int kZeroEnd = kA + 128 * 128;
  static constexpr int kW = kZeroEnd;                           // [128][128 B] adaptive weights (fp16)
  static constexpr int kBias = kW + 128 * 128;                  // [L][LS] fp32, pre-multiplied by log2(e)
  static constexpr int kBars = (kBias + L * LS * 4 + 127) & ~127;
  static constexpr int kTmemPtr = kBars + kNumBars * 8;
  static constexpr int kBytes = kTmemPtr + 16;
  static constexpr int kDynamic = kBytes + 1024;                // slack for 1024-B alignment of the base
  static_assert(CN * 65 * 4 <= 128 * 128, "omega staging must fit in the means tile");
  static_assert(kK % 1024 == 0 && kV % 1024 == 0 && kKbar % 1024 == 0 && kBeta % 1024 == 0 && kA % 1024 == 0 && kW % 1024 == 0,
                "UMMA tiles must be 1024-byte aligned");
  static_assert(2 * LP8 + 64 + 64 <= (int)kTmemCols, "TMEM column budget");
  static_assert((2 * LP8) % 16 == 0, "local S width must be a multiple of 16");
  // TMEM columns
  static constexpr uint32_t cSloc = 0, cSrfa = 2 * LP8, cO = 2 * LP8 + 64, cPloc = 0, cPrfa = LP8;
};

template <typename T> struct IoFmt;
template <> struct IoFmt<__half> {
  static constexpr uint32_t kUmma = ptx::kFmtF16;
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
};
template <> struct IoFmt<__nv_bfloat16> {
  static constexpr uint32_t kUmma = ptx::kFmtBF16;
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
};
__device__ __forceinline__ uint32_t pack2_f16(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// N consecutive TMEM columns -> registers, N decomposed into x16 / x1 loads (compile time)
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* r) {
#pragma unroll
  for (int g = 0; g < N / 16; ++g) ptx::tmem_ld16(taddr + 16 * g, r + 16 * g);
#pragma unroll
  for (int j = (N / 16) * 16; j < N; ++j) ptx::tmem_ld1(taddr + j, r[j]);
}
// registers -> N consecutive TMEM columns, N a multiple of 4
template <int N>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t* r) {
  static_assert(N % 4 == 0, "");
  int done = 0;
#pragma unroll
  for (int g = 0; g < N / 16; ++g) { ptx::tmem_st16(taddr + done, r + done); done += 16; }
  if constexpr ((N % 16) >= 8) { ptx::tmem_st8(taddr + done, r + done); done += 8; }
  if constexpr ((N % 8) >= 4) { ptx::tmem_st4(taddr + done, r + done); done += 4; }
}

// 16-byte store of 8 packed 16-bit values into row `row`, 16-byte chunk `chunk` of a 128-byte-swizzled tile
__device__ __forceinline__ void st_tile_chunk(uint8_t* tile, int row, int chunk, uint4 v) {
  *reinterpret_cast<uint4*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
}

template <typename T, int W, int CN>
__global__ void __launch_bounds__(kThreads, 2)
eva_fused_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const Params p) {
  constexpr int L = W * W;
  using Lay = Layout<L, CN>;
  constexpr int LP8 = Lay::LP8, LS = Lay::LS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* Qt = sm + Lay::kQ;
  uint8_t* Kt = sm + Lay::kK;
  uint8_t* Vt = sm + Lay::kV;
  uint8_t* KBt = sm + Lay::kKbar;
  uint8_t* BTt = sm + Lay::kBeta;
  uint8_t* At = sm + Lay::kA;
  uint8_t* Wt = sm + Lay::kW;
  float* omega = reinterpret_cast<float*>(At);           // [CN][65], valid after the Linear MMA has consumed At
  float* bias2 = reinterpret_cast<float*>(sm + Lay::kBias);
  const uint32_t bars = ptx::smem_u32(sm + Lay::kBars);
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + Lay::kTmemPtr);
  auto bar = [&](int i) { return bars + 8u * i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup --------------------------------------------------------------------------
  for (int i = tid; i < Lay::kZeroEnd / 16; i += kThreads) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  for (int idx = tid; idx < 128 * 8; idx += kThreads) {   // W tile: rows 0-63 = W_q, 64-127 = W_k, fp16
    const int row = idx >> 3, ch = idx & 7;
    const float* src = row < 64 ? p.w_q : p.w_k;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (src) {
      float f[8];
      load8<float>(src + (row & 63) * 64 + ch * 8, f);
      v = make_uint4(pack2_f16(f[0], f[1]), pack2_f16(f[2], f[3]), pack2_f16(f[4], f[5]), pack2_f16(f[6], f[7]));
    }
    st_tile_chunk(Wt, row, ch, v);
  }
  if (warp == 4 && lane == 0) {
    ptx::mbar_init(bar(kQkFull), 1);
    ptx::mbar_init(bar(kVFull), 1);
    ptx::mbar_init(bar(kQkFree), 1);
    ptx::mbar_init(bar(kVFree), 1);
    ptx::mbar_init(bar(kSFull), 1);
    ptx::mbar_init(bar(kPFull), kComputeThreads);
    ptx::mbar_init(bar(kOFull), 1);
    ptx::mbar_init(bar(kOFree), kComputeThreads);
    ptx::mbar_init(bar(kAFull), kComputeThreads);
    ptx::mbar_init(bar(kLinFull), 1);
    ptx::mbar_init(bar(kStatsFull), kComputeThreads);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tm_q);
    ptx::prefetch_tmap(&tm_k);
    ptx::prefetch_tmap(&tm_v);
  }
  if (warp == 5) ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr)), kTmemCols);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == 4) {
    // =================================== TMA producer ==========================================
    if (lane == 0) {
      uint32_t np = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int b = item / p.H, h = item % p.H;
        for (int pr = 0; pr < p.n_pairs; ++pr, ++np) {
          const int w0 = 2 * pr, w1 = w0 + 1;
          const bool two = w1 < p.n_windows;
          const int x0 = (w0 % p.nwx) * W, y0 = (w0 / p.nwx) * W;
          const int x1 = (w1 % p.nwx) * W, y1 = (w1 / p.nwx) * W;
          ptx::mbar_wait(bar(kQkFree), (np & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(bar(kQkFull), (two ? 4u : 2u) * L * 128u);
          ptx::tma_load_5d(ptx::smem_u32(Qt), &tm_q, bar(kQkFull), 0, h, x0, y0, b);
          ptx::tma_load_5d(ptx::smem_u32(Kt), &tm_k, bar(kQkFull), 0, h, x0, y0, b);
          if (two) {
            ptx::tma_load_5d(ptx::smem_u32(Qt + 64 * 128), &tm_q, bar(kQkFull), 0, h, x1, y1, b);
            ptx::tma_load_5d(ptx::smem_u32(Kt + LP8 * 128), &tm_k, bar(kQkFull), 0, h, x1, y1, b);
          }
          ptx::mbar_wait(bar(kVFree), (np & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(bar(kVFull), (two ? 2u : 1u) * L * 128u);
          ptx::tma_load_5d(ptx::smem_u32(Vt), &tm_v, bar(kVFull), 0, h, x0, y0, b);
          if (two) ptx::tma_load_5d(ptx::smem_u32(Vt + LP8 * 128), &tm_v, bar(kVFull), 0, h, x1, y1, b);
        }
      }
    }
  } else if (warp == 5) {
    // =================================== MMA issuer ============================================
    if (lane == 0) {
      constexpr uint32_t fmt = IoFmt<T>::kUmma;
      constexpr uint32_t id_lin = ptx::umma_idesc(ptx::kFmtF16, ptx::kFmtF16, 0, 0, 128, 128);
      constexpr uint32_t id_sl = ptx::umma_idesc(fmt, fmt, 0, 0, 128, 2 * LP8);
      constexpr uint32_t id_sr = ptx::umma_idesc(fmt, fmt, 0, 0, 128, 64);
      constexpr uint32_t id_pv = ptx::umma_idesc(fmt, fmt, 0, 1, 128, 64);
      const uint64_t dQ = ptx::umma_desc_sw128(ptx::smem_u32(Qt)), dK = ptx::umma_desc_sw128(ptx::smem_u32(Kt));
      const uint64_t dV = ptx::umma_desc_sw128(ptx::smem_u32(Vt)), dKB = ptx::umma_desc_sw128(ptx::smem_u32(KBt));
      const uint64_t dBT = ptx::umma_desc_sw128(ptx::smem_u32(BTt)), dA = ptx::umma_desc_sw128(ptx::smem_u32(At));
      const uint64_t dW = ptx::umma_desc_sw128(ptx::smem_u32(Wt));
      uint32_t ni = 0, np = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ni) {
        // adaptive Linear for all chunks at once: [q means ; k means] x [W_q ; W_k]^T (diagonal blocks used)
        ptx::mbar_wait(bar(kAFull), ni & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem, dA + 2 * ks, dW + 2 * ks, id_lin, ks > 0);
        ptx::umma_commit(bar(kLinFull));
        ptx::mbar_wait(bar(kStatsFull), ni & 1);
        ptx::tc_fence_after();
        for (int pr = 0; pr < p.n_pairs; ++pr, ++np) {
          ptx::mbar_wait(bar(kQkFull), np & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + Lay::cSloc, dQ + 2 * ks, dK + 2 * ks, id_sl, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + Lay::cSrfa, dQ + 2 * ks, dKB + 2 * ks, id_sr, ks > 0);
          ptx::umma_commit(bar(kSFull));
          ptx::umma_commit(bar(kQkFree));
          ptx::mbar_wait(bar(kPFull), np & 1);
          ptx::mbar_wait(bar(kVFull), np & 1);
          ptx::mbar_wait(bar(kOFree), (np & 1) ^ 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 2 * LP8 / 16; ++ks)
            ptx::umma_ts(tmem + Lay::cO, tmem + Lay::cPloc + 8 * ks, dV + 128 * ks, id_pv, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_ts(tmem + Lay::cO, tmem + Lay::cPrfa + 8 * ks, dBT + 128 * ks, id_pv, 1);
          ptx::umma_commit(bar(kOFull));
          ptx::umma_commit(bar(kVFree));
        }
      }
    }
  } else {
    // =================================== compute warps ==========================================
    const int ws = tid >> 6;           // which window of the pair this row belongs to (warp-uniform)
    const int i = tid & 63;            // query slot inside the window (valid if < L)
    const int ic = i < L ? i : L - 1;  // clamped, for shared-memory reads of idle rows
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const int oct = lane & 7, slot = lane >> 3;
    const float scale = 0.125f;                      // head_dim 64
    const float scale_log2 = scale * kLog2e;
    const float inv_cnt = 1.0f / (float)p.Jc;
    T* const out = reinterpret_cast<T*>(p.out);
    uint32_t ni = 0, np = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++ni) {
      const int b = item / p.H, h = item % p.H;
      // ---- per-head bias table (x log2 e) ------------------------------------------------------
      for (int idx = tid; idx < L * L; idx += kComputeThreads) {
        const int r = idx / L, c = idx % L;
        bias2[r * LS + c] = p.bias ? __ldg(p.bias + (long long)h * p.bias_sh + idx) * kLog2e : 0.f;
      }
      // ---- A1: chunk means of q and k -> fp16 tile rows c (q) and 64+c (k) ----------------------
      for (int c = warp; c < CN; c += 4) {
        const int ty0 = (c / p.ncx) * p.chunk, tx0 = (c % p.ncx) * p.chunk;
        float aq[8], ak[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) aq[e] = ak[e] = 0.f;
        for (int t0 = 0; t0 < p.Jc; t0 += 4) {
          const int t = t0 + slot;
          if (t < p.Jc) {
            const int tok = (ty0 + t / p.chunk) * p.gw + tx0 + t % p.chunk;
            float fq[8], fk[8];
            load8<T>(p.q.row<T>(b, tok, h) + oct * 8, fq);
            load8<T>(p.k.row<T>(b, tok, h) + oct * 8, fk);
#pragma unroll
            for (int e = 0; e < 8; ++e) { aq[e] += fq[e]; ak[e] += fk[e]; }
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          aq[e] += __shfl_xor_sync(0xffffffffu, aq[e], 8);
          aq[e] += __shfl_xor_sync(0xffffffffu, aq[e], 16);
          ak[e] += __shfl_xor_sync(0xffffffffu, ak[e], 8);
          ak[e] += __shfl_xor_sync(0xffffffffu, ak[e], 16);
          aq[e] *= inv_cnt;
          ak[e] *= inv_cnt;
        }
        if (slot == 0) {
          st_tile_chunk(At, c, oct, make_uint4(pack2_f16(aq[0], aq[1]), pack2_f16(aq[2], aq[3]), pack2_f16(aq[4], aq[5]), pack2_f16(aq[6], aq[7])));
          st_tile_chunk(At, 64 + c, oct, make_uint4(pack2_f16(ak[0], ak[1]), pack2_f16(ak[2], ak[3]), pack2_f16(ak[4], ak[5]), pack2_f16(ak[6], ak[7])));
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(bar(kAFull));
      // ---- Linear result -> bias, LayerNorm; rows 0-63 = q side, rows 64-127 = k side ------------
      ptx::mbar_wait(bar(kLinFull), ni & 1);
      ptx::tc_fence_after();
      {
        float y[64];
        tmem_ld_cols<64>(trow + (ws ? 64u : 0u), reinterpret_cast<uint32_t*>(y));
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        const float* lb = ws ? p.b_k : p.b_q;
        const float* gain = ws ? p.g_k : p.g_q;
        const float* lnb = ws ? p.beta_k : p.beta_q;
        if (lb) {
#pragma unroll
          for (int e = 0; e < 64; ++e) y[e] += __ldg(lb + e);
        }
        if (gain) {
          float s = 0.f;
#pragma unroll
          for (int e = 0; e < 64; ++e) s += y[e];
          const float mean = s * (1.0f / 64);
          float var = 0.f;
#pragma unroll
          for (int e = 0; e < 64; ++e) { const float d_ = y[e] - mean; var = fmaf(d_, d_, var); }
          const float inv = 1.0f / sqrtf(var * (1.0f / 64) + p.ln_eps);
#pragma unroll
          for (int e = 0; e < 64; ++e) y[e] = (y[e] - mean) * inv * __ldg(gain + e) + __ldg(lnb + e);
        }
        const bool valid = i < CN;
        if (ws == 1 && valid) {   // k side: k_bar tile (B operand of the chunk logits) + fp32 copy for mu
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            st_tile_chunk(KBt, i, ch, make_uint4(IoFmt<T>::pack2(y[8 * ch], y[8 * ch + 1]), IoFmt<T>::pack2(y[8 * ch + 2], y[8 * ch + 3]),
                                                  IoFmt<T>::pack2(y[8 * ch + 4], y[8 * ch + 5]), IoFmt<T>::pack2(y[8 * ch + 6], y[8 * ch + 7])));
#pragma unroll
          for (int e = 0; e < 64; ++e) omega[i * 65 + e] = y[e];
        }
        ptx::named_bar_sync(1, kComputeThreads);
        if (ws == 0 && valid) {   // q side: omega = mu_coeff (q_bar + k_bar) [+ noise]   (eva.py:182-190)
          const float* nz = p.noise ? p.noise + (((long long)b * p.H + h) * CN + i) * 64 : nullptr;
#pragma unroll
          for (int e = 0; e < 64; ++e) {
            float o = p.w_q ? p.mu_coeff * (y[e] + omega[i * 65 + e]) : 0.f;
            if (nz) o += __ldg(nz + e);
            omega[i * 65 + e] = o;
          }
        }
        ptx::named_bar_sync(1, kComputeThreads);
      }
      // ---- A2: beta_c = softmax_j(prm(k_j, omega_c)) . v_j  -> beta tile --------------------------
      for (int c = warp; c < CN; c += 4) {
        const int ty0 = (c / p.ncx) * p.chunk, tx0 = (c % p.ncx) * p.chunk;
        float om[8], acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { om[e] = omega[c * 65 + oct * 8 + e]; acc[e] = 0.f; }
        float m = kNegInf, l = 0.f;
        for (int t0 = 0; t0 < p.Jc; t0 += 4) {
          const int t = t0 + slot;
          float lg = kNegInf, fv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) fv[e] = 0.f;
          if (t < p.Jc) {
            const int tok = (ty0 + t / p.chunk) * p.gw + tx0 + t % p.chunk;
            float fk[8];
            load8<T>(p.k.row<T>(b, tok, h) + oct * 8, fk);
            load8<T>(p.v.row<T>(b, tok, h) + oct * 8, fv);
            float part = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) part = fmaf(fk[e], om[e] - 0.5f * fk[e], part);
            lg = part;
          }
          // the 8 lanes of a token slot hold partial dot products (invalid slots carry -inf in all 8)
          float tot = (t < p.Jc) ? lg : 0.f;
          tot += __shfl_xor_sync(0xffffffffu, tot, 1);
          tot += __shfl_xor_sync(0xffffffffu, tot, 2);
          tot += __shfl_xor_sync(0xffffffffu, tot, 4);
          lg = (t < p.Jc) ? tot * scale : kNegInf;
          const float mn = fmaxf(m, lg);
          if (mn != kNegInf) {
            const float corr = ex2((m - mn) * kLog2e), pj = ex2((lg - mn) * kLog2e);
            l = fmaf(l, corr, pj);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(acc[e], corr, pj * fv[e]);
            m = mn;
          }
        }
        // merge the 4 token slots
        float mg = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
        mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, 16));
        const float f = (m == kNegInf) ? 0.f : ex2((m - mg) * kLog2e);
        l *= f;
        l += __shfl_xor_sync(0xffffffffu, l, 8);
        l += __shfl_xor_sync(0xffffffffu, l, 16);
        const float inv_l = 1.0f / l;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          acc[e] *= f;
          acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
          acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
          acc[e] *= inv_l;
        }
        if (slot == 0)
          st_tile_chunk(BTt, c, oct, make_uint4(IoFmt<T>::pack2(acc[0], acc[1]), IoFmt<T>::pack2(acc[2], acc[3]),
                                                 IoFmt<T>::pack2(acc[4], acc[5]), IoFmt<T>::pack2(acc[6], acc[7])));
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(kStatsFull));

      // ---- phase B: pairs of windows ----------------------------------------------------------------
      for (int pr = 0; pr < p.n_pairs; ++pr, ++np) {
        const int wi = 2 * pr + ws;
        const bool valid = i < L && wi < p.n_windows;
        ptx::mbar_wait(bar(kSFull), np & 1);
        ptx::tc_fence_after();
        float sl[L], sr[CN];
        tmem_ld_cols<L>(trow + Lay::cSloc + (uint32_t)(ws * LP8), reinterpret_cast<uint32_t*>(sl));
        tmem_ld_cols<CN>(trow + Lay::cSrfa, reinterpret_cast<uint32_t*>(sr));
        ptx::tmem_ld_wait();
        const float* brow = bias2 + ic * LS;
        float mx = kNegInf;
#pragma unroll
        for (int j = 0; j < L; ++j) { sl[j] = fmaf(sl[j], scale_log2, brow[j]); mx = fmaxf(mx, sl[j]); }
#pragma unroll
        for (int c = 0; c < CN; ++c) { sr[c] *= scale_log2; mx = fmaxf(mx, sr[c]); }
        float sum = 0.f;
        uint32_t pl[LP8 / 2], prf[32], zeros[LP8 / 2];
#pragma unroll
        for (int j = 0; j < LP8 / 2; ++j) {
          const float a = (2 * j < L) ? ex2(sl[(2 * j < L) ? 2 * j : 0] - mx) : 0.f;
          const float c2 = (2 * j + 1 < L) ? ex2(sl[(2 * j + 1 < L) ? 2 * j + 1 : 0] - mx) : 0.f;
          sum += a + c2;
          pl[j] = IoFmt<T>::pack2(a, c2);
          zeros[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float a = (2 * j < CN) ? ex2(sr[(2 * j < CN) ? 2 * j : 0] - mx) : 0.f;
          const float c2 = (2 * j + 1 < CN) ? ex2(sr[(2 * j + 1 < CN) ? 2 * j + 1 : 0] - mx) : 0.f;
          sum += a + c2;
          prf[j] = IoFmt<T>::pack2(a, c2);
        }
        // P (16-bit, two per column) overwrites the S columns this thread has finished reading
        tmem_st_cols<LP8 / 2>(trow + Lay::cPloc + (uint32_t)(ws * (LP8 / 2)), pl);
        tmem_st_cols<LP8 / 2>(trow + Lay::cPloc + (uint32_t)((1 - ws) * (LP8 / 2)), zeros);
        tmem_st_cols<32>(trow + Lay::cPrfa, prf);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar(kPFull));
        // ---- epilogue: O / rowsum -> out[b, token, h, :] ---------------------------------------------
        ptx::mbar_wait(bar(kOFull), np & 1);
        ptx::tc_fence_after();
        float o[64];
        tmem_ld_cols<64>(trow + Lay::cO, reinterpret_cast<uint32_t*>(o));
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar(kOFree));
        if (valid) {
          const float inv = 1.0f / sum;
          const int tok = ((wi / p.nwx) * W + i / W) * p.gw + (wi % p.nwx) * W + i % W;
          uint4* dst = reinterpret_cast<uint4*>(out + (((long long)b * p.N + tok) * p.H + h) * kD);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            dst[ch] = make_uint4(IoFmt<T>::pack2(o[8 * ch] * inv, o[8 * ch + 1] * inv), IoFmt<T>::pack2(o[8 * ch + 2] * inv, o[8 * ch + 3] * inv),
                                 IoFmt<T>::pack2(o[8 * ch + 4] * inv, o[8 * ch + 5] * inv), IoFmt<T>::pack2(o[8 * ch + 6] * inv, o[8 * ch + 7] * inv));
        }
      }
    }
  }
  // ---- teardown ------------------------------------------------------------------------------------
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 5) ptx::tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

// [B, gh, gw, H, 64] view (strides from the caller's q/k/v view) with a (w x w x 64) box, 128-B swizzle
static bool make_window_map(CUtensorMap* tm, const View& v, const Geo& g, int io_dtype) {
  auto enc = get_encode();
  if (!enc) return false;
  const cuuint64_t dims[5] = {64, (cuuint64_t)g.H, (cuuint64_t)g.gw, (cuuint64_t)g.gh, (cuuint64_t)g.B};
  const cuuint64_t strides[4] = {(cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2, (cuuint64_t)v.sn * g.gw * 2, (cuuint64_t)v.sb * 2};
  const cuuint32_t box[5] = {64, 1, (cuuint32_t)g.window, (cuuint32_t)g.window, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(tm, io_dtype == EVA_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                         const_cast<void*>(v.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n > 0 ? n : 148;
}

template <typename T, int W, int CN>
static cudaError_t launch_t(const Geo& g, const View& q, const View& k, const View& v, const EvaAdaptive& ada,
                            const float* noise, const float* bias, long long bias_sh, void* out, cudaStream_t st,
                            const char** msg) {
  using Lay = Layout<W * W, CN>;
  CUtensorMap tq, tk, tv;
  const int io = sizeof(T) == 2 && std::is_same<T, __half>::value ? EVA_F16 : EVA_BF16;
  if (!make_window_map(&tq, q, g, io) || !make_window_map(&tk, k, g, io) || !make_window_map(&tv, v, g, io)) {
    *msg = "cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
  Params p{};
  p.B = g.B; p.H = g.H; p.N = g.N; p.gh = g.gh; p.gw = g.gw;
  p.nwx = g.gw / g.window; p.n_windows = g.n_windows; p.n_pairs = (g.n_windows + 1) / 2;
  p.chunk = g.chunk; p.ncx = g.gw / g.chunk; p.Jc = g.Jc;
  p.items = g.B * g.H;
  p.q = q; p.k = k; p.v = v;
  p.w_q = ada.w_q; p.b_q = ada.b_q; p.g_q = ada.ln_gain_q; p.beta_q = ada.ln_bias_q;
  p.w_k = ada.w_k; p.b_k = ada.b_k; p.g_k = ada.ln_gain_k; p.beta_k = ada.ln_bias_k;
  p.mu_coeff = ada.mu_coeff; p.ln_eps = ada.ln_eps;
  p.noise = noise; p.bias = bias; p.bias_sh = bias_sh; p.out = out;
  auto kern = eva_fused_kernel<T, W, CN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay::kDynamic);
  if (e != cudaSuccess) { *msg = "cudaFuncSetAttribute"; return e; }
  const int grid = p.items < 2 * sm_count() ? p.items : 2 * sm_count();
  kern<<<grid, kThreads, Lay::kDynamic, st>>>(tq, tk, tv, p);
  *msg = "kernel launch";
  return cudaGetLastError();
}

}  // namespace fused

static bool fused_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EVA_SM100_DISABLE_FUSED");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

bool fused_supported(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask,
                     const EvaAdaptive& ada, const float* bias, long long bias_sh) {
  (void)ada; (void)bias; (void)bias_sh;
  if (fused_disabled()) return false;
  if (g.dims != 2 || g.ext != 0 || g.chunk_ext != 0 || g.causal || g.D != 64 || mask != nullptr) return false;
  if (io_dtype != EVA_F16 && io_dtype != EVA_BF16) return false;
  if (g.window != 7 || g.n_chunks != 49 || g.chunk <= 0) return false;
  if (g.gw % g.chunk || g.gh % g.chunk) return false;
  for (const View* x : {&q, &k, &v}) {
    if (x->sh * 2 % 16 || x->sn * 2 % 16 || x->sb * 2 % 16) return false;
    if (x->sh <= 0 || x->sn <= 0 || x->sb <= 0) return false;
  }
  return fused::get_encode() != nullptr;
}

size_t fused_workspace_bytes(const Geo&) { return 0; }

cudaError_t launch_fused(const Geo& g, int io_dtype, const View& q, const View& k, const View& v,
                         const EvaAdaptive& ada, const float* noise, const float* bias, long long bias_sh,
                         void* out, void* workspace, cudaStream_t st, const char** msg) {
  (void)workspace;
  if (io_dtype == EVA_F16) return fused::launch_t<__half, 7, 49>(g, q, k, v, ada, noise, bias, bias_sh, out, st, msg);
  return fused::launch_t<__nv_bfloat16, 7, 49>(g, q, k, v, ada, noise, bias, bias_sh, out, st, msg);
}

}  // namespace eva
