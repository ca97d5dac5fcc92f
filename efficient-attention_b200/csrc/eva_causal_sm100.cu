// Causal EVA window attention for sm_100a (causal_eva.py:722-783) on tcgen05 / TMEM / TMA: stage B of the causal layer
// for window = 256, no halo, head_dim 64, 16-bit I/O, no padding mask, optional Toeplitz (T5) position bias (BASELINE c5).
// ONE PASS (round 2): when a window holds whole chunks (chunk in {64, 128, 256}) and there are <= 32 of them, the chunk statistics
// (causal_eva.py:676-719: means, adaptive Linear + LayerNorm, phi-logits, softmax, beta) of the window's own chunks are computed by
// the compute warps from the q / k / v tiles the window already has in shared memory and published to global memory with a
// per-window flag; windows are handed out in window-major order, so the chunks a window needs (c < its queries' chunk) come from
// CTAs that started earlier, and the producer warp waits for their flags before it converts the rows.  q, k, v are read from HBM
// once instead of twice (chunk_stats kernel + this one).  Other geometries keep the separate statistics kernel.
//
// Work item = one window of one (batch, head): 256 queries x (256 causal local keys + up to 64 chunk keys).
//   S   = Q [K_w ; k_bar]^T     M = 128 per row-block; row-block 0 only needs keys 0-127 (causality), row-block 1 all 256
//   P   = softmax(S) with the causal mask j <= i inside the diagonal block and the chunk mask c < (query chunk)
//         (masked logits are -5e4 in the reference, i.e. exactly 0 after the softmax: every row sees its own key)
//   O   = P [V_w ; beta]        A operand from TMEM
// One persistent CTA per SM, 10 warps: warps 0-3 own row-block 0 (TMEM lane = query row), warps 4-7 row-block 1,
// warp 8 = TMA producer (+ k_bar / beta fp32 -> 16-bit tiles), warp 9 = MMA issuer.  Two stages of {Q, K, V} (3 x 32 KB) so
// the loads of the next window run under the current one.  The logits stay in TMEM and are read twice (max, then exp) in
// 32-column pieces; P overwrites the first half of the columns it came from; O lands in columns of the same region that
// are dead by then, so each row-block needs exactly n_keys TMEM columns (192 + 320 = 512).  The output rows are staged in
// the (dead) Q tile of the stage and leave through one TMA store per row-block.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>

#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "launch.h"
#include "sm100_ptx.cuh"

namespace eva {
namespace causal {

constexpr int kWin = 256, kThreads = 448, kStatsThreads = 128, kStageBytes = 3 * 32768, kKbBytes = 16384;
constexpr int kSmemBars = 2 * kStageBytes + 2 * kKbBytes;
constexpr int kSmemBias = kSmemBars + 256;          // [256] fp32: Toeplitz position bias by distance i - j, x log2(e)
constexpr int kDynamic = kSmemBias + 1024 + 1024;

enum Bar { kFullQK0, kFullQK1, kFullV0, kFullV1, kFullKB0, kFullKB1, kFree0, kFree1,
           kSFull0, kSFull1, kPFull0, kPFull1, kOFull0, kOFull1, kOFree0, kOFree1, kStatsDone0, kStatsDone1, kNumBars };

struct Params {
  int B, H, N, n_win, items, n_chunks, cnp, chunk;
  int swap;                      // bit i: tensor map i (q, k, v) has its batch and token dimensions exchanged (time-major activations)
  const float *kbar, *beta;      // [B, H, n_chunks, 64] fp32 from chunk_stats_kernel (fuse == 0) or written by this kernel (fuse == 1)
  const float* bias;             // [256, 256] fp32 with bias[i][j] = f(i - j) (T5 bias shared by the heads), or NULL
  // one-pass mode
  int fuse, cpw;                 // chunks per window
  float *kbar_w, *beta_w;        // same arrays, writable
  unsigned int* flags;           // [B * H * n_win], zeroed by the launcher: 1 = the window's chunk statistics are in global memory
  const float *w_q, *b_q, *g_q, *beta_q, *w_k, *b_k, *g_k, *beta_k;
  float mu_coeff, ln_eps;
  const float* noise;            // [B, H, n_chunks, 64] or NULL
  float* lse_out;                // training: log2-domain log-sum-exp of every query row -> float32 [B, H, N] (the backward reads it), or NULL
};

template <typename T> struct Fmt;
template <> struct Fmt<__half> {
  static constexpr uint32_t kUmma = ptx::kFmtF16;
  static __device__ __forceinline__ uint32_t pack2(float a, float b) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
};
template <> struct Fmt<__nv_bfloat16> {
  static constexpr uint32_t kUmma = ptx::kFmtBF16;
  static __device__ __forceinline__ uint32_t pack2(float a, float b) { const __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
};
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
eva_causal_window_kernel(const __grid_constant__ CUtensorMap t_q, const __grid_constant__ CUtensorMap t_k,
                         const __grid_constant__ CUtensorMap t_v, const __grid_constant__ CUtensorMap t_o, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* const sm = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t bars = ptx::smem_u32(sm + kSmemBars);
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + kSmemBars + kNumBars * 8);
  auto bar = [&](int i) { return bars + 8u * i; };
  auto stage_ptr = [&](int s) { return sm + s * kStageBytes; };                       // Q | K | V, 32 KB each
  auto kb_ptr = [&](int s) { return sm + 2 * kStageBytes + s * kKbBytes; };           // k_bar | beta, 8 KB each
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < (2 * kKbBytes) / 16; i += kThreads) reinterpret_cast<uint4*>(sm + 2 * kStageBytes)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 8 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(bar(kFullQK0 + s), 1);
      ptx::mbar_init(bar(kFullV0 + s), 1);
      ptx::mbar_init(bar(kFullKB0 + s), 1);
      ptx::mbar_init(bar(kFree0 + s), 3);          // MMA commit (K, V, k_bar, beta read) + one output-store drain per row-block (Q tile)
      ptx::mbar_init(bar(kSFull0 + s), 1);
      ptx::mbar_init(bar(kPFull0 + s), 128);
      ptx::mbar_init(bar(kOFull0 + s), 1);
      ptx::mbar_init(bar(kOFree0 + s), 128);
      ptx::mbar_init(bar(kStatsDone0 + s), kStatsThreads);
    }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&t_q); ptx::prefetch_tmap(&t_k); ptx::prefetch_tmap(&t_v); ptx::prefetch_tmap(&t_o);
  }
  float* const dbias = reinterpret_cast<float*>(sm + kSmemBias);
  for (int n = tid; n < kWin; n += kThreads) dbias[n] = p.bias ? __ldg(p.bias + (long long)n * kWin) * kLog2e : 0.f;   // column 0: distance n
  if (warp == 9) ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr)), 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  // TMEM columns of row-block rb: logits [base, base + n_loc + cnp), P (16-bit pairs) from base, O in the last 64 columns
  // of [base, base + n_loc + 64): rb 0 -> [0, 192), rb 1 -> [192, 512)
  // work items: (batch * head)-major for the two-kernel path; WINDOW-major for the one-pass path (all windows 0 first), so that
  // the windows a window depends on were handed out earlier
  const int n_bh = p.B * p.H;
  auto decode = [&](int item, int& wi, int& bh) {
    if (p.fuse) { wi = item / n_bh; bh = item % n_bh; } else { wi = item % p.n_win; bh = item / p.n_win; }
  };
  auto base_col = [](int rb) -> uint32_t { return rb ? 192u : 0u; };
  auto o_col = [](int rb) -> uint32_t { return rb ? 448u : 128u; };

  if (warp == 8) {
    // =================================== TMA producer ==========================================
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int s = it & 1;
      int wi, bh;
      decode(item, wi, bh);
      const int h = bh % p.H, b = bh / p.H;
      if (it >= 2) ptx::mbar_wait(bar(kFree0 + s), ((it >> 1) - 1) & 1);
      const uint32_t st = ptx::smem_u32(stage_ptr(s));
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(bar(kFullQK0 + s), 65536);
        const int n0 = wi * kWin;
        ptx::tma_load_4d(st, &t_q, bar(kFullQK0 + s), 0, h, (p.swap & 1) ? b : n0, (p.swap & 1) ? n0 : b);
        ptx::tma_load_4d(st + 32768, &t_k, bar(kFullQK0 + s), 0, h, (p.swap & 2) ? b : n0, (p.swap & 2) ? n0 : b);
        ptx::mbar_arrive_expect_tx(bar(kFullV0 + s), 32768);
        ptx::tma_load_4d(st + 65536, &t_v, bar(kFullV0 + s), 0, h, (p.swap & 4) ? b : n0, (p.swap & 4) ? n0 : b);
      }
      // chunk keys / values of this (batch, head): fp32 rows -> 16-bit swizzled tiles (rows >= n_chunks stay zero)
      uint8_t* kb = kb_ptr(s);
      const float* src_k = p.kbar + (long long)bh * p.n_chunks * 64;
      const float* src_b = p.beta + (long long)bh * p.n_chunks * 64;
      int n_rows = p.n_chunks;
      if (p.fuse) {
        // rows of the chunks this window can see: those of earlier windows, and (several chunks per window) its own; wait until the
        // CTAs that own them have published them
        const int n_need = p.cpw > 1 ? wi + 1 : wi;
        n_rows = n_need * p.cpw;
        if (lane < n_need) {
          const unsigned int* f = p.flags + (long long)bh * p.n_win + lane;
          unsigned int v_ = 0, spins = 0;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v_) : "l"(f) : "memory");
            if (!v_ && ++spins > 20000000u) __trap();          // bounded: a scheduling bug must not hang the GPU
          } while (!v_);
        }
        __syncwarp();
      }
      for (int idx = lane; idx < p.cnp * 8; idx += 32) {
        const int row = idx >> 3, ch = idx & 7;
        const int off = row * 128 + ((ch ^ (row & 7)) << 4);
        if (row >= n_rows) {                       // not visible to this window (or not computed yet): exact zeros, never stale data
          if (p.fuse) {
            *reinterpret_cast<uint4*>(kb + off) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(kb + 8192 + off) = make_uint4(0, 0, 0, 0);
          }
          continue;
        }
        // plain (coherent) loads: in the one-pass mode these rows were written by other CTAs of this launch
        const float4 a0 = *reinterpret_cast<const float4*>(src_k + row * 64 + ch * 8);
        const float4 a1 = *(reinterpret_cast<const float4*>(src_k + row * 64 + ch * 8) + 1);
        const float4 b0 = *reinterpret_cast<const float4*>(src_b + row * 64 + ch * 8);
        const float4 b1 = *(reinterpret_cast<const float4*>(src_b + row * 64 + ch * 8) + 1);
        *reinterpret_cast<uint4*>(kb + off) = make_uint4(Fmt<T>::pack2(a0.x, a0.y), Fmt<T>::pack2(a0.z, a0.w), Fmt<T>::pack2(a1.x, a1.y), Fmt<T>::pack2(a1.z, a1.w));
        *reinterpret_cast<uint4*>(kb + 8192 + off) = make_uint4(Fmt<T>::pack2(b0.x, b0.y), Fmt<T>::pack2(b0.z, b0.w), Fmt<T>::pack2(b1.x, b1.y), Fmt<T>::pack2(b1.z, b1.w));
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (ptx::elect_one()) ptx::mbar_arrive(bar(kFullKB0 + s));
    }
  } else if (warp == 9) {
    // =================================== MMA issuer ============================================
    constexpr uint32_t fmt = Fmt<T>::kUmma;
    constexpr uint32_t id_s128 = ptx::umma_idesc(fmt, fmt, 0, 0, 128, 128), id_s256 = ptx::umma_idesc(fmt, fmt, 0, 0, 128, 256);
    constexpr uint32_t id_pv = ptx::umma_idesc(fmt, fmt, 0, 1, 128, 64);
    const uint32_t id_rfa = ptx::umma_idesc(fmt, fmt, 0, 0, 128, (uint32_t)p.cnp);
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const uint64_t dQ = ptx::umma_desc_sw128(ptx::smem_u32(stage_ptr(s))), dK = dQ + (32768 >> 4), dV = dQ + (65536 >> 4);
      const uint64_t dKB = ptx::umma_desc_sw128(ptx::smem_u32(kb_ptr(s))), dBT = dKB + (8192 >> 4);
      ptx::mbar_wait(bar(kFullQK0 + s), ph);
      ptx::mbar_wait(bar(kFullKB0 + s), ph);
#pragma unroll 1
      for (int rb = 0; rb < 2; ++rb) {
        if (it > 0) ptx::mbar_wait(bar(kOFree0 + rb), (it - 1) & 1);     // the epilogue has read O of the previous window
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint64_t dQr = dQ + (uint64_t)(rb * (16384 >> 4));
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + base_col(rb), dQr + 2 * ks, dK + 2 * ks, rb ? id_s256 : id_s128, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + base_col(rb) + 128u * (rb + 1), dQr + 2 * ks, dKB + 2 * ks, id_rfa, ks > 0);
          ptx::umma_commit(bar(kSFull0 + rb));
        }
      }
      ptx::mbar_wait(bar(kFullV0 + s), ph);
#pragma unroll 1
      for (int rb = 0; rb < 2; ++rb) {
        ptx::mbar_wait(bar(kPFull0 + rb), it & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const int n_ks = 8 * (rb + 1);
#pragma unroll 1
          for (int ks = 0; ks < n_ks; ++ks) ptx::umma_ts(tmem + o_col(rb), tmem + base_col(rb) + 8 * ks, dV + 128 * ks, id_pv, ks > 0);
#pragma unroll 1
          for (int ks = 0; ks < p.cnp / 16; ++ks) ptx::umma_ts(tmem + o_col(rb), tmem + base_col(rb) + 64u * (rb + 1) + 8 * ks, dBT + 128 * ks, id_pv, 1);
          ptx::umma_commit(bar(kOFull0 + rb));
          if (rb == 1) ptx::umma_commit(bar(kFree0 + s));
        }
      }
    }
  } else if (warp >= 10) {
    // =================================== chunk-statistics warpgroup (one-pass mode) =============
    // Runs on its own four warps, up to one window AHEAD of the softmax warps (on the prefetched stage), so the statistics of
    // window i + 1 are computed under the softmax of window i.  Scratch = rows 32-63 of the four k_bar / beta tile buffers
    // (never read when cnp <= 32).
    if (p.fuse) {
      const int t = tid - 320, w4 = warp - 10;
      float* const partQ = reinterpret_cast<float*>(kb_ptr(0) + 4096);           // [16][64] partial sums (q) / beta partials 0-15
      float* const partK = reinterpret_cast<float*>(kb_ptr(0) + 8192 + 4096);    // [16][64] partial sums (k) / beta partials 16-31
      float* const meanv = reinterpret_cast<float*>(kb_ptr(1) + 4096);           // [2][4][64] chunk means, then [2][4][64] Linear + LN
      float* const yv = meanv + 512;
      float* const om = reinterpret_cast<float*>(kb_ptr(1) + 8192 + 4096);       // [4][64] omega
      float* const lgv = om + 256;                                               // [256] logits, then softmax weights
      float* const red = lgv + 256;                                              // [4] chunk max | [4] chunk sum
      const int gpc = p.chunk >> 4;                                              // 16-token groups per chunk
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t phq = (it >> 1) & 1;
        int wi, bh;
        decode(item, wi, bh);
        const long long c0 = (long long)bh * p.n_chunks + (long long)wi * p.cpw;   // first chunk of this window
        ptx::mbar_wait(bar(kFullQK0 + s), phq);
        const uint8_t* Qs = stage_ptr(s);
        const uint8_t* Ks = Qs + 32768;
        const uint8_t* Vs = Qs + 65536;
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {             // column sums of q and k over 16-token groups
          const uint8_t* src = side ? Ks : Qs;
          const int c8 = t & 7, g = t >> 3;
          float acc[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 4
          for (int j = 0; j < 16; ++j) {
            const int row = 16 * g + j;
            const uint4 raw = *reinterpret_cast<const uint4*>(src + row * 128 + ((c8 ^ (row & 7)) << 4));
            const T* e8 = reinterpret_cast<const T*>(&raw);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += to_f32(e8[e]);
          }
          float* dst = (side ? partK : partQ) + g * 64 + 8 * c8;
          *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
        ptx::named_bar_sync(3, kStatsThreads);
        for (int idx = t; idx < 2 * p.cpw * 64; idx += kStatsThreads) {
          const int f = idx & 63, cc = (idx >> 6) % p.cpw, side = idx / (64 * p.cpw);
          const float* part = side ? partK : partQ;
          float a = 0.f;
          for (int g = cc * gpc; g < (cc + 1) * gpc; ++g) a += part[g * 64 + f];
          meanv[(side * 4 + cc) * 64 + f] = a / (float)p.chunk;
        }
        ptx::named_bar_sync(3, kStatsThreads);
        if (w4 < p.cpw) {
          // one warp per chunk, q side then k side: lane owns outputs lane and lane + 32; LayerNorm by shuffles
#pragma unroll 1
          for (int side = 0; side < 2; ++side) {
            const int cc = w4;
            const float* Wm = side ? p.w_k : p.w_q;
            const float* bv = side ? p.b_k : p.b_q;
            const float* gain = side ? p.g_k : p.g_q;
            const float* lb = side ? p.beta_k : p.beta_q;
            const float* mv = meanv + (side * 4 + cc) * 64;
            float y0 = 0.f, y1 = 0.f;
            if (Wm) {
              y0 = bv ? __ldg(bv + lane) : 0.f;
              y1 = bv ? __ldg(bv + lane + 32) : 0.f;
              const float4* w0 = reinterpret_cast<const float4*>(Wm + lane * 64);
              const float4* w1 = reinterpret_cast<const float4*>(Wm + (lane + 32) * 64);
#pragma unroll 4
              for (int i4 = 0; i4 < 16; ++i4) {
                const float4 a = __ldg(w0 + i4), c = __ldg(w1 + i4);
                const float4 m = *reinterpret_cast<const float4*>(mv + 4 * i4);
                y0 = fmaf(a.x, m.x, y0); y0 = fmaf(a.y, m.y, y0); y0 = fmaf(a.z, m.z, y0); y0 = fmaf(a.w, m.w, y0);
                y1 = fmaf(c.x, m.x, y1); y1 = fmaf(c.y, m.y, y1); y1 = fmaf(c.z, m.z, y1); y1 = fmaf(c.w, m.w, y1);
              }
              if (gain) {
                const float mu = warp_sum(y0 + y1) * (1.0f / 64);
                const float d0 = y0 - mu, d1 = y1 - mu;
                const float inv = 1.0f / sqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.0f / 64) + p.ln_eps);
                y0 = d0 * inv * __ldg(gain + lane) + __ldg(lb + lane);
                y1 = d1 * inv * __ldg(gain + lane + 32) + __ldg(lb + lane + 32);
              }
            }
            yv[(side * 4 + cc) * 64 + lane] = y0;
            yv[(side * 4 + cc) * 64 + lane + 32] = y1;
            if (side) {
              p.kbar_w[(c0 + cc) * 64 + lane] = y0;
              p.kbar_w[(c0 + cc) * 64 + lane + 32] = y1;
            }
          }
          __syncwarp();
          // omega of this chunk (same warp: no block barrier needed)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int f = lane + 32 * hh;
            float o = p.w_q ? p.mu_coeff * (yv[w4 * 64 + f] + yv[(4 + w4) * 64 + f]) : 0.f;
            if (p.noise) o += __ldg(p.noise + (c0 + w4) * 64 + f);
            om[w4 * 64 + f] = o;
          }
        }
        ptx::named_bar_sync(3, kStatsThreads);
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {                   // phi-logit of token rows t and t + 128
          const int row = t + 128 * hh;
          const float* omr = om + (row / p.chunk) * 64;
          float acc = 0.f;
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {
            const uint4 raw = *reinterpret_cast<const uint4*>(Ks + row * 128 + ((c8 ^ (row & 7)) << 4));
            const T* e8 = reinterpret_cast<const T*>(&raw);
            const float4 o0 = *reinterpret_cast<const float4*>(omr + 8 * c8), o1 = *reinterpret_cast<const float4*>(omr + 8 * c8 + 4);
            const float ov[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float f = to_f32(e8[e]); acc = fmaf(f, ov[e] - 0.5f * f, acc); }
          }
          lgv[row] = 0.125f * acc;
        }
        ptx::named_bar_sync(3, kStatsThreads);
        if (w4 < p.cpw) {                                  // one warp per chunk: max and sum over the chunk's tokens
          const float* lc = lgv + w4 * p.chunk;
          float mx = kNegInf;
          for (int j = lane; j < p.chunk; j += 32) mx = fmaxf(mx, lc[j]);
          mx = warp_max(mx);
          float sm_ = 0.f;
          for (int j = lane; j < p.chunk; j += 32) sm_ += exp_nonpos(lc[j] - mx);
          sm_ = warp_sum(sm_);
          if (lane == 0) { red[w4] = mx; red[4 + w4] = sm_; }
        }
        ptx::named_bar_sync(3, kStatsThreads);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int row = t + 128 * hh, cc = row / p.chunk;
          lgv[row] = exp_nonpos(lgv[row] - red[cc]) / red[4 + cc];
        }
        ptx::mbar_wait(bar(kFullV0 + s), phq);
        ptx::named_bar_sync(3, kStatsThreads);
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {                   // beta partial sums over 8-token groups
          const int c8 = t & 7, g = (t >> 3) + 16 * hh;
          float acc[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int row = 8 * g + j;
            const uint4 raw = *reinterpret_cast<const uint4*>(Vs + row * 128 + ((c8 ^ (row & 7)) << 4));
            const T* e8 = reinterpret_cast<const T*>(&raw);
            const float w_ = lgv[row];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(w_, to_f32(e8[e]), acc[e]);
          }
          float* dst = (hh ? partK : partQ) + (g & 15) * 64 + 8 * c8;
          *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
        ptx::mbar_arrive(bar(kStatsDone0 + s));            // this stage's tiles are not read again by this warpgroup
        ptx::named_bar_sync(3, kStatsThreads);
        for (int idx = t; idx < p.cpw * 64; idx += kStatsThreads) {
          const int f = idx & 63, cc = idx >> 6;
          const int g8 = p.chunk >> 3;                      // 8-token groups per chunk
          float a = 0.f;
          for (int g = cc * g8; g < (cc + 1) * g8; ++g) a += (g < 16 ? partQ[g * 64 + f] : partK[(g - 16) * 64 + f]);
          p.beta_w[(c0 + cc) * 64 + f] = a;
        }
        __threadfence();
        ptx::named_bar_sync(3, kStatsThreads);
        if (t == 0) {
          unsigned int* fl = p.flags + (long long)bh * p.n_win + wi;
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(fl), "r"(1u) : "memory");
        }
      }
    }
  } else {
    // =================================== softmax / epilogue warps ===============================
    const int rb = warp >> 2, wl = warp & 3;
    const int i = 32 * wl + lane;                        // query row inside the row-block
    const uint32_t trow = tmem + ((uint32_t)(32 * wl) << 16);
    const uint32_t cS = base_col(rb), cO = o_col(rb);
    const int n_blk = rb + 1;                            // 128-key blocks this row-block attends to
    const float scale_log2 = 0.125f * kLog2e;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int s = it & 1;
      int wi, bh;
      decode(item, wi, bh);
      const int h = bh % p.H, b = bh / p.H;
      const int qc = (wi * kWin + 128 * rb + i) / p.chunk;     // chunk keys c < qc are visible (causal_eva.py:725-739)
      ptx::mbar_wait(bar(kSFull0 + rb), it & 1);
      ptx::tc_fence_after();
      // ---- pass 1: row maximum over the visible keys ----
      float m0 = kNegInf, m1 = kNegInf;
#pragma unroll 1
      for (int g = 0; g < 4 * n_blk; ++g) {
        float v[32];
        ptx::tmem_ld16(trow + cS + 32 * g, reinterpret_cast<uint32_t*>(v));
        ptx::tmem_ld16(trow + cS + 32 * g + 16, reinterpret_cast<uint32_t*>(v) + 16);
        ptx::tmem_ld_wait();
        const int lim = (g >> 2) < rb ? 1 << 30 : i - 32 * (g & 3);       // key e of this piece is visible iff e <= lim
        if (p.bias) {
          const float* db = dbias + (128 * rb + i - 32 * g);              // distance of key e: 128 rb + i - (32 g + e) >= 0 when visible
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            m0 = fmaxf(m0, e <= lim ? fmaf(v[e], scale_log2, db[-e]) : kNegInf);
            m1 = fmaxf(m1, e + 1 <= lim ? fmaf(v[e + 1], scale_log2, db[-e - 1]) : kNegInf);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            m0 = fmaxf(m0, e <= lim ? v[e] * scale_log2 : kNegInf);
            m1 = fmaxf(m1, e + 1 <= lim ? v[e + 1] * scale_log2 : kNegInf);
          }
        }
      }
#pragma unroll 1
      for (int g = 0; g < p.cnp / 16; ++g) {
        float v[16];
        ptx::tmem_ld16(trow + cS + 128 * n_blk + 16 * g, reinterpret_cast<uint32_t*>(v));
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) m0 = fmaxf(m0, 16 * g + e < qc ? v[e] * scale_log2 : kNegInf);
      }
      const float nmx = -fmaxf(m0, m1);
      // ---- pass 2: P = exp2(scale * s - max), 16-bit pairs written over the first half of the columns just read ----
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
      for (int g = 0; g < 4 * n_blk; ++g) {
        float v[32];
        uint32_t pk[16];
        ptx::tmem_ld16(trow + cS + 32 * g, reinterpret_cast<uint32_t*>(v));
        ptx::tmem_ld16(trow + cS + 32 * g + 16, reinterpret_cast<uint32_t*>(v) + 16);
        ptx::tmem_ld_wait();
        const int lim = (g >> 2) < rb ? 1 << 30 : i - 32 * (g & 3);
        if (p.bias) {
          const float* db = dbias + (128 * rb + i - 32 * g);
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float a = e <= lim ? ex2(fmaf(v[e], scale_log2, db[-e] + nmx)) : 0.f;
            const float c = e + 1 <= lim ? ex2(fmaf(v[e + 1], scale_log2, db[-e - 1] + nmx)) : 0.f;
            s0 += a; s1 += c;
            pk[e >> 1] = Fmt<T>::pack2(a, c);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float a = e <= lim ? ex2(fmaf(v[e], scale_log2, nmx)) : 0.f;
            const float c = e + 1 <= lim ? ex2(fmaf(v[e + 1], scale_log2, nmx)) : 0.f;
            s0 += a; s1 += c;
            pk[e >> 1] = Fmt<T>::pack2(a, c);
          }
        }
        ptx::tmem_st16(trow + cS + 16 * g, pk);
      }
#pragma unroll 1
      for (int g = 0; g < p.cnp / 16; ++g) {
        float v[16];
        uint32_t pk[8];
        ptx::tmem_ld16(trow + cS + 128 * n_blk + 16 * g, reinterpret_cast<uint32_t*>(v));
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const float a = 16 * g + e < qc ? ex2(fmaf(v[e], scale_log2, nmx)) : 0.f;
          const float c = 16 * g + e + 1 < qc ? ex2(fmaf(v[e + 1], scale_log2, nmx)) : 0.f;
          s0 += a; s1 += c;
          pk[e >> 1] = Fmt<T>::pack2(a, c);
        }
        ptx::tmem_st8(trow + cS + 64 * n_blk + 8 * g, pk);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(kPFull0 + rb));
      // ---- epilogue: O / rowsum -> the (dead) Q rows of this row-block -> TMA store ----
      ptx::mbar_wait(bar(kOFull0 + rb), it & 1);
      ptx::tc_fence_after();
      float o[64];
#pragma unroll
      for (int g = 0; g < 4; ++g) ptx::tmem_ld16(trow + cO + 16 * g, reinterpret_cast<uint32_t*>(o) + 16 * g);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(kOFree0 + rb));
      const float inv = 1.0f / (s0 + s1);
      const int orow = 128 * rb + i;
      if (p.lse_out) p.lse_out[((long long)b * p.H + h) * p.N + wi * kWin + orow] = log2f(s0 + s1) - nmx;
      if (p.fuse) ptx::mbar_wait(bar(kStatsDone0 + s), (it >> 1) & 1);   // the statistics warpgroup has read this stage's q tile
      uint8_t* row = stage_ptr(s) + orow * 128;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        *reinterpret_cast<uint4*>(row + ((ch ^ (orow & 7)) << 4)) =
            make_uint4(Fmt<T>::pack2(o[8 * ch] * inv, o[8 * ch + 1] * inv), Fmt<T>::pack2(o[8 * ch + 2] * inv, o[8 * ch + 3] * inv),
                       Fmt<T>::pack2(o[8 * ch + 4] * inv, o[8 * ch + 5] * inv), Fmt<T>::pack2(o[8 * ch + 6] * inv, o[8 * ch + 7] * inv));
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(1 + rb, 128);
      if (wl == 0 && ptx::elect_one()) {
        ptx::tma_store_4d(&t_o, ptx::smem_u32(stage_ptr(s)) + rb * 16384, 0, h, wi * kWin + 128 * rb, b);
        ptx::bulk_commit_group();
        ptx::bulk_wait_read0();                          // the Q tile of this stage may be overwritten by the next load
        ptx::mbar_arrive(bar(kFree0 + s));
      }
    }
    if (wl == 0 && ptx::elect_one()) ptx::bulk_wait_all();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 9) ptx::tmem_dealloc(tmem, 512);
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

// [B, N, H, 64] view -> (64, H, N, B) tensor map with a (64, 1, rows, 1) box, 128-byte swizzle.  Tensor-map strides have to grow
// with the dimension, so a time-major activation (token stride > batch stride: fairseq's [T, B, C]) gets (64, H, B, N) and a
// (64, 1, 1, rows) box instead -- same tile in shared memory -- and *swapped tells the kernel to exchange the two coordinates.
static bool make_seq_map(CUtensorMap* tm, const void* ptr, long long sb, long long sn, long long sh, const Geo& g, int io_dtype, int rows,
                         bool* swapped = nullptr) {
  auto enc = get_encode();
  if (!enc) return false;
  const bool sw = swapped && sb < sn;
  if (swapped) *swapped = sw;
  const cuuint64_t dims[4] = {64, (cuuint64_t)g.H, (cuuint64_t)(sw ? g.B : g.N), (cuuint64_t)(sw ? g.N : g.B)};
  const cuuint64_t strides[3] = {(cuuint64_t)sh * 2, (cuuint64_t)(sw ? sb : sn) * 2, (cuuint64_t)(sw ? sn : sb) * 2};
  const cuuint32_t box[4] = {64, 1, (cuuint32_t)(sw ? 1 : rows), (cuuint32_t)(sw ? rows : 1)};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tm, io_dtype == EVA_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims,
             strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T>
static cudaError_t launch_t(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const float* kbar, const float* beta,
                            const float* bias, void* out, cudaStream_t st, const char** msg, const EvaAdaptive* ada, const float* noise,
                            unsigned int* flags, float* lse_out) {
  CUtensorMap tq, tk, tv, to;
  bool swq = false, swk = false, swv = false;
  if (!make_seq_map(&tq, q.ptr, q.sb, q.sn, q.sh, g, io_dtype, kWin, &swq) || !make_seq_map(&tk, k.ptr, k.sb, k.sn, k.sh, g, io_dtype, kWin, &swk) ||
      !make_seq_map(&tv, v.ptr, v.sb, v.sn, v.sh, g, io_dtype, kWin, &swv) ||
      !make_seq_map(&to, out, (long long)g.N * g.H * 64, (long long)g.H * 64, 64, g, io_dtype, 128)) {
    *msg = "cuTensorMapEncodeTiled failed";
    return cudaErrorInvalidValue;
  }
  Params p{};
  p.B = g.B; p.H = g.H; p.N = g.N; p.n_win = g.N / kWin; p.items = g.B * g.H * p.n_win;
  p.n_chunks = g.n_chunks; p.cnp = (g.n_chunks + 15) & ~15; p.chunk = g.chunk;
  p.kbar = kbar; p.beta = beta; p.bias = bias; p.lse_out = lse_out;
  p.swap = (swq ? 1 : 0) | (swk ? 2 : 0) | (swv ? 4 : 0);
  p.fuse = 0; p.cpw = kWin / (g.chunk > 0 ? g.chunk : kWin);
  if (ada && flags) {                                   // one-pass mode (see causal_one_pass_supported)
    p.fuse = 1;
    p.kbar_w = const_cast<float*>(kbar); p.beta_w = const_cast<float*>(beta); p.flags = flags;
    p.w_q = ada->w_q; p.b_q = ada->b_q; p.g_q = ada->ln_gain_q; p.beta_q = ada->ln_bias_q;
    p.w_k = ada->w_k; p.b_k = ada->b_k; p.g_k = ada->ln_gain_k; p.beta_k = ada->ln_bias_k;
    p.mu_coeff = ada->mu_coeff; p.ln_eps = ada->ln_eps; p.noise = noise;
    const cudaError_t em = cudaMemsetAsync(flags, 0, sizeof(unsigned int) * (size_t)g.B * g.H * p.n_win, st);
    if (em != cudaSuccess) { *msg = "cudaMemsetAsync(flags)"; return em; }
  }
  auto kern = eva_causal_window_kernel<T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynamic);
  if (e != cudaSuccess) { *msg = "cudaFuncSetAttribute"; return e; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.items < sms ? p.items : sms;
  kern<<<grid, kThreads, kDynamic, st>>>(tq, tk, tv, to, p);
  *msg = "kernel launch";
  return cudaGetLastError();
}

}  // namespace causal

bool causal_window_supported(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask,
                             const float* bias, long long bias_sh) {
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("EVA_SM100_DISABLE_FUSED"); disabled = (e && e[0] == '1') ? 1 : 0; }
  if (disabled) return false;
  if (g.dims != 1 || !g.causal || g.D != 64 || g.window != causal::kWin || g.ext != 0 || g.chunk_ext != 0 || mask) return false;
  if (bias && !(g.bias_toeplitz && bias_sh == 0)) return false;      // a dense per-head table would be 256 KB of uncoalesced reads per window
  if (io_dtype != EVA_F16 && io_dtype != EVA_BF16) return false;
  if (g.N % causal::kWin != 0 || g.n_chunks < 1 || g.n_chunks > 64 || g.chunk < 1) return false;
  for (const View* x : {&q, &k, &v}) {
    if (x->sh * 2 % 16 || x->sn * 2 % 16 || x->sb * 2 % 16) return false;
    if (x->sh <= 0 || x->sn <= 0 || x->sb <= 0) return false;
    if (reinterpret_cast<uintptr_t>(x->ptr) % 16) return false;
  }
  return causal::get_encode() != nullptr;
}

// One pass (chunk statistics inside the window kernel): whole chunks per window, at most four of them, at most 32 chunks in all
// (their tiles then leave rows 32-63 of the k_bar / beta buffers free as scratch), Linear on the k side present
// Measured (c5, T = 4096, h = 8, profiles/r02/README.md 8): batch 16: 0.160 ms one pass vs 0.162 ms two passes (statistics kernel 0.071
// + window kernel 0.086); batch 4: 0.061 vs 0.051 ms.  The statistics are SIMT work either way (TMEM is fully taken by the window's
// logits, so they cannot use the tensor cores here), and four extra warps per SM do them no faster than a full-chip kernel does;
// the flag dependencies add latency at small batches.  HBM traffic drops to 1.0x, time does not: OPT-IN
// (EVA_SM100_CAUSAL_ONE_PASS=1 or eva_debug_set_causal_one_pass(1)).
static int g_one_pass_mode = -1;           // -1: environment, 0: off, 1: on
extern "C" int eva_debug_set_causal_one_pass(int mode) {
  const int prev = g_one_pass_mode;
  g_one_pass_mode = mode;
  return prev;
}
bool causal_one_pass_supported(const Geo& g, const EvaAdaptive& ada) {
  static const bool env_on = [] { const char* e = getenv("EVA_SM100_CAUSAL_ONE_PASS"); return e && e[0] == '1'; }();
  if (g_one_pass_mode < 0 ? !env_on : g_one_pass_mode == 0) return false;
  if (g.chunk != 64 && g.chunk != 128 && g.chunk != 256) return false;
  if (((g.n_chunks + 15) & ~15) > 32) return false;
  if (g.n_chunks * g.chunk != g.N) return false;
  return ada.w_k != nullptr;
}

cudaError_t launch_causal_window(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const float* kbar,
                                 const float* beta, const float* bias, void* out, cudaStream_t st, const char** msg,
                                 const EvaAdaptive* ada, const float* noise, unsigned int* flags, float* lse_out) {
  if (io_dtype == EVA_F16) return causal::launch_t<__half>(g, io_dtype, q, k, v, kbar, beta, bias, out, st, msg, ada, noise, flags, lse_out);
  return causal::launch_t<__nv_bfloat16>(g, io_dtype, q, k, v, kbar, beta, bias, out, st, msg, ada, noise, flags, lse_out);
}

}  // namespace eva
