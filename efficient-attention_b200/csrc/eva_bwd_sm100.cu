// tcgen05 backward of the EVA window attention (gradient of eva.py:200-227) for the DeiT-style geometries: head_dim 64, 16-bit I/O,
// windows without halo of at most 64 tokens, at most 64 chunk keys, no padding mask, not causal.  Every other geometry keeps the
// CUDA-core kernel of eva_backward.cu; the chunk-statistics gradient (chunk_stats_bwd_kernel) follows either of them.
//
// One window per CTA iteration, 8 warps, two CTAs per SM (96 KB of tiles, 256 tensor-memory columns each):
//   load   q, k, v, grad_out rows of the window and the item's k_bar / beta rows -> 128-byte-swizzled 16-bit tiles;
//          delta_r = <grad_out_r, out_r>
//   MMA 1  S  [64 x 128] = Q  [64 x 64] . Kx^T       Kx = [local k rows (64) ; k_bar rows (64)]
//          dP [64 x 128] = dO [64 x 64] . Vx^T       Vx = [local v rows      ; beta rows     ]
//   E 1    joint softmax of the row over [local | chunk] logits, P and dS = P o (dP - delta) -> 16-bit tiles, row-major (dS) and
//          key-major (dS^T, P^T).  M = 64 accumulators sit on lanes 0-15 of each tensor-memory quarter: warps w and w + 4 read
//          the same quarter and split the columns (w < 4: local keys, w >= 4: chunk keys), the two (max, sum) pairs of a row meet
//          in shared memory
//   MMA 2  dQ  [64 x 64]  = dS   [64 x 128] . Kx     (B operand MN-major: the k tile as loaded)
//          dKx [128 x 64] = dS^T [128 x 64] . Q
//          dVx [128 x 64] = P^T  [128 x 64] . dO
//   E 2    dQ, dK, dV rows of the window -> float32 stores (rows belong to this window only); chunk-key rows -> atomics into
//          d k_bar / d beta
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "fused_common.cuh"
#include "launch.h"
#include "sm100_ptx.cuh"

namespace eva {
namespace bwdtc {

using fused::IoFmt;
using fused::tile_off;
using fused::tmem_ld_cols;
using fused::ex2;

constexpr int kThreads = 256;
constexpr int kQ = 0, kG = 8192, kKx = 16384, kVx = 32768, kdS = 49152, kP = 65536, kO = 81920, kMisc = 90112;
constexpr int kDelta = kMisc, kPm = kDelta + 256, kPl = kPm + 512, kBar = kPl + 512, kSlot = kBar + 16;
constexpr int kBias = kSlot + 16;                // [L][J] bias rows of the CTA's head, pre-multiplied by log2(e)
constexpr int kSmemFixed = kBias + 1024;         // + slack to align the tiles to 1024 B; + L * J floats when there is a bias
constexpr uint32_t kTmemCols = 256;

struct Params {
  Geo g;
  View q, k, v;
  const float* kbar; const float* beta; const float* bias;
  long long bias_sh;
  const void* out; const void* dout;
  float* dq; float* dk; float* dv; float* dkbar; float* dbeta; float* dbias;
  int total;
  int trace;   // EVA_SM100_TRACE=1: CTA 0 prints the phase clocks of its second window
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
eva_window_bwd_tc_kernel(const __grid_constant__ CUtensorMap t_q, const __grid_constant__ CUtensorMap t_k,
                         const __grid_constant__ CUtensorMap t_v, const __grid_constant__ CUtensorMap t_g,
                         const __grid_constant__ CUtensorMap t_o, const Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (ptx::smem_u32(raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Geo& g = p.g;
  const int L = g.L, C = g.n_chunks;
  const long long HD = (long long)g.H * 64;
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + kSlot);
  const uint32_t bar = ptx::smem_u32(sm + kBar), bar_in = bar + 8;
  float* delta = reinterpret_cast<float*>(sm + kDelta);
  float* pm = reinterpret_cast<float*>(sm + kPm);     // [2][64] row maxima of the two column halves (log2 domain)
  float* pl = reinterpret_cast<float*>(sm + kPl);     // [2][64] row sums
  float* sbias = reinterpret_cast<float*>(sm + kBias);
  // CTA -> one head (its bias table stays in shared memory), windows of that head strided over the CTAs of the head
  const int h = blockIdx.x % g.H, cta_h = blockIdx.x / g.H, ctas_h = gridDim.x / g.H;
  const int per_head = g.B * g.n_windows;
  if (p.bias)
    for (int i = tid; i < L * L; i += kThreads) sbias[i] = __ldg(p.bias + (long long)h * p.bias_sh + i) * kLog2e;

  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(slot), kTmemCols);
  if (tid == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_init(bar_in, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&t_q); ptx::prefetch_tmap(&t_k); ptx::prefetch_tmap(&t_v); ptx::prefetch_tmap(&t_g); ptx::prefetch_tmap(&t_o);
  }
  // rows L .. 63 of the row tiles are never written by the boxes: zero once
  for (int i = tid; i < (64 - L) * 8; i += kThreads) {
    const int off = tile_off(L + (i >> 3), 8 * (i & 7));
    const uint4 z = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(sm + kQ + off) = z;  *reinterpret_cast<uint4*>(sm + kG + off) = z;
    *reinterpret_cast<uint4*>(sm + kKx + off) = z; *reinterpret_cast<uint4*>(sm + kVx + off) = z;
    *reinterpret_cast<uint4*>(sm + kO + off) = z;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;

  constexpr uint32_t fmt = IoFmt<T>::kUmma;
  constexpr uint32_t id_s = ptx::umma_idesc(fmt, fmt, 0, 0, 64, 128);
  constexpr uint32_t id_dq = ptx::umma_idesc(fmt, fmt, 0, 1, 64, 64);
  constexpr uint32_t id_dk = ptx::umma_idesc(fmt, fmt, 1, 1, 128, 64);   // A = dS / P read MN-major (keys contiguous)
  const float scale = 0.125f, scale2 = 0.125f * kLog2e;

  const int qr = warp & 3, hf = warp >> 2;
  const uint32_t trow = tmem + ((uint32_t)(32 * qr) << 16);
  const int r = 16 * qr + (lane & 15);            // row of the M = 64 accumulators this thread reads (lanes 0-15)
  const bool act = lane < 16;

  // contiguous runs of windows per CTA: the item's chunk rows (k_bar / beta) are re-staged only when the image changes
  const int lo = (int)((long long)per_head * cta_h / ctas_h), hi = (int)((long long)per_head * (cta_h + 1) / ctas_h);
  const int nwx = g.gw / g.window;                                 // windows per grid row (1-D: gw = N)
  auto issue_loads = [&](int item) {                               // one thread: five boxes of L rows x 128 B
    const int win = item % g.n_windows, b = item / g.n_windows;
    const int x0 = (win % nwx) * g.window, y0 = g.dims == 2 ? (win / nwx) * g.window : 0;
    ptx::mbar_arrive_expect_tx(bar_in, 5u * (uint32_t)L * 128u);
    ptx::tma_load_5d(ptx::smem_u32(sm + kQ), &t_q, bar_in, 0, h, x0, y0, b);
    ptx::tma_load_5d(ptx::smem_u32(sm + kKx), &t_k, bar_in, 0, h, x0, y0, b);
    ptx::tma_load_5d(ptx::smem_u32(sm + kVx), &t_v, bar_in, 0, h, x0, y0, b);
    ptx::tma_load_5d(ptx::smem_u32(sm + kG), &t_g, bar_in, 0, h, x0, y0, b);
    ptx::tma_load_5d(ptx::smem_u32(sm + kO), &t_o, bar_in, 0, h, x0, y0, b);
  };
  if (tid == 0 && lo < hi) issue_loads(lo);
  long long tk[8];
  int it = 0, b_staged = -1;
#define EVA_MARK(i) if (p.trace && it == 1) tk[i] = clock64();
  for (int item = lo; item < hi; ++item, ++it) {
    const int win = item % g.n_windows, b = item / g.n_windows;
    const int bh = b * g.H + h;
    EVA_MARK(0)
    // ---------------------------------------------- loads ------------------------------------------------------------
    if (b != b_staged) {                        // chunk keys / values of the item: float32 statistics -> the I/O format
      b_staged = b;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int row = pass * 32 + (tid >> 3), ch = tid & 7;
        uint4 ck = make_uint4(0, 0, 0, 0), cv = ck;
        if (row < C) {
          const long long base = ((long long)bh * C + row) * 64 + 8 * ch;
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.kbar + base)), a1 = __ldg(reinterpret_cast<const float4*>(p.kbar + base) + 1);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + base)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + base) + 1);
          ck = make_uint4(IoFmt<T>::pack2(a0.x, a0.y), IoFmt<T>::pack2(a0.z, a0.w), IoFmt<T>::pack2(a1.x, a1.y), IoFmt<T>::pack2(a1.z, a1.w));
          cv = make_uint4(IoFmt<T>::pack2(b0.x, b0.y), IoFmt<T>::pack2(b0.z, b0.w), IoFmt<T>::pack2(b1.x, b1.y), IoFmt<T>::pack2(b1.z, b1.w));
        }
        const int offc = tile_off(64 + row, 8 * ch);
        *reinterpret_cast<uint4*>(sm + kKx + offc) = ck;
        *reinterpret_cast<uint4*>(sm + kVx + offc) = cv;
      }
    }
    ptx::mbar_wait(bar_in, it & 1);
    {                                           // delta_r = <grad_out_r, out_r>: four threads per row, two 16-byte pieces each
      const int row = tid >> 2, c0 = 2 * (tid & 3);
      float part = 0.f;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int off = tile_off(row, 8 * (c0 + cc));
        const uint4 zg = *reinterpret_cast<const uint4*>(sm + kG + off), zo = *reinterpret_cast<const uint4*>(sm + kO + off);
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
      if ((tid & 3) == 0) delta[row] = part;
    }
    EVA_MARK(1)
    ptx::fence_proxy_async_smem();
    __syncthreads();
    const uint64_t dQd = ptx::umma_desc_sw128(ptx::smem_u32(sm + kQ)), dGd = ptx::umma_desc_sw128(ptx::smem_u32(sm + kG));
    const uint64_t dKd = ptx::umma_desc_sw128(ptx::smem_u32(sm + kKx)), dVd = ptx::umma_desc_sw128(ptx::smem_u32(sm + kVx));
    if (tid == 0) {
      ptx::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem, dQd + 2 * ks, dKd + 2 * ks, id_s, ks > 0);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + 128, dGd + 2 * ks, dVd + 2 * ks, id_s, ks > 0);
      ptx::umma_commit(bar);
    }
    ptx::mbar_wait(bar, 0);
    ptx::tc_fence_after();
    EVA_MARK(2)
    // ---------------------------------------------- E 1 --------------------------------------------------------------
    {
      const bool row_live = act && r < L;
      const int ncol = hf ? C : L;
      float s[64];
      tmem_ld_cols<64>(trow + 64 * hf, reinterpret_cast<uint32_t*>(s));
      ptx::tmem_ld_wait();
      const float* brow = (p.bias && !hf && row_live) ? sbias + r * L : nullptr;
      float mloc = kNegInf;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        float x = s[j] * scale2;
        if (brow && j < ncol) x += brow[j];
        x = j < ncol ? x : kNegInf;
        s[j] = x;
        mloc = fmaxf(mloc, x);
      }
      float lloc = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) { s[j] = ex2(s[j] - mloc); lloc += s[j]; }      // s <- exp2(logit - half maximum)
      if (act) { pm[hf * 64 + r] = mloc; pl[hf * 64 + r] = lloc; }
      __syncthreads();
      const float m0 = pm[r], m1 = pm[64 + r];
      const float mm = fmaxf(m0, m1);
      const float linv = 1.0f / (pl[r] * ex2(m0 - mm) + pl[64 + r] * ex2(m1 - mm));
      const float pscale = row_live ? ex2(mloc - mm) * linv : 0.f;                  // dead rows: P = dS = 0
      const float dl = delta[r];
      float* dbrow = (p.dbias && !hf && row_live) ? p.dbias + (long long)h * p.bias_sh + (long long)r * g.J : nullptr;
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        float dp[16];
        ptx::tmem_ld16(trow + 128 + 64 * hf + 16 * blk, reinterpret_cast<uint32_t*>(dp));
        ptx::tmem_ld_wait();
        uint32_t pk[8], pp[8];
#pragma unroll
        for (int jj = 0; jj < 16; jj += 2) {
          const int j = 16 * blk + jj;
          const float p0 = s[j] * pscale, p1 = s[j + 1] * pscale;                  // dead columns: s = exp2(-inf) = 0
          const float d0 = p0 * (dp[jj] - dl), d1 = p1 * (dp[jj + 1] - dl);
          if (dbrow) {
            if (j < ncol) atomicAdd(dbrow + j, d0);
            if (j + 1 < ncol) atomicAdd(dbrow + j + 1, d1);
          }
          pk[jj >> 1] = IoFmt<T>::pack2(d0, d1);
          pp[jj >> 1] = IoFmt<T>::pack2(p0, p1);
        }
        if (act) {
          const int o0 = hf * 8192 + tile_off(r, 16 * blk), o1 = hf * 8192 + tile_off(r, 16 * blk + 8);
          *reinterpret_cast<uint4*>(sm + kdS + o0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(sm + kdS + o1) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          *reinterpret_cast<uint4*>(sm + kP + o0) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
          *reinterpret_cast<uint4*>(sm + kP + o1) = make_uint4(pp[4], pp[5], pp[6], pp[7]);
        }
      }
    }
    EVA_MARK(3)
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    __syncthreads();
    EVA_MARK(4)
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint64_t dSd = ptx::umma_desc_sw128(ptx::smem_u32(sm + kdS));
      // the same row-major [64 rows][2 x 64 keys] tiles read as A^T: M = 128 keys (two 64-key atoms 8192 B apart), K = rows
      const uint64_t dStd = ptx::umma_desc_sw128_mn(ptx::smem_u32(sm + kdS), 8192), dPtd = ptx::umma_desc_sw128_mn(ptx::smem_u32(sm + kP), 8192);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        ptx::umma_ss(tmem, dSd + (uint64_t)((ks >> 2) * (8192 >> 4) + 2 * (ks & 3)), dKd + 128 * ks, id_dq, ks > 0);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + 64, dStd + 128 * ks, dQd + 128 * ks, id_dk, ks > 0);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + 128, dPtd + 128 * ks, dGd + 128 * ks, id_dk, ks > 0);
      ptx::umma_commit(bar);
    }
    ptx::mbar_wait(bar, 1);
    ptx::tc_fence_after();
    if (tid == 0 && item + 1 < hi) issue_loads(item + 1);    // every tile is free again: the next window lands under E 2
    EVA_MARK(5)
    // ---------------------------------------------- E 2 --------------------------------------------------------------
    {
      float y[32];
      tmem_ld_cols<32>(trow + 32 * hf, reinterpret_cast<uint32_t*>(y));
      ptx::tmem_ld_wait();
      if (act && r < L) {
        const int tok = group_token(g, win, r, g.window, 0);
        float4* dst = reinterpret_cast<float4*>(p.dq + (((long long)b * g.N + tok) * g.H + h) * 64 + 32 * hf);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_float4(y[4 * i] * scale, y[4 * i + 1] * scale, y[4 * i + 2] * scale, y[4 * i + 3] * scale);
      }
      float z[32];
      tmem_ld_cols<32>(trow + 64 + 32 * hf, reinterpret_cast<uint32_t*>(y));
      tmem_ld_cols<32>(trow + 128 + 32 * hf, reinterpret_cast<uint32_t*>(z));
      ptx::tmem_ld_wait();
      const int key = 32 * qr + lane;
      if (key < 64) {
        if (key < L) {
          const int tok = group_token(g, win, key, g.window, 0);
          const long long base = (((long long)b * g.N + tok) * g.H + h) * 64 + 32 * hf;
          float4* dk4 = reinterpret_cast<float4*>(p.dk + base);
          float4* dv4 = reinterpret_cast<float4*>(p.dv + base);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            dk4[i] = make_float4(y[4 * i] * scale, y[4 * i + 1] * scale, y[4 * i + 2] * scale, y[4 * i + 3] * scale);
            dv4[i] = make_float4(z[4 * i], z[4 * i + 1], z[4 * i + 2], z[4 * i + 3]);
          }
        }
      } else if (key - 64 < C) {
        const long long base = ((long long)bh * C + (key - 64)) * 64 + 32 * hf;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.dkbar + base + i), "f"(y[i] * scale), "f"(y[i + 1] * scale),
                       "f"(y[i + 2] * scale), "f"(y[i + 3] * scale) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.dbeta + base + i), "f"(z[i]), "f"(z[i + 1]), "f"(z[i + 2]),
                       "f"(z[i + 3]) : "memory");
        }
      }
    }
    EVA_MARK(6)
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    EVA_MARK(7)
    if (p.trace && it == 1 && blockIdx.x == 0 && (tid == 0 || tid == 128 || tid == 16))
      printf("bwd tc trace tid %d: load %lld | sync+mma1 %lld | E1 %lld | sync %lld | mma2 %lld | E2 %lld | sync %lld | total %lld\n", tid,
             tk[1] - tk[0], tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3], tk[5] - tk[4], tk[6] - tk[5], tk[7] - tk[6], tk[7] - tk[0]);
  }
#undef EVA_MARK
  if (warp == 0) ptx::tmem_dealloc(tmem, kTmemCols);
}

template <typename T>
static cudaError_t launch_t(const Params& p, cudaStream_t st) {
  auto kern = eva_window_bwd_tc_kernel<T>;
  const int smem = kSmemFixed + (p.bias ? p.g.L * p.g.L * (int)sizeof(float) : 0);
  const Geo& g = p.g;
  const int io = std::is_same<T, __half>::value ? EVA_F16 : EVA_BF16;
  const int bw = g.window, bh_ = g.dims == 2 ? g.window : 1;
  View vo, vg;
  vo.ptr = p.out;  vo.sb = (long long)g.N * g.H * 64; vo.sn = (long long)g.H * 64; vo.sh = 64;
  vg = vo; vg.ptr = p.dout;
  CUtensorMap tq, tk_, tv, tg, to;
  if (!fused::make_box_map(&tq, p.q, g, io, bw, bh_) || !fused::make_box_map(&tk_, p.k, g, io, bw, bh_) ||
      !fused::make_box_map(&tv, p.v, g, io, bw, bh_) || !fused::make_box_map(&tg, vg, g, io, bw, bh_) ||
      !fused::make_box_map(&to, vo, g, io, bw, bh_))
    return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // a whole number of CTAs per head, two CTAs per SM
  int per_head = (2 * sms) / p.g.H;
  if (per_head < 1) per_head = 1;
  if (per_head > p.g.B * p.g.n_windows) per_head = p.g.B * p.g.n_windows;
  kern<<<per_head * p.g.H, kThreads, smem, st>>>(tq, tk_, tv, tg, to, p);
  return cudaGetLastError();
}

}  // namespace bwdtc

static int g_bwd_tc_count = 0;
static int g_bwd_tc_mode = -1;   // -1: environment (EVA_SM100_BWD_SIMT=1 disables), 0: off, 1: on

bool bwd_tc_enabled() {
  static const bool env_off = [] { const char* e = getenv("EVA_SM100_BWD_SIMT"); return e && e[0] == '1'; }();
  return !(g_bwd_tc_mode == 0 || (g_bwd_tc_mode < 0 && env_off));
}
void note_bwd_tc_launch() { ++g_bwd_tc_count; }

bool window_bwd_tc_supported(const Geo& g, int io_dtype, const uint8_t* mask) {
  if (!bwd_tc_enabled()) return false;
  return g.D == 64 && (io_dtype == EVA_F16 || io_dtype == EVA_BF16) && !mask && !g.causal && g.ext == 0 && g.chunk_ext == 0 &&
         g.L <= 64 && g.J == g.L && g.n_chunks >= 1 && g.n_chunks <= 64 &&
         (long long)g.B * g.H * g.n_windows <= 0x7fffffffLL;
}

cudaError_t launch_window_bwd_tc(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const float* kbar,
                                 const float* beta, const float* bias, long long bias_sh, const void* out, const void* dout,
                                 float* dq, float* dk, float* dv, float* dkbar, float* dbeta, float* dbias, cudaStream_t st) {
  bwdtc::Params p;
  p.g = g; p.q = q; p.k = k; p.v = v;
  p.kbar = kbar; p.beta = beta; p.bias = bias; p.bias_sh = bias_sh;
  p.out = out; p.dout = dout;
  p.dq = dq; p.dk = dk; p.dv = dv; p.dkbar = dkbar; p.dbeta = dbeta; p.dbias = dbias;
  p.total = g.B * g.H * g.n_windows;
  static const int trace = [] { const char* e = getenv("EVA_SM100_TRACE"); return (e && e[0] == '1') ? 1 : 0; }();
  p.trace = trace;
  ++g_bwd_tc_count;
  return io_dtype == EVA_F16 ? bwdtc::launch_t<__half>(p, st) : bwdtc::launch_t<__nv_bfloat16>(p, st);
}

}  // namespace eva

// diagnostics (not part of the public ABI): how many backward calls took the tcgen05 window kernel; force it off / on
extern "C" int eva_debug_bwd_tc_count(void) { return eva::g_bwd_tc_count; }
extern "C" void eva_debug_set_bwd_tc(int mode) { eva::g_bwd_tc_mode = mode; }
