// The key draw of randomized attention (randomized_attention.py:36-40: one key index per query from pi = softmax(scale q k^T),
// `torch.multinomial` there) on tcgen05 for head_dim 64 / 16-bit I/O, WITHOUT the [N, N] probabilities: the Gumbel-max trick --
// argmax_m (scale q_n . k_m + G_nm), G i.i.d. standard Gumbel, is distributed as pi_n -- so one pass S = Q K^T per key tile with a
// running (max, argmax) per query row replaces softmax + multinomial.  G comes from a counter-based hash of (seed, item, n, m), or
// from an explicit [B, H, N, N] tensor (tests).  256 threads: warps w and w + 4 share a TMEM lane quarter and split the 128 key
// columns of a tile; K tiles are double-buffered with cp.async.
#include <stdlib.h>

#include "common.cuh"
#include "fused_common.cuh"
#include "launch.h"
#include "sm100_ptx.cuh"

namespace eva {
namespace rasample {

constexpr int kThreads = 256;
constexpr int kQ = 0, kK = 16384, kPm = 49152, kPi = kPm + 1024, kBar = kPi + 1024, kTmemPtr = kBar + 16, kSmemBytes = kTmemPtr + 16;

struct Params {
  int B, H, N, items;           // items = B * H * query blocks of 128
  unsigned long long seed;
  const float* gumbel;          // explicit noise [B * H, N, N] or NULL
  long long* k_ind;             // [B * H, N]
};

template <typename T> struct Fmt;
template <> struct Fmt<__half> { static constexpr uint32_t kUmma = ptx::kFmtF16; };
template <> struct Fmt<__nv_bfloat16> { static constexpr uint32_t kUmma = ptx::kFmtBF16; };

template <typename T>
__device__ __forceinline__ void load_tile_async(const View& x, int b, int h, int n0, int N, uint8_t* dst) {
#pragma unroll
  for (int it = 0; it < 1024 / kThreads; ++it) {
    const int idx = it * kThreads + threadIdx.x, row = idx >> 3, ch = idx & 7;
    const bool ok = n0 + row < N;
    const uint4* src = reinterpret_cast<const uint4*>(x.row<T>(b, ok ? n0 + row : 0, h)) + ch;
    const uint32_t d = ptx::smem_u32(dst + row * 128 + ((ch ^ (row & 7)) << 4));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(ok ? 16 : 0) : "memory");
  }
}

// two standard Gumbel variates from one 64-bit hash of the counter (murmur3 finaliser)
__device__ __forceinline__ void gumbel2(unsigned long long seed, unsigned long long ctr, float& g0, float& g1) {
  unsigned long long x = seed ^ (ctr * 0x9E3779B97F4A7C15ull);
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  const float u0 = ((float)(unsigned)(x & 0xffffffu) + 0.5f) * (1.0f / 16777216.0f);
  const float u1 = ((float)(unsigned)((x >> 32) & 0xffffffu) + 0.5f) * (1.0f / 16777216.0f);
  g0 = -__logf(-__logf(u0));
  g1 = -__logf(-__logf(u1));
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 3) ra_sample_tc_kernel(const View q, const View k, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* const sm = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qr = warp & 3, hf = warp >> 2, r = 32 * qr + lane;
  float* const pm = reinterpret_cast<float*>(sm + kPm);      // [2][128] best value of each column half
  int* const pi = reinterpret_cast<int*>(sm + kPi);          // [2][128] its key index
  const uint32_t bar = ptx::smem_u32(sm + kBar);
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + kTmemPtr);
  constexpr uint32_t fmt = Fmt<T>::kUmma;
  constexpr uint32_t id_s = ptx::umma_idesc(fmt, fmt, 0, 0, 128, 128);
  if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr)), 128);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t trow = tmem + ((uint32_t)(32 * qr) << 16);
  const uint64_t dQ = ptx::umma_desc_sw128(ptx::smem_u32(sm + kQ)), dK = ptx::umma_desc_sw128(ptx::smem_u32(sm + kK));
  const float scale = 0.125f;
  const int qblocks = (p.N + 127) >> 7, tiles = qblocks;
  uint32_t phase = 0, gs = 0;
  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    const int qb = item % qblocks, bh = item / qblocks, b = bh / p.H, h = bh % p.H;
    const int n = 128 * qb + r;
    load_tile_async<T>(q, b, h, 128 * qb, p.N, sm + kQ);
    load_tile_async<T>(k, b, h, 0, p.N, sm + kK + (gs & 1) * 16384);
    asm volatile("cp.async.commit_group;" ::: "memory");
    float best = kNegInf;
    int best_m = 0;
    for (int t = 0; t < tiles; ++t, ++gs) {
      const int buf = (int)(gs & 1);
      if (t + 1 < tiles) load_tile_async<T>(k, b, h, 128 * (t + 1), p.N, sm + kK + (buf ^ 1) * 16384);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncthreads();
      if (warp == 0 && ptx::elect_one()) {
        ptx::tc_fence_after();
        const uint64_t dKb = dK + (uint64_t)(buf * 1024);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem, dQ + 2 * ks, dKb + 2 * ks, id_s, ks > 0);
        ptx::umma_commit(bar);
      }
      ptx::mbar_wait(bar, phase & 1);
      ++phase;
      ptx::tc_fence_after();
      const int m0 = 128 * t + 64 * hf;                    // my 64 key columns of the tile
#pragma unroll 1
      for (int g16 = 0; g16 < 4; ++g16) {
        float s[16];
        ptx::tmem_ld16(trow + 64 * hf + 16 * g16, reinterpret_cast<uint32_t*>(s));
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const int m = m0 + 16 * g16 + e;
          float g0, g1;
          if (p.gumbel) {
            const float* gp = p.gumbel + ((long long)bh * p.N + (n < p.N ? n : 0)) * p.N;
            g0 = m < p.N ? __ldg(gp + m) : 0.f;
            g1 = m + 1 < p.N ? __ldg(gp + m + 1) : 0.f;
          } else {
            gumbel2(p.seed, ((unsigned long long)bh * p.N + n) * (unsigned long long)((p.N + 1) >> 1) + (unsigned long long)(m >> 1), g0, g1);
          }
          const float v0 = m < p.N ? fmaf(scale, s[e], g0) : kNegInf, v1 = m + 1 < p.N ? fmaf(scale, s[e + 1], g1) : kNegInf;
          if (v0 > best) { best = v0; best_m = m; }
          if (v1 > best) { best = v1; best_m = m + 1; }
        }
      }
      ptx::tc_fence_before();                              // the accumulator and the other K buffer are rewritten by the next tile
    }
    pm[128 * hf + r] = best;
    pi[128 * hf + r] = best_m;
    __syncthreads();
    if (hf == 0 && n < p.N) p.k_ind[(long long)bh * p.N + n] = pm[128 + r] > pm[r] ? pi[128 + r] : pi[r];
    __syncthreads();                                       // pm / pi / the Q tile are rewritten by the next item
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 0) ptx::tmem_dealloc(tmem, 128);
}

}  // namespace rasample

bool ra_sample_tc_supported(int D, int io_dtype, const View& q, const View& k) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("EVA_SM100_DISABLE_FUSED"); off = (e && e[0] == '1') ? 1 : 0; }
  if (off || D != 64 || (io_dtype != EVA_F16 && io_dtype != EVA_BF16)) return false;
  for (const View* x : {&q, &k})
    if (x->sh * 2 % 16 || x->sn * 2 % 16 || x->sb * 2 % 16 || reinterpret_cast<uintptr_t>(x->ptr) % 16) return false;
  return true;
}

cudaError_t launch_ra_sample_tc(int B, int H, int N, int io_dtype, const View& q, const View& k, unsigned long long seed,
                                const float* gumbel, long long* k_ind, cudaStream_t st) {
  rasample::Params p{B, H, N, B * H * ((N + 127) / 128), seed, gumbel, k_ind};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int dyn = rasample::kSmemBytes + 1024;
  const int grid = p.items < 3 * sms ? p.items : 3 * sms;
  auto go = [&](auto kern) -> cudaError_t {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return e;
    kern<<<grid, rasample::kThreads, dyn, st>>>(q, k, p);
    return cudaGetLastError();
  };
  return io_dtype == EVA_F16 ? go(rasample::ra_sample_tc_kernel<__half>) : go(rasample::ra_sample_tc_kernel<__nv_bfloat16>);
}

}  // namespace eva
