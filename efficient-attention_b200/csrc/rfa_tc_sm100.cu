// Performer (FAVOR+) linear attention on tcgen05 / TMEM for sm_100a: the default configuration of the reference's 'performer'
// (kernelized_attention.py:21-55, 115-120, 301-320: proj_method 'favorp', 64 random features) at head_dim 64 with 16-bit I/O.
// Everything else (other feature maps / widths, float32 I/O, cosFormer re-weighting) stays on the CUDA-core kernels of
// rfa_kernels.cu.
//
// One (batch, head) item per CTA iteration, 256 threads (TMEM lane = token of the 128-token tile; warps w and w + 4 share a lane
// quarter and split the 64 feature columns of every epilogue), two CTAs per SM (100 KB of tiles, 256 TMEM columns).  Per item, with W' = d^-1/4 W in 16 bits ([features][d], K-major):
//   keys     per key tile:    DD = K W'^T (M = 128 tokens, N = 64 features); the key stabiliser of the reference (max over every token
//                             and feature of DD) is a RUNNING maximum: phi~(k) = exp(DD - |k|^2 d^-1/2 / 2 - s_run) (0 for padding) ->
//                             16-bit tile F [tokens][features];  KVe (+)= F^T V  (A and B both MN-major, M = 64 features),
//                             KSe (+)= F^T 1  (the same MMA against a constant tile of ones: column sums without a reduction),
//                             SV (+)= 1^T V;  when a tile raises the maximum the accumulators are rescaled in tensor memory
//            KV = m^-1/2 KVe + 1e-4 SV -> 16-bit tile [features][d], ksum = m^-1/2 KSe + 1e-4 n_live (float32)
//            (kLogF, ScatterBrain's statistics: a first pass W' K^T for the per-feature maxima instead, see below)
//   phase Q  per query tile:  DD = Q W'^T; phi(q) with the row's own max; den = phi(q) . ksum (float32, thread-local);
//                             O = F KV (B MN-major) -> / max(den, 1e-2) -> 128-byte row stores
// Loads: cp.async, 16 bytes per lane (8 lanes per 128-byte row), into 128-byte-swizzled tiles, two buffers: the tiles of an item form
// a list of stages (pass 1 | pass 2 | queries) and the successor stage -- also across items -- is requested before the current one is
// consumed.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "fused_common.cuh"
#include "launch.h"
#include "sm100_ptx.cuh"

namespace eva {
namespace rfatc {

using fused::tmem_ld_cols;
using fused::ex2;

constexpr int kThreads = 256;       // warps w and w + 4 share a TMEM lane quarter and split the columns of every epilogue
constexpr int kTile = 128;
constexpr uint32_t cDD = 0, cKV = 64, cKS = 128, cSV = 144, cO = 0;   // TMEM columns (O overlays DD: DD is in registers by then)
// shared memory (bytes from the 1024-aligned base)
constexpr int kX = 0, kV = 32768, kF = 65536, kW = 81920, kKVt = 90112, kOnes = 98304, kKsum = 100352, kRed = kKsum + 256,   // X / V: two buffers of 16 KB
              kHs = kRed + 64, kMxs = kHs + 512, kPm = kMxs + 256, kPd = kPm + 1024, kMx2 = kPd + 1024, kBar = kMx2 + 512, kTmemPtr = kBar + 16, kSmemBytes = kTmemPtr + 16;

struct Params {
  int B, H, N, items;
  const float* proj;        // [H, 64, 64]
  const uint8_t* mask;      // [B, N] or NULL
  float* stabv;             // kLogF: [items][64] per-feature maxima of the keys' log-features
  float* part;              // kLogF: [items][64 * 64 + 64] KV | ksum
  int trace;                // EVA_SM100_TRACE=1: phase clocks of one key stage and one query stage (block 0, thread 0)
};

template <typename T> struct Fmt;
template <> struct Fmt<__half> { static constexpr uint32_t kUmma = ptx::kFmtF16; };
template <> struct Fmt<__nv_bfloat16> { static constexpr uint32_t kUmma = ptx::kFmtBF16; };

// rows n0 .. n0 + 128 of (b, h) -> swizzled [128][128 B] tile, asynchronously (cp.async; rows past the sequence are zero-filled)
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

// |x_row|^2 of the thread's own row of a swizzled tile
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

// features [32 hf, 32 hf + 32) of a row -> chunks 4 hf .. 4 hf + 3 of the swizzled 16-bit tile
template <typename T>
__device__ __forceinline__ void store_half16(uint8_t* tile, int row, int hf, const float (&f)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(tile + row * 128 + (((4 * hf + c) ^ (row & 7)) << 4)) =
        make_uint4(Pair16<T>::pk(f[8 * c], f[8 * c + 1]), Pair16<T>::pk(f[8 * c + 2], f[8 * c + 3]),
                   Pair16<T>::pk(f[8 * c + 4], f[8 * c + 5]), Pair16<T>::pk(f[8 * c + 6], f[8 * c + 7]));
}

// kLogF: the key statistics of ScatterBrain instead (scatterbrain_attention.py:10-44, 107-121): features exp(log phi(k)_c - max_n log
// phi(k_n)_c) with a PER-FEATURE maximum (pass 1 computes W' K^T, features on the TMEM lanes, so that the maximum over the tokens is
// thread-local), no query phase; KV / ksum / the maxima go to global memory (float32) for sb_window_tc_kernel.
template <typename T, bool kLogF>
__global__ void __launch_bounds__(kThreads, 2) rfa_favorp_tc_kernel(const View q, const View k, const View v, T* __restrict__ out, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* const sm = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qr = warp & 3, hf = warp >> 2, r = 32 * qr + lane;     // my TMEM lane quarter, column half, row of the tile
  float* const pm = reinterpret_cast<float*>(sm + kPm);          // [2][128] partial row maxima of the two column halves
  float* const pd = reinterpret_cast<float*>(sm + kPd);          // [2][128] partial denominators
  float* const mx2 = reinterpret_cast<float*>(sm + kMx2);        // kLogF: [2][64] partial per-feature maxima
  const uint32_t bar = ptx::smem_u32(sm + kBar);
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + kTmemPtr);
  float* const ksum = reinterpret_cast<float*>(sm + kKsum);
  float* const red = reinterpret_cast<float*>(sm + kRed);
  float* const hs_s = reinterpret_cast<float*>(sm + kHs);       // kLogF: [128] |k|^2 term of the tile's tokens (-inf: dead token)
  float* const mxs = reinterpret_cast<float*>(sm + kMxs);       // kLogF: [64] per-feature maxima
  constexpr uint32_t id_ddt = ptx::umma_idesc(Fmt<T>::kUmma, Fmt<T>::kUmma, 0, 0, 64, 128);   // W' [64 x 64] . K^T -> [features x tokens]
  const float hlm = 0.5f * 4.1588830833596715f;                 // log(64) / 2
  constexpr uint32_t fmt = Fmt<T>::kUmma;
  constexpr uint32_t id_dd = ptx::umma_idesc(fmt, fmt, 0, 0, 128, 64);     // X [128 x 64] . W'^T
  constexpr uint32_t id_kv = ptx::umma_idesc(fmt, fmt, 1, 1, 64, 64);      // F^T [64 x 128] . V [128 x 64]
  constexpr uint32_t id_ks = ptx::umma_idesc(fmt, fmt, 1, 1, 64, 16);      // F^T . ones
  constexpr uint32_t id_o = ptx::umma_idesc(fmt, fmt, 0, 1, 128, 64);      // F [128 x 64] . KV [64 x 64]
  {
    const uint32_t one2 = Pair16<T>::pk(1.0f, 1.0f);
    for (int i = tid; i < 2048 / 16; i += kThreads) reinterpret_cast<uint4*>(sm + kOnes)[i] = make_uint4(one2, one2, one2, one2);   // 16 rows: every K step reads the same ones
  }
  if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr)), 256);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t trow = tmem + ((uint32_t)(32 * qr) << 16);
  const uint64_t dX = ptx::umma_desc_sw128(ptx::smem_u32(sm + kX)), dV = ptx::umma_desc_sw128(ptx::smem_u32(sm + kV));
  const uint64_t dF = ptx::umma_desc_sw128(ptx::smem_u32(sm + kF)), dOnes = ptx::umma_desc_sw128(ptx::smem_u32(sm + kOnes));
  const uint64_t dW = ptx::umma_desc_sw128(ptx::smem_u32(sm + kW)), dKV = ptx::umma_desc_sw128(ptx::smem_u32(sm + kKVt));
  const float dn = 0.35355339059327373f;          // 64^-1/4
  const float half_dn2 = 0.5f * dn * dn, ratio = 0.125f;   // m^-1/2, m = 64
  uint32_t phase = 0;
  auto mma_wait = [&]() { ptx::mbar_wait(bar, phase & 1); ++phase; ptx::tc_fence_after(); };
  // smem tiles written by this thread -> visible to the tensor core; TMEM reads of this thread retired; then everyone
  auto hand_over = [&]() { ptx::fence_proxy_async_smem(); ptx::tc_fence_before(); __syncthreads(); };
  const int tiles = (p.N + kTile - 1) / kTile;
  // The tiles of an item form a list of stages: [K tiles (pass 1)] [K + V tiles (pass 2)] [Q tiles].  Stage g of this CTA lives in
  // buffer g & 1; its successor (possibly the first stage of the CTA's next item) is requested with cp.async before it is consumed.
  const int stages_per_item = 2 * tiles;           // kLogF: [K (maxima)] [K + V]; else: [K + V (online stabiliser)] [Q]
  uint32_t gs = 0;                                  // stages consumed so far
  long long tk[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) tk[i] = 0;
  const bool tr = p.trace && blockIdx.x == 0 && tid == 0;
#define RFA_MARK(i, cond) if (tr && (cond)) tk[i] = clock64();
  int buf = 0;
  auto stage_load = [&](int item_, int s_, int buf_) {
    const int b_ = item_ / p.H, h_ = item_ % p.H, pass = s_ / tiles, t_ = s_ - pass * tiles;
    const bool is_q = !kLogF && pass == 1, with_v = kLogF ? pass == 1 : pass == 0;
    load_tile_async<T>(is_q ? q : k, b_, h_, t_ * kTile, p.N, sm + kX + buf_ * 16384);
    if (with_v) load_tile_async<T>(v, b_, h_, t_ * kTile, p.N, sm + kV + buf_ * 16384);
  };
  // make stage (item_, s_) current: request its successor, wait for its own tiles
  auto acquire = [&](int item_, int s_) {
    buf = (int)(gs & 1);
    int ni = item_, ns = s_ + 1;
    if (ns == stages_per_item) { ni = item_ + (int)gridDim.x; ns = 0; }
    if (ni < p.items) stage_load(ni, ns, buf ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    ++gs;
  };
  auto issue_dd = [&]() {
    if (warp == 0 && ptx::elect_one()) {
      ptx::tc_fence_after();
      const uint64_t dXb = dX + (uint64_t)(buf * 1024);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cDD, dXb + 2 * ks, dW + 2 * ks, id_dd, ks > 0);
      ptx::umma_commit(bar);
    }
  };
  if ((int)blockIdx.x < p.items) stage_load(blockIdx.x, 0, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  int h_loaded = -1;
  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    const int b = item / p.H, h = item % p.H;
    RFA_MARK(10, item == (int)blockIdx.x)
    if (h != h_loaded) {                            // W' = d^-1/4 W of this head, 16-bit, [feature][d] rows of 128 bytes
      const float* W = p.proj + (long long)h * 4096;
      for (int idx = tid; idx < 512; idx += kThreads) {
        const int row = idx >> 3, ch = idx & 7;
        const float4 a = __ldg(reinterpret_cast<const float4*>(W + row * 64 + 8 * ch)), c = __ldg(reinterpret_cast<const float4*>(W + row * 64 + 8 * ch) + 1);
        *reinterpret_cast<uint4*>(sm + kW + row * 128 + ((ch ^ (row & 7)) << 4)) =
            make_uint4(Pair16<T>::pk(dn * a.x, dn * a.y), Pair16<T>::pk(dn * a.z, dn * a.w), Pair16<T>::pk(dn * c.x, dn * c.y), Pair16<T>::pk(dn * c.z, dn * c.w));
      }
      h_loaded = h;
    }
    if constexpr (kLogF) {
      // ---- pass 1: per-feature maximum over the live tokens of log phi(k)_c = DD - |k|^2 term - log(m) / 2 ----
      float mxc = kNegInf;
      for (int t = 0; t < tiles; ++t) {
        acquire(item, t);
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncthreads();
        if (tid < 128) {
          const int n = t * kTile + tid;
          const bool dead = n >= p.N || (p.mask && p.mask[(long long)b * p.N + n]);
          hs_s[tid] = dead ? __int_as_float(0x7f800000) : half_dn2 * row_sq<T>(sm + kX + buf * 16384, tid);
        }
        if (warp == 0 && ptx::elect_one()) {
          ptx::tc_fence_after();
          const uint64_t dXb = dX + (uint64_t)(buf * 1024);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cKV, dW + 2 * ks, dXb + 2 * ks, id_ddt, ks > 0);
          ptx::umma_commit(bar);
        }
        __syncthreads();                            // hs_s complete
        mma_wait();
#pragma unroll 1
        for (int g8 = 0; g8 < 4; ++g8) {            // lanes < 16 of each warp hold a feature row (warp-collective loads); my half of
          float dd[16];                             // the 128 token columns
          ptx::tmem_ld16(trow + cKV + 64 * hf + 16 * g8, reinterpret_cast<uint32_t*>(dd));
          ptx::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) mxc = fmaxf(mxc, dd[e] - hs_s[64 * hf + 16 * g8 + e]);
        }
        ptx::tc_fence_before();
        __syncthreads();                            // hs_s / the accumulator columns are rewritten by the next tile
      }
      if (lane < 16) mx2[64 * hf + 16 * qr + lane] = mxc;
      __syncthreads();
      if (tid < 64) mxs[tid] = fmaxf(mx2[tid], mx2[64 + tid]) - hlm;
      __syncthreads();
    }
    // favorp: NO separate stabiliser pass.  The key stabiliser (max over every token and feature of DD, padded keys included,
    // reference :48-51) is tracked as a running maximum s_run: features are exp(DD - |k|^2 term - s_run) WITHOUT the m^-1/2 factor and
    // the 1e-4 (which must not be rescaled); when a tile raises the maximum the accumulators in tensor memory are multiplied by
    // exp(s_old - s_new); at the end KV = m^-1/2 KVe + 1e-4 sum_n v_n (a third MMA, ones^T V), ksum = m^-1/2 KSe + 1e-4 n_live.
    float s_run = kNegInf, n_live = (float)p.N;
    if constexpr (!kLogF) {
      if (p.mask) {
        float c = 0.f;
        for (int i = tid; i < p.N; i += kThreads) c += p.mask[(long long)b * p.N + i] ? 1.f : 0.f;
        c = warp_sum(c);
        if (lane == 0) red[8 + warp] = c;
        __syncthreads();
        n_live = (float)p.N - (red[8] + red[9] + red[10] + red[11] + red[12] + red[13] + red[14] + red[15]);
      }
    }
    // ---- pass 2: KV = phi(K)^T V, KS = phi(K)^T 1 ----
    for (int t = 0; t < tiles; ++t) {
      RFA_MARK(0, t == 2 && item == (int)blockIdx.x)
      if (t > 0) mma_wait();                        // the previous tile's KV / KS MMAs have read F and its V buffer (the one the
      acquire(item, kLogF ? tiles + t : t);         // successor stage is loaded into)
      if constexpr (!kLogF) {
        if (p.mask) {                               // padded keys must not enter sum_n v_n: their v rows become zero
          __syncthreads();                          // (a row's 16-byte pieces were copied by OTHER threads: their cp.async must have landed)
          const int n = t * kTile + r;
          if (n < p.N && p.mask[(long long)b * p.N + n]) {
#pragma unroll
            for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(sm + kV + buf * 16384 + r * 128 + (((4 * hf + c) ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
          }
        }
      }
      RFA_MARK(1, t == 2 && item == (int)blockIdx.x)
      hand_over();
      issue_dd();
      mma_wait();
      RFA_MARK(2, t == 2 && item == (int)blockIdx.x)
      {
        const int n = t * kTile + r;
        const bool dead = n >= p.N || (p.mask && p.mask[(long long)b * p.N + n]);
        float f[32];
        tmem_ld_cols<32>(trow + cDD + 32 * hf, reinterpret_cast<uint32_t*>(f));
        ptx::tmem_ld_wait();
        if constexpr (!kLogF) {                     // running stabiliser: this tile's maximum (every row of the sequence), rescale if it grew
          float tmx = kNegInf;
#pragma unroll
          for (int j = 0; j < 32; ++j) tmx = fmaxf(tmx, f[j]);
          tmx = warp_max(n < p.N ? tmx : kNegInf);
          if (lane == 0) red[warp] = tmx;
          __syncthreads();
          const float s_new = fmaxf(s_run, fmaxf(fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])), fmaxf(fmaxf(red[4], red[5]), fmaxf(red[6], red[7]))));
          if (t > 0 && s_new > s_run) {             // (block-uniform) accumulators of the earlier tiles: x exp(s_old - s_new)
            const float c = __expf(s_run - s_new);
            float kv[32];
            uint32_t ks4[4];
            tmem_ld_cols<32>(trow + cKV + 32 * hf, reinterpret_cast<uint32_t*>(kv));
            tmem_ld_cols<4>(trow + cKS, ks4);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) kv[j] *= c;
#pragma unroll
            for (int j = 0; j < 4; ++j) ks4[j] = __float_as_uint(__uint_as_float(ks4[j]) * c);
            fused::tmem_st_cols<32>(trow + cKV + 32 * hf, reinterpret_cast<const uint32_t*>(kv));
            if (hf == 0) fused::tmem_st_cols<4>(trow + cKS, ks4);
            ptx::tmem_st_wait();
          }
          s_run = s_new;
        }
        const float sub = half_dn2 * row_sq<T>(sm + kX + buf * 16384, r) + (kLogF ? hlm : s_run);
        const float live = dead ? 0.f : 1.f;
        const float sub2 = dead ? kNegInf : -sub * kLog2e;
#pragma unroll
        for (int j = 0; j < 32; ++j) {              // branch-free (a select around every exponential compiles to a divergence region)
          if constexpr (kLogF) f[j] = live * __expf(fminf(f[j] - sub - mxs[32 * hf + j], 0.f));   // live rows: <= 0 by the definition of mxs
          else f[j] = ex2(fmaf(f[j], kLog2e, sub2));        // one FMA + one MUFU per feature; <= 0: s_run covers this tile; dead rows: -inf
        }
        store_half16<T>(sm + kF, r, hf, f);
      }
      RFA_MARK(3, t == 2 && item == (int)blockIdx.x)
      hand_over();
      RFA_MARK(4, t == 2 && item == (int)blockIdx.x)
      if (warp == 0 && ptx::elect_one()) {
        ptx::tc_fence_after();
        const uint64_t dVb = dV + (uint64_t)(buf * 1024);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) ptx::umma_ss(tmem + cKV, dF + 128 * ks, dVb + 128 * ks, id_kv, (t > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) ptx::umma_ss(tmem + cKS, dF + 128 * ks, dOnes, id_ks, (t > 0 || ks > 0) ? 1u : 0u);
        if constexpr (!kLogF) {                     // SV = ones^T V: every accumulator row holds sum_n v_n
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) ptx::umma_ss(tmem + cSV, dOnes, dVb + 128 * ks, id_kv, (t > 0 || ks > 0) ? 1u : 0u);
        }
        ptx::umma_commit(bar);
      }
    }
    mma_wait();
    {                                               // M = 64 accumulators: feature 16 qr + l sits on lane l < 16 of quarter qr;
      const int j = 16 * qr + (lane & 15);          // my half of its 64 columns (tcgen05.ld is warp-collective: every lane loads)
      float kv[32];
      uint32_t ks0;
      tmem_ld_cols<32>(trow + cKV + 32 * hf, reinterpret_cast<uint32_t*>(kv));
      ptx::tmem_ld1(trow + cKS, ks0);
      ptx::tmem_ld_wait();
      if constexpr (kLogF) {
        if (lane < 16) {
          float* dst = p.part + (long long)item * (64 * 64 + 64);
#pragma unroll
          for (int d4 = 0; d4 < 8; ++d4) reinterpret_cast<float4*>(dst + j * 64 + 32 * hf)[d4] = make_float4(kv[4 * d4], kv[4 * d4 + 1], kv[4 * d4 + 2], kv[4 * d4 + 3]);
          if (hf == 0) { dst[64 * 64 + j] = __uint_as_float(ks0); p.stabv[(long long)item * 64 + j] = mxs[j]; }
        }
      } else {
        float sv[32];
        tmem_ld_cols<32>(trow + cSV + 32 * hf, reinterpret_cast<uint32_t*>(sv));
        ptx::tmem_ld_wait();
        if (lane < 16) {
#pragma unroll
          for (int d = 0; d < 32; ++d) kv[d] = fmaf(ratio, kv[d], 1e-4f * sv[d]);
          store_half16<T>(sm + kKVt, j, hf, kv);
          if (hf == 0) ksum[j] = fmaf(ratio, __uint_as_float(ks0), 1e-4f * n_live);
        }
      }
    }
    if constexpr (kLogF) {
      ptx::tc_fence_before();
      __syncthreads();
      continue;
    }
    // ---- phase Q ----
    for (int t = 0; t < tiles; ++t) {
      RFA_MARK(5, t == 2 && item == (int)blockIdx.x)
      acquire(item, tiles + t);
      hand_over();
      issue_dd();
      mma_wait();
      RFA_MARK(6, t == 2 && item == (int)blockIdx.x)
      {
        float f[32];
        tmem_ld_cols<32>(trow + cDD + 32 * hf, reinterpret_cast<uint32_t*>(f));
        ptx::tmem_ld_wait();
        float rmx = kNegInf;
#pragma unroll
        for (int j = 0; j < 32; ++j) rmx = fmaxf(rmx, f[j]);
        pm[128 * hf + r] = rmx;
        __syncthreads();                            // the two column halves of a row meet
        rmx = fmaxf(pm[r], pm[128 + r]);
        const float sub2 = -(half_dn2 * row_sq<T>(sm + kX + buf * 16384, r) + rmx) * kLog2e;
        float den = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j] = fmaf(ratio, ex2(fmaf(f[j], kLog2e, sub2)), 1e-4f);
          den = fmaf(f[j], ksum[32 * hf + j], den);
        }
        pd[128 * hf + r] = den;
        store_half16<T>(sm + kF, r, hf, f);
      }
      RFA_MARK(7, t == 2 && item == (int)blockIdx.x)
      hand_over();
      if (warp == 0 && ptx::elect_one()) {
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cO, dF + 2 * ks, dKV + 128 * ks, id_o, ks > 0);
        ptx::umma_commit(bar);
      }
      mma_wait();
      RFA_MARK(8, t == 2 && item == (int)blockIdx.x)
      {
        float o[32];
        tmem_ld_cols<32>(trow + cO + 32 * hf, reinterpret_cast<uint32_t*>(o));
        ptx::tmem_ld_wait();
        const int n = t * kTile + r;
        if (n < p.N) {
          const float inv = 1.0f / fmaxf(pd[r] + pd[128 + r], 1e-2f);
          uint4* dst = reinterpret_cast<uint4*>(out + ((long long)b * p.N + n) * ((long long)p.H * 64) + (long long)h * 64) + 4 * hf;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            dst[ch] = make_uint4(Pair16<T>::pk(o[8 * ch] * inv, o[8 * ch + 1] * inv), Pair16<T>::pk(o[8 * ch + 2] * inv, o[8 * ch + 3] * inv),
                                 Pair16<T>::pk(o[8 * ch + 4] * inv, o[8 * ch + 5] * inv), Pair16<T>::pk(o[8 * ch + 6] * inv, o[8 * ch + 7] * inv));
        }
      }
    }
    RFA_MARK(9, item == (int)blockIdx.x)
    if (tr && item == (int)blockIdx.x)
      printf("rfa trace (cycles): key stage: wait KV MMA + acquire %lld | sync + DD MMA %lld | epilogue %lld | sync %lld ; query stage: acquire + sync + DD MMA %lld | epilogue %lld | sync + O MMA %lld ; item total %lld\n",
             tk[1] - tk[0], tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3], tk[6] - tk[5], tk[7] - tk[6], tk[8] - tk[7], tk[9] - tk[10]);
    ptx::tc_fence_before();
    __syncthreads();                                // ksum / KV tile / W tile are rewritten by the next item
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

static int g_launches = 0;

}  // namespace rfatc

extern "C" int eva_debug_rfa_tc_launches(void) { return rfatc::g_launches; }

bool rfa_tc_supported(int method, int D, int m, int cosw, int io_dtype, const View& q, const View& k, const View& v) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("EVA_SM100_DISABLE_FUSED"); off = (e && e[0] == '1') ? 1 : 0; }
  if (off) return false;
  if (method != RFA_FAVORP || D != 64 || m != 64 || cosw) return false;
  if (io_dtype != EVA_F16 && io_dtype != EVA_BF16) return false;
  for (const View* x : {&q, &k, &v})
    if (x->sh * 2 % 16 || x->sn * 2 % 16 || x->sb * 2 % 16 || reinterpret_cast<uintptr_t>(x->ptr) % 16) return false;
  return true;
}

cudaError_t launch_rfa_tc(int B, int H, int N, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask,
                          const float* proj, void* out, cudaStream_t st, float* sb_stabv, float* sb_part) {
  static const int trace = [] { const char* e = getenv("EVA_SM100_TRACE"); return (e && e[0] == '1') ? 1 : 0; }();
  rfatc::Params p{B, H, N, B * H, proj, mask, sb_stabv, sb_part, trace};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int dyn = rfatc::kSmemBytes + 1024;
  const int grid = p.items < 2 * sms ? p.items : 2 * sms;
  ++rfatc::g_launches;
  auto go = [&](auto kern, auto* o) -> cudaError_t {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return e;
    kern<<<grid, rfatc::kThreads, dyn, st>>>(q, k, v, o, p);
    return cudaGetLastError();
  };
  if (io_dtype == EVA_F16) {
    __half* o = reinterpret_cast<__half*>(out);
    return sb_part ? go(rfatc::rfa_favorp_tc_kernel<__half, true>, o) : go(rfatc::rfa_favorp_tc_kernel<__half, false>, o);
  }
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  return sb_part ? go(rfatc::rfa_favorp_tc_kernel<__nv_bfloat16, true>, o) : go(rfatc::rfa_favorp_tc_kernel<__nv_bfloat16, false>, o);
}

}  // namespace eva
