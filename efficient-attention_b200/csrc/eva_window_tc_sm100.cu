// Window attention on tcgen05 for EVERY geometry with head_dim 64 and 16-bit I/O (eva.py:200-227, causal_eva.py:722-783,
// local_attention.py:134-182, abstract_attention.py:115-133): 1-D / 2-D, halos, padding masks, causal, chunk keys or none, any
// window length.  The fused kernels (eva_fused_sm100.cu, eva_causal_sm100.cu) keep the geometries they are specialised for; this
// kernel takes over what used to fall to the CUDA-core window_attn_kernel of eva_generic.cu.  Same arguments, same semantics.
//
// CTA iteration = (batch x head, window, block of 128 query rows); keys = [local window slots | chunk keys] in tiles of 128.
//   rows -> Q tile (gathered with the window index arithmetic, 128-byte swizzle); per key tile: K / V rows (or k_bar / beta rows,
//   float32 -> the I/O format) -> tiles, flags (live / masked / absent) -> shared memory
//   per tile: S = Q K^T, finished logits, online softmax (running row maximum; the O row in tensor memory and the row sum are
//   rescaled when it moves), P = exp2(logit - max) as 16-bit pairs -> tensor memory, O += P V (A operand from tensor memory,
//   V MN-major); epilogue: O / rowsum -> out
// TMEM lane = query row (M = 128); warps w and w + 4 share a lane quarter and split the 128 columns of a tile; the two partial
// maxima / sums of a row meet in shared memory once per pass.  Bias, padding, causal and chunk-visibility rules are applied by
// finish_logit (common.cuh), the same function the CUDA-core kernels use.
#include <stdlib.h>

#include "common.cuh"
#include "fused_common.cuh"
#include "launch.h"
#include "sm100_ptx.cuh"

namespace eva {
namespace wintc {

using fused::IoFmt;
using fused::tile_off;
using fused::tmem_ld_cols;
using fused::ex2;

constexpr int kThreads = 256;
constexpr int kQ = 0, kK = 16384, kV = 32768, kBuf = 32768, kMisc = 81920;   // K / V tiles of buffer 1 at + kBuf
constexpr int kFlag = kMisc, kFac = kFlag + 1024, kQtok = kFac + 4096, kQpad = kQtok + 512, kPm = kQpad + 512, kPl = kPm + 1024, kBiasT = kPl + 1024, kBar = kBiasT + 8 * 32 * 17 * 4,
              kSlot = kBar + 16, kSmemBytes = kSlot + 16 + 1024;
constexpr uint32_t kTmemCols = 256, cS = 0, cP = 128, cO = 192;

struct Params {
  Geo g;
  View q, k, v;
  const uint8_t* mask;
  const float* kbar; const float* beta; const float* bias;
  long long bias_sh;
  void* out;
  float* lse_out;                // training: log2-domain log-sum-exp per query row -> float32 [B, H, N], or NULL
  const float* key_bias;         // float32 [B, H, N] added to the (scaled) logit of every live local key, or NULL (randomized attention)
  long long total;
  int trace;
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
eva_window_tc_kernel(const Params p) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (ptx::smem_u32(raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Geo& g = p.g;
  int* kflag_all = reinterpret_cast<int*>(sm + kFlag); // [2][128] per buffer: 0 live, 1 masked, 2 absent
  // [2][128] per key: logit = s * mul + add with (1, 0) live | (0, mask_fill) masked | (0, -inf) absent, and the causal rule as two
  // thresholds: overwritten by -5e4 when row index < thr_row (local keys: key slot - halo) or row chunk < thr_chunk (chunk keys: c + 1)
  float4* kfac_all = reinterpret_cast<float4*>(sm + kFac);
  int* qtok = reinterpret_cast<int*>(sm + kQtok);      // [128] token of the row, -1: no such row
  int* qpad = reinterpret_cast<int*>(sm + kQpad);
  float* pm = reinterpret_cast<float*>(sm + kPm);      // [2][128]
  float* pl = reinterpret_cast<float*>(sm + kPl);      // [2][128]
  float* biasT = reinterpret_cast<float*>(sm + kBiasT);   // [8 warps][32][17] bias transposition
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
  constexpr uint32_t id_pv = ptx::umma_idesc(fmt, fmt, 0, 1, 128, 64);
  const uint64_t dQ = ptx::umma_desc_sw128(ptx::smem_u32(sm + kQ)), dK0 = ptx::umma_desc_sw128(ptx::smem_u32(sm + kK)),
                 dV0 = ptx::umma_desc_sw128(ptx::smem_u32(sm + kV));
  const int qr = warp & 3, hf = warp >> 2;
  const int r = 32 * qr + lane;                        // query row of the block = TMEM lane
  const uint32_t trow = tmem + ((uint32_t)(32 * qr) << 16);
  const float scale = 0.125f;
  const int n_rb = (g.L + 127) / 128;
  const int n_keys = g.J + g.n_chunks;
  const int n_tiles = (n_keys + 127) / 128;
  const long long HD = (long long)g.H * 64;
  uint32_t ph = 0;                                      // completed phases of `bar`

  for (long long item = blockIdx.x; item < p.total; item += gridDim.x) {
    const int rb = (int)(item % n_rb), win = (int)((item / n_rb) % g.n_windows);
    const int bh = (int)(item / ((long long)n_rb * g.n_windows));
    const int b = bh / g.H, h = bh % g.H;
    const long long bias_off = (long long)h * p.bias_sh;
    __syncthreads();                                   // the previous iteration is done with every tile and table
    if (tid < 128) {
      const int li = rb * 128 + tid;
      const int tok = li < g.L ? group_token(g, win, li, g.window, 0) : -1;
      qtok[tid] = tok;
      qpad[tid] = (tok >= 0 && p.mask) ? (int)p.mask[(long long)b * g.N + tok] : 0;
    }
    // The eight lanes that share a row compute its token ONCE: lane piece p < 4 of a group takes the group's row of iteration p
    // (the window arithmetic has integer divisions; it used to run once per 16-byte piece), the others get it by shuffle.
    int qtok_mine = -1;
    if ((tid & 7) < 4) {
      const int li = rb * 128 + (tid >> 3) + 32 * (tid & 7);
      qtok_mine = li < g.L ? group_token(g, win, li, g.window, 0) : -1;
    }
    for (int idx = tid, it = 0; idx < 128 * 8; idx += kThreads, ++it) {
      const int row = idx >> 3, piece = idx & 7;
      const int tok = __shfl_sync(0xffffffffu, qtok_mine, (lane & 24) + it);
      uint4 z = make_uint4(0, 0, 0, 0);
      if (tok >= 0) z = __ldg(reinterpret_cast<const uint4*>(p.q.row<T>(b, tok, h)) + piece);
      *reinterpret_cast<uint4*>(sm + kQ + tile_off(row, 8 * piece)) = z;
    }
    const int last_visible = rb * 128 + 127 + g.ext;   // causal: local key tiles entirely above the diagonal are skipped
    auto skip_tile = [&](int kt0) { return g.causal && kt0 > last_visible && kt0 + 128 <= g.J; };
    // K / V rows of one key tile -> buffer `buf`, asynchronously (cp.async, zero-filled where there is no row); the chunk rows are
    // converted from float32 on the way.  Returns the OR of the flags this thread wrote.
    auto load_tile = [&](int kt0, int buf) -> int {
      int flags = 0;
      int* kflag = kflag_all + 128 * buf;
      int ktok_mine = -1;                            // token of the group's row of iteration p, computed by lane piece p < 4
      if ((tid & 7) < 4) {
        const int gj = kt0 + (tid >> 3) + 32 * (tid & 7);
        if (gj < g.J) ktok_mine = group_token(g, win, gj, g.window, g.ext);
      }
      for (int idx = tid, it = 0; idx < 128 * 8; idx += kThreads, ++it) {
        const int j = idx >> 3, piece = idx & 7;
        const int gj = kt0 + j;
        const uint32_t dk = ptx::smem_u32(sm + kK + buf * kBuf + tile_off(j, 8 * piece)), dv = dk + (kV - kK);
        const int tok_row = __shfl_sync(0xffffffffu, ktok_mine, (lane & 24) + it);
        int flag = 0;
        if (gj < g.J) {
          const int tok = tok_row;
          const int sz = tok >= 0 ? 16 : 0;
          const int tk_ = tok >= 0 ? tok : 0;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dk), "l"(reinterpret_cast<const uint4*>(p.k.row<T>(b, tk_, h)) + piece), "r"(sz) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dv), "l"(reinterpret_cast<const uint4*>(p.v.row<T>(b, tk_, h)) + piece), "r"(sz) : "memory");
          flag = tok >= 0 ? ((p.mask && p.mask[(long long)b * g.N + tok]) ? 1 : 0) : 1;
        } else if (gj < n_keys) {
          const long long base = ((long long)bh * g.n_chunks + (gj - g.J)) * 64 + 8 * piece;
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.kbar + base)), a1 = __ldg(reinterpret_cast<const float4*>(p.kbar + base) + 1);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + base)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + base) + 1);
          *reinterpret_cast<uint4*>(sm + kK + buf * kBuf + tile_off(j, 8 * piece)) =
              make_uint4(IoFmt<T>::pack2(a0.x, a0.y), IoFmt<T>::pack2(a0.z, a0.w), IoFmt<T>::pack2(a1.x, a1.y), IoFmt<T>::pack2(a1.z, a1.w));
          *reinterpret_cast<uint4*>(sm + kV + buf * kBuf + tile_off(j, 8 * piece)) =
              make_uint4(IoFmt<T>::pack2(b0.x, b0.y), IoFmt<T>::pack2(b0.z, b0.w), IoFmt<T>::pack2(b1.x, b1.y), IoFmt<T>::pack2(b1.z, b1.w));
        } else {
          const uint4 z = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(sm + kK + buf * kBuf + tile_off(j, 8 * piece)) = z;
          *reinterpret_cast<uint4*>(sm + kV + buf * kBuf + tile_off(j, 8 * piece)) = z;
          flag = 2;
        }
        if (piece == 0) {
          kflag[j] = flag;
          const int never = -2147483647 - 1;
          const int thr_row = (g.causal && gj < g.J) ? gj - g.ext : never;
          const int thr_chunk = (g.causal && gj >= g.J && gj < n_keys) ? gj - g.J + 1 : never;
          float add = flag == 0 ? 0.f : (flag == 1 ? g.mask_fill : kNegInf);
          if (p.key_bias && flag == 0 && gj < g.J) {     // (flag 0 and gj < J: the slot holds a token)
            add = __ldg(p.key_bias + (long long)bh * g.N + tok_row);
            flags |= 4;                                  // the tile needs the per-key table
          }
          kfac_all[128 * buf + j] = make_float4(flag == 0 ? 1.f : 0.f, add, __int_as_float(thr_row), __int_as_float(thr_chunk));
        }
        flags |= flag;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      return flags;
    };
    // S = Q K^T of the staged tiles; every thread returns once it is in TMEM.  Returns whether ANY key of the tile needs more than
    // the scale (a mask flag; the block-wide OR rides on the barrier the MMA issue needs anyway)
    auto mma_s = [&](int my_flags, int buf) -> bool {
      const uint64_t dK = dK0 + (uint64_t)((buf * kBuf) >> 4);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      const int any = __syncthreads_or(my_flags);
      if (tid == 0) {
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cS, dQ + 2 * ks, dK + 2 * ks, id_s, ks > 0);
        ptx::umma_commit(bar);
      }
      ptx::mbar_wait(bar, ph & 1);
      ++ph;
      ptx::tc_fence_after();
      return any != 0;
    };
    __syncthreads();                                   // qtok / qpad are visible
    // finished logits of my 64 columns of the tile, log2 domain
    const int tq_row = qtok[r], li_row = rb * 128 + r, qp_row = qpad[r];
    const int tq_chunk = (g.causal && g.chunk > 0 && tq_row >= 0) ? tq_row / g.chunk : 0;
    const float* brow = (p.bias && tq_row >= 0) ? p.bias + bias_off + (long long)li_row * g.J : nullptr;
    const bool rules = p.bias != nullptr || g.causal || g.mask_queries;       // anything beyond per-key flags
    auto logits = [&](int kt0, bool tile_flags, int buf, float (&x)[64]) {
      const int* kflag = kflag_all + 128 * buf;
      tmem_ld_cols<64>(trow + cS + 64 * hf, reinterpret_cast<uint32_t*>(x));
      ptx::tmem_ld_wait();
      if (!rules && !tile_flags) {                     // the common tile: every key live, no bias, no causal rule
#pragma unroll
        for (int j = 0; j < 64; ++j) x[j] *= scale * kLog2e;
        return;
      }
      const int c0 = kt0 + 64 * hf;
      const float4* kf = kfac_all + 128 * buf + 64 * hf;
      const bool row_masked = g.mask_queries && qp_row;
      if (!rules) {                                    // flags only (warp-uniform choice): one broadcast read and two FMAs per column
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const float4 f = kf[j];
          x[j] = fmaf(x[j] * scale, f.x, f.y) * kLog2e;
        }
        return;
      }
      const bool any_row_masked = __any_sync(0xffffffffu, row_masked);
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
          if (any_row_masked) sv = (row_masked && (c0 + j < g.J)) ? g.mask_fill : sv;       // padded query rows of the causal layer
          x[j] = (cm ? kMaskVal : sv) * kLog2e;
        }
      }
    };
    float mrow = kNegInf;                              // running row maximum (both column halves), log2 domain
    float lrow = 0.f;
    bool first = true;
    long long tk[6];
    bool first_print = true;
    const bool tr_on = p.trace && item == blockIdx.x + gridDim.x;
    auto next_live = [&](int kt) { while (kt < n_keys && skip_tile(kt)) kt += 128; return kt; };
    int kt0 = next_live(0), buf = 0;
    int flags_cur = load_tile(kt0, 0);                 // (every item has at least one live tile: the diagonal / the chunk keys)
    for (; kt0 < n_keys;) {
      const int kt_next = next_live(kt0 + 128);
      if (tr_on) tk[0] = clock64();
      int flags_next = 0;
      if (kt_next < n_keys) {                          // the next tile travels while this one is computed
        flags_next = load_tile(kt_next, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      if (tr_on) tk[1] = clock64();
      const bool tf = mma_s(flags_cur, buf);
      if (tr_on) tk[2] = clock64();
      float x[64];
      logits(kt0, tf, buf, x);
      if (tr_on) tk[3] = clock64();
      // online softmax: new running maximum, the accumulated O row and sum are rescaled when it moves
      {
        float mloc = kNegInf;
#pragma unroll
        for (int j = 0; j < 64; ++j) mloc = fmaxf(mloc, x[j]);
        pm[hf * 128 + r] = mloc;
        __syncthreads();
        const float mnew = fmaxf(mrow, fmaxf(pm[r], pm[128 + r]));
        const float alpha = (mnew == kNegInf || mrow == mnew) ? 1.f : ex2(mrow - mnew);      // exp2(-inf) = 0 for the first live tile
        mrow = mnew;
        lrow *= alpha;
        if (!first && __any_sync(0xffffffffu, alpha != 1.f)) {
          float o[32];
          tmem_ld_cols<32>(trow + cO + 32 * hf, reinterpret_cast<uint32_t*>(o));
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] *= alpha;
          ptx::tmem_st16(trow + cO + 32 * hf, reinterpret_cast<const uint32_t*>(o));
          ptx::tmem_st16(trow + cO + 32 * hf + 16, reinterpret_cast<const uint32_t*>(o) + 16);
        }
      }
      const bool row_ok = qtok[r] >= 0 && mrow != kNegInf;
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float a = row_ok ? ex2(x[32 * blk + e] - mrow) : 0.f;
          const float c = row_ok ? ex2(x[32 * blk + e + 1] - mrow) : 0.f;
          lrow += a + c;
          pk[e >> 1] = IoFmt<T>::pack2(a, c);
        }
        ptx::tmem_st16(trow + cP + 32 * hf + 16 * blk, pk);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        ptx::tc_fence_after();
#pragma unroll
        const uint64_t dV = dV0 + (uint64_t)((buf * kBuf) >> 4);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) ptx::umma_ts(tmem + cO, tmem + cP + 8 * ks, dV + 128 * ks, id_pv, (first ? 0u : 1u) | (ks > 0 ? 1u : 0u));
        ptx::umma_commit(bar);
      }
      first = false;
      if (tr_on) tk[4] = clock64();
      ptx::mbar_wait(bar, ph & 1);                     // the tiles and the P columns are free again
      ++ph;
      ptx::tc_fence_after();
      if (tr_on && blockIdx.x == 0 && (tid == 0 || tid == 200) && first_print)
        printf("window tc trace tid %d: issue next + wait %lld | sync+S mma %lld | logits %lld | exp+P+sync %lld | O mma %lld\n", tid, tk[1] - tk[0],
               tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3], clock64() - tk[4]);
      first_print = false;
      kt0 = kt_next; buf ^= 1; flags_cur = flags_next;
    }
    pl[hf * 128 + r] = lrow;
    __syncthreads();
    {
      const float inv = 1.0f / (pl[r] + pl[128 + r]);
      if (p.lse_out && hf == 0 && qtok[r] >= 0) p.lse_out[(long long)bh * g.N + qtok[r]] = mrow + log2f(pl[r] + pl[128 + r]);
      float o[32];
      tmem_ld_cols<32>(trow + cO + 32 * hf, reinterpret_cast<uint32_t*>(o));
      ptx::tmem_ld_wait();
      const int tq = qtok[r];
      if (tq >= 0) {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<T*>(p.out) + ((long long)b * g.N + tq) * HD + (long long)h * 64 + 32 * hf);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
          dst[c4] = make_uint4(IoFmt<T>::pack2(o[8 * c4] * inv, o[8 * c4 + 1] * inv), IoFmt<T>::pack2(o[8 * c4 + 2] * inv, o[8 * c4 + 3] * inv),
                               IoFmt<T>::pack2(o[8 * c4 + 4] * inv, o[8 * c4 + 5] * inv), IoFmt<T>::pack2(o[8 * c4 + 6] * inv, o[8 * c4 + 7] * inv));
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, kTmemCols);
}

}  // namespace wintc

static int g_window_tc_count = 0;
static int g_window_tc_mode = -1;   // -1: environment (EVA_SM100_WINDOW_SIMT=1 disables), 0: off, 1: on

bool window_tc_supported(const Geo& g, int io_dtype) {
  static const bool env_off = [] {
    const char* e = getenv("EVA_SM100_WINDOW_SIMT");
    const char* f = getenv("EVA_SM100_DISABLE_FUSED");
    return (e && e[0] == '1') || (f && f[0] == '1');
  }();
  if (g_window_tc_mode == 0 || (g_window_tc_mode < 0 && env_off)) return false;
  return g.D == 64 && (io_dtype == EVA_F16 || io_dtype == EVA_BF16) && g.J + g.n_chunks >= 1;
}

cudaError_t launch_window_tc(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask,
                             const float* kbar, const float* beta, const float* bias, long long bias_sh, void* out, cudaStream_t st,
                             float* lse_out, const float* key_bias) {
  wintc::Params p;
  p.key_bias = key_bias;
  p.g = g; p.q = q; p.k = k; p.v = v; p.mask = mask;
  p.kbar = kbar; p.beta = beta; p.bias = bias; p.bias_sh = bias_sh; p.out = out; p.lse_out = lse_out;
  p.total = (long long)((g.L + 127) / 128) * g.n_windows * g.B * g.H;
  static const int trace = [] { const char* e = getenv("EVA_SM100_TRACE"); return (e && e[0] == '1') ? 1 : 0; }();
  p.trace = trace;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long grid = p.total < 2LL * sms ? p.total : 2LL * sms;
  ++g_window_tc_count;
  cudaError_t e;
  if (io_dtype == EVA_F16) {
    e = cudaFuncSetAttribute(wintc::eva_window_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, wintc::kSmemBytes);
    if (e != cudaSuccess) return e;
    wintc::eva_window_tc_kernel<__half><<<(unsigned)grid, wintc::kThreads, wintc::kSmemBytes, st>>>(p);
  } else {
    e = cudaFuncSetAttribute(wintc::eva_window_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, wintc::kSmemBytes);
    if (e != cudaSuccess) return e;
    wintc::eva_window_tc_kernel<__nv_bfloat16><<<(unsigned)grid, wintc::kThreads, wintc::kSmemBytes, st>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace eva

// diagnostics (not part of the public ABI)
extern "C" int eva_debug_window_tc_count(void) { return eva::g_window_tc_count; }
extern "C" void eva_debug_set_window_tc(int mode) { eva::g_window_tc_mode = mode; }
