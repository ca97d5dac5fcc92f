// Fused EVA forward for sm_100a (v2): one persistent kernel; per (batch, head) work item everything runs
// off TMA tiles and tcgen05 MMAs, the CUDA cores only do LayerNorm, |k|^2, two small softmaxes and packing.
//
//   pass 1   chunk means:      [feat x chunk]   = Q_r^T . Pool^T, K_r^T . Pool^T       (M=64 MMAs, A MN-major)
//            adaptive Linear:  [chunk x feat]   = [mq ; mk] . [W_q ; W_k]^T            (M=128 MMA) -> LayerNorm
//            -> k_bar tile, omega tile                                                  (eva.py:155-190)
//   pass 2   phi-logits:       [token x chunk]  = K_r . (Qbar_r + Kbar_r)^T            (M=128 MMAs, all rows first)
//            16-token softmax per chunk (SIMT, one batch for the item) -> P tiles;
//            beta^T = V_r^T . P_r^T                                                     (M=64 MMAs)
//            -> beta tile                                                               (eva.py:192-196)
//   phase B  per pair of windows: S = Q [K_w ; k_bar]^T, joint row softmax (TMEM -> RF), P -> TMEM,
//            O = P [V_w ; beta] (A operand from TMEM), normalise, store                 (eva.py:200-227)
//
// q/k/v reach the SM only through TMA: chunk-row boxes (pass 1: q,k from HBM; pass 2: all k rows from L2, then
// all v rows) and window boxes (phase B, L2 hits).  All tiles travel through one ring of four 16-KB
// shared-memory slots with full/free mbarriers; the means tile of the Linear step borrows a ring slot.
// Chunks are indexed c' = 8 r + cx everywhere (chunk-row r, column cx < 7; every 8th row / column is padding),
// so that the k_bar tile serves both as the second B operand of the phi-logits and as the chunk keys of phase B.
//
// CTA = 6 warps: warps 0-3 compute (thread t <-> TMEM lane t), warp 4 = TMA producer, warp 5 = MMA issuer.
// Two CTAs per SM (256 TMEM columns, ~105 KB shared memory each).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>

#include <string.h>

#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "launch.h"
#include "sm100_ptx.cuh"
#include "fused_common.cuh"

namespace eva {
namespace fused {

constexpr int kD = 64;
constexpr int kThreads = 192;
constexpr int kComputeThreads = 128;
constexpr uint32_t kTmemCols = 256;
constexpr int kSlots = 4;
constexpr int kSlotBytes = 16384;

enum Bar {
  kFull0 = 0, kFree0 = kFull0 + kSlots, kPoolFull = kFree0 + kSlots, kAFull, kLinFull, kOmFull,
  kD2Full, kP2Full, kP2FullB, kBetaFull, kStatsFull,
  kSFull, kPFull, kOFull0, kOFull1, kOFree0, kOFree1, kBiasFull, kBiasFree, kItem0, kItem1, kNormDone0, kNumBars = kNormDone0 + 7
};

struct Params {
  int B, H, N, gh, gw;
  int nwx, n_windows, n_pairs;   // windows per grid row, total, pairs per item
  int items;
  const float *b_q, *g_q, *beta_q, *b_k, *g_k, *beta_k;
  int has_q;                     // 0: adaptive_proj == 'none' (mu = 0)
  float mu_coeff, inv_mu_coeff, ln_eps;
  const float* noise;
  unsigned int* next_item;       // global work counter (starts at gridDim.x): items are handed out dynamically, SMs differ by > 10 %
  const float* bias2;            // [H][kBiasSlab/4] fp32 bias x log2(e), row stride LS (packed by pack_params), or NULL
  void* out;
  int trace;
  int prefetch_rows;             // chunk-rows of the NEXT item to warm in L2 during phase B (tuning knob)
  int prefetch_v;                // warm L2 with v during pass 1
  float* kbar_out;               // training: k_bar / beta of every item -> float32 [B, H, chunks, 64] (the backward reads them); or NULL
  float* beta_out;
};

// G = chunk-rows per pass-1 / pass-2 tile: 1 on the 28-wide grid (112-token rows); 4 on the 14-wide grid, where a 28-token row
// per TMA box would leave the kernel bound by the per-warp box issue rate (34 boxes per 196-token item)
template <int W, int GW, int CH, int NR, int G> struct Cfg {
  static constexpr int L = W * W;
  static constexpr int LP8 = (L + 7) & ~7;
  static constexpr int LS = (L + 4) & ~3;       // bias row stride: 16-byte rows (LDS.128, conflict-free for stride 52); column L = row max
  static constexpr int NCX = GW / CH;           // chunks per chunk-row
  static constexpr int CN = NR * NCX;           // chunks per item
  static constexpr int CNP = 8 * NR;            // padded chunk index space: c' = 8 r + cx
  static constexpr int TOK = GW * CH;           // tokens per chunk-row
  static constexpr int NT = (NR + G - 1) / G;   // tiles per operand and pass
  static constexpr int TOKT = G * TOK;          // tokens per tile
  static constexpr int CW = 8 * G;              // chunk columns (c' slots) per tile
  static constexpr int D2N = CW < 16 ? 16 : CW; // N of the phi-logit MMA (M = 128 needs a multiple of 16)
  static constexpr int KS = (TOKT + 15) / 16;   // k-steps over the tokens of a tile
  static constexpr int KBLK = (TOKT + 63) / 64; // 64-token K blocks of the Pool / P2 tiles
  static constexpr int kBlk = CW * 128;         // bytes of one 64-token block of a K-major [CW][TOKT] tile
  static constexpr bool kSplitIssue = G == 1;   // pass-1 / pass-2 loads shared between the TMA and MMA warps (schedule written for NT = 7)
  static constexpr int kP2Split = NT > 4 ? 4 : (NT + 1) / 2;   // pass 2: tiles [0, kP2Split) and the rest are handed to the beta MMAs separately
  __host__ __device__ static constexpr int rows_in(int R) { return (R + 1) * G <= NR ? G : NR - R * G; }   // chunk-rows of tile R
  static constexpr int JC = CH * CH;
  static constexpr int kQOff = 64 - L, kKOff = LP8 - L;   // tile row of window a's first token in the q slot / the k, v slots
  static constexpr int kPairBytes = 2 * L * 128;
  static_assert(kQOff >= 0 && kQOff + 2 * L <= 128 && kKOff + 2 * L <= 2 * LP8, "stacked pair layout");
  static_assert(GW % CH == 0 && NCX <= 7 && NR <= 7 && TOKT <= 128 && L <= 64 && CNP <= 64 && CW <= 32, "geometry");
  static_assert(G == 1 || NT * G * 8 <= 64, "tile columns");
  __host__ __device__ static constexpr bool chunk_ok(int c) { return c < CNP && (c & 7) < NCX; }
  // shared memory map (bytes from the 1024-aligned base)
  static constexpr int kSlot = 0;
  static constexpr int kKbar = kSlots * kSlotBytes;    // [64][128 B]  k_bar, row = c'
  static constexpr int kBeta = kKbar + 8192;           // [64][128 B]  beta,  row = c'
  static constexpr int kOm = kBeta + 8192;             // [64][128 B]  q_bar (+ noise), row = c'
  // Three things share [kOm, kLbuf), one after the other: the q_bar tile (until the phi-logit MMAs are done), the
  // NR softmax tiles P2 of pass 2 (until the beta MMAs are done), the output staging of phase B
  // (window a at +0, window b at +LP8*128, swizzled rows).
  static constexpr int kP2 = kOm;                      // NT x KBLK x [CW][128 B]
  static constexpr int kP2Bytes = NT * KBLK * kBlk;
  static constexpr int kOStage = kOm;
  static constexpr int kOStageBytes = LP8 * 128 + L * 128;
  static constexpr int kShared = kP2Bytes > kOStageBytes ? (kP2Bytes > 8192 ? kP2Bytes : 8192) : (kOStageBytes > 8192 ? kOStageBytes : 8192);
  static constexpr int kLbuf = (kOm + kShared + 1023) & ~1023;   // NT x [128] fp32 logit exchange
  static constexpr int kPool = kLbuf + ((NT * 128 * 4 + 1023) & ~1023);   // KBLK x [CW][128 B] pooling weights (constant)
  // bias slab: [L][LS] x log2(e), float32 -- except with four chunk-rows per tile, where a 16-bit slab frees the 4.8 KB the larger
  // pooling / P2 tiles need (values ~1e-1 rounded to 11 bits: far below the 16-bit P the logits feed)
  static constexpr bool kHalfBias = G == 4;
  static constexpr int kBias = kPool + KBLK * kBlk;
  static constexpr int kBiasSlab = (L * LS * (kHalfBias ? 2 : 4) + 15) & ~15;
  static constexpr int kZeroEnd = kBias + kBiasSlab;
  static constexpr int kLn = kZeroEnd;                 // [6][64] fp32: b_q, gain_q, beta_q, b_k, gain_k, beta_k
  static constexpr int kBars = (kLn + 6 * 64 * 4 + 7) & ~7;
  static constexpr int kTmemPtr = kBars + kNumBars * 8;
  static constexpr int kItemIds = kTmemPtr + 16;       // 2 x int: work item of the current / next round (dynamic scheduling)
  static constexpr int kBytes = kItemIds + 16;
  static constexpr int kDynamic = kBytes + 1024;
  static_assert(kDynamic <= 115712, "two CTAs per SM need <= 113 KB each");
  // TMEM columns
  static constexpr uint32_t cPoolQ = 0, cPoolK = 56;   // pass 1: [feat x c']; stays clear of the X buffers
  static constexpr uint32_t cLin = 0;                  // Linear: [c' x 128]
  static constexpr uint32_t cBetaT = 0;                // pass 2: [feat x c']
  static constexpr uint32_t cD2 = 128;                 // pass 2: NT x [token x D2N]
  static_assert(cD2 + D2N * NT <= 256, "TMEM budget (pass 2)");
  // phase B: local logits / P in [0, 2*LP8); chunk logits of pair p and later its O share buffer X[p&1]
  static constexpr uint32_t cSloc = 0, cX0 = 2 * LP8, cPloc = 0, cPrfa = LP8;
  static_assert(2 * LP8 + 128 <= 256 && (2 * LP8) % 16 == 0, "TMEM budget");
  // loads per item, in ring order: NT x (q_r, k_r) | means (borrowed) | W | NT x k_r | NT x v_r | pairs x (q, k, v)
  static constexpr int nAt = 2 * NT, nW = 2 * NT + 1, nPass2 = 2 * NT + 2, nPairs = 4 * NT + 2;
};

__device__ __forceinline__ uint32_t slot_of(uint32_t n) { return n & (kSlots - 1); }
__device__ __forceinline__ uint32_t par_of(uint32_t n) { return (n >> 2) & 1u; }

// optional phase trace of CTA 0 (EVA_SM100_TRACE=1): [0] compute thread 0, [1] MMA thread; pairs (event, clock64)
constexpr int kTraceLen = 8192;
__device__ unsigned long long g_trace[3][kTraceLen];   // [2] = TMA thread: (1000 + ring position in item, clock) at issue
template <bool kOn> struct Tracer {       // kOn = false (production instantiation): no code at all -- the kernel is instruction-cache bound
  unsigned long long* buf;
  int n;
  __device__ __forceinline__ void operator()(int ev) {
    if constexpr (kOn) {
      if (buf && n + 2 <= kTraceLen) { buf[n] = (unsigned long long)ev; buf[n + 1] = (unsigned long long)clock64(); n += 2; }
    }
  }
  __device__ __forceinline__ void finish() {
    if constexpr (kOn) { if (buf) buf[kTraceLen - 1] = n; }
  }
};

template <typename T, int W, int GW, int CH, int NR, int G, bool TR>
__global__ void __launch_bounds__(kThreads, 2)
eva_fused_kernel(const __grid_constant__ CUtensorMap tw_q, const __grid_constant__ CUtensorMap tw_k,
                 const __grid_constant__ CUtensorMap tw_v, const __grid_constant__ CUtensorMap tr_q,
                 const __grid_constant__ CUtensorMap tr_k, const __grid_constant__ CUtensorMap tr_v,
                 const __grid_constant__ CUtensorMap t_w, const __grid_constant__ CUtensorMap t_o, const Params p) {
  using C = Cfg<W, GW, CH, NR, G>;
  constexpr int L = C::L, LP8 = C::LP8, LS = C::LS, CN = C::CN, CNP = C::CNP, NCX = C::NCX, TOK = C::TOK, KS = C::KS;
  constexpr int NT = C::NT, TOKT = C::TOKT, CW = C::CW;
  extern __shared__ uint8_t smem_raw[];
  // align by offset arithmetic (not by integer round trip) so the compiler keeps the shared address space
  uint8_t* const sm = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* KBt = sm + C::kKbar;
  uint8_t* BTt = sm + C::kBeta;
  uint8_t* OMt = sm + C::kOm;
  uint8_t* PoolT = sm + C::kPool;
  uint8_t* P2t = sm + C::kP2;
  float* bias2 = reinterpret_cast<float*>(sm + C::kBias);
  float* lbuf = reinterpret_cast<float*>(sm + C::kLbuf);
  static_assert(NT <= 7 && (!C::kSplitIssue || (NR & 1) == 1), "one kNormDone barrier per tile; the split pass-2 issue schedule assumes odd NR");
  const uint32_t bars = ptx::smem_u32(sm + C::kBars);
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + C::kTmemPtr);
  auto bar = [&](int i) { return bars + 8u * i; };
  auto slot_ptr = [&](uint32_t s) { return sm + C::kSlot + s * kSlotBytes; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_per_item = C::nPairs + 3 * p.n_pairs;
  volatile int* const item_ids = reinterpret_cast<volatile int*>(sm + C::kItemIds);
  // consumers (MMA warp, compute warps): the TMA warp publishes the item of round ni in item_ids[ni & 1]
  auto item_of_round = [&](uint32_t ni) -> int {
    ptx::mbar_wait(bar(kItem0 + (ni & 1)), (ni >> 1) & 1);
    return item_ids[ni & 1];
  };

  // ---- one-time setup --------------------------------------------------------------------------
  for (int i = tid; i < C::kZeroEnd / 16; i += kThreads) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int t = tid; t < TOKT; t += kThreads)   // Pool^T[n][t] = 1/Jc for the chunk column n = 8 (row in tile) + cx that owns token t
    *reinterpret_cast<uint16_t*>(PoolT + ktile_off(8 * (t / TOK) + (t % GW) / CH, t, C::kBlk)) = IoFmt<T>::one(1.0f / C::JC);
  {
    float* ln = reinterpret_cast<float*>(sm + C::kLn);
    const float* src[6] = {p.b_q, p.g_q, p.beta_q, p.b_k, p.g_k, p.beta_k};
    for (int idx = tid; idx < 6 * 64; idx += kThreads) ln[idx] = src[idx >> 6] ? __ldg(src[idx >> 6] + (idx & 63)) : 0.f;
  }
  if (warp == 4 && lane == 0) {
    for (int s = 0; s < kSlots; ++s) { ptx::mbar_init(bar(kFull0 + s), 1); ptx::mbar_init(bar(kFree0 + s), 1); }
    ptx::mbar_init(bar(kPoolFull), 1);
    ptx::mbar_init(bar(kAFull), kComputeThreads);
    ptx::mbar_init(bar(kLinFull), 1);
    ptx::mbar_init(bar(kOmFull), kComputeThreads);
    ptx::mbar_init(bar(kD2Full), 1);
    ptx::mbar_init(bar(kP2Full), kComputeThreads);
    ptx::mbar_init(bar(kP2FullB), kComputeThreads);
    ptx::mbar_init(bar(kBetaFull), 1);
    ptx::mbar_init(bar(kStatsFull), kComputeThreads);
    ptx::mbar_init(bar(kSFull), 1);
    ptx::mbar_init(bar(kPFull), kComputeThreads);
    ptx::mbar_init(bar(kOFull0), 1);
    ptx::mbar_init(bar(kOFull1), 1);
    ptx::mbar_init(bar(kOFree0), kComputeThreads);
    ptx::mbar_init(bar(kOFree1), kComputeThreads);
    ptx::mbar_init(bar(kBiasFull), 1);
    ptx::mbar_init(bar(kBiasFree), kComputeThreads);
    ptx::mbar_init(bar(kItem0), 1);
    ptx::mbar_init(bar(kItem1), 1);
    for (int r = 0; r < NT; ++r) ptx::mbar_init(bar(kNormDone0 + r), kComputeThreads);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tw_q); ptx::prefetch_tmap(&tw_k); ptx::prefetch_tmap(&tw_v);
    ptx::prefetch_tmap(&tr_q); ptx::prefetch_tmap(&tr_k); ptx::prefetch_tmap(&tr_v);
    ptx::prefetch_tmap(&t_w);
    ptx::prefetch_tmap(&t_o);
  }
  if (warp == 5) ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr)), kTmemCols);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == 4) {
    // =================================== TMA producer ==========================================
    {
      // A warp gets one TMA box through the engine every ~0.55 us however deep its ring is (tools/tma_stream.cu), so the
      // loads of an item are split between this warp (q rows, W, bias, K rows, Q/K windows) and the MMA warp
      // (k rows, V rows, V windows).  Ring positions stay globally ordered; each issuer acquires its own positions.
      uint32_t nb = 0, ni = 0;
      const uint64_t keep = ptx::policy_evict_last(), stream = ptx::policy_evict_first();
      Tracer<TR> tr{(p.trace && blockIdx.x == 0 && lane == 0) ? g_trace[2] : nullptr, 0};
      // With two issuers the previous use of a slot may belong to the other warp.  A parity wait can only tell the last two
      // phases apart, so at every acquire the use BEFORE the previous one must already be known to be consumed.  The
      // assignment (passes 1 and 2: even ring positions here, odd ones in the MMA warp, except k_0 and k_1; phase B: q and k
      // here, v there) makes that hold by causality: commits complete in order, and each warp's own earlier acquires or
      // waits imply it.
      auto acquire = [&](uint32_t n, uint32_t bytes) -> uint32_t {   // returns the slot of ring position n; arms its full barrier
        const uint32_t s = slot_of(n);
        ptx::mbar_wait(bar(kFree0 + s), par_of(n) ^ 1);
        tr(1000 + (int)(n - nb));
        if (ptx::elect_one()) ptx::mbar_arrive_expect_tx(bar(kFull0 + s), bytes);
        return s;
      };
      auto publish = [&](uint32_t round, int it) {       // safe to overwrite: round - 2 was read before any load of round - 1 was consumed
        if (ptx::elect_one()) { item_ids[round & 1] = it; ptx::mbar_arrive(bar(kItem0 + (round & 1))); }
      };
      int item = blockIdx.x;
      publish(0, item);
      for (; item < p.items; ++ni, nb += n_per_item) {
        const int b = item / p.H, h = item % p.H;
        int item_next = 0;                               // fetched now, needed in phase B: the atomic's latency hides under pass 1
        if (ni > 0 && lane == 0) item_next = (int)atomicAdd(p.next_item, 1u);
#pragma unroll 1
        for (int r = 0; r < NT; ++r) {                   // pass 1: q tiles (first touch: HBM) ...
          uint32_t s = acquire(nb + 2 * r, TOKT * 128);
          if (ptx::elect_one()) ptx::tma_load_5d_hint(ptx::smem_u32(slot_ptr(s)), &tr_q, bar(kFull0 + s), 0, h, 0, r * CH * G, b, keep);
          if (!C::kSplitIssue || r < 2) {                // ... and the first two k rows: the MMA warp is still finishing the previous item
            s = acquire(nb + 2 * r + 1, TOKT * 128);
            if (ptx::elect_one()) ptx::tma_load_5d_hint(ptx::smem_u32(slot_ptr(s)), &tr_k, bar(kFull0 + s), 0, h, 0, r * CH * G, b, keep);
          }
        }
        if (ni == 0) {
          // programmatic dependent launch: this grid may start while pack_params (weight tile, bias slabs, work counter)
          // is still running; everything it produced is first touched below, by this warp only
          asm volatile("griddepcontrol.wait;" ::: "memory");
          if (lane == 0) item_next = (int)atomicAdd(p.next_item, 1u);
        }
        if (p.bias2) {                                   // this head's bias table, once the previous item's softmax is done
          ptx::mbar_wait(bar(kBiasFree), (ni & 1) ^ 1);
          if (ptx::elect_one()) ptx::mbar_arrive_expect_tx(bar(kBiasFull), C::kBiasSlab);
          if (ptx::elect_one()) ptx::bulk_load(ptx::smem_u32(sm + C::kBias), reinterpret_cast<const uint8_t*>(p.bias2) + (size_t)h * C::kBiasSlab,
                         C::kBiasSlab, bar(kBiasFull));
        }
        {
          const uint32_t s = acquire(nb + C::nW, 128 * 128);
          if (ptx::elect_one()) ptx::tma_load_2d(ptx::smem_u32(slot_ptr(s)), &t_w, bar(kFull0 + s), 0, 0);
        }
#pragma unroll 1
        for (int q = 0; q < 2 * NT; q += C::kSplitIssue ? 2 : 1) {   // pass 2: K_0 .. K_{NT-1}, V_0 .. V_{NT-1}; split issue: the even positions
          const uint32_t s = acquire(nb + C::nPass2 + q, TOKT * 128);
          const bool is_k = q < NT;
          if (ptx::elect_one()) ptx::tma_load_5d_hint(ptx::smem_u32(slot_ptr(s)), is_k ? &tr_k : &tr_v, bar(kFull0 + s), 0, h, 0, (is_k ? q : q - NT) * CH * G, b, keep);
        }
        item_next = __shfl_sync(0xffffffffu, item_next, 0);
        publish(ni + 1, item_next);
#pragma unroll 1
        for (int pr = 0; pr < p.n_pairs; ++pr) {         // phase B: q and k of the window pairs (L2)
          if (item_next < p.items) {                     // warm L2 with the next item's q/k chunk-rows, a few per pair
            const int bn = item_next / p.H, hn = item_next % p.H;
            for (int r = pr * p.prefetch_rows / p.n_pairs; r < (pr + 1) * p.prefetch_rows / p.n_pairs; ++r) {
              if (ptx::elect_one()) ptx::tma_prefetch_5d_hint(&tr_q, 0, hn, 0, r * CH, bn, keep);
              if (ptx::elect_one()) ptx::tma_prefetch_5d_hint(&tr_k, 0, hn, 0, r * CH, bn, keep);
            }
          }
          // pair pr = windows (wx, 2 wyp) and (wx, 2 wyp + 1), stacked vertically: ONE box of 7 x 14 tokens per operand
          // (the TMA engine charges ~300 cycles per box).  The q box lands 15 rows into its slot so that window a sits in
          // tile rows 15-63 and window b in rows 64-112 (TMEM lane = tile row: b starts on a warp boundary); the k and v
          // boxes land 7 rows in (a: rows 7-55, b: rows 56-104 of the 112 key columns).  128-byte swizzling is a function
          // of the shared-memory address, so an offset destination stays consistent with descriptors based at the slot.
          const int x0 = (pr % p.nwx) * W, y0 = (pr / p.nwx) * 2 * W;
          uint32_t s = acquire(nb + C::nPairs + 3 * pr, C::kPairBytes);      // last use of these lines: let L2 drop them first
          if (ptx::elect_one()) ptx::tma_load_5d_hint(ptx::smem_u32(slot_ptr(s)) + C::kQOff * 128, &tw_q, bar(kFull0 + s), 0, h, x0, y0, b, stream);
          s = acquire(nb + C::nPairs + 3 * pr + 1, C::kPairBytes);
          if (ptx::elect_one()) ptx::tma_load_5d_hint(ptx::smem_u32(slot_ptr(s)) + C::kKOff * 128, &tw_k, bar(kFull0 + s), 0, h, x0, y0, b, stream);
        }
        item = item_next;
      }
      // drain: the last hand-back of every ring slot is observed before the CTA exits (nothing depends on it, but a completed
      // mbarrier phase that nobody waited for is what compute-sanitizer's synccheck reports as "missing wait")
      if (nb >= (uint32_t)kSlots) {
#pragma unroll 1
        for (uint32_t n = nb - kSlots; n < nb; ++n) ptx::mbar_wait(bar(kFree0 + slot_of(n)), par_of(n));
      }
      tr.finish();
    }
  } else if (warp == 5) {
    // =================================== MMA issuer ============================================
    {
      constexpr uint32_t fmt = IoFmt<T>::kUmma;
      constexpr uint32_t id_pool = ptx::umma_idesc(fmt, fmt, 1, 0, 64, CW);     // A MN-major (feat), B K-major; N = chunk columns of a full tile ...
      constexpr uint32_t id_pool_last = ptx::umma_idesc(fmt, fmt, 1, 0, 64, 8 * C::rows_in(NT - 1));   // ... and of the last one
      constexpr uint32_t id_lin = ptx::umma_idesc(ptx::kFmtF16, ptx::kFmtF16, 0, 0, 128, 128);
      constexpr uint32_t id_d2 = ptx::umma_idesc(fmt, fmt, 0, 0, 128, C::D2N);
      constexpr uint32_t id_sl = ptx::umma_idesc(fmt, fmt, 0, 0, 128, 2 * LP8);
      constexpr uint32_t id_sr = ptx::umma_idesc(fmt, fmt, 0, 0, 128, 64);
      constexpr uint32_t id_pv = ptx::umma_idesc(fmt, fmt, 0, 1, 128, 64);
      const uint64_t dSlot0 = ptx::umma_desc_sw128(ptx::smem_u32(slot_ptr(0)));
      auto dSlot = [&](uint32_t s) { return dSlot0 + (uint64_t)(s * (kSlotBytes >> 4)); };
      const uint64_t dKB = ptx::umma_desc_sw128(ptx::smem_u32(KBt)), dBT = ptx::umma_desc_sw128(ptx::smem_u32(BTt));
      const uint64_t dOM = ptx::umma_desc_sw128(ptx::smem_u32(OMt)), dPool = ptx::umma_desc_sw128(ptx::smem_u32(PoolT));
      const uint64_t dP2 = ptx::umma_desc_sw128(ptx::smem_u32(P2t));
      // k-step ks over tokens: A (MN-major, rows = tokens) advances 16 rows; B (K-major [CW][TOKT]) 32 B inside a 64-token block
      auto tokB = [](int ks) { return (uint64_t)((ks >> 2) * (C::kBlk >> 4) + (ks & 3) * 2); };
      uint32_t nb = 0, ni = 0, np = 0;
      Tracer<TR> tr{(p.trace && blockIdx.x == 0 && lane == 0) ? g_trace[1] : nullptr, 0};
      auto wait_full = [&](uint32_t n) { ptx::mbar_wait(bar(kFull0 + slot_of(n)), par_of(n)); tr(2000 + (int)(n - nb)); };
      // one elected region per group of MMAs and their commits: every separate elect costs ~10 instructions of code
      auto free_raw = [&](uint32_t n) { ptx::umma_commit(bar(kFree0 + slot_of(n))); };
      // this warp's share of the loads (see the TMA warp): always issued after the MMAs that read the slot's previous tile
      const uint64_t keep = ptx::policy_evict_last(), stream = ptx::policy_evict_first();
      auto acquire = [&](uint32_t n, uint32_t bytes) -> uint32_t {      // see the TMA warp
        const uint32_t s = slot_of(n);
        ptx::mbar_wait(bar(kFree0 + s), par_of(n) ^ 1);
        tr(1000 + (int)(n - nb));
        if (ptx::elect_one()) ptx::mbar_arrive_expect_tx(bar(kFull0 + s), bytes);
        return s;
      };
      for (;; ++ni, nb += n_per_item) {
        const int item = item_of_round(ni);
        if (item >= p.items) break;
        const int b = item / p.H, h = item % p.H;
        auto load_row = [&](uint32_t n, const CUtensorMap* tm, int r) {
          const uint32_t s = acquire(n, TOKT * 128);
          if (ptx::elect_one()) ptx::tma_load_5d_hint(ptx::smem_u32(slot_ptr(s)), tm, bar(kFull0 + s), 0, h, 0, r * CH * G, b, keep);
        };
        auto load_v_pair = [&](int pr) {
          const uint32_t s = acquire(nb + C::nPairs + 3 * pr + 2, C::kPairBytes);
          if (ptx::elect_one()) ptx::tma_load_5d_hint(ptx::smem_u32(slot_ptr(s)) + C::kKOff * 128, &tw_v, bar(kFull0 + s), 0, h, (pr % p.nwx) * W, (pr / p.nwx) * 2 * W, b, stream);
        };
        // ---- pass 1: chunk means ------------------------------------------------------------
        tr(101);
        auto load_pass2 = [&](int q) {                        // relative position in pass 2: K_q (q < NR) or V_{q-NR}
          load_row(nb + C::nPass2 + q, q < NT ? &tr_k : &tr_v, q < NT ? q : q - NT);
        };
#pragma unroll 1
        for (int r = 0; r < NT; ++r) {
          const uint32_t nq = nb + 2 * r, nk = nq + 1;
          wait_full(nq);
          wait_full(nk);
          ptx::tc_fence_after();
          const uint32_t idp = (G > 1 && r == NT - 1) ? id_pool_last : id_pool;
          if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
              ptx::umma_ss(tmem + C::cPoolQ + CW * r, dSlot(slot_of(nq)) + 128 * ks, dPool + tokB(ks), idp, ks > 0);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
              ptx::umma_ss(tmem + C::cPoolK + CW * r, dSlot(slot_of(nk)) + 128 * ks, dPool + tokB(ks), idp, ks > 0);
            free_raw(nq);
            free_raw(nk);
            if (r == NT - 1) ptx::umma_commit(bar(kPoolFull));
          }
          if constexpr (C::kSplitIssue) {
            if (r + 2 < NT) load_row(nk + 4, &tr_k, r + 2);    // same slot as k_r: waits for the MMAs just issued
          }
        }
        tr(102);
        // ---- adaptive Linear: [q means ; k means] x [W_q ; W_k]^T (diagonal blocks used) ----------
        ptx::mbar_wait(bar(kAFull), ni & 1);
        // the borrowed slot was filled by the compute warps, not by TMA: complete its `full` phase by hand
        // so that the slot's phase count keeps matching the ring counter
        if (ptx::elect_one()) ptx::mbar_arrive(bar(kFull0 + slot_of(nb + C::nAt)));
        wait_full(nb + C::nW);
        ptx::tc_fence_after();
        // both tiles are dead once the Linear MMA has completed: hand the slots back right away so that the first k rows of
        // pass 2 are requested while the compute warps are still in the LayerNorm
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            ptx::umma_ss(tmem + C::cLin, dSlot(slot_of(nb + C::nAt)) + 2 * ks, dSlot(slot_of(nb + C::nW)) + 2 * ks, id_lin, ks > 0);
          ptx::umma_commit(bar(kLinFull));
          free_raw(nb + C::nW);
          free_raw(nb + C::nAt);
        }
        tr(103);
        if constexpr (C::kSplitIssue) {
          load_pass2(1);                                      // slot of k_{NR-1}
          load_pass2(3);                                      // slot of W: waits for the Linear MMA just issued
        }
        ptx::mbar_wait(bar(kOmFull), ni & 1);
        tr(104);      // q_bar / k_bar tiles written
        ptx::tc_fence_after();
        // ---- pass 2: phi-logits of every chunk-row, D2_r = K_r (Qbar_r + Kbar_r)^T ------------------------
#pragma unroll 1
        for (int r = 0; r < NT; ++r) {
          const uint32_t nk = nb + C::nPass2 + r;
          wait_full(nk);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_ss(tmem + C::cD2 + C::D2N * r, dSlot(slot_of(nk)) + 2 * ks, dOM + (uint64_t)(r * (C::kBlk >> 4)) + 2 * ks, id_d2, ks > 0);
            if (p.has_q) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                ptx::umma_ss(tmem + C::cD2 + C::D2N * r, dSlot(slot_of(nk)) + 2 * ks, dKB + (uint64_t)(r * (C::kBlk >> 4)) + 2 * ks, id_d2, 1);
            }
          }
          // the compute warps take |k|^2 from the same tile: hand it back when both are done
          ptx::mbar_wait(bar(kNormDone0 + r), ni & 1);
          if (ptx::elect_one()) {
            free_raw(nk);
            if (r == NT - 1) ptx::umma_commit(bar(kD2Full));
          }
          if constexpr (C::kSplitIssue) {
            if (r & 1) load_pass2(r + 4);                   // odd positions are this warp's: K_5, V_0, V_2 for NR = 7
          }
        }
        tr(110);
        // ---- beta^T += V_r^T P_r^T once the softmax batch is in shared memory ----------------------------
        ptx::mbar_wait(bar(kP2Full), ni & 1);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int r = 0; r < NT; ++r) {
          if (r == C::kP2Split) { ptx::mbar_wait(bar(kP2FullB), ni & 1); ptx::tc_fence_after(); }
          const uint32_t nv = nb + C::nPass2 + NT + r;
          wait_full(nv);
          ptx::tc_fence_after();
          const uint32_t idp = (G > 1 && r == NT - 1) ? id_pool_last : id_pool;
          if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
              ptx::umma_ss(tmem + C::cBetaT + CW * r, dSlot(slot_of(nv)) + 128 * ks,
                           dP2 + (uint64_t)(r * C::KBLK * (C::kBlk >> 4)) + tokB(ks), idp, ks > 0);
            free_raw(nv);
            if (r == NT - 1) ptx::umma_commit(bar(kBetaFull));
          }
          if constexpr (C::kSplitIssue) {
            if ((r & 1) == 0 && r + 4 < NT) load_pass2(NT + r + 4);   // V_4, V_6
          }
        }
        load_v_pair(0);                                      // slot of V_{NT-2}
        tr(111);
        ptx::mbar_wait(bar(kStatsFull), ni & 1);
        tr(120);
        ptx::tc_fence_after();
        // ---- phase B --------------------------------------------------------------------------------
#pragma unroll 1
        for (int pr = 0; pr < p.n_pairs; ++pr, ++np) {
          const uint32_t nq = nb + C::nPairs + 3 * pr, nk = nq + 1, nv = nq + 2;
          wait_full(nq);
          wait_full(nk);
          const uint32_t cX = C::cX0 + 64 * (np & 1);
          ptx::tc_fence_after();
          const uint64_t dQ = dSlot(slot_of(nq)), dK = dSlot(slot_of(nk)), dV = dSlot(slot_of(nv));
          if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + C::cSloc, dQ + 2 * ks, dK + 2 * ks, id_sl, ks > 0);
          }
          // X[np&1] last held O of pair np-2: its epilogue must have read it (per-buffer barrier: no lapping).
          // Only the chunk logits live there, so the local logits above are already in flight.
          if (np >= 2) ptx::mbar_wait(bar(kOFree0 + (np & 1)), ((np >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) ptx::umma_ss(tmem + cX, dQ + 2 * ks, dKB + 2 * ks, id_sr, ks > 0);
            ptx::umma_commit(bar(kSFull));
            free_raw(nq);
            free_raw(nk);
          }
          if (pr + 1 < p.n_pairs) load_v_pair(pr + 1);     // slot of K(pr): waits for the S MMAs just issued, under the softmax
          tr(130 + 2 * pr);
          ptx::mbar_wait(bar(kPFull), np & 1);
          wait_full(nv);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 2 * LP8 / 16; ++ks)
              ptx::umma_ts(tmem + cX, tmem + C::cPloc + 8 * ks, dV + 128 * ks, id_pv, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_ts(tmem + cX, tmem + C::cPrfa + 8 * ks, dBT + 128 * ks, id_pv, 1);
            ptx::umma_commit(bar(kOFull0 + (np & 1)));   // per-buffer barrier: the MMA warp may run two pairs ahead of the epilogue
            free_raw(nv);
          }
          tr(131 + 2 * pr);
        }
      }
      tr.finish();
    }
  } else {
    // =================================== compute warps ==========================================
    const int ws = tid >> 6;           // phase B: which window of the pair this row belongs to (warp-uniform)
    const int i = tid & 63;            // phase B: query slot inside the window (valid if < L)
    const int iq = ws ? (tid - 64) : (tid - C::kQOff);   // phase B: query slot inside my window (valid if 0 <= iq < L)
    const int ic = iq < 0 ? 0 : (iq < L ? iq : L - 1);
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const float scale = 0.125f;        // head_dim 64
    const float scale_log2 = scale * kLog2e;
    // pass 1 / beta readback: lanes 0-15 of each warp hold feature 16*warp + lane of an M=64 accumulator
    const int feat = 16 * warp + (lane & 15);
    const bool feat_lane = lane < 16;
    // pass 2: token of the tile owned by this thread: chunk-row trr inside the tile, chunk column tcx
    const int tcx = (tid % GW) / CH;
    const int trr = G > 1 ? tid / TOK : 0;
    uint32_t nb = 0, ni = 0, np_s = 0, np_e = 0;
    uint8_t* const ostage = sm + C::kOStage;
    const uint64_t out_policy = ptx::policy_evict_first();
    Tracer<TR> tr{(p.trace && blockIdx.x == 0 && tid == 0) ? g_trace[0] : nullptr, 0};
    for (;; ++ni, nb += n_per_item) {
      const int item = item_of_round(ni);
      if (item >= p.items) break;
      const int b = item / p.H, h = item % p.H;
      tr(1);
      // ---- pass 1 readback: means^T (TMEM) -> fp16 means tile [c'][feat] in the borrowed slot ----------
      const uint32_t nat = nb + C::nAt;
      uint8_t* At = slot_ptr(slot_of(nat));
      ptx::mbar_wait(bar(kFree0 + slot_of(nat)), par_of(nat) ^ 1);
      ptx::mbar_wait(bar(kPoolFull), ni & 1);
      ptx::tc_fence_after();
      tr(2);
#pragma unroll 1
      for (int r = 0; r < NR; ++r) {          // rolled (instruction-cache footprint): one chunk-row of means per iteration
        float mq[8], mk[8];
        ptx::tmem_ld8(trow + C::cPoolQ + 8 * r, reinterpret_cast<uint32_t*>(mq));
        ptx::tmem_ld8(trow + C::cPoolK + 8 * r, reinterpret_cast<uint32_t*>(mk));
        ptx::tmem_ld_wait();
        if (feat_lane) {   // padding rows (cx >= NCX) keep whatever the slot held: their Linear rows are never used
#pragma unroll
          for (int cx = 0; cx < NCX; ++cx) {
            const int c = 8 * r + cx;
            *reinterpret_cast<uint16_t*>(At + tile_off(c, feat)) = f16_bits(mq[cx]);
            *reinterpret_cast<uint16_t*>(At + tile_off(64 + c, feat)) = f16_bits(mk[cx]);
          }
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(kAFull));
      tr(3);
      // ---- Linear result -> bias, LayerNorm; lanes 0-63 = q side, lanes 64-127 = k side, lane & 63 = c' ----
      ptx::mbar_wait(bar(kLinFull), ni & 1);
      ptx::tc_fence_after();
      tr(4);
      {
        float y[64];
        tmem_ld_cols<64>(trow + C::cLin + (ws ? 64u : 0u), reinterpret_cast<uint32_t*>(y));
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        tr(240);
        const float* lnp = reinterpret_cast<const float*>(sm + C::kLn) + (ws ? 192 : 0);   // bias | gain | beta
        const bool has_lin_bias = ws ? (p.b_k != nullptr) : (p.b_q != nullptr);
        const bool has_ln = ws ? (p.g_k != nullptr) : (p.g_q != nullptr);
        if (has_lin_bias) {
#pragma unroll
          for (int e4 = 0; e4 < 16; ++e4) {
            const float4 bb = *reinterpret_cast<const float4*>(lnp + 4 * e4);
            y[4 * e4] += bb.x; y[4 * e4 + 1] += bb.y; y[4 * e4 + 2] += bb.z; y[4 * e4 + 3] += bb.w;
          }
        }
        if (has_ln) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int e = 0; e < 64; e += 4) { s0 += y[e]; s1 += y[e + 1]; s2 += y[e + 2]; s3 += y[e + 3]; }
          const float mean = ((s0 + s1) + (s2 + s3)) * (1.0f / 64);
          float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
          for (int e = 0; e < 64; e += 4) {
            const float d0 = y[e] - mean, d1 = y[e + 1] - mean, d2_ = y[e + 2] - mean, d3 = y[e + 3] - mean;
            v0 = fmaf(d0, d0, v0); v1 = fmaf(d1, d1, v1); v2 = fmaf(d2_, d2_, v2); v3 = fmaf(d3, d3, v3);
          }
          const float inv = 1.0f / sqrtf(((v0 + v1) + (v2 + v3)) * (1.0f / 64) + p.ln_eps);
#pragma unroll
          for (int e4 = 0; e4 < 16; ++e4) {
            const float4 gg = *reinterpret_cast<const float4*>(lnp + 64 + 4 * e4);
            const float4 bb = *reinterpret_cast<const float4*>(lnp + 128 + 4 * e4);
            y[4 * e4] = (y[4 * e4] - mean) * inv * gg.x + bb.x;
            y[4 * e4 + 1] = (y[4 * e4 + 1] - mean) * inv * gg.y + bb.y;
            y[4 * e4 + 2] = (y[4 * e4 + 2] - mean) * inv * gg.z + bb.z;
            y[4 * e4 + 3] = (y[4 * e4 + 3] - mean) * inv * gg.w + bb.w;
          }
        }
        tr(241);
        if (C::chunk_ok(i)) {
          uint8_t* const dst = (ws ? KBt : OMt) + i * 128;
          if (ws == 1 && p.kbar_out) {
            float4* kd = reinterpret_cast<float4*>(p.kbar_out + (((long long)b * p.H + h) * CN + (i >> 3) * NCX + (i & 7)) * 64);
#pragma unroll
            for (int e4 = 0; e4 < 16; ++e4) kd[e4] = make_float4(y[4 * e4], y[4 * e4 + 1], y[4 * e4 + 2], y[4 * e4 + 3]);
          }
          if (ws == 0) {
            // q side: omega = mu_coeff (q_bar + k_bar) + noise is applied as  mu_coeff ((q_bar + noise / mu_coeff) + k_bar):
            // the phi-logit MMAs accumulate K_r q'^T and K_r k_bar^T, so the two halves never have to meet in a thread
            const float* nz = p.noise ? p.noise + (((long long)b * p.H + h) * CN + (i >> 3) * NCX + (i & 7)) * 64 : nullptr;
#pragma unroll
            for (int e4 = 0; e4 < 16; ++e4) {
              float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
              if (nz) z = __ldg(reinterpret_cast<const float4*>(nz) + e4);
              if (p.has_q) {
                y[4 * e4] = fmaf(z.x, p.inv_mu_coeff, y[4 * e4]); y[4 * e4 + 1] = fmaf(z.y, p.inv_mu_coeff, y[4 * e4 + 1]);
                y[4 * e4 + 2] = fmaf(z.z, p.inv_mu_coeff, y[4 * e4 + 2]); y[4 * e4 + 3] = fmaf(z.w, p.inv_mu_coeff, y[4 * e4 + 3]);
              } else {
                y[4 * e4] = z.x; y[4 * e4 + 1] = z.y; y[4 * e4 + 2] = z.z; y[4 * e4 + 3] = z.w;
              }
            }
          }
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<uint4*>(dst + ((ch ^ (i & 7)) << 4)) =
                make_uint4(IoFmt<T>::pack2(y[8 * ch], y[8 * ch + 1]), IoFmt<T>::pack2(y[8 * ch + 2], y[8 * ch + 3]),
                           IoFmt<T>::pack2(y[8 * ch + 4], y[8 * ch + 5]), IoFmt<T>::pack2(y[8 * ch + 6], y[8 * ch + 7]));
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(bar(kOmFull));
      tr(5);
      // ---- pass 2: |k|^2 of my token in every chunk-row while the rows stream through the ring ---------
#pragma unroll 1
      for (int r = 0; r < NT; ++r) {          // rolled (instruction-cache footprint); |k|^2 waits in the logit exchange buffer
        const uint32_t nk = nb + C::nPass2 + r;
        const uint8_t* Kr = slot_ptr(slot_of(nk));
        ptx::mbar_wait(bar(kFull0 + slot_of(nk)), par_of(nk));
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (tid < C::rows_in(r) * TOK) {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint4 raw = *reinterpret_cast<const uint4*>(Kr + tid * 128 + ((ch ^ (tid & 7)) << 4));
            const float2 a = IoFmt<T>::unpack2(raw.x), b2 = IoFmt<T>::unpack2(raw.y), c2 = IoFmt<T>::unpack2(raw.z), d2 = IoFmt<T>::unpack2(raw.w);
            a0 = fmaf(a.x, a.x, a0); a1 = fmaf(a.y, a.y, a1); a2 = fmaf(b2.x, b2.x, a2); a3 = fmaf(b2.y, b2.y, a3);
            a0 = fmaf(c2.x, c2.x, a0); a1 = fmaf(c2.y, c2.y, a1); a2 = fmaf(d2.x, d2.x, a2); a3 = fmaf(d2.y, d2.y, a3);
          }
        }
        lbuf[r * 128 + tid] = (a0 + a1) + (a2 + a3);
        ptx::mbar_arrive(bar(kNormDone0 + r));
      }
      tr(10);
      // ---- all phi-logits are in TMEM: one exchange, one batch of 16-token softmaxes, NR P2 tiles -------
      ptx::mbar_wait(bar(kD2Full), ni & 1);
      ptx::tc_fence_after();
      tr(11);
      {
        const float dcoef = p.has_q ? p.mu_coeff : 1.0f;
        {
          uint32_t dd[NT][CW];
#pragma unroll
          for (int r = 0; r < NT; ++r) tmem_ld_cols<CW>(trow + C::cD2 + C::D2N * r, dd[r]);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int r = 0; r < NT; ++r) {
            uint32_t d8[8];                    // the 8 chunk columns of my chunk-row inside the tile
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              d8[c] = dd[r][c];
#pragma unroll
              for (int g = 1; g < G; ++g) d8[c] = (trr == g) ? dd[r][8 * g + c] : d8[c];
            }
            float dsel = __uint_as_float(d8[0]);
#pragma unroll
            for (int c = 1; c < NCX; ++c) dsel = (tcx == c) ? __uint_as_float(d8[c]) : dsel;
            const bool ok = tid < C::rows_in(r) * TOK;
            lbuf[r * 128 + tid] = ok ? scale_log2 * fmaf(dcoef, dsel, -0.5f * lbuf[r * 128 + tid]) : kNegInf;   // log2 units
          }
        }
        // the q_bar tile is dead (every phi-logit MMA has completed): clear the P2 tiles that overlay it
        for (int z = tid; z < C::kP2Bytes / 16; z += kComputeThreads) reinterpret_cast<uint4*>(P2t)[z] = make_uint4(0, 0, 0, 0);
        tr(12);
        ptx::named_bar_sync(1, kComputeThreads);
        tr(13);
        const int t0 = tid < TOKT ? trr * CH * GW + tcx * CH : 0;   // first token of my chunk inside the tile
        // two sub-batches: the beta MMAs of the first rows (and with them the hand-back of their V slots, i.e. the
        // request of the last V rows) start while the softmax of the remaining rows is still being computed
#pragma unroll 1
        for (int r = 0; r < NT; ++r) {                       // not unrolled: instruction-cache footprint
          if (r == C::kP2Split) {
            ptx::fence_proxy_async_smem();
            ptx::tc_fence_before();
            ptx::mbar_arrive(bar(kP2Full));
          }
          const float* lb_ = lbuf + r * 128;
          float lv[CH * CH];
#pragma unroll
          for (int yy = 0; yy < CH; ++yy) {
            if constexpr (CH == 4) {
              const float4 v4 = *reinterpret_cast<const float4*>(lb_ + yy * GW + t0);
              lv[4 * yy] = v4.x; lv[4 * yy + 1] = v4.y; lv[4 * yy + 2] = v4.z; lv[4 * yy + 3] = v4.w;
            } else if constexpr (CH == 2) {
              const float2 v2 = *reinterpret_cast<const float2*>(lb_ + yy * GW + t0);
              lv[2 * yy] = v2.x; lv[2 * yy + 1] = v2.y;
            } else {
#pragma unroll
              for (int xx = 0; xx < CH; ++xx) lv[CH * yy + xx] = lb_[yy * GW + t0 + xx];
            }
          }
          float mx = kNegInf;
#pragma unroll
          for (int j = 0; j < CH * CH; ++j) mx = fmaxf(mx, lv[j]);
          float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
          for (int j = 0; j < CH * CH; j += 2) { sum0 += ex2(lv[j] - mx); sum1 += ex2(lv[j + 1] - mx); }
          const float pt = __fdividef(ex2(lb_[tid] - mx), sum0 + sum1);
          if (tid < C::rows_in(r) * TOK) *reinterpret_cast<uint16_t*>(P2t + r * C::KBLK * C::kBlk + ktile_off(8 * trr + tcx, tid, C::kBlk)) = IoFmt<T>::one(pt);
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(kP2FullB));
      tr(14);
      // ---- beta^T (TMEM) -> beta tile [c'][feat] -------------------------------------------------------
      ptx::mbar_wait(bar(kBetaFull), ni & 1);
      ptx::tc_fence_after();
      tr(20);
#pragma unroll 1
      for (int r = 0; r < NR; ++r) {
        float bt[8];
        ptx::tmem_ld8(trow + C::cBetaT + 8 * r, reinterpret_cast<uint32_t*>(bt));
        ptx::tmem_ld_wait();
        if (feat_lane) {
#pragma unroll
          for (int cx = 0; cx < NCX; ++cx)
            *reinterpret_cast<uint16_t*>(BTt + tile_off(8 * r + cx, feat)) = IoFmt<T>::one(bt[cx]);
          if (p.beta_out) {
            float* bd = p.beta_out + (((long long)b * p.H + h) * CN + r * NCX) * 64 + feat;
#pragma unroll
            for (int cx = 0; cx < NCX; ++cx) bd[cx * 64] = bt[cx];
          }
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar(kStatsFull));
      tr(21);
      if (p.bias2) ptx::mbar_wait(bar(kBiasFull), ni & 1);

      // ---- phase B: pairs of windows; softmax of pair p+1 is done before the epilogue of pair p so that
      //      the PV MMA of pair p and the S MMA of pair p+1 run under SIMT work ---------------------------
      auto softmax_pair = [&](int pr) -> float {
        ptx::mbar_wait(bar(kSFull), np_s & 1);
        ptx::tc_fence_after();
        tr(30 + 4 * pr);
        float sl[L], sr[CNP];    // chunk logits by c' = 8 r + cx; every 8th column is padding (statically skipped)
        tmem_ld_cols<L>(trow + C::cSloc + (uint32_t)(ws ? LP8 : C::kKOff), reinterpret_cast<uint32_t*>(sl));   // my window's key columns
        tmem_ld_cols<CNP>(trow + C::cX0 + 64 * (np_s & 1), reinterpret_cast<uint32_t*>(sr));
        ptx::tmem_ld_wait();
        if (pr == 1) tr(250);
        // Shift by M = scale*max_j(raw) + max_j(bias) >= true row max (softmax is shift invariant; M only has to
        // prevent overflow and sits within max|bias| of the true max, far inside fp16's range for P).
        const float* brow = bias2 + ic * LS;
        const __half* hrow = reinterpret_cast<const __half*>(bias2) + ic * LS;       // (kHalfBias instantiations)
        float m0 = kNegInf, m1 = kNegInf, m2 = kNegInf, m3 = kNegInf;
#pragma unroll
        for (int j = 0; j < L; ++j) {
          if ((j & 3) == 0) m0 = fmaxf(m0, sl[j]); else if ((j & 3) == 1) m1 = fmaxf(m1, sl[j]);
          else if ((j & 3) == 2) m2 = fmaxf(m2, sl[j]); else m3 = fmaxf(m3, sl[j]);
        }
        float r0 = kNegInf, r1 = kNegInf, r2 = kNegInf, r3 = kNegInf;
#pragma unroll
        for (int c = 0; c < CNP; ++c) {
          if (!C::chunk_ok(c)) continue;
          if ((c & 3) == 0) r0 = fmaxf(r0, sr[c]); else if ((c & 3) == 1) r1 = fmaxf(r1, sr[c]);
          else if ((c & 3) == 2) r2 = fmaxf(r2, sr[c]); else r3 = fmaxf(r3, sr[c]);
        }
        const float bmax = p.bias2 ? (C::kHalfBias ? __half2float(hrow[L]) : brow[L]) : 0.f;
        const float mloc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * scale_log2 + bmax;
        const float mrfa = fmaxf(fmaxf(r0, r1), fmaxf(r2, r3)) * scale_log2;
        const float mx = fmaxf(mloc, mrfa);
        if (pr == 1) tr(251);
        uint32_t pl[LP8 / 2], prf[32], zeros[LP8 / 2];
        // 16-bit P, two keys per TMEM column: word m = (e[2m], e[2m+1]).  Window b's keys start at the even position LP8
        // (words go left-aligned into the second LP8/2 columns as they are); window a's start at the odd position
        // LP8 - L, so its words are funnel-shifted by 16 bits and right-aligned in the first LP8/2 columns (below).
        // The exponent arguments and the row sum are computed two at a time (FADD2 / FFMA2).
        static_assert((L & 1) == 1 && LP8 - L == 7 && LS % 4 == 0, "P packing assumes window 7");
        const uint64_t sc2 = pk2(scale_log2, scale_log2), nmx2 = pk2(-mx, -mx);
        uint64_t acc0 = pk2(0.f, 0.f), acc1 = acc0;
#pragma unroll
        for (int m4 = 0; m4 < (L + 3) / 4; ++m4) {          // four keys per 16-byte bias load
          float4 bq;
          if constexpr (C::kHalfBias) {
            const uint2 hb = *reinterpret_cast<const uint2*>(hrow + 4 * m4);
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&hb.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&hb.y));
            bq = make_float4(lo.x, lo.y, hi.x, hi.y);
          } else {
            bq = *reinterpret_cast<const float4*>(brow + 4 * m4);
          }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int m = 2 * m4 + hh;                      // P word: keys 2m, 2m+1
            if (2 * m < L) {
              const uint64_t t = fma2(pk2(sl[2 * m], sl[(2 * m + 1 < L) ? 2 * m + 1 : 2 * m]), sc2,
                                      add2(hh ? pk2(bq.z, bq.w) : pk2(bq.x, bq.y), nmx2));
              float x0, x1;
              upk2(t, x0, x1);
              const float e0 = ex2(x0), e1 = (2 * m + 1 < L) ? ex2(x1) : 0.f;
              if (m & 1) acc1 = add2(acc1, pk2(e0, e1)); else acc0 = add2(acc0, pk2(e0, e1));
              pl[m] = IoFmt<T>::pack2(e0, e1);
            }
          }
        }
#pragma unroll
        for (int m = (L + 1) / 2; m < LP8 / 2; ++m) pl[m] = 0u;
#pragma unroll
        for (int m = 0; m < LP8 / 2; ++m) zeros[m] = 0u;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (C::chunk_ok(2 * j)) {                          // 2j is never a padding column when 2j+1 is a chunk
            const uint64_t t = fma2(pk2(sr[2 * j], sr[C::chunk_ok(2 * j + 1) ? 2 * j + 1 : 2 * j]), sc2, nmx2);
            float x0, x1;
            upk2(t, x0, x1);
            const float e0 = ex2(x0), e1 = C::chunk_ok(2 * j + 1) ? ex2(x1) : 0.f;
            if (j & 1) acc1 = add2(acc1, pk2(e0, e1)); else acc0 = add2(acc0, pk2(e0, e1));
            prf[j] = IoFmt<T>::pack2(e0, e1);
          } else {
            prf[j] = 0u;
          }
        }
        float s0, s1, s2, s3;
        upk2(acc0, s0, s1);
        upk2(acc1, s2, s3);
        if (pr == 1) tr(252);
        // P (16-bit, two per column) overwrites the S columns this thread has finished reading
        if (ws) {
          tmem_st_cols<LP8 / 2>(trow + C::cPloc + LP8 / 2, pl);
        } else {
          constexpr int kLead = LP8 / 2 - (L + 1) / 2;      // 3 leading zero words, then (e[2m-1], e[2m]) for m = 0 .. (L-1)/2
          uint32_t ps[LP8 / 2];
#pragma unroll
          for (int m = 0; m < LP8 / 2; ++m)
            ps[m] = m < kLead ? 0u : __funnelshift_l(m - kLead > 0 ? pl[m - kLead - 1] : 0u, pl[m - kLead], 16);
          tmem_st_cols<LP8 / 2>(trow + C::cPloc, ps);
        }
        tmem_st_cols<LP8 / 2>(trow + C::cPloc + (uint32_t)(ws ? 0 : LP8 / 2), zeros);
        tmem_st_cols<32>(trow + C::cPrfa, prf);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar(kPFull));
        if (p.bias2 && pr == p.n_pairs - 1) ptx::mbar_arrive(bar(kBiasFree));   // bias table no longer needed for this item
        tr(31 + 4 * pr);
        ++np_s;
        return (s0 + s1) + (s2 + s3);
      };
      // epilogue: O / rowsum -> swizzled staging rows (window a: rows 0..L-1, window b: rows L..2L-1) -> one TMA store per pair
      auto epilogue_pair = [&](int pr, float sum) {
        ptx::mbar_wait(bar(kOFull0 + (np_e & 1)), (np_e >> 1) & 1);
        ptx::tc_fence_after();
        tr(32 + 4 * pr);
        float o[64];
        tmem_ld_cols<64>(trow + C::cX0 + 64 * (np_e & 1), reinterpret_cast<uint32_t*>(o));
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar(kOFree0 + (np_e & 1)));
        if (pr == 1) tr(253);
        if (warp == 0 && ptx::elect_one()) ptx::bulk_wait_read0();   // the previous store has drained the staging rows
        if (pr == 1) tr(270);
        ptx::named_bar_sync(2, kComputeThreads);
        if (pr == 1) tr(271);
        if (iq >= 0 && iq < L) {
          const float inv = 1.0f / sum;
          const int orow = ws * L + iq;
          uint8_t* row = ostage + orow * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<uint4*>(row + ((ch ^ (orow & 7)) << 4)) =
                make_uint4(IoFmt<T>::pack2(o[8 * ch] * inv, o[8 * ch + 1] * inv), IoFmt<T>::pack2(o[8 * ch + 2] * inv, o[8 * ch + 3] * inv),
                           IoFmt<T>::pack2(o[8 * ch + 4] * inv, o[8 * ch + 5] * inv), IoFmt<T>::pack2(o[8 * ch + 6] * inv, o[8 * ch + 7] * inv));
        }
        if (pr == 1) tr(272);
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(2, kComputeThreads);
        if (pr == 1) tr(273);
        if (warp == 0 && ptx::elect_one()) {
          // the output is never re-read here: evict-first keeps L2 for the q/k/v lines that are
          // (a direct st.global of each thread's 128-byte row was measured 14 % slower: 16-byte pieces of 32 different lines per instruction)
          ptx::tma_store_5d_hint(&t_o, ptx::smem_u32(ostage), 0, h, (pr % p.nwx) * W, (pr / p.nwx) * 2 * W, b, out_policy);
          ptx::bulk_commit_group();
        }
        tr(33 + 4 * pr);
        ++np_e;
      };
      float sum_prev = 0.f;
#pragma unroll 1
      for (int pr = 0; pr <= p.n_pairs; ++pr) {            // one call site each: the kernel is instruction-cache sensitive
        float sum_new = 0.f;
        if (pr < p.n_pairs) sum_new = softmax_pair(pr);
        if (pr > 0) epilogue_pair(pr - 1, sum_prev);
        sum_prev = sum_new;
      }
      if (warp == 0 && ptx::elect_one()) ptx::bulk_wait_read0();   // staging rows alias phase-A tiles of the next item
    }
    if (warp == 0 && ptx::elect_one()) ptx::bulk_wait_all();
    tr.finish();
  }
  // ---- teardown ------------------------------------------------------------------------------------
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 5) ptx::tmem_dealloc(tmem, kTmemCols);
}

static int trace_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EVA_SM100_TRACE");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

// Tensor maps depend only on the pointers / strides / sizes of a call: a model that runs the same layer on the same
// activation buffers (every layer of a network under the caching allocator, every replay of a CUDA graph) re-encodes the same
// eight maps on each forward.  Small per-thread cache, keyed by everything the encode reads (cached tensor-map state).
struct MapKey {
  const void *q, *k, *v, *out, *w16;
  long long qs[3], ks[3], vs[3];
  int B, H, gh, gw, io, variant;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapSet { CUtensorMap twq, twk, twv, trq, trk, trv, tw, to; };
struct MapCache {
  static constexpr int kEntries = 16;
  MapKey key[kEntries];
  MapSet val[kEntries];
  int used = 0, next = 0;
  MapSet* find(const MapKey& k) {
    for (int i = 0; i < used; ++i) if (key[i] == k) return &val[i];
    return nullptr;
  }
  MapSet* insert(const MapKey& k) {
    const int i = used < kEntries ? used++ : (next++ % kEntries);
    key[i] = k;
    return &val[i];
  }
};

template <typename T, int W, int GW, int CH, int NR, int G>
static cudaError_t launch_t(const Geo& g, const View& q, const View& k, const View& v, const EvaAdaptive& ada,
                            const float* noise, const float* bias, long long bias_sh, void* out, void* workspace,
                            cudaStream_t st, const char** msg, float* kbar_out, float* beta_out) {
  using C = Cfg<W, GW, CH, NR, G>;
  constexpr int io = std::is_same<T, __half>::value ? EVA_F16 : EVA_BF16;
  __half* w16 = reinterpret_cast<__half*>(workspace);
  float* bias2 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + 128 * 64 * sizeof(__half));
  unsigned int* next_item = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(bias2) + (size_t)g.H * C::kBiasSlab);
  const int items = g.B * g.H;
  const int dev = current_device();
  static const int ctas_per_sm = env_int("EVA_SM100_CTAS_PER_SM", 2);   // tuning knob, read once; 2 = as many as fit
  const int max_ctas = ctas_per_sm * sm_count(dev);
  const int grid = items < max_ctas ? items : max_ctas;
  pack_params<<<32, 256, 0, st>>>(ada.w_q, ada.w_k, w16, bias, bias_sh, bias2, g.H, C::L, C::LS, C::kBiasSlab / (C::kHalfBias ? 2 : 4), next_item,
                                  (unsigned)grid, C::kHalfBias ? 1 : 0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { *msg = "pack_params launch"; return e; }
  static thread_local MapCache cache;
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.q = q.ptr; key.k = k.ptr; key.v = v.ptr; key.out = out; key.w16 = w16;
  key.qs[0] = q.sb; key.qs[1] = q.sn; key.qs[2] = q.sh; key.ks[0] = k.sb; key.ks[1] = k.sn; key.ks[2] = k.sh;
  key.vs[0] = v.sb; key.vs[1] = v.sn; key.vs[2] = v.sh;
  key.B = g.B; key.H = g.H; key.gh = g.gh; key.gw = g.gw; key.io = io; key.variant = GW * 16 + G;
  MapSet* ms = cache.find(key);
  if (!ms) {
    MapSet fresh;
    View ov;
    ov.ptr = out; ov.sh = 64; ov.sn = (long long)g.H * 64; ov.sb = (long long)g.N * g.H * 64;
    if (!make_box_map(&fresh.twq, q, g, io, W, 2 * W) || !make_box_map(&fresh.twk, k, g, io, W, 2 * W) || !make_box_map(&fresh.twv, v, g, io, W, 2 * W) ||
        !make_box_map(&fresh.trq, q, g, io, GW, CH * G) || !make_box_map(&fresh.trk, k, g, io, GW, CH * G) || !make_box_map(&fresh.trv, v, g, io, GW, CH * G) ||
        !make_weight_map(&fresh.tw, w16) || !make_box_map(&fresh.to, ov, g, io, W, 2 * W)) {
      *msg = "cuTensorMapEncodeTiled failed";
      return cudaErrorInvalidValue;
    }
    ms = cache.insert(key);
    *ms = fresh;
  }
  Params p{};
  p.B = g.B; p.H = g.H; p.N = g.N; p.gh = g.gh; p.gw = g.gw;
  p.nwx = g.gw / g.window; p.n_windows = g.n_windows; p.n_pairs = (g.n_windows + 1) / 2;
  p.items = g.B * g.H;
  p.b_q = ada.b_q; p.g_q = ada.ln_gain_q; p.beta_q = ada.ln_bias_q;
  p.b_k = ada.b_k; p.g_k = ada.ln_gain_k; p.beta_k = ada.ln_bias_k;
  p.has_q = ada.w_q != nullptr;
  p.mu_coeff = ada.mu_coeff; p.inv_mu_coeff = ada.mu_coeff != 0.f ? 1.0f / ada.mu_coeff : 0.f; p.ln_eps = ada.ln_eps;
  p.noise = noise; p.bias2 = bias ? bias2 : nullptr; p.out = out; p.next_item = next_item;
  p.trace = trace_enabled();
  static const int prefetch_rows = env_int("EVA_SM100_PREFETCH_ROWS", 0);   // measured: warming L2 with the next item's rows no longer pays (0.5 % slower)
  static const int prefetch_v = env_int("EVA_SM100_PREFETCH_V", 0);         // measured: warming L2 with v during pass 1 costs 5 % (L2 is already full)
  p.prefetch_rows = prefetch_rows;
  p.prefetch_v = prefetch_v;
  p.kbar_out = kbar_out; p.beta_out = beta_out;
  auto kern = p.trace ? eva_fused_kernel<T, W, GW, CH, NR, G, true> : eva_fused_kernel<T, W, GW, CH, NR, G, false>;
  static bool attr_set[2][kMaxDevices] = {};         // per (instantiation, traced or not, device): the attribute is sticky
  if (dev < 0 || dev >= kMaxDevices || !attr_set[p.trace ? 1 : 0][dev]) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kDynamic);
    if (e != cudaSuccess) { *msg = "cudaFuncSetAttribute"; return e; }
    if (dev >= 0 && dev < kMaxDevices) attr_set[p.trace ? 1 : 0][dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = C::kDynamic; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;           // overlap with pack_params (see griddepcontrol.wait in the TMA warp)
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, ms->twq, ms->twk, ms->twv, ms->trq, ms->trk, ms->trv, ms->tw, ms->to, p);
  if (e != cudaSuccess) { *msg = "kernel launch"; return e; }
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

// geometry families the fused kernel is instantiated for: window 7, 49 chunks on a 28-wide (chunk 4) or
// 14-wide (chunk 2) grid -- DeiT-tiny/small p8 and p16 (BASELINE configs c2, c3)
// chunk-rows per pass-1 / pass-2 tile on the 14-wide grid (see Cfg): G = 4 (one 112-token box per 4 rows); its [32][112] P2 / pooling
// tiles need 4.8 KB more than G = 2 -- paid for by the 16-bit bias slab of this instantiation (Cfg::kHalfBias).  Round 1 tried to
// make room with 15 KB ring slots instead: they cannot hold the 16 KB [W_q ; W_k] tile (the overflow corrupted the next slot)
constexpr int kG14 = 4;
static int fused_variant(const Geo& g) {
  if (g.window != 7 || g.n_chunks != 49) return 0;
  if (g.gw == 28 && g.gh == 28 && g.chunk == 4) return 1;
  if (g.gw == 14 && g.gh == 14 && g.chunk == 2) return 2;
  return 0;
}

bool fused_supported(const Geo& g, int io_dtype, const View& q, const View& k, const View& v, const uint8_t* mask,
                     const EvaAdaptive& ada, const float* bias, long long bias_sh) {
  (void)ada; (void)bias; (void)bias_sh;
  if (fused_disabled()) return false;
  if (g.dims != 2 || g.ext != 0 || g.chunk_ext != 0 || g.causal || g.D != 64 || mask != nullptr) return false;
  if (io_dtype != EVA_F16 && io_dtype != EVA_BF16) return false;
  if (fused_variant(g) == 0) return false;
  for (const View* x : {&q, &k, &v}) {
    if (x->sh * 2 % 16 || x->sn * 2 % 16 || x->sb * 2 % 16) return false;
    if (x->sh <= 0 || x->sn <= 0 || x->sb <= 0) return false;
  }
  return fused::get_encode() != nullptr;
}

size_t fused_workspace_bytes(const Geo& g) {
  const int L = g.window * g.window, LS = (L + 4) & ~3;
  return 128 * 64 * sizeof(__half) + (size_t)g.H * ((L * LS * 4 + 15) & ~15) + 16;   // W tile | bias slabs | work counter
}

// diagnostic (not part of the public ABI): copy the phase trace of CTA 0 to the host
extern "C" int eva_debug_read_trace(unsigned long long* dst, int which, int n) {
  if (which < 0 || which > 2 || n > fused::kTraceLen) return -22;
  return cudaMemcpyFromSymbol(dst, fused::g_trace, (size_t)n * 8, (size_t)which * fused::kTraceLen * 8) == cudaSuccess ? 0 : -5;
}

cudaError_t launch_fused(const Geo& g, int io_dtype, const View& q, const View& k, const View& v,
                         const EvaAdaptive& ada, const float* noise, const float* bias, long long bias_sh,
                         void* out, void* workspace, cudaStream_t st, const char** msg, float* kbar_out, float* beta_out) {
  const int var = fused_variant(g);
  if (io_dtype == EVA_F16) {
    if (var == 1) return fused::launch_t<__half, 7, 28, 4, 7, 1>(g, q, k, v, ada, noise, bias, bias_sh, out, workspace, st, msg, kbar_out, beta_out);
    return fused::launch_t<__half, 7, 14, 2, 7, kG14>(g, q, k, v, ada, noise, bias, bias_sh, out, workspace, st, msg, kbar_out, beta_out);
  }
  if (var == 1) return fused::launch_t<__nv_bfloat16, 7, 28, 4, 7, 1>(g, q, k, v, ada, noise, bias, bias_sh, out, workspace, st, msg, kbar_out, beta_out);
  return fused::launch_t<__nv_bfloat16, 7, 14, 2, 7, kG14>(g, q, k, v, ada, noise, bias, bias_sh, out, workspace, st, msg, kbar_out, beta_out);
}

}  // namespace eva
