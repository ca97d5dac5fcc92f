// Helpers shared by the two fused EVA kernels (eva_fused_sm100.cu: one item per CTA, tiles streamed through a ring;
// eva_cluster_sm100.cu: one item per 2-CTA cluster, k/v resident in shared memory): 16-bit I/O formats, TMEM <-> register
// column helpers, 128-byte-swizzle tile addressing, the parameter-packing kernel and the tensor-map encoders.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "launch.h"
#include "sm100_ptx.cuh"

namespace eva {
namespace fused {

template <typename T> struct IoFmt;
template <> struct IoFmt<__half> {
  static constexpr uint32_t kUmma = ptx::kFmtF16;
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack2(uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }
  static __device__ __forceinline__ uint16_t one(float a) { return __half_as_ushort(__float2half_rn(a)); }
};
template <> struct IoFmt<__nv_bfloat16> {
  static constexpr uint32_t kUmma = ptx::kFmtBF16;
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  static __device__ __forceinline__ float2 unpack2(uint32_t v) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v)); }
  static __device__ __forceinline__ uint16_t one(float a) { return __bfloat16_as_ushort(__float2bfloat16_rn(a)); }
};
__device__ __forceinline__ uint16_t f16_bits(float a) { return __half_as_ushort(__float2half_rn(a)); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

using ptx::pk2; using ptx::upk2; using ptx::fma2; using ptx::add2;

// N consecutive TMEM columns -> registers (x16 / x8 / x1 pieces, compile time)
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* r) {
#pragma unroll
  for (int g = 0; g < N / 16; ++g) ptx::tmem_ld16(taddr + 16 * g, r + 16 * g);
  constexpr int done = (N / 16) * 16;
  if constexpr ((N % 16) >= 8) ptx::tmem_ld8(taddr + done, r + done);
  constexpr int done2 = done + (((N % 16) >= 8) ? 8 : 0);
#pragma unroll
  for (int j = done2; j < N; ++j) ptx::tmem_ld1(taddr + j, r[j]);
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

// byte offset of 16-bit element (row, col) inside a [rows][64] tile with 128-byte swizzle
__device__ __forceinline__ int tile_off(int row, int col) {
  return row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1));
}
// byte offset of 16-bit element (row n, token t) in a K-major [rows][tokens] tile made of 64-token blocks of `blk` bytes (rows x 128)
__device__ __forceinline__ int ktile_off(int n, int t, int blk) { return (t >> 6) * blk + tile_off(n, t & 63); }


// workspace packing: [W_q ; W_k] (fp32 [64][64] each, row-major [out][in]) -> fp16 [128][64];
// bias [H or 1][L][L] -> per-head slabs [L][LS] fp32 pre-multiplied by log2(e)
static __global__ void pack_params(const float* __restrict__ wq, const float* __restrict__ wk, __half* __restrict__ w16,
                            const float* __restrict__ bias, long long bias_sh, float* __restrict__ bias2, int H, int L,
                            int LS, int slab_floats, unsigned int* next_item, unsigned int first_free_item, int half_bias = 0) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the fused kernel may begin its prologue and first loads now
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0) *next_item = first_free_item;          // items 0 .. grid-1 are taken by blockIdx, the rest are handed out dynamically
  if (idx < 128 * 64) {
    const int row = idx >> 6, col = idx & 63;
    const float* src = row < 64 ? wq : wk;
    w16[idx] = __float2half_rn(src ? src[(row & 63) * 64 + col] : 0.f);
  }
  if (bias) {
    for (int j = idx; j < H * slab_floats; j += gridDim.x * blockDim.x) {
      const int h = j / slab_floats, o = j % slab_floats, r = o / LS, c = o % LS;
      float val = 0.f;
      if (r < L && c < L) {
        val = bias[(long long)h * bias_sh + r * L + c] * kLog2e;
      } else if (r < L && c == L) {      // row maximum: lets the softmax bound its max without touching the bias
        val = -INFINITY;
        for (int cc = 0; cc < L; ++cc) {
          const float bv = bias[(long long)h * bias_sh + r * L + cc] * kLog2e;
          val = fmaxf(val, half_bias ? __half2float(__float2half_rn(bv)) : bv);      // the maximum of the values as they are stored
        }
      }
      if (half_bias) reinterpret_cast<__half*>(bias2)[j] = __float2half_rn(val);     // slab_floats then counts 16-bit elements
      else bias2[j] = val;
    }
  }
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

// [B, gh, gw, H, 64] view (strides from the caller's q/k/v view) with a (box_h x box_w x 64) box, 128-B swizzle
static bool make_box_map(CUtensorMap* tm, const View& v, const Geo& g, int io_dtype, int box_w, int box_h) {
  auto enc = get_encode();
  if (!enc) return false;
  const cuuint64_t dims[5] = {64, (cuuint64_t)g.H, (cuuint64_t)g.gw, (cuuint64_t)g.gh, (cuuint64_t)g.B};
  const cuuint64_t strides[4] = {(cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2, (cuuint64_t)v.sn * g.gw * 2, (cuuint64_t)v.sb * 2};
  const cuuint32_t box[5] = {64, 1, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(tm, io_dtype == EVA_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                         const_cast<void*>(v.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static bool make_weight_map(CUtensorMap* tm, const void* w16) {
  auto enc = get_encode();
  if (!enc) return false;
  const cuuint64_t dims[2] = {64, 128};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w16), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}


constexpr int kMaxDevices = 64;
static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
// per device: a process may drive several different GPUs
static int sm_count(int dev) {
  static int n[kMaxDevices] = {};
  if (dev < 0 || dev >= kMaxDevices) { int v = 0; cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); return v > 0 ? v : 148; }
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}


}  // namespace fused
}  // namespace eva
