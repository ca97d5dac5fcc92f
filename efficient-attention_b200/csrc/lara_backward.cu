// LARA backward, the [landmarks x tokens]-sized part of the mis-opt estimator (gradient of lara.py:201-246): three kernels that fuse
// everything BETWEEN the batched GEMMs of the explicit formulas (efficient_attention/_recompute.py: _lara_stage2_backward_fused;
// the GEMMs themselves are plain library calls on [BH, C, N] / [BH, N, d] tensors).  Matrices are row-major [BH][rows][N] in the
// activation format T (float, half, bfloat16); all arithmetic is float32.
//   lara_bwd_rows1      per (item, landmark) row:  Bk = s B_raw - s |k|^2 / 2 -> lse_B, Pk = softmax_m;  T = s T_raw -> t = softmax_n
//   lara_bwd_cols       per (item, token) column:  alpha, log-weights, W = softmax_c, dlw = W o (dW - <W, dW>), d alpha, dt; and the
//                       per-landmark sums  d lse_B = sum_n dlw,  d bh = sum_n d alpha,  R = sum_n t dt
//   lara_bwd_row_affine per row:  X <- Y o (X - a_c + b_c)      (dT = t o (dt - R);  dBk = Pk o (dPk - <dkv, kv> + d lse_B))
#include "common.cuh"
#include "launch.h"

namespace eva {

template <typename T>
__global__ void __launch_bounds__(256)
lara_bwd_rows1_kernel(T* __restrict__ X, T* __restrict__ Bm, const float* __restrict__ k2s, float* __restrict__ lseB,
                      float* __restrict__ lseT, int C, int N, float s, long long rows_total) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // (item, landmark)
  if (row >= rows_total) return;
  const long long item = row / C;
  const int c = (int)(row % C);
  T* brow = Bm + row * N;
  T* trow = X + (item * 2 * C + C + c) * N;
  const float* kk = k2s + item * N;
  float mb = kNegInf, mt = kNegInf;
  for (int n = lane; n < N; n += 32) {
    mb = fmaxf(mb, fmaf(s, to_f32(brow[n]), -kk[n]));
    mt = fmaxf(mt, s * to_f32(trow[n]));
  }
  mb = warp_max(mb); mt = warp_max(mt);
  float sb = 0.f, st = 0.f;
  for (int n = lane; n < N; n += 32) {
    sb += expf(fmaf(s, to_f32(brow[n]), -kk[n]) - mb);
    st += expf(s * to_f32(trow[n]) - mt);
  }
  sb = warp_sum(sb); st = warp_sum(st);
  const float lb = mb + logf(sb), lt = mt + logf(st);
  for (int n = lane; n < N; n += 32) {
    brow[n] = from_f32<T>(expf(fmaf(s, to_f32(brow[n]), -kk[n]) - lb));
    trow[n] = from_f32<T>(expf(s * to_f32(trow[n]) - lt));
  }
  if (lane == 0) { lseB[row] = lb; lseT[row] = lt; }
}

// X rows [0, C): A_raw in, W out; rows [C, 2C): t.  dW: kv dO^T.  M2 rows [0, C): dlw out; rows [C, 2C): dt out.
template <typename T>
__global__ void __launch_bounds__(128)
lara_bwd_cols_kernel(T* __restrict__ X, const T* __restrict__ dW, T* __restrict__ M2, const float* __restrict__ q2s,
                     const float* __restrict__ bh, const float* __restrict__ lp, const float* __restrict__ lseB,
                     float* __restrict__ dlseB, float* __restrict__ dbh, float* __restrict__ R, int C, int N, float s, float coeff) {
  __shared__ float red[3][4];
  const long long item = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = n < N;
  const int nn = live ? n : N - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* Ar = X + item * 2 * C * N + nn;                    // + c * N
  const T* tr = Ar + (long long)C * N;
  const T* dWr = dW + item * C * N + nn;
  T* dlw_o = M2 + item * 2 * C * N + nn;
  T* dt_o = dlw_o + (long long)C * N;
  const float* bhv = bh + item * C;
  const float* lpv = lp + item * C;
  const float* lbv = lseB + item * C;
  const float qq = q2s[item * N + nn];
  float tbar = 0.f;
  for (int c = 0; c < C; ++c) tbar += to_f32(tr[(long long)c * N]);
  tbar /= (float)C;
  auto alpha_of = [&](int c) { return bhv[c] + coeff * (to_f32(tr[(long long)c * N]) - tbar); };
  auto logw_of = [&](int c) { return logf(fmaxf(alpha_of(c), 1e-8f)) + fmaf(s, to_f32(Ar[(long long)c * N]), -qq) + lbv[c] - lpv[c]; };
  float mx = kNegInf;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, logw_of(c));
  float Z = 0.f;
  for (int c = 0; c < C; ++c) Z += expf(logw_of(c) - mx);
  const float invZ = 1.0f / Z;
  float D = 0.f;
  for (int c = 0; c < C; ++c) D = fmaf(expf(logw_of(c) - mx) * invZ, to_f32(dWr[(long long)c * N]), D);
  // dlw, d alpha; W replaces A_raw only after the last use of A_raw for this landmark
  float sum_da = 0.f;
  for (int c = 0; c < C; ++c) {
    const float w = expf(logw_of(c) - mx) * invZ;
    const float dl = w * (to_f32(dWr[(long long)c * N]) - D);
    const float al = alpha_of(c);
    sum_da += al > 1e-8f ? dl / al : 0.f;
    if (live) { dlw_o[(long long)c * N] = from_f32<T>(dl); Ar[(long long)c * N] = from_f32<T>(w); }
  }
  const float mean_da = sum_da / (float)C;
  for (int c = 0; c < C; ++c) {
    const float dl = to_f32(dlw_o[(long long)c * N]);     // this thread's own store
    const float al = alpha_of(c);
    const float da = al > 1e-8f ? dl / al : 0.f;
    const float dt = coeff * (da - mean_da);
    if (live) dt_o[(long long)c * N] = from_f32<T>(dt);
    // per-landmark sums over the tokens of this block
    float a0 = live ? dl : 0.f, a1 = live ? da : 0.f, a2 = live ? to_f32(tr[(long long)c * N]) * dt : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; red[2][warp] = a2; }
    __syncthreads();
    if (threadIdx.x < 3) {
      const float v = red[threadIdx.x][0] + red[threadIdx.x][1] + red[threadIdx.x][2] + red[threadIdx.x][3];
      float* dst = threadIdx.x == 0 ? dlseB : (threadIdx.x == 1 ? dbh : R);
      atomicAdd(dst + item * C + c, v);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
lara_bwd_row_affine_kernel(T* __restrict__ Xm, long long x_item_stride, const T* __restrict__ Ym, long long y_item_stride,
                           const float* __restrict__ a, const float* __restrict__ b, int C, int N, long long rows_total) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows_total) return;
  const long long item = row / C;
  const int c = (int)(row % C);
  T* x = Xm + item * x_item_stride + (long long)c * N;
  const T* y = Ym + item * y_item_stride + (long long)c * N;
  const float off = (b ? b[row] : 0.f) - a[row];
  for (int n = lane; n < N; n += 32) x[n] = from_f32<T>(to_f32(y[n]) * (to_f32(x[n]) + off));
}

template <typename T>
static cudaError_t lara_bwd_launch(int which, void* X, void* Bm, const void* dW, void* M2, const float* v0, const float* v1, const float* v2,
                                   const float* v3, float* o0, float* o1, float* o2, long long xs, long long ys, int BH, int C, int N,
                                   float s, float coeff, cudaStream_t st) {
  const long long rows = (long long)BH * C;
  if (which == 0) {
    lara_bwd_rows1_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(reinterpret_cast<T*>(X), reinterpret_cast<T*>(Bm), v0, o0, o1, C, N, s, rows);
  } else if (which == 1) {
    dim3 grid((N + 127) / 128, BH);
    lara_bwd_cols_kernel<T><<<grid, 128, 0, st>>>(reinterpret_cast<T*>(X), reinterpret_cast<const T*>(dW), reinterpret_cast<T*>(M2), v0, v1, v2, v3,
                                                  o0, o1, o2, C, N, s, coeff);
  } else {
    lara_bwd_row_affine_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(reinterpret_cast<T*>(X), xs, reinterpret_cast<const T*>(Bm), ys, v0, v1, C,
                                                                                N, rows);
  }
  return cudaGetLastError();
}

cudaError_t launch_lara_bwd(int which, int io_dtype, void* X, void* Bm, const void* dW, void* M2, const float* v0, const float* v1,
                            const float* v2, const float* v3, float* o0, float* o1, float* o2, long long xs, long long ys, int BH, int C,
                            int N, float s, float coeff, cudaStream_t st) {
  if (io_dtype == EVA_F32) return lara_bwd_launch<float>(which, X, Bm, dW, M2, v0, v1, v2, v3, o0, o1, o2, xs, ys, BH, C, N, s, coeff, st);
  if (io_dtype == EVA_F16) return lara_bwd_launch<__half>(which, X, Bm, dW, M2, v0, v1, v2, v3, o0, o1, o2, xs, ys, BH, C, N, s, coeff, st);
  return lara_bwd_launch<__nv_bfloat16>(which, X, Bm, dW, M2, v0, v1, v2, v3, o0, o1, o2, xs, ys, BH, C, N, s, coeff, st);
}

}  // namespace eva
