// Per-(batch, head) float32 workspace shared by the LARA kernels:
//   qbar[C,D] mu[C,D] omega[S,D] kv[S,D] lp[S] bh[S] lse_k[S] lse_t[C]
#pragma once
#include <stddef.h>

namespace eva {

struct LaraWs {
  float *qbar, *mu, *omega, *kv, *lp, *bh, *lse_k, *lse_t;
};

__host__ __device__ inline size_t lara_ws_floats_per_bh(int C, int S, int D) {
  return (size_t)(2 * C + 2 * S) * D + 3 * (size_t)S + C;
}
__host__ __device__ inline LaraWs lara_ws_at(float* base, long long bh, int C, int S, int D) {
  float* p = base + bh * (long long)lara_ws_floats_per_bh(C, S, D);
  LaraWs w;
  w.qbar = p; p += (size_t)C * D;
  w.mu = p; p += (size_t)C * D;
  w.omega = p; p += (size_t)S * D;
  w.kv = p; p += (size_t)S * D;
  w.lp = p; p += S;
  w.bh = p; p += S;
  w.lse_k = p; p += S;
  w.lse_t = p;
  return w;
}

}  // namespace eva
