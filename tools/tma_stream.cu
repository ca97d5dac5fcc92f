// Micro-benchmark (development tool): what HBM bandwidth can TMA deliver for the fused kernel's access pattern?
// Every CTA streams the q/k/v tiles of (batch, head) items of a packed qkv tensor [B, N, 3, H, 64] fp16 into a ring of
// D shared-memory slots and drops them (no compute).  Modes:
//   0  chunk-row boxes (64, 1, 28, 4, 1): 112 fragments of 128 B, token stride 1152 B   (what pass 1 / pass 2 issue)
//   1  window boxes    (64, 1, 7, 7, 1):  49 fragments of 128 B                          (what phase B issues)
//   2  flat bulk copies of 14336 contiguous bytes                                        (ideal streaming reference)
//   3  chunk-row boxes over all three heads (192, 1, 28, 4): 112 fragments of 384 B      (one CTA per image)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_stream tools/tma_stream.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../efficient-attention_b200/csrc/sm100_ptx.cuh"
using namespace eva;

constexpr int kSlot = 16384 * 3;   // mode 3 needs 43 KB tiles

struct P {
  int items, H, mode, D, tile_bytes, wait_mode, warps, reps, sub;   // sub: 0 row boxes, 1 window boxes (mode 6)
  const uint8_t* flat;
  long long flat_bytes;
};

__global__ void __launch_bounds__(256, 1) stream_kernel(const __grid_constant__ CUtensorMap t_row, const __grid_constant__ CUtensorMap t_win,
                                                        const __grid_constant__ CUtensorMap t_row3, const __grid_constant__ CUtensorMap t_row4d,
                                                        const __grid_constant__ CUtensorMap t_half, const __grid_constant__ CUtensorMap t_win2,
                                                        const P p, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* const sm = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bars[128];
  const int wid = threadIdx.x >> 5;
  if (wid >= (p.mode == 6 ? p.warps : 1)) return;
  if (p.mode == 6 && p.sub == 2 && p.reps > 1 && blockIdx.x >= 49) return;
  const uint32_t bar0 = ptx::smem_u32(bars) + 128 * wid;
  const int slot_bytes = (p.mode == 3 || (p.mode == 6 && p.sub == 2)) ? kSlot : 16384;
  if ((threadIdx.x & 31) == 0) {
    for (int i = 0; i < p.D; ++i) ptx::mbar_init(bar0 + 8 * i, 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();
  uint32_t n_issue = 0, n_wait = 0;
  long long t_wait = 0, t_issue = 0;
  const long long t0 = clock64();
  auto slot = [&](uint32_t n) { return ptx::smem_u32(sm) + ((n % p.D) + wid * p.D) * slot_bytes; };
  auto wait = [&](uint32_t bar, uint32_t parity) {
    if (p.wait_mode == 0) {
      ptx::mbar_wait(bar, parity);                       // 2 x try_wait, then try_wait with a 2000 ns suspend hint
    } else if (p.wait_mode == 1) {
      while (!ptx::mbar_try_wait(bar, parity)) {}        // try_wait, default time limit
    } else if (p.wait_mode == 2) {
      uint32_t ok = 0;                                   // test_wait: never suspends
      do {
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.test_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
      } while (!ok);
    } else {
      while (!ptx::mbar_try_wait_suspend(bar, parity, 20)) {}   // short suspend hint
    }
  };
  auto acquire = [&]() {
    if (n_issue - n_wait == (uint32_t)p.D) {
      wait(bar0 + 8 * (n_wait % p.D), (n_wait / p.D) & 1);
      ++n_wait;
    }
    if (ptx::elect_one()) ptx::mbar_arrive_expect_tx(bar0 + 8 * (n_issue % p.D), p.tile_bytes);
  };
  for (int rep = 0; rep < p.reps; ++rep)
  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    const int b = p.mode == 3 ? item : item / p.H, h = p.mode == 3 ? 0 : item % p.H;
    if (p.mode == 4) {
      for (int t = 0; t < 3; ++t)
        for (int r = 0; r < 7; ++r) {
          acquire();
          if (ptx::elect_one())
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(slot(n_issue)), "l"(reinterpret_cast<uint64_t>(&t_row4d)), "r"(bar0 + 8 * (n_issue % p.D)), "r"(0), "r"(t * p.H + h), "r"(112 * r), "r"(b) : "memory");
          ++n_issue;
        }
    } else if (p.mode == 5) {
      for (int t = 0; t < 3; ++t)
        for (int r = 0; r < 14; ++r) {
          acquire();
          if (ptx::elect_one()) ptx::tma_load_5d(slot(n_issue), &t_half, bar0 + 8 * (n_issue % p.D), 0, t * p.H + h, 0, 2 * r, b);
          ++n_issue;
        }
    } else if (p.mode == 7) {
      for (int t = 0; t < 3; ++t)
        for (int w = 0; w < 8; ++w) {
          acquire();
          if (ptx::elect_one()) ptx::tma_load_5d(slot(n_issue), &t_win2, bar0 + 8 * (n_issue % p.D), 0, t * p.H + h, 7 * (w & 3), 14 * (w >> 2), b);
          ++n_issue;
        }
    } else if (p.mode == 6 && p.sub == 1) {
      for (int t = 0; t < 3; ++t)
        for (int w = 0; w < 16; ++w) {
          if ((t * 16 + w) % p.warps != wid) continue;
          acquire();
          if (ptx::elect_one()) ptx::tma_load_5d(slot(n_issue), &t_win, bar0 + 8 * (n_issue % p.D), 0, t * p.H + h, 7 * (w & 3), 7 * (w >> 2), b);
          ++n_issue;
        }
    } else if (p.mode == 6 && p.sub == 3) {
      const int lane = threadIdx.x & 31;
      for (int t = 0; t < 3; ++t)
        for (int r = 0; r < 7; ++r) {
          if ((t * 7 + r) % p.warps != wid) continue;
          if (n_issue - n_wait == (uint32_t)p.D) {
            wait(bar0 + 8 * (n_wait % p.D), (n_wait / p.D) & 1);
            ++n_wait;
          }
          // tile = 112 tokens x 128 B; token stride 3*H*128 B; lane -> (row = 4 j + lane / 8, 16-byte chunk = lane % 8)
          const uint8_t* src = p.flat + ((long long)(b * 784 + 112 * r) * 3 * p.H + (t * p.H + h)) * 128 + (long long)(lane >> 3) * 3 * p.H * 128 + (lane & 7) * 16;
          const uint32_t dst0 = slot(n_issue) + (lane >> 3) * 128;
          const uint32_t sw0 = (uint32_t)(((lane & 7) ^ (lane >> 3)) << 4), sw1 = (uint32_t)(((lane & 7) ^ (4 + (lane >> 3))) << 4);
#pragma unroll 4
          for (int j = 0; j < 28; ++j) {
            const uint32_t dst = dst0 + j * 512 + ((j & 1) ? sw1 : sw0);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            src += (long long)4 * 3 * p.H * 128;
          }
          asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];" ::"r"(bar0 + 8 * (n_issue % p.D)) : "memory");
          __syncwarp();
          if (ptx::elect_one()) ptx::mbar_arrive(bar0 + 8 * (n_issue % p.D));
          ++n_issue;
        }
    } else if (p.mode == 6 && (p.sub == 4 || p.sub == 5)) {      // stacked window pairs (7 x 14 tokens), aligned / 15 rows into the slot
      for (int t = 0; t < 3; ++t)
        for (int w = 0; w < 8; ++w) {
          if ((t * 8 + w) % p.warps != wid) continue;
          acquire();
          if (ptx::elect_one()) ptx::tma_load_5d(slot(n_issue) + (p.sub == 5 ? 15 * 128 : 0), &t_win2, bar0 + 8 * (n_issue % p.D), 0, t * p.H + h, 7 * (w & 3), 14 * (w >> 2), b);
          ++n_issue;
        }
    } else if (p.mode == 6 && p.sub == 2) {
      if (h == 0)
        for (int t = 0; t < 3; ++t)
          for (int r = 0; r < 7; ++r) {
            if ((t * 7 + r) % p.warps != wid) continue;
            acquire();
            if (ptx::elect_one()) ptx::tma_load_5d(slot(n_issue), &t_row3, bar0 + 8 * (n_issue % p.D), 0, t, 0, 4 * r, b);
            ++n_issue;
          }
    } else if (p.mode == 0 || p.mode == 6) {
      for (int t = 0; t < 3; ++t)            // q rows, k rows, v rows (the tensor map covers [.., 3*H, ..] heads)
        for (int r = 0; r < 7; ++r) {
          if (p.mode == 6 && (t * 7 + r) % p.warps != wid) continue;
          const long long ta = clock64();
          acquire();
          const long long tb = clock64();
          if (ptx::elect_one()) ptx::tma_load_5d(slot(n_issue), &t_row, bar0 + 8 * (n_issue % p.D), 0, t * p.H + h, 0, 4 * r, b);
          const long long tc = clock64();
          t_wait += tb - ta; t_issue += tc - tb;
          ++n_issue;
        }
    } else if (p.mode == 1) {
      for (int t = 0; t < 3; ++t)
        for (int w = 0; w < 16; ++w) {
          acquire();
          if (ptx::elect_one()) ptx::tma_load_5d(slot(n_issue), &t_win, bar0 + 8 * (n_issue % p.D), 0, t * p.H + h, 7 * (w & 3), 7 * (w >> 2), b);
          ++n_issue;
        }
    } else if (p.mode == 2) {
      for (int t = 0; t < 21; ++t) {
        acquire();
        const long long off = (((long long)item * 21 + t) * 14336) % (p.flat_bytes - 14336);
        if (ptx::elect_one()) ptx::bulk_load(slot(n_issue), p.flat + (off & ~15LL), 14336, bar0 + 8 * (n_issue % p.D));
        ++n_issue;
      }
    } else {
      if (h == 0)
        for (int t = 0; t < 3; ++t)
          for (int r = 0; r < 7; ++r) {
            acquire();
            if (ptx::elect_one()) ptx::tma_load_5d(slot(n_issue), &t_row3, bar0 + 8 * (n_issue % p.D), 0, t, 0, 4 * r, b);
            ++n_issue;
          }
    }
  }
  while (n_wait < n_issue) {
    wait(bar0 + 8 * (n_wait % p.D), (n_wait / p.D) & 1);
    ++n_wait;
  }
  if (threadIdx.x == 0) { cycles[blockIdx.x] = clock64() - t0; cycles[2048 + blockIdx.x] = t_wait; cycles[3072 + blockIdx.x] = t_issue; cycles[1024 + blockIdx.x] = n_issue; }
  if (p.mode == 8) cycles[1024 + blockIdx.x] = (unsigned long long)(&t_row4d) + (unsigned long long)(&t_half) + (unsigned long long)(&t_win2);
}

static PFN_cuTensorMapEncodeTiled_v12000 enc() {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 1024, H = 3, N = 784;
  const size_t bytes = (size_t)B * N * 3 * H * 64 * 2;
  uint8_t* qkv;
  cudaMalloc(&qkv, bytes);
  cudaMemset(qkv, 0, bytes);
  unsigned long long* cyc;
  cudaMalloc(&cyc, 4096 * 8);
  CUtensorMap t_row, t_win, t_row3, t_row4d, t_half, t_win2;
  {  // dims (64, 3H, 28, 28, B): head index t*H + h selects q/k/v and the head
    const cuuint64_t dims[5] = {64, (cuuint64_t)(3 * H), 28, 28, (cuuint64_t)B};
    const cuuint64_t str[4] = {128, (cuuint64_t)3 * H * 128, (cuuint64_t)3 * H * 128 * 28, (cuuint64_t)3 * H * 128 * N};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    const cuuint32_t box_r[5] = {64, 1, 28, 4, 1}, box_w[5] = {64, 1, 7, 7, 1};
    CUresult r1 = enc()(&t_row, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, qkv, dims, str, box_r, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc()(&t_win, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, qkv, dims, str, box_w, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    // all heads of q (or k, v) in one box: dims (192, 3, 28, 28, B), no swizzle (inner box 384 B)
    const cuuint64_t dims3[5] = {192, 3, 28, 28, (cuuint64_t)B};
    const cuuint64_t str3[4] = {384, (cuuint64_t)3 * H * 128, (cuuint64_t)3 * H * 128 * 28, (cuuint64_t)3 * H * 128 * N};
    const cuuint32_t box3[5] = {192, 1, 28, 4, 1};
    CUresult r3 = enc()(&t_row3, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, qkv, dims3, str3, box3, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const cuuint32_t box_h[5] = {64, 1, 28, 2, 1}, box_w2[5] = {64, 1, 7, 14, 1};
    CUresult r4 = enc()(&t_half, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, qkv, dims, str, box_h, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r5 = enc()(&t_win2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, qkv, dims, str, box_w2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const cuuint64_t dims4[4] = {64, (cuuint64_t)(3 * H), (cuuint64_t)N, (cuuint64_t)B};
    const cuuint64_t str4[3] = {128, (cuuint64_t)3 * H * 128, (cuuint64_t)3 * H * 128 * N};
    const cuuint32_t box4[4] = {64, 1, 112, 1};
    CUresult r6 = enc()(&t_row4d, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, qkv, dims4, str4, box4, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 || r2 || r3 || r4 || r5 || r6) { printf("encode failed %d %d %d %d %d %d\n", r1, r2, r3, r4, r5, r6); return 1; }
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int tile_bytes[8] = {14336, 6272, 14336, 43008, 14336, 7168, 14336, 12544};
  for (int Bsub : {49, B})            // 49 images = 147 items (one per CTA), 43 MB: L2-resident after the first pass
    for (int sub : {1, 4, 5})
      for (int warps : {1, 2})
        for (int per_sm = 1; per_sm <= (Bsub == B ? 2 : 1); ++per_sm)
          for (int D : {2, 4}) {
            const size_t smem = (size_t)D * (sub == 2 ? kSlot : 16384) * warps + 1024;
            if (smem * per_sm > 225 * 1024 || smem > 220 * 1024) continue;
            
            const int reps = Bsub == B ? 1 : 40;
            P p{Bsub * H, H, 6, D, sub == 2 ? 43008 : sub == 1 ? 6272 : sub >= 4 ? 12544 : 14336, 0, warps, reps, sub, qkv, (long long)bytes};
            const size_t smem_req = per_sm == 1 ? (smem > 116 * 1024 ? smem : 116 * 1024) : smem;
            float best = 1e9f;
            for (int it = 0; it < 3; ++it) {
              cudaEventRecord(e0);
              stream_kernel<<<sms * per_sm, 256, smem_req>>>(t_row, t_win, t_row3, t_row4d, t_half, t_win2, p, cyc);
              cudaEventRecord(e1);
              if (cudaEventSynchronize(e1) != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
              float ms;
              cudaEventElapsedTime(&ms, e0, e1);
              if (it > 0 && ms < best) best = ms;
            }
            unsigned long long hc[4096];
            cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
            if (false) printf("   CTA 0 warp 0: %llu ops, total %llu cyc, in acquire (wait) %llu cyc/op, in the TMA instruction %llu cyc/op\n", hc[1024], hc[0],
                                 hc[2048] / (hc[1024] ? hc[1024] : 1), hc[3072] / (hc[1024] ? hc[1024] : 1));
            const double moved = (double)Bsub * N * 3 * H * 128 * reps;
            const double active = Bsub == B ? sms * per_sm : (Bsub * (sub == 2 ? 1 : H) < sms ? Bsub * (sub == 2 ? 1 : H) : sms);
            const double ops = (double)Bsub * (sub == 2 ? 21 : sub == 1 ? 48 * H : sub >= 4 ? 24 * H : 21 * H) * reps;
            printf("%s %s  warps/CTA %d  CTAs/SM %d  depth %d: %7.1f us  %6.0f GB/s  %.2f ops/us per active SM (%.0f active)\n", Bsub == B ? "HBM" : "L2 ",
                   sub == 5 ? "stacked +15 rows" : sub == 4 ? "stacked aligned " : sub == 3 ? "cp.async rows" : sub == 2 ? "384B-row box" : sub ? "window boxes" : "row boxes   ", warps, per_sm, D, best * 1e3, moved / (best * 1e-3) / 1e9,
                   ops / (best * 1e3) / (Bsub == B ? sms : active), active);
            fflush(stdout);
          }
  return 0;
}
