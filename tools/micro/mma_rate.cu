// Micro-benchmark (dev tool): issue cost and completion time of tcgen05.mma (M=128, N, K=16, bf16, both operands in shared
// memory, K-major, 128B swizzle) issued back to back by ONE elected thread, and by TWO warps at once (different accumulators).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate tools/micro/mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../yolov5_obb_b200/csrc/ptx.cuh"
using namespace y5obb;

__global__ void __launch_bounds__(384, 1) k(int N, int n_batches, int n_warps, int reps, long long* out, int mode = 0, uint32_t a_step = 0,
                                            uint32_t b_step = 0, int noise = 0, float* sink = nullptr) {
  __shared__ volatile int stop_noise;
  if (threadIdx.x == 0) stop_noise = 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ __align__(8) uint64_t bar2[8];
  __shared__ __align__(8) uint64_t bar3;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&bar2[i], 1);
    ptx::mbar_init(&bar3, 1);
    ptx::fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    ptx::tmem_alloc(&tmem_base_s, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (noise && warp >= 4) {  // epilogue-like instruction pressure on every scheduler: FFMA + MUFU.TANH chains until the issuer is done
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    while (!stop_noise) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a0 = fmaf(a0, 0.5f, a1);
        a1 = fmaf(a1, 0.25f, a2);
        asm volatile("tanh.approx.f32 %0, %1;" : "=f"(a2) : "f"(a3));
        a3 = fmaf(a3, 0.5f, a0);
      }
    }
    if (sink) sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  }
  if (warp < n_warps) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, N);
      const uint32_t a_u32 = ptx::smem_u32(smem) + (uint32_t)warp * 32768u, b_u32 = ptx::smem_u32(smem + 64 * 1024);
      const uint64_t hi = ptx::make_kmajor_desc(0u, 128u);
      const uint64_t da0 = hi | (uint64_t)((a_u32 & 0x3FFFFu) >> 4);
      const uint64_t db0 = hi | (uint64_t)((b_u32 & 0x3FFFFu) >> 4);
      const uint32_t d = tmem + (uint32_t)warp * 256u;
      uint32_t ph = 0;
      long long t_issue = 0, t_done = 0;
      for (int r = 0; r < reps; ++r) {
        const long long t0 = clock64();
        if (mode == 0) {
          for (int bt = 0; bt < n_batches; ++bt) {  // 8 MMAs per batch, compile-time operand offsets, no index arithmetic
#pragma unroll
            for (int j = 0; j < 8; ++j)
              ptx::umma_bf16(d, da0 + (uint64_t)((j >> 2) * 1024 + (j & 3) * 2), db0 + (uint64_t)((j & 3) * 2), idesc, (bt | j) ? 1u : 0u);
          }
        } else if (mode == 1) {  // the conv kernel's issue_unit<2, 4>: operand offsets from RUNTIME steps (64-bit descriptor adds)
          for (int bt = 0; bt < n_batches; ++bt) {
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                ptx::umma_bf16(d, da0 + (uint64_t)(u * a_step + 2 * j), db0 + (uint64_t)(u * b_step + 2 * j), idesc, (bt | u | j) ? 1u : 0u);
          }
        } else if (mode == 5) {  // 8 MMAs, then a tcgen05.commit to a (never waited) barrier, repeated: does a commit stall the MMAs behind it?
          for (int bt = 0; bt < n_batches; ++bt) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              ptx::umma_bf16(d, da0 + (uint64_t)((j >> 2) * 1024 + (j & 3) * 2), db0 + (uint64_t)((j & 3) * 2), idesc, (bt | j) ? 1u : 0u);
            ptx::umma_commit(&bar2[bt & 7]);
          }
        } else if (mode == 6) {  // ... and with a (successful) mbarrier.try_wait on another barrier before every batch
          for (int bt = 0; bt < n_batches; ++bt) {
            ptx::mbar_wait(&bar3, 1u);  // a fresh barrier: the wait for the preceding phase returns at once
#pragma unroll
            for (int j = 0; j < 8; ++j)
              ptx::umma_bf16(d, da0 + (uint64_t)((j >> 2) * 1024 + (j & 3) * 2), db0 + (uint64_t)((j & 3) * 2), idesc, (bt | j) ? 1u : 0u);
            ptx::umma_commit(&bar2[bt & 7]);
          }
        } else if (mode == 3) {  // descriptors that really change per batch (ring position): low word + register-pair move, as the conv kernel does
          const uint32_t hi32 = (uint32_t)(hi >> 32);
          uint32_t a_lo = (uint32_t)da0, b_lo = (uint32_t)db0;
          for (int bt = 0; bt < n_batches; ++bt) {
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint64_t da, db;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(da) : "r"(a_lo + u * a_step + 2 * j), "r"(hi32));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(db) : "r"(b_lo + u * b_step + 2 * j), "r"(hi32));
                ptx::umma_bf16(d, da, db, idesc, (bt | u | j) ? 1u : 0u);
              }
            a_lo ^= a_step;  // alternate between two ring positions: nothing can be hoisted
            b_lo ^= a_step;
          }
        } else if (mode == 4) {  // the same, 64-bit descriptors advanced incrementally (one add per descriptor per MMA)
          uint64_t da = da0, db = db0;
          for (int bt = 0; bt < n_batches; ++bt) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              ptx::umma_bf16(d, da, db, idesc, (bt | j) ? 1u : 0u);
              da += (j == 3) ? (uint64_t)(a_step - 6) : ((j == 7) ? (uint64_t)0 : 2ull);
              db += (j == 3) ? (uint64_t)(b_step - 6) : ((j == 7) ? (uint64_t)0 : 2ull);
            }
            da = da0 ^ (uint64_t)((bt & 1) ? 0 : a_step);
            db = db0 ^ (uint64_t)((bt & 1) ? 0 : a_step);
          }
        } else {  // same offsets, the low descriptor word advanced by 32-bit adds, high word constant
          const uint32_t hi32 = (uint32_t)(hi >> 32), a_lo = (uint32_t)da0, b_lo = (uint32_t)db0;
          for (int bt = 0; bt < n_batches; ++bt) {
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint64_t da, db;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(da) : "r"(a_lo + u * a_step + 2 * j), "r"(hi32));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(db) : "r"(b_lo + u * b_step + 2 * j), "r"(hi32));
                ptx::umma_bf16(d, da, db, idesc, (bt | u | j) ? 1u : 0u);
              }
          }
        }
        const long long t1 = clock64();
        ptx::umma_commit(&bar[warp]);
        ptx::mbar_wait(&bar[warp], ph);
        ph ^= 1u;
        const long long t2 = clock64();
        if (r > 0) {
          t_issue += t1 - t0;
          t_done += t2 - t0;
        }
      }
      if (blockIdx.x == 0 && warp == 0) {
        out[0] = t_issue / (reps - 1);
        out[1] = t_done / (reps - 1);
      }
      if (warp == 0) stop_noise = 1;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("  N  n_mma warps | issue cycles (per mma) | until complete (per mma)   [tensor floor per mma = N/2 cycles]\n");
  const int only_mode = getenv("MMA_RATE_MODES") ? 1 : 0;   // set: compare the three descriptor-arithmetic modes (1 warp)
  if (getenv("MMA_RATE_NOISE")) {  // issue rate with 8 busy warps (2 per scheduler) beside the issuer(s), as in the conv kernel
    float* sink;
    cudaMalloc(&sink, 148 * 384 * 4);
    for (int noise : {0, 1})
      for (int nw : {1, 2})
        for (int N : {32, 64, 128})
          for (int nb : {2, 8}) {
            k<<<148, 384, 200 * 1024>>>(N, nb, nw, 6, d, 0, 0u, 0u, noise, sink);
            long long h[2];
            if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
            printf("busy warps %d issuers %d N %3d n_mma %3d | issue %6lld (%5.1f per mma) | complete %6lld (%5.1f)\n", noise ? 8 : 0, nw, N, nb * 8, h[0],
                   (double)h[0] / (nb * 8), h[1], (double)h[1] / (nb * 8));
          }
    return 0;
  }
  if (only_mode) {
    for (int mode : {0, 3, 4, 5, 6})
      for (int N : {32, 64, 256})
        for (int nb : {2, 8}) {
          k<<<148, 128, 200 * 1024>>>(N, nb, 1, 6, d, mode, 64u, 0u);
          long long h[2];
          if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
          printf("mode %d N %3d n_mma %3d | issue %6lld (%5.1f per mma) | complete %6lld (%5.1f)\n", mode, N, nb * 8, h[0], (double)h[0] / (nb * 8), h[1],
                 (double)h[1] / (nb * 8));
        }
    return 0;
  }
  for (int N : {16, 32, 64, 128, 256})
    for (int nb : {1, 2, 8, 32})
      for (int nw : {1, 2}) {
        k<<<148, 128, 200 * 1024>>>(N, nb, nw, 6, d);
        long long h[2];
        cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
          printf("error %s\n", cudaGetErrorString(e));
          return 1;
        }
        printf("%4d %5d %4d | %7lld (%6.1f) | %7lld (%6.1f)\n", N, nb * 8, nw, h[0], (double)h[0] / (nb * 8), h[1], (double)h[1] / (nb * 8));
      }
  return 0;
}
