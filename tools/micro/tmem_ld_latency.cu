// Micro-benchmark (dev tool): latency of tcgen05.ld 32x32b.{x16,x32} + wait::ld seen by epilogue warps, with the tensor pipe idle and
// with one thread issuing tcgen05.mma back to back into OTHER TMEM columns (the conv kernel's situation: the epilogue of tile i reads
// while the MMAs of tile i+1 run).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_ld_latency.bin tools/micro/tmem_ld_latency.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../yolov5_obb_b200/csrc/ptx.cuh"
using namespace y5obb;

// warps 0: MMA issuer; warps 4..4+n_ld-1: loaders (TMEM lane quarter = warp % 4)
__global__ void __launch_bounds__(384, 1) k(int N, int mma_on, int n_ld_warps, int x16, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
    stop = 0;
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
  if (warp == 0) {
    if (mma_on && ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, N);
      const uint64_t hi = ptx::make_kmajor_desc(0u, 128u);
      const uint64_t da0 = hi | (uint64_t)((ptx::smem_u32(smem) & 0x3FFFFu) >> 4);
      const uint64_t db0 = hi | (uint64_t)((ptx::smem_u32(smem + 64 * 1024) & 0x3FFFFu) >> 4);
      uint32_t ph = 0;
      while (!stop) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          ptx::umma_bf16(tmem, da0 + (uint64_t)((j >> 2) * 1024 + (j & 3) * 2), db0 + (uint64_t)((j & 3) * 2), idesc, 1u);
        ptx::umma_commit(&bar);
        ptx::mbar_wait(&bar, ph);
        ph ^= 1u;
      }
    }
  } else if (warp >= 4 && warp < 4 + n_ld_warps) {
    const uint32_t taddr = tmem + 256u + (uint32_t)(((warp - 4) >> 2) * 64) + ((uint32_t)((warp & 3) * 32) << 16);
    long long tot = 0, mx = 0;
    float acc = 0.f;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (x16) {
        uint32_t v[16];
        ptx::tmem_ld_32x32b_x16(taddr, v);
        ptx::tmem_ld_wait();
        acc += __uint_as_float(v[0]) + __uint_as_float(v[15]);
      } else {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr, v);
        ptx::tmem_ld_wait();
        acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);
      }
      const long long t1 = clock64();
      if (r >= 4) {
        tot += t1 - t0;
        mx = max(mx, t1 - t0);
      }
      __nanosleep(200);
    }
    if (blockIdx.x == 0 && warp == 4 && (threadIdx.x & 31) == 0) {
      out[0] = tot / (reps - 4);
      out[1] = mx;
      out[2] = (long long)acc;
    }
    __syncwarp();
    if (warp == 4 && (threadIdx.x & 31) == 0) stop = 1;
  }
  if (warp >= 4 + n_ld_warps || (warp > 0 && warp < 4)) {
    // idle warps
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
  cudaMalloc(&d, 32);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("tcgen05.ld 32x32b + wait::ld latency in cycles (mean / max over 60 loads), loader warps on their own TMEM columns\n");
  for (int x16 : {0, 1})
    for (int nld : {1, 4, 8})
      for (int mma : {0, 1})
        for (int N : {32, 64, 256}) {
          if (!mma && N != 32) continue;
          k<<<148, 384, 200 * 1024>>>(N, mma, nld, x16, 64, d);
          long long h[3];
          cudaError_t e = cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) {
            printf("error %s\n", cudaGetErrorString(e));
            return 1;
          }
          printf("x%-2d loader warps %d  mma %s  | mean %5lld  max %5lld\n", x16 ? 16 : 32, nld, mma ? (N == 32 ? "N=32 " : (N == 64 ? "N=64 " : "N=256")) : "off  ",
                 h[0], h[1]);
        }
  return 0;
}
