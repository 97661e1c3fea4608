// Measures the issue rate of the MUFU ops the conv epilogues use (dev tool): tanh.approx / ex2.approx / rcp.approx per clock per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_rate tools/micro/mufu_rate.cu && /tmp/mufu_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ float f(float x) {
  float y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 3) {  // sigmoid by ex2 + rcp
    float e;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(1.0f + e));
  }
  if (OP == 4) {  // sigmoid by tanh
    float t;
    asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    y = fmaf(0.5f, t, 0.5f);
  }
  if (OP == 5) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(*reinterpret_cast<unsigned*>(&y)) : "r"(__float_as_uint(x)));
  return y;
}

template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = f<OP>(a[i]);
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int threads) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<OP><<<148, threads>>>(out, iters, cyc);
  k<OP><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-18s %4d threads/SM: %.2f results/clk/SM\n", name, threads, (double)iters * 8 * threads / h);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int th : {128, 256, 512, 1024}) {
    run<0>("tanh.approx.f32", th);
    run<1>("ex2.approx.f32", th);
    run<2>("rcp.approx.f32", th);
    run<3>("sigmoid ex2+rcp", th);
    run<4>("sigmoid tanh", th);
    run<5>("tanh.approx.f16x2", th);
  }
  return 0;
}
