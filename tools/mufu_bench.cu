// Microbenchmark: MUFU (ex2 / lg2 / rcp / tanh) issue rate per SM sub-partition on B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__device__ __forceinline__ float f(float x) {
  float y;
  if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 1) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 3) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 4) asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = 1.0f + threadIdx.x * 1e-3f + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = f<OP>(a[i]);
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<OP><<<148, threads>>>(out, cyc, iters);
  k<OP><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double warp_instr_per_smsp = (double)iters * 8 * (threads / 32) / 4.0;
  printf("%-6s %4d threads/SM: %.0f cycles, %.2f cycles per warp-instruction per SMSP\n", name, threads, c, c / warp_instr_per_smsp);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int th : {128, 256, 512, 1024}) {
    run<0>("ex2", th); run<1>("lg2", th); run<2>("rcp", th); run<3>("tanh", th); run<4>("ffma", th);
  }
  return 0;
}
