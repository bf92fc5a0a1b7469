// FFMA vs FFMA2 (packed f32x2, sm_100) issue rate per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, long long* cyc, int iters, float b, float c) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(1.0f + threadIdx.x * 1e-3f + i, 0.5f + i);
  const float2 b2 = make_float2(b, b * 1.01f), c2 = make_float2(c, c * 0.99f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) { a[i].x = fmaf(a[i].x, b2.x, c2.x); a[i].y = fmaf(a[i].y, b2.y, c2.y); }
      else a[i] = __ffma2_rn(a[i], b2, c2);
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  k<OP><<<148, threads>>>(out, cyc, iters, 0.999f, 0.001f);
  k<OP><<<148, threads>>>(out, cyc, iters, 0.999f, 0.001f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("%-6s %4d threads/SM: %.0f cycles for %d x 16 FMAs per thread -> %.2f FMA lanes per clock per SM\n", name, threads, c, iters,
         (double)iters * 16 * threads / c);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int th : {128, 256, 512, 1024}) { run<0>("ffma", th); run<1>("ffma2", th); }
  return 0;
}
