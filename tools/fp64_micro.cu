// FP64 pipe micro-benchmark (tuning aid, not product code): cycles per DFMA per warp as a function of
// resident warps per SM and independent chains per thread.   nvcc -O3 -arch=sm_100a tools/fp64_micro.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void __launch_bounds__(32) k(double *out, int iters, long long *cyc) {
  double a[C];
  const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * (blockIdx.x + 1);
#pragma unroll
  for (int c = 0; c < C; ++c) a[c] = 0.1 * (c + 1);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int c = 0; c < C; ++c) a[c] = fma(a[c], x, y);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += a[c];
  out[blockIdx.x * 32 + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int C>
void run(int wps, double *out, long long *cyc, long long *h) {
  const int blocks = 148 * wps, iters = 2000;
  k<C><<<blocks, 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < blocks; ++i) m += (double)h[i];
  m /= blocks;
  printf("warps/SM %2d chains %d: %.2f cycles per DFMA per warp, %.3f DFMA/cycle/SM\n", wps, C, m / (iters * 16.0 * C),
         wps * iters * 16.0 * C / m);
}
int main() {
  double *out; long long *cyc; long long *h = new long long[148 * 64];
  cudaMalloc(&out, 148 * 64 * 32 * 8); cudaMalloc(&cyc, 148 * 64 * 8);
  for (int wps : {1, 2, 4, 8, 16, 32}) { run<1>(wps, out, cyc, h); run<2>(wps, out, cyc, h); run<4>(wps, out, cyc, h); run<8>(wps, out, cyc, h); }
  return 0;
}
