#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(double *err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double x = 1.0 + 3.0 * (i + 0.5) / (gridDim.x * blockDim.x);   // [1, 4): one octave pair
  double r, q;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(x));
  const double e1 = fabs(r * sqrt(x) - 1.0), e2 = fabs(q * x - 1.0);
  // block max
  __shared__ double m1[256], m2[256];
  m1[threadIdx.x] = e1; m2[threadIdx.x] = e2;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) { m1[threadIdx.x] = fmax(m1[threadIdx.x], m1[threadIdx.x + s]); m2[threadIdx.x] = fmax(m2[threadIdx.x], m2[threadIdx.x + s]); } __syncthreads(); }
  if (threadIdx.x == 0) { err[2 * blockIdx.x] = m1[0]; err[2 * blockIdx.x + 1] = m2[0]; }
}
int main() {
  const int nb = 4096;
  double *d; cudaMalloc(&d, nb * 16);
  k<<<nb, 256>>>(d);
  static double h[2 * nb];
  cudaMemcpy(h, d, nb * 16, cudaMemcpyDeviceToHost);
  double a = 0, b = 0;
  for (int i = 0; i < nb; i++) { a = fmax(a, h[2 * i]); b = fmax(b, h[2 * i + 1]); }
  printf("max rel err: rsqrt.approx.f64 %.3e (2^%.1f)   rcp.approx.f64 %.3e (2^%.1f)\n", a, log2(a), b, log2(b));
  return 0;
}
