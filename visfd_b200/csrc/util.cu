// util.cu -- benchmark bookkeeping: the exact (receiver, voter) pair count behind the
// voting roofline and a register-resident FFMA microbenchmark that measures the FP32
// CUDA-core peak of the device the numbers are quoted against.
#include "common.cuh"
#include "kernels.cuh"
#include <cmath>

namespace visfd_cuda {

struct PairArgs {
  const float *sal, *mask_src, *mask_dst;
  float thr;
  int nx, ny, nz, hw;
  int rz0, rz1;                       // receiver planes
  unsigned long long interior_count;  // lattice points with r^2 <= hw^2
};

__global__ void __launch_bounds__(256) pair_count_kernel(PairArgs a, unsigned long long *out) {
  const int ix = blockIdx.x * 64 + (threadIdx.x & 63);
  const int iy = blockIdx.y * 4 + (threadIdx.x >> 6);
  const int iz = blockIdx.z;
  unsigned long long c = 0;
  if (ix < a.nx && iy < a.ny) {
    const size_t i = ((size_t)iz * a.ny + iy) * a.nx + ix;
    float s = __ldg(a.sal + i);
    bool voter = (s >= a.thr) && s != 0.0f && !(a.mask_src && __ldg(a.mask_src + i) == 0.0f);
    if (voter) {
      const int hw = a.hw;
      bool interior = ix >= hw && ix + hw < a.nx && iy >= hw && iy + hw < a.ny && iz - hw >= a.rz0 && iz + hw < a.rz1;
      if (interior && !a.mask_dst) {
        c = a.interior_count;
      } else {
        for (int dz = -hw; dz <= hw; dz++) {
          int z = iz + dz;
          if (z < a.rz0 || z >= a.rz1) continue;
          for (int dy = -hw; dy <= hw; dy++) {
            int y = iy + dy;
            if (y < 0 || y >= a.ny) continue;
            int rem = hw * hw - dz * dz - dy * dy;
            if (rem < 0) continue;
            int w = (int)sqrtf((float)rem);
            while ((w + 1) * (w + 1) <= rem) w++;
            while (w * w > rem) w--;
            int x0 = max(0, ix - w), x1 = min(a.nx - 1, ix + w);
            if (!a.mask_dst) {
              c += (unsigned long long)(x1 - x0 + 1);
            } else {
              const float *m = a.mask_dst + ((size_t)z * a.ny + y) * a.nx;
              for (int x = x0; x <= x1; x++) c += (__ldg(m + x) != 0.0f);
            }
          }
        }
      }
    }
  }
  for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

i64 tv_count_pairs_device(visfd_ctx *ctx, i64 nx, i64 ny, i64 nz, const float *sal, float thr,
                          const float *mask_src, const float *mask_dst, int hw, int recv_z0, int recv_z1) {
  VREQUIRE(nx > 0 && ny > 0 && nz > 0 && nz <= 65535 && hw >= 0, "bad arguments to the pair count");
  unsigned long long V = 0;
  for (int dz = -hw; dz <= hw; dz++)
    for (int dy = -hw; dy <= hw; dy++)
      for (int dx = -hw; dx <= hw; dx++) V += (dx * dx + dy * dy + dz * dz <= hw * hw);
  Scratch<unsigned long long> d(ctx, 1);
  VCK(cudaMemsetAsync(d.get(), 0, sizeof(unsigned long long), ctx->stream));
  PairArgs a{sal, mask_src, mask_dst, thr, (int)nx, (int)ny, (int)nz, hw, recv_z0, recv_z1, V};
  dim3 grid(div_up(nx, 64), div_up(ny, 4), (unsigned)nz);
  pair_count_kernel<<<grid, 256, 0, ctx->stream>>>(a, d.get());
  VCK(cudaGetLastError());
  ctx->count_launch();
  unsigned long long h = 0;
  VCK(cudaMemcpyAsync(&h, d.get(), sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  VCK(cudaStreamSynchronize(ctx->stream));
  return (i64)h;
}

// 16 independent FFMA chains per thread; 2 FLOP per FFMA.
constexpr int PEAK_CHAINS = 16, PEAK_ITERS = 4096;
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *out, float a, float b) {
  float v[PEAK_CHAINS];
#pragma unroll
  for (int k = 0; k < PEAK_CHAINS; k++) v[k] = (float)(threadIdx.x + k);
  for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) v[k] = fmaf(v[k], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < PEAK_CHAINS; k++) s += v[k];
  if (s == 12345.678f) out[0] = s;  // keep the chains alive
}

// the same with packed FFMA2 (fma.rn.f32x2, Blackwell): 2 FMAs per lane per instruction
__global__ void __launch_bounds__(256) fp32x2_peak_kernel(float *out, float a, float b) {
  float2 v[PEAK_CHAINS / 2];
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
  for (int k = 0; k < PEAK_CHAINS / 2; k++) v[k] = make_float2((float)(threadIdx.x + k), (float)k);
  for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS / 2; k++) v[k] = __ffma2_rn(v[k], a2, b2);
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < PEAK_CHAINS / 2; k++) s += v[k].x + v[k].y;
  if (s == 12345.678f) out[0] = s;
}

// FFMA with three DISTINCT register operands (acc += x*y, the shape of a real
// accumulation): the register file delivers two 32-bit operands per lane per clock, so
// this form runs at ~2/3 of the chain peak above -- the practical ceiling of any kernel
// whose FMAs read three registers.
__global__ void __launch_bounds__(256) fp32_3op_peak_kernel(float *out, float a, float b) {
  float v[PEAK_CHAINS], x[4], y[4];
#pragma unroll
  for (int k = 0; k < PEAK_CHAINS; k++) v[k] = (float)(threadIdx.x + k);
#pragma unroll
  for (int k = 0; k < 4; k++) { x[k] = a + k * 1e-3f * threadIdx.x; y[k] = b - k * 1e-3f * threadIdx.x; }
  for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; k++) v[k] = fmaf(x[k & 3], y[(k >> 2) & 3], v[k]);
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < PEAK_CHAINS; k++) s += v[k];
  if (s == 12345.678f) out[0] = s;
}

double fp32_peak_device(visfd_ctx *ctx, double ms_target, int mode) {
  Scratch<float> d(ctx, 1);
  const int grid = ctx->sm_count * 8;
  cudaEvent_t e0, e1;
  VCK(cudaEventCreate(&e0));
  VCK(cudaEventCreate(&e1));
  // warm-up, then as many launches as fit the target time
  auto launch = [&]() {
    if (mode == 1) fp32x2_peak_kernel<<<grid, 256, 0, ctx->stream>>>(d.get(), 0.999f, 0.001f);
    else if (mode == 2) fp32_3op_peak_kernel<<<grid, 256, 0, ctx->stream>>>(d.get(), 0.999f, 0.001f);
    else fp32_peak_kernel<<<grid, 256, 0, ctx->stream>>>(d.get(), 0.999f, 0.001f);
  };
  for (int k = 0; k < 3; k++) launch();
  VCK(cudaStreamSynchronize(ctx->stream));
  const double flop_per_launch = 2.0 * PEAK_CHAINS * (double)PEAK_ITERS * 256.0 * grid;
  int launches = std::max(1, (int)(ms_target * 1e-3 * 60e12 / flop_per_launch));
  VCK(cudaEventRecord(e0, ctx->stream));
  for (int k = 0; k < launches; k++) launch();
  VCK(cudaEventRecord(e1, ctx->stream));
  VCK(cudaStreamSynchronize(ctx->stream));
  VCK(cudaGetLastError());
  float ms = 0.0f;
  VCK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ctx->count_launch(launches + 3);
  return flop_per_launch * launches / (ms * 1e-3) / 1e12;
}

}  // namespace visfd_cuda
