// threshold.cu -- the per-voxel maps downstream of the detectors: HandleThresholds
// (bin/filter_mrc/handlers.cpp:1037-1080) fused with the mask fill of
// bin/filter_mrc/filter_mrc.cpp:771-776, and the mean / standard deviation needed by
// "-cl" (AverageArr / StdDevArr, lib/visfd/visfd_utils.hpp:685-790).
// One streaming pass: 4 B read + 4 B written per voxel (+4 B with a mask).
#include "common.cuh"
#include "kernels.cuh"

namespace visfd_cuda {

struct ThreshParams {
  int kind;
  float t0, t1, t2, t3, outA, outB;
  int use_masked_value;
  float masked_value;
};

// lib/threshold/threshold.hpp:10-12
__device__ __forceinline__ bool is_between(float x, float a, float b) {
  return ((a <= x) && (x < b)) || ((b < x) && (x <= a));
}
// lib/threshold/threshold.hpp:52-77
__device__ __forceinline__ float thresh2(float x, float a, float b, float outA, float outB) {
  float g;
  if (is_between(x, a, b))
    g = __fdiv_rn(__fsub_rn(x, a), __fsub_rn(b, a));
  else if (__fmul_rn(__fsub_rn(x, a), __fsub_rn(b, a)) > 0.0f)
    g = 1.0f;
  else
    g = 0.0f;
  return __fadd_rn(outA, __fmul_rn(g, __fsub_rn(outB, outA)));
}
// lib/threshold/threshold.hpp:117-169
__device__ __forceinline__ float thresh4(float x, float a01, float b01, float a10, float b10,
                                         float outA, float outB) {
  float g = thresh2(x, a01, b01, 0.0f, 1.0f);
  if ((b01 == a10) && (b01 == b10)) return g;  // :131-133 (returns g unscaled)
  if (is_between(x, a01, b01))
    g = thresh2(x, a01, b01, 0.0f, 1.0f);
  else if (is_between(x, a10, b10))
    g = thresh2(x, a10, b10, 0.0f, 1.0f);
  else if (b01 <= a10)
    g = is_between(x, b01, a10) ? 1.0f : 0.0f;
  else if (b10 <= a01)
    g = is_between(x, b10, a01) ? 0.0f : 1.0f;
  return __fadd_rn(outA, __fmul_rn(g, __fsub_rn(outB, outA)));
}

__device__ __forceinline__ float thresh_map(float x, float prev_out, const ThreshParams &p) {
  switch (p.kind) {
    case VISFD_THRESH_SINGLE: return (x > p.t0) ? p.outB : p.outA;   // handlers.cpp:1049-1053
    case VISFD_THRESH_2: return thresh2(x, p.t0, p.t1, p.outA, p.outB);
    case VISFD_THRESH_4: return thresh4(x, p.t0, p.t1, p.t2, p.t3, p.outA, p.outB);
    case VISFD_THRESH_GAUSS: {  // threshold.hpp:248-258 (mixed float/double as written there)
      float dx = __fsub_rn(x, p.t0);
      float xr = __fdiv_rn(dx, p.t1);
      double e = exp(-0.5 * (double)xr * (double)xr);
      return (float)((double)p.outA + (double)__fsub_rn(p.outB, p.outA) * e);
    }
    case VISFD_RESCALE: return __fadd_rn(__fmul_rn(prev_out, p.t0), p.t1);  // handlers.cpp:1040-1043
    default: return x;  // kind 0: copy (mask fill only)
  }
}

__global__ void __launch_bounds__(256)
threshold_kernel(const float *__restrict__ in, float *out, const float *__restrict__ mask, i64 n,
                 ThreshParams p) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float prev = (p.kind == VISFD_RESCALE) ? out[i] : 0.0f;
    float x = (p.kind == VISFD_RESCALE) ? 0.0f : __ldg(in + i);
    float v = thresh_map(x, prev, p);
    if (mask && p.use_masked_value && __ldg(mask + i) == 0.0f) v = p.masked_value;
    out[i] = v;
  }
}

void threshold_device(visfd_ctx *ctx, i64 n, const float *in, float *out, int kind,
                      const float t[4], float outA, float outB, const float *mask,
                      int use_masked_value, float masked_value) {
  VREQUIRE(kind == 0 || kind == VISFD_THRESH_SINGLE || kind == VISFD_THRESH_2 || kind == VISFD_THRESH_4 ||
               kind == VISFD_THRESH_GAUSS || kind == VISFD_RESCALE,
           "unknown threshold kind");
  if (n == 0) return;
  StageTimer timer(ctx, "threshold");
  ThreshParams p{kind, t ? t[0] : 0.f, t ? t[1] : 0.f, t ? t[2] : 0.f, t ? t[3] : 0.f, outA, outB,
                 use_masked_value, masked_value};
  int grid = (int)std::min<i64>((n + 255) / 256, (i64)ctx->sm_count * 16);
  threshold_kernel<<<grid, 256, 0, ctx->stream>>>(in, out, mask, n, p);
  VCK(cudaGetLastError());
  ctx->count_launch();
}

// sums[0] += w*h (or h), sums[1] += w (or 1), sums[2] += w*(h-ave)^2 -- in double.
__global__ void __launch_bounds__(256)
moments_kernel(const float *__restrict__ in, const float *__restrict__ w, i64 n, double ave, int pass,
               double *__restrict__ sums) {
  __shared__ double sh[2][8];
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  double a = 0.0, b = 0.0;
  for (; i < n; i += stride) {
    double h = __ldg(in + i);
    double wt = w ? (double)__ldg(w + i) : 1.0;
    if (pass == 1) { h -= ave; h *= h; }
    a += h * wt;
    b += wt;
  }
  for (int o = 16; o; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][wp] = a; sh[1][wp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tb = 0;
    for (int k = 0; k < 8; k++) { ta += sh[0][k]; tb += sh[1][k]; }
    atomicAdd(&sums[0], ta);
    atomicAdd(&sums[1], tb);
  }
}

// AverageArr / StdDevArr.  The reference accumulates serially in float
// (visfd_utils.hpp:691-707, :770-789), which stops being accurate (its unit-weight
// denominator saturates at 2^24 voxels); the device reduction is in double, i.e. it
// agrees with the reference to float rounding on small volumes and is deliberately
// the exact value on large ones.
// sums[0] = sum w*h (squared: sum w*(h-center)^2), sums[1] = sum w (w = 1 without weights), in double
void moment_sums_device(visfd_ctx *ctx, i64 n, const float *in, const float *w, double center, bool squared,
                        double sums[2]) {
  VREQUIRE(n >= 0, "negative length");
  sums[0] = sums[1] = 0.0;
  if (n == 0) return;
  Scratch<double> d(ctx, 2);
  int grid = (int)std::min<i64>((n + 255) / 256, (i64)ctx->sm_count * 16);
  {
    StageTimer timer(ctx, "threshold");
    VCK(cudaMemsetAsync(d.get(), 0, 2 * sizeof(double), ctx->stream));
    moments_kernel<<<grid, 256, 0, ctx->stream>>>(in, w, n, center, squared ? 1 : 0, d.get());
    VCK(cudaGetLastError());
    ctx->count_launch();
  }
  VCK(cudaMemcpyAsync(sums, d.get(), 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  VCK(cudaStreamSynchronize(ctx->stream));
}

void mean_stddev_device(visfd_ctx *ctx, i64 n, const float *in, const float *w, float *mean,
                        float *stddev) {
  VREQUIRE(n > 0, "mean/stddev of an empty volume");
  double h[2];
  moment_sums_device(ctx, n, in, w, 0.0, false, h);
  const float ave = (float)(h[0] / h[1]);
  if (mean) *mean = ave;
  if (stddev) {
    moment_sums_device(ctx, n, in, w, (double)ave, true, h);
    *stddev = (float)sqrt(h[0] / h[1]);
  }
}

}  // namespace visfd_cuda
