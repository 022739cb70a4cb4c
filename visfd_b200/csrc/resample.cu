// resample.cu -- BinArray3D / UnbinArray3D (lib/visfd/resample.hpp:53-166): the reduction
// filter_mrc applies before the membrane path when the feature is wide (-bin N, or the
// automatic binning of bin/filter_mrc/filter_mrc.cpp:130-210, handlers.cpp:2361-2425) and the
// nearest-voxel expansion it applies to the result afterwards (handlers.cpp:2321-2355).
//
// Both stream: bin reads every source voxel once and writes one voxel per bin
// (4 B * (N_src + N_dst)); unbin reads N_src (from L2 mostly) and writes N_dst.
// bin: one thread per destination voxel, lanes along x; the bx*by*bz sources are summed in
// the reference's order (dz outer, dy, dx inner; float accumulator) and divided by the
// integer bin volume with an IEEE division, so the result is bit-identical.
#include "common.cuh"
#include "kernels.cuh"

namespace visfd_cuda {

struct ResampleArgs {
  int snx, sny, snz;  // source size
  int dnx, dny, dnz;  // destination size
  int bx, by, bz;     // bin size = floor(larger / smaller) per axis
  int ox, oy, oz;     // offset of the binning window
};

__global__ void __launch_bounds__(256) bin3d_kernel(ResampleArgs a, const float *__restrict__ src, float *__restrict__ dst) {
  const int X = blockIdx.x * 64 + (threadIdx.x & 63);
  const int Y = blockIdx.y * 4 + (threadIdx.x >> 6);
  const int Z = blockIdx.z;
  if (X >= a.dnx || Y >= a.dny) return;
  float sum = 0.0f;
  for (int dz = 0; dz < a.bz; dz++)
    for (int dy = 0; dy < a.by; dy++) {
      const float *row = src + ((size_t)(Z * a.bz + dz + a.oz) * a.sny + (Y * a.by + dy + a.oy)) * (size_t)a.snx +
                         (X * a.bx + a.ox);
      for (int dx = 0; dx < a.bx; dx++) sum = __fadd_rn(sum, __ldg(row + dx));
    }
  // resample.hpp:100: sum / (bin_size[0]*bin_size[1]*bin_size[2]) (int product -> float)
  dst[((size_t)Z * a.dny + Y) * (size_t)a.dnx + X] = __fdiv_rn(sum, (float)(a.bx * a.by * a.bz));
}

// four consecutive x per thread: one integer division per axis and thread, the other three
// x indices follow by counting the remainder up (VEC: float4 store, needs dnx % 4 == 0)
template <bool VEC>
__global__ void __launch_bounds__(256) unbin3d_kernel(ResampleArgs a, const float *__restrict__ src, float *__restrict__ dst) {
  const int X0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
  const int Y = blockIdx.y * 4 + (threadIdx.x >> 6);
  const int Z = blockIdx.z;
  if (X0 >= a.dnx || Y >= a.dny) return;
  // resample.hpp:151-160: C integer division (truncates toward zero), then clamp
  const int iy = min(max((Y - a.oy) / a.by, 0), a.sny - 1);
  const int iz = min(max((Z - a.oz) / a.bz, 0), a.snz - 1);
  const float *row = src + ((size_t)iz * a.sny + iy) * (size_t)a.snx;
  int q = (X0 - a.ox) / a.bx, r = (X0 - a.ox) - q * a.bx;   // r < 0 only while X < ox, where q == 0 as well
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    v[e] = __ldg(row + min(max(q, 0), a.snx - 1));
    if (++r == a.bx) { r = 0; q++; }
  }
  float *o = dst + ((size_t)Z * a.dny + Y) * (size_t)a.dnx + X0;
  if (VEC) {
    *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int e = 0; e < 4; e++)
      if (X0 + e < a.dnx) o[e] = v[e];
  }
}

static ResampleArgs resample_args(const i64 small[3], const i64 large[3], const int *offset, bool binning) {
  ResampleArgs a;
  int b[3], o[3] = {0, 0, 0};
  for (int d = 0; d < 3; d++) {
    VREQUIRE(small[d] > 0 && large[d] > 0 && large[d] < 2147483647LL, "bad image size");
    VREQUIRE(large[d] >= small[d], binning ? "BinArray3D: the destination must not be larger than the source"
                                           : "UnbinArray3D: the destination must not be smaller than the source");
    b[d] = (int)(large[d] / small[d]);
    if (offset) {
      // resample.hpp:62-70 / :128-136
      VREQUIRE(offset[d] >= 0 && offset[d] < b[d], "offset[d] should lie between 0 and floor(size ratio)-1");
      o[d] = offset[d];
    }
  }
  a.bx = b[0]; a.by = b[1]; a.bz = b[2];
  a.ox = o[0]; a.oy = o[1]; a.oz = o[2];
  if (binning) {
    a.snx = (int)large[0]; a.sny = (int)large[1]; a.snz = (int)large[2];
    a.dnx = (int)small[0]; a.dny = (int)small[1]; a.dnz = (int)small[2];
    // the last window must lie inside the source (the reference asserts it, resample.hpp:89-94)
    for (int d = 0; d < 3; d++)
      VREQUIRE(small[d] * b[d] + o[d] <= large[d], "BinArray3D: the shifted binning window leaves the source image");
  } else {
    a.snx = (int)small[0]; a.sny = (int)small[1]; a.snz = (int)small[2];
    a.dnx = (int)large[0]; a.dny = (int)large[1]; a.dnz = (int)large[2];
  }
  return a;
}

void bin3d_device(visfd_ctx *ctx, const i64 size_src[3], const i64 size_dst[3], const float *src, float *dst,
                  const int *offset) {
  StageTimer t(ctx, "bin");
  ResampleArgs a = resample_args(size_dst, size_src, offset, true);
  VREQUIRE(a.dnz <= 65535 && div_up(a.dny, 4) <= 65535, "image too large in y/z for one launch");
  dim3 grid(div_up(a.dnx, 64), div_up(a.dny, 4), (unsigned)a.dnz);
  bin3d_kernel<<<grid, 256, 0, ctx->stream>>>(a, src, dst);
  VCK(cudaGetLastError());
  ctx->count_launch();
}

void unbin3d_device(visfd_ctx *ctx, const i64 size_src[3], const i64 size_dst[3], const float *src, float *dst,
                    const int *offset) {
  StageTimer t(ctx, "unbin");
  ResampleArgs a = resample_args(size_src, size_dst, offset, false);
  VREQUIRE(a.dnz <= 65535 && div_up(a.dny, 4) <= 65535, "image too large in y/z for one launch");
  dim3 grid(div_up(a.dnx, 256), div_up(a.dny, 4), (unsigned)a.dnz);
  if (a.dnx % 4 == 0 && ((uintptr_t)dst & 15) == 0) unbin3d_kernel<true><<<grid, 256, 0, ctx->stream>>>(a, src, dst);
  else unbin3d_kernel<false><<<grid, 256, 0, ctx->stream>>>(a, src, dst);
  VCK(cudaGetLastError());
  ctx->count_launch();
}

}  // namespace visfd_cuda
