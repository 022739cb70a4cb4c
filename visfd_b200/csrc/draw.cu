// draw.cu -- DrawRegions (lib/visfd/draw.hpp:90-237): the rasterisation filter_mrc uses to build
// a mask from "-mask-rect" / "-mask-sphere" arguments (bin/filter_mrc/filter_mrc.cpp:213-285)
// before the membrane / blob path reads it.
//
// The reference paints the regions one after the other, so a voxel ends up with the value of
// the LAST region that covers it (value >= 0: overwrite; value < 0 with
// negative_means_subtract: clear the voxel if it is positive).  Here every voxel replays the
// region list for itself, in order: one thread per voxel, lanes along x, the prepared list
// (integer bounds and R*R computed on the host exactly as the reference rounds them) in
// shared memory.  One read and at most one write per voxel.
#include <cmath>

#include "../../include/visfd_cuda.h"
#include "common.cuh"
#include "kernels.cuh"

namespace visfd_cuda {

struct PreparedRegion {
  int type;            // 0 box, 1 sphere
  int lo[3], hi[3];    // inclusive voxel bounds x, y, z (box: rounded corners clipped to the image;
                       // sphere: centre -/+ Ri)
  int c[3];            // sphere centre (voxels)
  float r2;            // R*R in float
  float value;
};

__device__ __forceinline__ bool covers(const PreparedRegion &r, int x, int y, int z) {
  if (x < r.lo[0] || x > r.hi[0] || y < r.lo[1] || y > r.hi[1] || z < r.lo[2] || z > r.hi[2]) return false;
  if (r.type == 0) return true;
  // draw.hpp:148-153: descr = R*R - (jy*jy + jz*jz), skipped if negative; |jx| <= floor(sqrt(descr))
  const int jx = x - r.c[0], jy = y - r.c[1], jz = z - r.c[2];
  const float descr = __fsub_rn(r.r2, (float)(jy * jy + jz * jz));
  if (descr < 0.0f) return false;
  const int xrange = (int)floorf(__fsqrt_rn(descr));
  return abs(jx) <= xrange;
}

// draw.hpp:110-134: "is every unmasked voxel zero?"  (only asked when the first region subtracts)
__global__ void __launch_bounds__(256) any_nonzero_kernel(const float *__restrict__ img, const float *__restrict__ mask,
                                                          i64 n, int *flag) {
  bool found = false;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
    if (!(mask && __ldg(mask + i) == 0.0f) && __ldg(img + i) != 0.0f) found = true;
  if (__syncthreads_or(found) && threadIdx.x == 0) atomicOr(flag, 1);
}

__global__ void __launch_bounds__(256) draw_regions_kernel(float *__restrict__ img, const float *__restrict__ mask,
                                                           const PreparedRegion *__restrict__ regions, int n_regions,
                                                           int nx, int ny, int subtract, const int *any_nonzero) {
  extern __shared__ __align__(16) unsigned char draw_smem[];
  PreparedRegion *sr = reinterpret_cast<PreparedRegion *>(draw_smem);
  for (int i = threadIdx.x; i < n_regions * (int)(sizeof(PreparedRegion) / 4); i += blockDim.x)
    reinterpret_cast<int *>(sr)[i] = reinterpret_cast<const int *>(regions)[i];
  __syncthreads();
  const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6), z = blockIdx.z;
  if (x >= nx || y >= ny) return;
  const size_t at = ((size_t)z * ny + y) * (size_t)nx + x;
  if (mask && __ldg(mask + at) == 0.0f) return;  // masked voxels are never touched
  const float before = img[at];
  float v = before;
  // an all-zero image whose first region subtracts starts from ones (draw.hpp:98-134)
  if (any_nonzero && *any_nonzero == 0) v = 1.0f;
  for (int i = 0; i < n_regions; i++) {
    if (!covers(sr[i], x, y, z)) continue;
    const float value = sr[i].value;
    if (value < 0.0f) {
      if (subtract && v > 0.0f) v = 0.0f;
    } else {
      v = value;
    }
  }
  if (v != before || (v != v) != (before != before)) img[at] = v;
}

void draw_regions_device(visfd_ctx *ctx, i64 nx, i64 ny, i64 nz, float *img, const float *mask,
                         const visfd_region *regions, int n_regions, bool negative_means_subtract) {
  std::vector<PreparedRegion> prep((size_t)n_regions);
  const i64 size[3] = {nx, ny, nz};
  for (int i = 0; i < n_regions; i++) {
    const visfd_region &r = regions[i];
    PreparedRegion &p = prep[(size_t)i];
    p.value = r.value;
    p.r2 = 0.0f;
    for (int d = 0; d < 3; d++) p.c[d] = 0;
    if (r.type == VISFD_REGION_SPHERE) {
      p.type = 1;
      const float R = r.p[3];
      const int Ri = (int)std::ceil((double)R - 0.5);                       // draw.hpp:141
      for (int d = 0; d < 3; d++) {
        p.c[d] = (int)std::floor((double)r.p[d] + 0.5);                      // draw.hpp:143-145
        p.lo[d] = p.c[d] - Ri;
        p.hi[d] = p.c[d] + Ri;
      }
      p.r2 = R * R;
    } else {
      VREQUIRE(r.type == VISFD_REGION_RECT, "unknown region type");
      p.type = 0;
      for (int d = 0; d < 3; d++) {
        // draw.hpp:178-196: corners rounded in double, stored as float, clipped with float min/max,
        // the loop variable an int compared against the float upper bound
        const float fmin_ = (float)std::floor((double)r.p[2 * d] + 0.5), fmax_ = (float)std::floor((double)r.p[2 * d + 1] + 0.5);
        const float lo = std::max<float>(fmin_, 0.0f), hi = std::min<float>(fmax_, (float)(size[d] - 1));
        p.lo[d] = (int)lo;
        p.hi[d] = hi < 0.0f ? -1 : (int)std::floor(hi);
      }
    }
  }
  const bool from_ones = negative_means_subtract && n_regions > 0 && regions[0].value < 0.0f;
  const i64 n = nx * ny * nz;
  Scratch<int> flag;
  if (from_ones) {
    flag.reset(ctx, 1);
    VCK(cudaMemsetAsync(flag.get(), 0, sizeof(int), ctx->stream));
    any_nonzero_kernel<<<(unsigned)std::min<i64>(div_up(n, 256), 148 * 16), 256, 0, ctx->stream>>>(img, mask, n, flag.get());
    VCK(cudaGetLastError());
    ctx->count_launch();
  }
  if (n_regions == 0) return;
  const size_t bytes = (size_t)n_regions * sizeof(PreparedRegion);
  VREQUIRE(bytes <= 48 * 1024, "too many regions for one pass (limit 930)");
  Scratch<PreparedRegion> d_regions(ctx, (size_t)n_regions);
  VCK(cudaMemcpyAsync(d_regions.get(), prep.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
  dim3 grid(div_up(nx, 64), div_up(ny, 4), (unsigned)nz);
  draw_regions_kernel<<<grid, 256, bytes, ctx->stream>>>(img, mask, d_regions.get(), n_regions, (int)nx, (int)ny,
                                                         negative_means_subtract ? 1 : 0, flag.get());
  VCK(cudaGetLastError());
  ctx->count_launch();
  VCK(cudaStreamSynchronize(ctx->stream));  // prep[] is pageable host memory; the scratch buffers go back to the pool
}

}  // namespace visfd_cuda
