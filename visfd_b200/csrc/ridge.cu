// ridge.cu -- finite-difference Hessian of the smoothed image fused with the 3x3
// symmetric eigensolve and the ridge score, so that the 6-component tensor never
// travels to HBM (CalcHessian's second half, lib/visfd/feature.hpp:1271-1345, plus the
// per-voxel loop of HandleTV, bin/filter_mrc/handlers.cpp:1645-1746).
//
// Traffic: 4 B read (smoothed, neighbours come from L1/L2) + 4 B saliency
// [+ 12 B direction] written per voxel.
#include "common.cuh"
#include "kernels.cuh"
#include "eigen3.cuh"
#include "stencil.cuh"

namespace visfd_cuda {

__global__ void __launch_bounds__(256)
hessian_fd_kernel(const float *__restrict__ sm, const float *__restrict__ mask, int nx, int ny,
                  i64 nz_local, i64 z_offset, i64 nz_global, float sigma,
                  float *__restrict__ grad, float *__restrict__ hess) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x;
  const int iy = blockIdx.y * blockDim.y + threadIdx.y;
  const i64 iz = blockIdx.z;
  if (ix >= nx || iy >= ny) return;
  const i64 i = (iz * ny + iy) * (i64)nx + ix;
  if (mask && __ldg(mask + i) == 0.0f) return;
  Stencil s = make_stencil(sm, nx, ny, z_offset, nz_global, ix, iy, iz);
  if (grad) {
    grad[3 * i + 0] = __fmul_rn(__fmul_rn(0.5f, __fsub_rn(s.at(1, 0, 0), s.at(-1, 0, 0))), sigma);
    grad[3 * i + 1] = __fmul_rn(__fmul_rn(0.5f, __fsub_rn(s.at(0, 1, 0), s.at(0, -1, 0))), sigma);
    grad[3 * i + 2] = __fmul_rn(__fmul_rn(0.5f, __fsub_rn(s.at(0, 0, 1), s.at(0, 0, -1))), sigma);
  }
  if (hess) {
    float h[6];
    fd_hessian(s, __fmul_rn(sigma, sigma), h);
#pragma unroll
    for (int k = 0; k < 6; k++) hess[6 * i + k] = h[k];
  }
}

// One thread per voxel of planes [z0, z1) (blockIdx.z + z0).
__global__ void __launch_bounds__(256)
ridge_kernel(const float *__restrict__ sm, const float *__restrict__ mask, int nx, int ny,
             i64 z0, i64 z_offset, i64 nz_global, float sigma, int order, int score_kind,
             float *__restrict__ sal, float *__restrict__ dir) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x;
  const int iy = blockIdx.y * blockDim.y + threadIdx.y;
  const i64 iz = z0 + blockIdx.z;
  if (ix >= nx || iy >= ny) return;
  const i64 i = (iz * ny + iy) * (i64)nx + ix;
  if (mask && __ldg(mask + i) == 0.0f) {
    sal[i] = 0.0f;  // tomo_out is zero-initialised for masked voxels (handlers.cpp:1640-1643)
    return;
  }
  Stencil s = make_stencil(sm, nx, ny, z_offset, nz_global, ix, iy, iz);
  float h[6];
  fd_hessian(s, __fmul_rn(sigma, sigma), h);
  Sym3d m = {h[0], h[1], h[2], h[3], h[4], h[5]};
  double ev[3], e0[3];
  if (dir) {
    sym3_eigen_first(m, order, ev, e0);
    dir[3 * i + 0] = (float)e0[0];
    dir[3 * i + 1] = (float)e0[1];
    dir[3 * i + 2] = (float)e0[2];
  } else {
    sym3_eigenvalues(m, order, ev);
  }
  sal[i] = score_from_eivals(ev, score_kind, 0);
}

__global__ void __launch_bounds__(256)
tensor_score_kernel(const float *__restrict__ tensor, const float *__restrict__ mask, i64 n,
                    int order, int score_kind, int is_vote, float *__restrict__ score,
                    float *__restrict__ eivals, float *__restrict__ dir) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && __ldg(mask + i) == 0.0f) {
    if (!is_vote && score) score[i] = 0.0f;
    return;
  }
  const float *t = tensor + 6 * i;
  Sym3d m = {__ldg(t + 0), __ldg(t + 1), __ldg(t + 2), __ldg(t + 3), __ldg(t + 4), __ldg(t + 5)};
  double ev[3], e0[3];
  if (dir) {
    sym3_eigen_first(m, order, ev, e0);
    dir[3 * i + 0] = (float)e0[0];
    dir[3 * i + 1] = (float)e0[1];
    dir[3 * i + 2] = (float)e0[2];
  } else {
    sym3_eigenvalues(m, order, ev);
  }
  if (eivals) {
    eivals[3 * i + 0] = (float)ev[0];
    eivals[3 * i + 1] = (float)ev[1];
    eivals[3 * i + 2] = (float)ev[2];
  }
  if (score) score[i] = score_from_eivals(ev, score_kind, is_vote);
}

static void check_fd_dims(i64 nx, i64 ny, i64 nz_local, i64 z_offset, i64 nz_global, i64 z0, i64 z1) {
  // feature.hpp:1260-1264
  VREQUIRE(nx >= 3 && ny >= 3 && nz_global >= 3,
           "CalcHessian() requires an image that is at least 3 voxels wide in the x,y,z directions");
  VREQUIRE(z0 >= 0 && z1 <= nz_local && z0 <= z1, "plane range outside the slab");
  VREQUIRE(z_offset >= 0 && z_offset + nz_local <= nz_global, "slab outside the volume");
  // the clamped stencil of plane z needs planes z-1..z+1 (global clamp) inside the slab
  i64 lo = std::max<i64>(z_offset + z0 - 1, 0), hi = std::min<i64>(z_offset + z1, nz_global - 1);
  if (z1 > z0)
    VREQUIRE(lo >= z_offset && hi < z_offset + nz_local,
             "slab lacks the 1-plane halo needed by the finite-difference stencil");
  VREQUIRE(nx < (1 << 30) && ny < (1 << 30) && (z1 - z0) <= 65535, "volume too large for one launch");
}

void hessian_fd_device(visfd_ctx *ctx, i64 nx, i64 ny, i64 nz_local, i64 z_offset, i64 nz_global,
                       const float *smoothed, const float *mask, float sigma, float *gradient,
                       float *hessian) {
  check_fd_dims(nx, ny, nz_local, z_offset, nz_global, 0, nz_local);
  StageTimer t(ctx, "ridge");
  dim3 block(64, 4, 1);
  dim3 grid(div_up(nx, 64), div_up(ny, 4), (unsigned)nz_local);
  hessian_fd_kernel<<<grid, block, 0, ctx->stream>>>(smoothed, mask, (int)nx, (int)ny, nz_local,
                                                     z_offset, nz_global, sigma, gradient, hessian);
  VCK(cudaGetLastError());
  ctx->count_launch();
}

void ridge_device(visfd_ctx *ctx, i64 nx, i64 ny, i64 nz_local, i64 z_offset, i64 nz_global,
                  i64 z0, i64 z1, const float *smoothed, const float *mask, float sigma,
                  int eival_order, int score_kind, float *saliency, float *direction) {
  check_fd_dims(nx, ny, nz_local, z_offset, nz_global, z0, z1);
  if (z1 == z0) return;
  StageTimer t(ctx, "ridge");
  dim3 block(64, 4, 1);
  dim3 grid(div_up(nx, 64), div_up(ny, 4), (unsigned)(z1 - z0));
  ridge_kernel<<<grid, block, 0, ctx->stream>>>(smoothed, mask, (int)nx, (int)ny, z0, z_offset,
                                                nz_global, sigma, eival_order, score_kind, saliency,
                                                direction);
  VCK(cudaGetLastError());
  ctx->count_launch();
}

void tensor_score_device(visfd_ctx *ctx, i64 n, const float *tensor, const float *mask,
                         int eival_order, int score_kind, int is_vote_tensor, float *score,
                         float *eivals, float *direction) {
  if (n == 0) return;
  StageTimer t(ctx, "ridge");
  tensor_score_kernel<<<div_up(n, 256), 256, 0, ctx->stream>>>(tensor, mask, n, eival_order, score_kind,
                                                               is_vote_tensor, score, eivals, direction);
  VCK(cudaGetLastError());
  ctx->count_launch();
}

}  // namespace visfd_cuda
