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
template <bool DIR>
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
  double ev[3];
  if (DIR) {
    double e0[3];
    sym3_eigen_first(m, order, ev, e0);
    dir[3 * i + 0] = (float)e0[0];
    dir[3 * i + 1] = (float)e0[1];
    dir[3 * i + 2] = (float)e0[2];
  } else {
    sym3_eigenvalues(m, order, ev);
  }
  sal[i] = score_from_eivals(ev, score_kind, 0);
}

// ---- saliency only: z-marching ---------------------------------------------------------------------
// One thread per (x, y) column of a chunk of RIDGE_ZC output planes.  The 3x3 neighbourhoods of the three
// planes around the stencil centre live in registers and rotate as the thread marches along z, so a voxel
// costs 9 loads (the incoming plane) instead of 19, with one pointer increment instead of 19 index
// computations; the Hessian keeps the reference's float operations and their order (fd_hessian), the
// eigenvalues come from sym3_eigenvalues_newton.  What bounds it: ~58 FP64 instructions per voxel on a pipe
// that retires 64 per clock and SM, and 8 float<->double conversions + 3 MUFU on one that retires 16.
constexpr int RIDGE_ZC = 32;

struct Rows3 {   // rows y-1, y, y+1 of the plane being loaded, at the column of the stencil centre
  const float *m, *c, *p;
  __device__ __forceinline__ void next(i64 sz) { m += sz; c += sz; p += sz; }
};
__device__ __forceinline__ void load9(float w[9], const Rows3 &r) {
  w[0] = __ldg(r.m - 1); w[1] = __ldg(r.m); w[2] = __ldg(r.m + 1);
  w[3] = __ldg(r.c - 1); w[4] = __ldg(r.c); w[5] = __ldg(r.c + 1);
  w[6] = __ldg(r.p - 1); w[7] = __ldg(r.p); w[8] = __ldg(r.p + 1);
}

// fd_hessian (stencil.cuh) on a register window: A = plane below the centre, B = centre plane, C = above;
// w[(dy + 1) * 3 + dx + 1]
__device__ __forceinline__ void fd_hessian_window(const float A[9], const float B[9], const float C[9], float s2, float h[6]) {
  const float c2 = __fmul_rn(2.0f, B[4]);
  h[0] = __fmul_rn(__fsub_rn(__fadd_rn(B[5], B[3]), c2), s2);
  h[1] = __fmul_rn(__fsub_rn(__fadd_rn(B[7], B[1]), c2), s2);
  h[2] = __fmul_rn(__fsub_rn(__fadd_rn(C[4], A[4]), c2), s2);
  const float xy = __fsub_rn(__fsub_rn(__fadd_rn(B[8], B[0]), B[2]), B[6]);
  const float yz = __fsub_rn(__fsub_rn(__fadd_rn(C[7], A[1]), A[7]), C[1]);
  const float xz = __fsub_rn(__fsub_rn(__fadd_rn(C[5], A[3]), C[3]), A[5]);
  h[3] = __fmul_rn(__fmul_rn(0.25f, xy), s2);
  h[4] = __fmul_rn(__fmul_rn(0.25f, yz), s2);
  h[5] = __fmul_rn(__fmul_rn(0.25f, xz), s2);
}

template <bool LINEAR>
__device__ __forceinline__ float ridge_score(const float h[6], int order) {
  Sym3d m = {h[0], h[1], h[2], h[3], h[4], h[5]};
  double ev[3];
  if (LINEAR) {
    sym3_eigenvalues(m, order, ev);
    return score_from_eivals(ev, 1, 0);
  }
  // ScoreHessianPlanar (feature.hpp:1529-1545) on the FLOAT eigenvalues: (l1^2 - l2^2)^2
  sym3_eigenvalues_newton(m, ev);   // ascending; "decreasing" puts the largest first, the middle one stays
  const float l1 = (float)(order == 1 ? ev[2] : ev[0]), l2 = (float)ev[1];
  const float d = (l1 - l2) * (l1 + l2);
  return d * d;
}

template <bool LINEAR>
__global__ void __launch_bounds__(256)
ridge_march_kernel(const float *__restrict__ sm, const float *__restrict__ mask, int nx, int ny, i64 z0, i64 z1,
                   i64 z_offset, i64 nz_global, float sigma, int order, float *__restrict__ sal) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x;
  const int iy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ix >= nx || iy >= ny) return;
  const i64 za = z0 + (i64)blockIdx.z * RIDGE_ZC, zb = min(z1, za + RIDGE_ZC);   // output planes of this thread
  // stencil centres move one voxel inward at the GLOBAL border (visfd_utils.hpp:597-610)
  const int xs = min(max(ix, 1), nx - 2), ys = min(max(iy, 1), ny - 2);
  const i64 c_first = min(max(z_offset + za, (i64)1), nz_global - 2) - z_offset;
  const i64 c_last = min(max(z_offset + zb - 1, (i64)1), nz_global - 2) - z_offset;
  const i64 sy = nx, sz = (i64)nx * ny;
  const float s2 = __fmul_rn(sigma, sigma);
  Rows3 p;
  p.c = sm + ((c_first - 1) * ny + ys) * sy + xs;
  p.m = p.c - sy;
  p.p = p.c + sy;
  // four register windows: while the stencil of centre c (planes c-1, c, c+1) is evaluated, plane c+2 is in flight
  float W0[9], W1[9], W2[9], W3[9];
  load9(W0, p);
  p.next(sz);
  load9(W1, p);
  p.next(sz);
  load9(W2, p);
  int c = (int)c_first;
  const int ia = (int)za, ib = (int)zb, ic_last = (int)c_last;
  float *out = sal + (c_first * sz + (i64)iy * nx + ix);           // voxel (ix, iy, c)
  const float *mk = mask ? mask + (out - sal) : nullptr;
  const int c_plane0 = (int)(1 - z_offset), c_planeN = (int)(nz_global - 2 - z_offset);   // centres that also serve a border plane
  auto put = [&](i64 dz, float v) {
    // tomo_out is zero-initialised for masked voxels (handlers.cpp:1640-1643)
    out[dz] = (mk && __ldg(mk + dz) == 0.0f) ? 0.0f : v;
  };
  auto step = [&](const float A[9], const float B[9], const float C[9], float N[9]) {
    if (c < ic_last) {   // plane c + 2, for the next centre (inside the slab: check_fd_dims)
      p.next(sz);
      load9(N, p);
    }
    float h[6];
    fd_hessian_window(A, B, C, s2, h);
    const float v = ridge_score<LINEAR>(h, order);
    if (c >= ia && c < ib) put(0, v);
    if (c == c_plane0 && c - 1 >= ia) put(-sz, v);                 // global plane 0 uses the stencil of plane 1
    if (c == c_planeN && c + 1 < ib) put(sz, v);                   // the last plane that of the one before it
    out += sz;
    if (mk) mk += sz;
    ++c;
  };
  while (c <= ic_last) {
    step(W0, W1, W2, W3); if (c > ic_last) break;
    step(W1, W2, W3, W0); if (c > ic_last) break;
    step(W2, W3, W0, W1); if (c > ic_last) break;
    step(W3, W0, W1, W2);
  }
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
  if (direction) {
    ridge_kernel<true><<<grid, block, 0, ctx->stream>>>(smoothed, mask, (int)nx, (int)ny, z0, z_offset, nz_global, sigma,
                                                        eival_order, score_kind, saliency, direction);
  } else {
    dim3 mgrid(div_up(nx, 64), div_up(ny, 4), (unsigned)div_up(z1 - z0, RIDGE_ZC));
    if (score_kind == 1)
      ridge_march_kernel<true><<<mgrid, block, 0, ctx->stream>>>(smoothed, mask, (int)nx, (int)ny, z0, z1, z_offset,
                                                                 nz_global, sigma, eival_order, saliency);
    else
      ridge_march_kernel<false><<<mgrid, block, 0, ctx->stream>>>(smoothed, mask, (int)nx, (int)ny, z0, z1, z_offset,
                                                                  nz_global, sigma, eival_order, saliency);
  }
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
