// stencil.cuh -- finite-difference stencils shared by ridge.cu and connect.cu
// (lib/visfd/visfd_utils.hpp:530-616), float arithmetic in the reference's order.
#pragma once
#include "common.cuh"

namespace visfd_cuda {

// Stencil access with the reference's border rule: the stencil CENTRE is moved one voxel
// inward at the (global) image border (lib/visfd/visfd_utils.hpp:597-610, :649-660).
struct Stencil {
  const float *p;  // pointer at the (clamped) centre voxel
  i64 sy, sz;
  __device__ __forceinline__ float at(int dx, int dy, int dz) const {
    return __ldg(p + dx + dy * sy + dz * sz);
  }
};

__device__ __forceinline__ Stencil make_stencil(const float *sm, int nx, int ny, i64 z_offset,
                                                i64 nz_global, int ix, int iy, i64 iz_local) {
  int x = ix, y = iy;
  i64 zg = z_offset + iz_local;
  if (x == 0) x++; else if (x == nx - 1) x--;
  if (y == 0) y++; else if (y == ny - 1) y--;
  if (zg == 0) zg++; else if (zg == nz_global - 1) zg--;
  Stencil s;
  s.sy = nx;
  s.sz = (i64)nx * ny;
  s.p = sm + ((zg - z_offset) * ny + y) * (i64)nx + x;
  return s;
}

// 19-point Hessian, flat order xx,yy,zz,xy,yz,xz, scaled by sigma^2
// (visfd_utils.hpp:530-565, feature.hpp:1331-1333); float arithmetic as the reference.
__device__ __forceinline__ void fd_hessian(const Stencil &s, float s2, float h[6]) {
  float c = s.at(0, 0, 0);
  float xp = s.at(1, 0, 0), xm = s.at(-1, 0, 0);
  float yp = s.at(0, 1, 0), ym = s.at(0, -1, 0);
  float zp = s.at(0, 0, 1), zm = s.at(0, 0, -1);
  float c2 = __fmul_rn(2.0f, c);
  h[0] = __fmul_rn(__fsub_rn(__fadd_rn(xp, xm), c2), s2);
  h[1] = __fmul_rn(__fsub_rn(__fadd_rn(yp, ym), c2), s2);
  h[2] = __fmul_rn(__fsub_rn(__fadd_rn(zp, zm), c2), s2);
  float xy = __fsub_rn(__fsub_rn(__fadd_rn(s.at(1, 1, 0), s.at(-1, -1, 0)), s.at(1, -1, 0)), s.at(-1, 1, 0));
  float yz = __fsub_rn(__fsub_rn(__fadd_rn(s.at(0, 1, 1), s.at(0, -1, -1)), s.at(0, 1, -1)), s.at(0, -1, 1));
  float xz = __fsub_rn(__fsub_rn(__fadd_rn(s.at(1, 0, 1), s.at(-1, 0, -1)), s.at(-1, 0, 1)), s.at(1, 0, -1));
  h[3] = __fmul_rn(__fmul_rn(0.25f, xy), s2);
  h[4] = __fmul_rn(__fmul_rn(0.25f, yz), s2);
  h[5] = __fmul_rn(__fmul_rn(0.25f, xz), s2);
}

}  // namespace visfd_cuda
