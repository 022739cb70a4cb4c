// surface.cu -- the oriented point cloud of a detected surface: the block of HandleTV that runs when
// `-normals-file` is given (bin/filter_mrc/handlers.cpp:2039-2309), SURVEY 8f rank 4.
//
// The reference visits the voxels of the selected cluster one after the other; every visit is independent of
// the others, so here each voxel of the cluster gets a thread:
//   1. the curve along the surface normal (step ds) is followed in both directions while it stays inside the
//      cluster (:2097-2167); instead of storing the samples the thread walks twice -- once to accumulate the
//      saliency-weighted mean arc length, once more (on the side where the mean lies) to recover the two samples
//      next to it.  The walk is a deterministic float recurrence, so the second pass revisits the same samples.
//      (The reference sums in float from the far end of the backward branch; the sums here are double and in
//      walking order: the mean differs by ~1e-7, below the six digits of the PLY file.)
//   2. the point is moved onto the ridge of the saliency: finite-difference gradient and Hessian at the rounded
//      position, eigenvector of the largest |eigenvalue| (DECREASING_ABS_EIVALS), Newton step along it
//      (:2224-2295); points too far from the ridge, with a vanishing gradient component or outside the image
//      are dropped.
// Survivors are appended to a list with their voxel index and sorted by it on the host: the reference's raster
// order.  Without labels every un-masked voxel is listed as it is (:2053-2066).
#include "common.cuh"
#include "kernels.cuh"
#include "eigen3.cuh"
#include <algorithm>
#include <numeric>
#include <vector>

namespace visfd_cuda {

struct SurfaceArgs {
  const float *sal, *dir, *labels, *mask;
  int nx, ny, nz;
  float select;           // (float) select_cluster: the reference compares an int with the float label image
  float vw[3];
  float ds;
  int find_ridge;
  float max_distance;
  int max_steps;
  unsigned long long capacity;
  unsigned long long *count;
  long long *index;       // [capacity] source voxel
  float *rows;            // [capacity][6]
};

struct Walker {
  const SurfaceArgs &g;
  float lab0;
  float r[3];
  int p[3];
  __device__ Walker(const SurfaceArgs &g_, int ix, int iy, int iz, float lab) : g(g_), lab0(lab) { reset(ix, iy, iz); }
  __device__ void reset(int ix, int iy, int iz) {
    r[0] = (float)ix; r[1] = (float)iy; r[2] = (float)iz;
    p[0] = ix; p[1] = iy; p[2] = iz;
  }
  __device__ i64 idx() const { return ((i64)p[2] * g.ny + p[1]) * g.nx + p[0]; }
  __device__ bool in_cluster() const {   // the loop condition of :2112-2123 / the breaks of :2141-2157
    if (p[0] < 0 || p[0] >= g.nx || p[1] < 0 || p[1] >= g.ny || p[2] < 0 || p[2] >= g.nz) return false;
    const i64 j = idx();
    if (g.mask && __ldg(g.mask + j) == 0.0f) return false;
    return __ldg(g.labels + j) == lab0;
  }
  // one step of length sgn * ds along the unit normal of the voxel the walker stands on (:2129-2135, :2144-2149)
  __device__ void step(float sgn_ds) {
    const float *v = g.dir + 3 * idx();
    const float vx = __ldg(v), vy = __ldg(v + 1), vz = __ldg(v + 2);
    const float norm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
    const float d[3] = {__fdiv_rn(vx, norm), __fdiv_rn(vy, norm), __fdiv_rn(vz, norm)};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      r[k] = __fadd_rn(r[k], __fmul_rn(sgn_ds, d[k]));
      p[k] = (int)roundf(r[k]);
    }
  }
};

__device__ __forceinline__ void unit_normal(const SurfaceArgs &g, i64 j, float n[3]) {
  const float *v = g.dir + 3 * j;
  const float vx = __ldg(v), vy = __ldg(v + 1), vz = __ldg(v + 2);
  const float norm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
  n[0] = __fdiv_rn(vx, norm); n[1] = __fdiv_rn(vy, norm); n[2] = __fdiv_rn(vz, norm);
}

__global__ void __launch_bounds__(128) surface_points_kernel(SurfaceArgs g) {
  const i64 N = (i64)g.nx * g.ny * g.nz;
  const i64 i0 = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= N) return;
  if (g.mask && __ldg(g.mask + i0) == 0.0f) return;
  const int ix = (int)(i0 % g.nx), iy = (int)((i0 / g.nx) % g.ny), iz = (int)(i0 / ((i64)g.nx * g.ny));
  float xyz[3], normal[3];
  bool keep = true;
  if (!g.labels) {
    xyz[0] = __fmul_rn((float)ix, g.vw[0]); xyz[1] = __fmul_rn((float)iy, g.vw[1]); xyz[2] = __fmul_rn((float)iz, g.vw[2]);
#pragma unroll
    for (int d = 0; d < 3; d++) normal[d] = __ldg(g.dir + 3 * i0 + d);
  } else {
    const float lab0 = __ldg(g.labels + i0);
    if (g.select != lab0) return;
    const float s0 = __ldg(g.sal + i0);
    xyz[0] = (float)ix; xyz[1] = (float)iy; xyz[2] = (float)iz;
    unit_normal(g, i0, normal);
    if (g.ds > 0.0f) {
      // ---- pass 1: weighted mean arc length over both branches ----
      double sum_s = 0.0, sum_w = 0.0;
      int n_fwd = 0, n_bwd = 0;
      Walker w(g, ix, iy, iz, lab0);
      float s = 0.0f;
      while (n_fwd < g.max_steps && w.in_cluster()) {
        const float wt = __ldg(g.sal + w.idx());
        sum_s += (double)__fmul_rn(wt, s);
        sum_w += (double)wt;
        n_fwd++;
        w.step(g.ds);
        s = __fadd_rn(s, g.ds);
      }
      w.reset(ix, iy, iz);
      s = 0.0f;
      while (n_bwd < g.max_steps) {
        w.step(-g.ds);
        s = __fsub_rn(s, g.ds);
        if (!w.in_cluster()) break;
        const float wt = __ldg(g.sal + w.idx());
        sum_s += (double)__fmul_rn(wt, s);
        sum_w += (double)wt;
        n_bwd++;
      }
      const float ave = (float)(sum_s / sum_w);
      const int n = n_bwd + n_fwd;
      // ---- the sample k the reference stops at (:2183-2188): the first k >= 1 with S[k-1] <= ave <= S[k], else n-1;
      // S[m] = (m - n_bwd) steps from the start.  Pass 2 regenerates S and the positions on the side it needs. ----
      // samples k and k+1: positions A (used for the normal and as interpolation base) and B
      float A[3] = {(float)ix, (float)iy, (float)iz}, B[3] = {0.f, 0.f, 0.f}, SA = 0.0f, SB = 0.0f;
      bool haveB = false;
      int k = 0;
      if (n > 1) {
        // arc lengths by the same recurrences as pass 1
        // search over m = 1 .. n-1
        k = n - 1;
        bool found = false;
        // backward side, m = 1 .. n_bwd: S[m] = -T(n_bwd - m), T(j) = j-fold recurrence of ds
        if (n_bwd > 0) {
          // walk the backward branch outwards keeping the arc lengths in registers is impossible in reverse,
          // so test the condition from the far end by regenerating T(j): O(n_bwd^2) adds, n_bwd is a few dozen
          for (int m = 1; m <= n_bwd && !found; m++) {
            float Sm1 = 0.0f, Sm = 0.0f;
            for (int j = 0; j < n_bwd - m + 1; j++) { Sm = Sm1; Sm1 = __fsub_rn(Sm1, g.ds); }
            // after the loop: Sm1 = S[m-1] = -T(n_bwd-m+1), Sm = S[m] = -T(n_bwd-m)
            if (Sm1 <= ave && ave <= Sm) { k = m; found = true; }
          }
        }
        if (!found) {
          float Sm1 = 0.0f;   // S[n_bwd] = 0
          for (int m = n_bwd + 1; m <= n - 1; m++) {
            const float Sm = __fadd_rn(Sm1, g.ds);
            if (Sm1 <= ave && ave <= Sm) { k = m; found = true; break; }
            Sm1 = Sm;
          }
        }
      }
      // positions and arc lengths of samples k and k+1
      {
        auto sample = [&](int m, float X[3], float &S) {
          Walker q(g, ix, iy, iz, lab0);
          float sv = 0.0f;
          if (m >= n_bwd) {
            for (int j = 0; j < m - n_bwd; j++) { q.step(g.ds); sv = __fadd_rn(sv, g.ds); }
          } else {
            for (int j = 0; j < n_bwd - m; j++) { q.step(-g.ds); sv = __fsub_rn(sv, g.ds); }
          }
          X[0] = q.r[0]; X[1] = q.r[1]; X[2] = q.r[2];
          S = sv;
        };
        sample(k, A, SA);
        if (k + 1 < n) { sample(k + 1, B, SB); haveB = true; }
      }
      int q[3];
      q[0] = min(max((int)roundf(A[0]), 0), g.nx - 1);
      q[1] = min(max((int)roundf(A[1]), 0), g.ny - 1);
      q[2] = min(max((int)roundf(A[2]), 0), g.nz - 1);
      unit_normal(g, ((i64)q[2] * g.ny + q[1]) * g.nx + q[0], normal);
#pragma unroll
      for (int d = 0; d < 3; d++) {
        if (haveB) xyz[d] = __fadd_rn(A[d], __fmul_rn(__fsub_rn(B[d], A[d]), __fdiv_rn(__fsub_rn(ave, SA), __fsub_rn(SB, SA))));
        else xyz[d] = A[d];
      }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) normal[d] = __fmul_rn(normal[d], s0);   // magnitude = saliency of the ORIGINAL voxel
    if (g.find_ridge) {
      int q[3];
      q[0] = min(max((int)roundf(xyz[0]), 0), g.nx - 1);
      q[1] = min(max((int)roundf(xyz[1]), 0), g.ny - 1);
      q[2] = min(max((int)roundf(xyz[2]), 0), g.nz - 1);
      int x = q[0], y = q[1], z = q[2];
      if (x == 0) x++; else if (x == g.nx - 1) x--;
      if (y == 0) y++; else if (y == g.ny - 1) y--;
      if (z == 0) z++; else if (z == g.nz - 1) z--;
      const i64 sy = g.nx, sz = (i64)g.nx * g.ny;
      const float *c = g.sal + ((i64)z * g.ny + y) * g.nx + x;
#define F(a, b, cc) __ldg(c + (a) + (b) * sy + (cc) * sz)
      const float ctr = F(0, 0, 0), c2 = __fmul_rn(2.0f, ctr);
      const float gr[3] = {__fmul_rn(0.5f, __fsub_rn(F(1, 0, 0), F(-1, 0, 0))), __fmul_rn(0.5f, __fsub_rn(F(0, 1, 0), F(0, -1, 0))),
                           __fmul_rn(0.5f, __fsub_rn(F(0, 0, 1), F(0, 0, -1)))};
      const float hxx = __fsub_rn(__fadd_rn(F(1, 0, 0), F(-1, 0, 0)), c2);
      const float hyy = __fsub_rn(__fadd_rn(F(0, 1, 0), F(0, -1, 0)), c2);
      const float hzz = __fsub_rn(__fadd_rn(F(0, 0, 1), F(0, 0, -1)), c2);
      const float hxy = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(F(1, 1, 0), F(-1, -1, 0)), F(1, -1, 0)), F(-1, 1, 0)));
      const float hyz = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(F(0, 1, 1), F(0, -1, -1)), F(0, 1, -1)), F(0, -1, 1)));
      const float hxz = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(F(1, 0, 1), F(-1, 0, -1)), F(-1, 0, 1)), F(1, 0, -1)));
#undef F
      const Sym3d m = {hxx, hyy, hzz, hxy, hyz, hxz};
      // DECREASING_ABS_EIVALS (eigen3_simple.hpp:252-264): increasing, then first and last exchanged when
      // |first| < |last|: the first eigenpair is the extreme eigenvalue of larger magnitude
      double ev[3], e_lo[3];
      sym3_eigen_first(m, 0, ev, e_lo);
      double v1d[3];
      float l1;
      if (fabs(ev[0]) < fabs(ev[2])) {
        double evd[3];
        sym3_eigen_first(m, 1, evd, v1d);
        l1 = (float)evd[0];
      } else {
        v1d[0] = e_lo[0]; v1d[1] = e_lo[1]; v1d[2] = e_lo[2];
        l1 = (float)ev[0];
      }
      float v1[3] = {(float)v1d[0], (float)v1d[1], (float)v1d[2]};
      float along = __fadd_rn(__fadd_rn(__fmul_rn(gr[0], v1[0]), __fmul_rn(gr[1], v1[1])), __fmul_rn(gr[2], v1[2]));
      if (along < 0.0f) {
        along = -along;
        v1[0] = -v1[0]; v1[1] = -v1[1]; v1[2] = -v1[2];
      } else if (along == 0.0f) {
        keep = false;
      }
      const float dist = (l1 != 0.0f) ? __fdiv_rn(along, l1) : __int_as_float(0x7f800000);
      if (g.max_distance > 0.0f && fabsf(dist) > g.max_distance) keep = false;
#pragma unroll
      for (int d = 0; d < 3; d++) xyz[d] = __fsub_rn((float)q[d], __fmul_rn(dist, v1[d]));
      if (xyz[0] < 0.0f || (float)g.nx < xyz[0] || xyz[1] < 0.0f || (float)g.ny < xyz[1] || xyz[2] < 0.0f || (float)g.nz < xyz[2])
        keep = false;
#pragma unroll
      for (int d = 0; d < 3; d++) xyz[d] = __fmul_rn(xyz[d], g.vw[d]);
    }
  }
  if (!keep) return;
  const unsigned long long slot = atomicAdd(g.count, 1ULL);
  if (slot < g.capacity) {
    g.index[slot] = i0;
    float *o = g.rows + 6 * slot;
    o[0] = xyz[0]; o[1] = xyz[1]; o[2] = xyz[2];
    o[3] = normal[0]; o[4] = normal[1]; o[5] = normal[2];
  }
}

}  // namespace visfd_cuda

using namespace visfd_cuda;

extern "C" int visfd_cuda_surface_points(visfd_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, const float *saliency,
                                         const float *direction, const float *labels, const float *mask,
                                         int select_cluster, const float voxel_width[3], float curve_ds, int find_ridge,
                                         float max_distance, float *rows, int64_t capacity, int64_t *n_points) {
  try {
    VREQUIRE(ctx != nullptr, "context is NULL");
    VCK(cudaSetDevice(ctx->device));
    drop_pending_stage_events(ctx);
    VREQUIRE(nx > 0 && ny > 0 && nz > 0 && nx < (1LL << 31) && ny < (1LL << 31) && nz < (1LL << 31), "bad image dimensions");
    VREQUIRE(direction && n_points && voxel_width && capacity >= 0 && (rows || capacity == 0), "NULL argument");
    VREQUIRE(!labels || saliency, "a selected cluster needs the saliency image");
    VREQUIRE(!(labels && find_ridge) || (nx >= 3 && ny >= 3 && nz >= 3), "ridge refinement needs an image at least 3 voxels wide");
    const size_t N = (size_t)nx * ny * nz;
    const bool host = !is_device_pointer(direction);
    Staged<float> s(ctx, saliency, N, Dir::In, host), d(ctx, direction, 3 * N, Dir::In, host), l(ctx, labels, N, Dir::In, host),
        m(ctx, mask, N, Dir::In, host);
    size_t cap = std::max<size_t>((size_t)capacity, 1);
    Scratch<long long> index;
    Scratch<float> drows;
    Scratch<unsigned long long> count(ctx, 1);
    SurfaceArgs g;
    g.sal = s.get(); g.dir = d.get(); g.labels = l.get(); g.mask = m.get();
    g.nx = (int)nx; g.ny = (int)ny; g.nz = (int)nz;
    g.select = (float)select_cluster;
    for (int k = 0; k < 3; k++) g.vw[k] = voxel_width[k];
    g.ds = curve_ds; g.find_ridge = find_ridge; g.max_distance = max_distance;
    // a zero or NaN direction keeps the reference walking on the spot for ever: bounded here
    g.max_steps = curve_ds > 0.0f ? (int)std::min(4.0 * (double)(nx + ny + nz) / curve_ds + 16.0, 1.0e9) : 0;
    g.count = count.get();
    unsigned long long n = 0;
    // The list is appended in no particular order and sorted afterwards; when more points are found than the
    // caller has room for, the pass is repeated with room for all of them so that the rows returned are the
    // FIRST `capacity` in the reference's order.
    for (int pass = 0; pass < 2; pass++) {
      index.reset(ctx, cap);
      drows.reset(ctx, 6 * cap);
      g.capacity = (unsigned long long)cap;
      g.index = index.get(); g.rows = drows.get();
      VCK(cudaMemsetAsync(count.get(), 0, sizeof(unsigned long long), ctx->stream));
      surface_points_kernel<<<(unsigned)((N + 127) / 128), 128, 0, ctx->stream>>>(g);
      VCK(cudaGetLastError());
      ctx->count_launch();
      VCK(cudaMemcpyAsync(&n, count.get(), sizeof(n), cudaMemcpyDeviceToHost, ctx->stream));
      VCK(cudaStreamSynchronize(ctx->stream));
      if (n <= cap) break;
      cap = (size_t)n;
    }
    *n_points = (int64_t)n;
    const size_t keep = (size_t)std::min<unsigned long long>(n, (unsigned long long)capacity);
    const size_t k = keep ? (size_t)n : 0;
    if (k) {
      std::vector<long long> hi(k);
      std::vector<float> hr(6 * k);
      VCK(cudaMemcpyAsync(hi.data(), index.get(), k * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
      VCK(cudaMemcpyAsync(hr.data(), drows.get(), 6 * k * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
      VCK(cudaStreamSynchronize(ctx->stream));
      // raster order of the source voxel = the order of the reference's loops (:2050-2052)
      std::vector<size_t> order(k);
      std::iota(order.begin(), order.end(), (size_t)0);
      std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return hi[a] < hi[b]; });
      std::vector<float> sorted(6 * keep);
      for (size_t j = 0; j < keep; j++) std::copy(hr.begin() + 6 * order[j], hr.begin() + 6 * order[j] + 6, sorted.begin() + 6 * j);
      if (is_device_pointer(rows))
        VCK(cudaMemcpyAsync(rows, sorted.data(), 6 * keep * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
      else
        std::copy(sorted.begin(), sorted.end(), rows);
      VCK(cudaStreamSynchronize(ctx->stream));
    }
    resolve_stage_times(ctx);
    return 0;
  } catch (const std::exception &ex) {
    set_last_error(ex.what());
    if (ctx) { cudaStreamSynchronize(ctx->stream); cudaGetLastError(); }
    return 1;
  }
}
