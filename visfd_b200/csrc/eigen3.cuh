// eigen3.cuh -- 3x3 symmetric eigen-decomposition on the device, in double.
//
// Same closed form as lib/visfd/eigen3_simple.hpp: shift by trace/3 and scale by
// max|a_ij| (:150-163), trigonometric roots of the characteristic cubic (computeRoots3,
// :49-82), eigenvectors as the larger cross product of two columns of A - lambda*I
// (extract_kernel3, :88-133), starting from the better separated end of the spectrum
// (:196-243), then a first/last swap for the requested order (:252-264).  The caller
// (filter_mrc) only ever consumes the eigenvalues and the first eigenvector, so those are
// what the entry points below return; the float Shoemake round trip of
// eigen3_simple.hpp:316-340 / lin3_utils.hpp:567-584 only adds ~1e-7 noise and a sign
// convention (normals are compared up to sign) and is not reproduced.
#pragma once
#include <cuda_runtime.h>

namespace visfd_cuda {

// flat order xx,yy,zz,xy,yz,xz (lib/visfd/lin3_utils.hpp:400-406)
struct Sym3d {
  double xx, yy, zz, xy, yz, xz;
};

__device__ __forceinline__ void cross3d(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3d(const double a[3], const double b[3]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

// Shift/scale + roots.  On return ev[0] <= ev[1] <= ev[2] are the eigenvalues of the
// SCALED, SHIFTED matrix s (returned through s, shift, scale).
__device__ __forceinline__ void sym3_roots(const Sym3d &m, Sym3d &s, double &shift, double &scale,
                                           double ev[3]) {
  shift = (m.xx + m.yy + m.zz) / 3.0;
  s = m;
  s.xx -= shift;
  s.yy -= shift;
  s.zz -= shift;
  scale = fmax(fmax(fabs(s.xx), fabs(s.yy)), fmax(fabs(s.zz), fmax(fabs(s.xy), fmax(fabs(s.yz), fabs(s.xz)))));
  if (scale > 0.0) {
    double si = 1.0 / scale;
    s.xx *= si; s.yy *= si; s.zz *= si; s.xy *= si; s.yz *= si; s.xz *= si;
  }
  const double inv3 = 1.0 / 3.0;
  const double sqrt3 = 1.7320508075688772;
  // m[1][0]=xy, m[2][0]=xz, m[2][1]=yz
  double c0 = s.xx * s.yy * s.zz + 2.0 * s.xy * s.xz * s.yz - s.xx * s.yz * s.yz -
              s.yy * s.xz * s.xz - s.zz * s.xy * s.xy;
  double c1 = s.xx * s.yy - s.xy * s.xy + s.xx * s.zz - s.xz * s.xz + s.yy * s.zz - s.yz * s.yz;
  double c2 = s.xx + s.yy + s.zz;
  double c2_3 = c2 * inv3;
  double a_3 = (c2 * c2_3 - c1) * inv3;
  a_3 = fmax(a_3, 0.0);
  double half_b = 0.5 * (c0 + c2_3 * (2.0 * c2_3 * c2_3 - c1));
  double q = a_3 * a_3 * a_3 - half_b * half_b;
  q = fmax(q, 0.0);
  double rho = sqrt(a_3);
  double theta = atan2(sqrt(q), half_b) * inv3;
  double st, ct;
  sincos(theta, &st, &ct);
  ev[0] = c2_3 - rho * (ct + sqrt3 * st);
  ev[1] = c2_3 - rho * (ct - sqrt3 * st);
  ev[2] = c2_3 + 2.0 * rho * ct;
}

// extract_kernel3 (eigen3_simple.hpp:88-133) for the symmetric matrix t.
// res: unit null vector estimate; rep: the column with the largest diagonal entry.
__device__ __forceinline__ void sym3_kernel(const Sym3d &t, double res[3], double rep[3]) {
  double col[3][3] = {{t.xx, t.xy, t.xz}, {t.xy, t.yy, t.yz}, {t.xz, t.yz, t.zz}};
  int i0 = 0;
  double md = fabs(t.xx);
  if (fabs(t.yy) > md) { i0 = 1; md = fabs(t.yy); }
  if (fabs(t.zz) > md) { i0 = 2; }
  int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
  rep[0] = col[i0][0]; rep[1] = col[i0][1]; rep[2] = col[i0][2];
  double a[3] = {col[i1][0], col[i1][1], col[i1][2]};
  double b[3] = {col[i2][0], col[i2][1], col[i2][2]};
  double c0[3], c1[3];
  cross3d(rep, a, c0);
  cross3d(rep, b, c1);
  double n0 = dot3d(c0, c0), n1 = dot3d(c1, c1);
  if (n0 > n1) {
    double si = 1.0 / sqrt(n0);
    res[0] = c0[0] * si; res[1] = c0[1] * si; res[2] = c0[2] * si;
  } else {
    double si = 1.0 / sqrt(n1);
    res[0] = c1[0] * si; res[1] = c1[1] * si; res[2] = c1[2] * si;
  }
}

__device__ __forceinline__ void normalize3d(double a[3]) {  // lin3_utils.hpp:142-155
  double L = sqrt(dot3d(a, a));
  if (L > 0.0) {
    L = 1.0 / L;
    a[0] *= L; a[1] *= L; a[2] *= L;
  } else {
    a[0] = 1.0; a[1] = 0.0; a[2] = 0.0;
  }
}

// ---- eigenvalues without transcendental functions ------------------------------------------------
// The closed form above spends ~500 instructions per matrix in double-precision atan2 / sincos / sqrt /
// division sequences.  The same numbers (to ~1e-12 of the matrix scale) come from the characteristic
// polynomial itself: with s = m - (trace/3) I, A = -(sum of principal 2x2 minors of s) >= 0 and det = |s|,
// the shifted eigenvalues are the roots of y^3 - A y - det = 0.  Writing y = 2 rho t (rho^2 = A/3) gives
// 4 t^3 - 3 t = k, k = det / (2 rho^3) in [-1, 1].  For k >= 0 the LARGEST root is simple and well
// separated (t in [sqrt(3)/2, 1], derivative >= 6); for k < 0 the same holds for the smallest root, which
// is the largest root of the problem with det -> -det.  So:
//   1. t0 = P4(|k|), a degree-4 fit of cos(acos(k)/3) on [0, 1] (max error 4.5e-6), with 1/rho from
//      MUFU.RSQ64H (rsqrt.approx.f64: no float<->double conversion, which costs as much as four DFMAs here);
//   2. two Newton steps on y^3 - A y - |det| in double, the reciprocal of the derivative once from
//      MUFU.RCP64H (error 5e-6 -> ~1e-10 -> rounding level);
//   3. the other two roots from the deflated quadratic, y = (-z -+ sqrt(4A - 3 z^2)) / 2, its square root
//      by MUFU.RSQ64H and one Heron step.  (A nearly double root at this end is as ill-conditioned in the
//      reference's formula -- sqrt(q), q = a^3 - b^2 -- as it is here: ~1e-8 of the scale in double.)
// ~58 FP64 instructions and 3 MUFU; no division, no branch.  Measured against the closed form: see
// tests/test_gpu_parity.py::test_newton_eigenvalues_match_closed_form.
__device__ __forceinline__ double rsqrt_approx64(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double rcp_approx64(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}

// ev[0] <= ev[1] <= ev[2]
__device__ __forceinline__ void sym3_eigenvalues_newton(const Sym3d &m, double ev[3]) {
  const double TINY = 1e-300;   // keeps 1/rho, 1/g' and 1/sqrt(D) finite for multiples of the identity
  const double shift = (m.xx + m.yy + m.zz) * (1.0 / 3.0);
  const double sx = m.xx - shift, sy = m.yy - shift, sz = m.zz - shift;
  const double A = fma(0.5, fma(sz, sz, fma(sy, sy, sx * sx)), fma(m.xz, m.xz, fma(m.yz, m.yz, m.xy * m.xy)));
  const double det = fma(sx, fma(sy, sz, -m.yz * m.yz),
                         fma(m.xy, fma(m.yz, m.xz, -m.xy * sz), m.xz * fma(m.xy, m.yz, -sy * m.xz)));
  const double A3 = fma(A, 1.0 / 3.0, TINY);               // rho^2
  const double rinv = rsqrt_approx64(A3);                  // 1/rho, 20 bits
  const double ad = fabs(det);
  const double k = (ad * (0.5 * rinv)) * (rinv * rinv);    // in [0, 1 + 3e-6]
  // 2 cos(acos(k)/3), k in [0, 1]
  const double t2 = fma(fma(fma(fma(-0.008198810912827108, k, 0.03528472977563763), k, -0.09201052271579164), k,
                            0.33285803676124636), k, 1.7320597067184762);
  double z = (A3 * rinv) * t2;                             // rho * 2t: within 5e-6 of the largest root
  // Two Newton steps with the same reciprocal derivative (g' >= 6 rho^2 near the root): 5e-6 -> 1e-10 -> rounding.
  // The second one matters when the OTHER two roots nearly coincide: their split sqrt(D) amplifies the error
  // of z by z / sqrt(D).
  double z2 = z * z;
  const double r = rcp_approx64(fma(3.0, z2, TINY - A));
  z = fma(-fma(z2 - A, z, -ad), r, z);
  z2 = z * z;
  z = fma(-fma(z2 - A, z, -ad), r, z);
  const double D = fabs(fma(-3.0 * z, z, 4.0 * A)) + TINY; // (v - u)^2
  const double rs = rsqrt_approx64(D);
  double x = D * rs;
  x = fma(0.5 * rs, fma(-x, x, D), x);                     // sqrt(D): one Heron step, 2^-40
  // roots of the |det| problem: u = -(z + x)/2 <= v = (x - z)/2 <= z; det < 0 mirrors them
  const bool neg = det < 0.0;
  const double sg = neg ? -1.0 : 1.0, hs = neg ? -0.5 : 0.5;
  const double ez = fma(sg, z, shift), ev_ = fma(hs, x - z, shift), eu = fma(-hs, z + x, shift);
  ev[0] = neg ? ez : eu;
  ev[1] = ev_;
  ev[2] = neg ? eu : ez;
}

// Eigenvalues only, in the requested order (0 increasing, 1 decreasing).
__device__ __forceinline__ void sym3_eigenvalues_closed_form(const Sym3d &m, int order, double ev[3]) {
  Sym3d s;
  double shift, scale;
  sym3_roots(m, s, shift, scale, ev);
  for (int d = 0; d < 3; d++) ev[d] = ev[d] * scale + shift;
  if ((order == 0 && ev[0] > ev[2]) || (order == 1 && ev[0] < ev[2])) {
    double t = ev[0]; ev[0] = ev[2]; ev[2] = t;
  }
}
__device__ __forceinline__ void sym3_eigenvalues(const Sym3d &m, int order, double ev[3]) {
  sym3_eigenvalues_newton(m, ev);   // ascending
  // decreasing order = first and last exchanged, the middle one stays (eigen3_simple.hpp:252-264)
  if (order == 1) {
    double t = ev[0]; ev[0] = ev[2]; ev[2] = t;
  }
}

// Eigenvalues in the requested order and the eigenvector that ends up FIRST
// (eivects[0] of DiagonalizeSym3, eigen3_simple.hpp:139-266).
__device__ __forceinline__ void sym3_eigen_first(const Sym3d &m, int order, double ev[3],
                                                 double first[3]) {
  const double EPS = 2.220446049250313e-16;
  Sym3d s;
  double shift, scale;
  sym3_roots(m, s, shift, scale, ev);
  double E0[3], E2[3];  // eigenvectors of ev[0] (smallest) and ev[2] (largest)
  if ((ev[2] - ev[0]) <= EPS) {
    E0[0] = 1.0; E0[1] = 0.0; E0[2] = 0.0;
    E2[0] = 0.0; E2[1] = 0.0; E2[2] = 1.0;
  } else {
    double d0 = ev[2] - ev[1], d1 = ev[1] - ev[0];
    bool top_first = d0 > d1;  // k = 2, l = 0
    if (top_first) d0 = d1;
    // after this, d0 = smaller gap, d1 = (original) ev[1]-ev[0]
    double *Ek = top_first ? E2 : E0;
    double *El = top_first ? E0 : E2;
    double evk = top_first ? ev[2] : ev[0];
    double evl = top_first ? ev[0] : ev[2];
    Sym3d t = s;
    t.xx -= evk; t.yy -= evk; t.zz -= evk;
    sym3_kernel(t, Ek, El);
    if (d0 <= 2.0 * EPS * d1) {
      // eigen3_simple.hpp:220-223 (sic: subtracts a multiple of itself)
      double kl = dot3d(Ek, El);
      El[0] -= kl * El[0]; El[1] -= kl * El[1]; El[2] -= kl * El[2];
      normalize3d(El);
    } else {
      t = s;
      t.xx -= evl; t.yy -= evl; t.zz -= evl;
      double dummy[3];
      sym3_kernel(t, El, dummy);
    }
  }
  for (int d = 0; d < 3; d++) ev[d] = ev[d] * scale + shift;
  bool swap = (order == 0 && ev[0] > ev[2]) || (order == 1 && ev[0] < ev[2]);
  if (swap) {
    double t = ev[0]; ev[0] = ev[2]; ev[2] = t;
    first[0] = E2[0]; first[1] = E2[1]; first[2] = E2[2];
  } else {
    first[0] = E0[0]; first[1] = E0[1]; first[2] = E0[2];
  }
}

// The eigenvector that DiagonalizeSym3 returns FIRST (largest eigenvalue for decreasing order, smallest for
// increasing), without the closed form: eigenvalue by sym3_eigenvalues_newton, then extract_kernel3's
// construction (eigen3_simple.hpp:88-133) on m - lambda I -- the column with the largest diagonal entry crossed
// with the other two, the longer product normalised (1/sqrt by MUFU.RSQ64H and one Newton step).  Branch-free,
// no local arrays.  Used for the ~5 % of voxels that vote; where the eigenvalue is not separated the vector is
// as arbitrary as the reference's.
__device__ __forceinline__ void sym3_first_eigenvector_newton(const Sym3d &m, int order, double first[3]) {
  double ev[3];
  sym3_eigenvalues_newton(m, ev);
  const double lam = order == 1 ? ev[2] : ev[0];
  const double xx = m.xx - lam, yy = m.yy - lam, zz = m.zz - lam;
  const double ax = fabs(xx), ay = fabs(yy), az = fabs(zz);
  const bool use_z = az > fmax(ax, ay), use_y = !use_z && ay > ax;   // first maximum wins, as the reference's loop
  // rep = column i0, a = column i0+1, b = column i0+2 (cyclic)
  const double c0[3] = {xx, m.xy, m.xz}, c1[3] = {m.xy, yy, m.yz}, c2[3] = {m.xz, m.yz, zz};
  double rep[3], a[3], b[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    rep[d] = use_z ? c2[d] : (use_y ? c1[d] : c0[d]);
    a[d] = use_z ? c0[d] : (use_y ? c2[d] : c1[d]);
    b[d] = use_z ? c1[d] : (use_y ? c0[d] : c2[d]);
  }
  double p[3], q[3];
  cross3d(rep, a, p);
  cross3d(rep, b, q);
  const double np = dot3d(p, p), nq = dot3d(q, q);
  const bool take_p = np > nq;
  const double n = (take_p ? np : nq) + 1e-300;
  double y = rsqrt_approx64(n);
  y = y * fma(-0.5 * n, y * y, 1.5);
#pragma unroll
  for (int d = 0; d < 3; d++) first[d] = (take_p ? p[d] : q[d]) * y;
}

// ScoreHessianPlanar / Linear (feature.hpp:1529-1581) and ScoreTensorPlanar / Linear
// (:1593-1612) from the FLOAT eigenvalues, evaluated in double, returned as float.
__device__ __forceinline__ float score_from_eivals(const double ev[3], int score_kind,
                                                   int is_vote_tensor) {
  double l1 = (double)(float)ev[0], l2 = (double)(float)ev[1], l3 = (double)(float)ev[2];
  double sc;
  if (score_kind == 1) {
    sc = l1 * l2 - l3 * l3;
  } else if (is_vote_tensor) {
    sc = l1 - l2;
  } else {
    sc = l1 * l1 - l2 * l2;
    sc *= sc;
  }
  return (float)sc;
}

}  // namespace visfd_cuda
