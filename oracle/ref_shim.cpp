// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" adapter that instantiates the UNMODIFIED reference templates
// (header-only library under $VISFD_REFERENCE/lib/visfd, compiled from where it
// lies) on flat float arrays, so that Python tests / the bench "reference" arm can
// call the reference's own CPU/OpenMP implementation of the hot path.
//
// Built by oracle/Makefile into oracle/_ref/libvisfd_ref.so (git-ignored).
// Nothing here re-implements reference arithmetic: every numeric result comes
// from a call into the reference headers.  The only code of our own is the
// pointer-table glue (float*** views over caller-owned flat memory) and the
// replay of the few loops that live in bin/filter_mrc/handlers.cpp (HandleTV),
// which cannot be linked because they sit inside the CLI handler; those loops
// are restated here calling the same reference functions in the same order
// (handlers.cpp:1645-1746 eigen loop, :1751-1797 cut, :1870-1892 post-vote score).

#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <array>
#include <algorithm>
#include <iostream>
#include <sstream>
using namespace std;

#include <visfd.hpp>
#include <threshold.hpp>
#include <mrc_simple.hpp>
using namespace visfd;
#include "../include/visfd_mrc.h"   // only for the plain header struct the tests pass around

namespace {

// float*** view over caller-owned contiguous memory (x fastest), matching the
// a[iz][iy][ix] contract of Alloc3D (lib/visfd/alloc3d.hpp:25-67).
template <typename T>
struct View3 {
  vector<T*>  rows;
  vector<T**> planes;
  T ***p = nullptr;
  View3(T *base, int nx, int ny, int nz) {
    if (!base) return;
    rows.resize(size_t(nz) * ny);
    planes.resize(nz);
    for (int iz = 0; iz < nz; iz++) {
      for (int iy = 0; iy < ny; iy++)
        rows[size_t(iz) * ny + iy] = base + (size_t(iz) * ny + iy) * size_t(nx);
      planes[iz] = rows.data() + size_t(iz) * ny;
    }
    p = planes.data();
  }
};

typedef array<float, 3> Vec3;

selfadjoint_eigen3::EigenOrderType order_from_int(int o) {
  return o == 0 ? selfadjoint_eigen3::INCREASING_EIVALS
                : selfadjoint_eigen3::DECREASING_EIVALS;
}

} // namespace

extern "C" {

int ref_version() { return 1; }

// lib/visfd/filter1d.hpp:411-460
void ref_gen_gauss1d(float sigma, int hw, float *taps /* 2*hw+1 */) {
  Filter1D<float, int> f = GenFilterGauss1D(sigma, hw);
  for (int i = -hw; i <= hw; i++) taps[i + hw] = f.afH[i];
}

// lib/visfd/filter3d.hpp:1088-1124 (-> ApplySeparable :688-1050)
float ref_apply_gauss(int nx, int ny, int nz, const float *src, float *dst,
                      const float *mask, const float sigma[3], const int hw[3],
                      int normalize) {
  int size[3] = {nx, ny, nz};
  View3<const float> s(src, nx, ny, nz), m(mask, nx, ny, nz);
  View3<float> d(dst, nx, ny, nz);
  return ApplyGauss(size, s.p, d.p, m.p, sigma, hw, normalize != 0,
                    (ostream *)nullptr);
}

// lib/visfd/filter3d.hpp:1340-1402
void ref_apply_dog(int nx, int ny, int nz, const float *src, float *dst,
                   const float *mask, const float sigma_a[3],
                   const float sigma_b[3], const int hw[3], float *pA,
                   float *pB) {
  int size[3] = {nx, ny, nz};
  View3<const float> s(src, nx, ny, nz), m(mask, nx, ny, nz);
  View3<float> d(dst, nx, ny, nz);
  ApplyDog(size, s.p, d.p, m.p, sigma_a, sigma_b, hw, pA, pB,
           (ostream *)nullptr);
}

// lib/visfd/filter3d.hpp:1430-1507
void ref_apply_log(int nx, int ny, int nz, const float *src, float *dst,
                   const float *mask, const float sigma[3], float delta,
                   float truncate_ratio, float *pA, float *pB) {
  int size[3] = {nx, ny, nz};
  View3<const float> s(src, nx, ny, nz), m(mask, nx, ny, nz);
  View3<float> d(dst, nx, ny, nz);
  ApplyLog(size, s.p, d.p, m.p, sigma, delta, truncate_ratio, pA, pB,
           (ostream *)nullptr);
}

// lib/visfd/feature.hpp:1210-1348.  Dense outputs: grad[N][3], hess[N][6]
// (flat order xx,yy,zz,xy,yz,xz, lin3_utils.hpp:400-406).  Entries of voxels
// with mask==0 are left as the caller initialised them.
int ref_calc_hessian(int nx, int ny, int nz, const float *src, const float *mask,
                     float sigma, float truncate_ratio, float *grad /*N*3*/,
                     float *hess /*N*6*/) {
  int size[3] = {nx, ny, nz};
  View3<const float> s(src, nx, ny, nz), m(mask, nx, ny, nz);
  View3<Vec3> g(reinterpret_cast<Vec3 *>(grad), nx, ny, nz);
  size_t N = size_t(nx) * ny * nz;
  vector<float *> hp(N);
  for (size_t i = 0; i < N; i++) hp[i] = hess + 6 * i;
  View3<float *> h(hp.data(), nx, ny, nz);
  try {
    CalcHessian(size, s.p, g.p, h.p, m.p, sigma, truncate_ratio,
                (ostream *)nullptr);
  } catch (const std::exception &e) {
    return 1;
  }
  return 0;
}

// Replays bin/filter_mrc/handlers.cpp:1645-1746 for filter_type SURFACE_RIDGE
// (score_kind 0) or CURVE (score_kind 1): per voxel ConvertFlatSym2Evects3
// (eigen3_simple.hpp:392) + ScoreHessianPlanar/Linear (feature.hpp:1529/1572).
void ref_hessian_eigen_score(int64_t N, const float *hess /*N*6*/,
                             const float *mask, int eival_order, int score_kind,
                             float *saliency /*N*/, float *direction /*N*3*/,
                             float *eivals_out /*N*3 or NULL*/) {
  auto order = order_from_int(eival_order);
#pragma omp parallel for
  for (int64_t i = 0; i < N; i++) {
    saliency[i] = 0.0;
    if (mask && mask[i] == 0.0f) continue;
    float eivals[3];
    float eivects[3][3];
    selfadjoint_eigen3::ConvertFlatSym2Evects3(hess + 6 * i, eivals, eivects,
                                               order);
    float score;
    if (score_kind == 1)
      score = ScoreHessianLinear(eivals, (float *)nullptr);
    else
      score = ScoreHessianPlanar(eivals, (float *)nullptr);
    float peak_height = 1.0;
    score *= peak_height;
    saliency[i] = score;
    direction[3 * i + 0] = eivects[0][0];
    direction[3 * i + 1] = eivects[0][1];
    direction[3 * i + 2] = eivects[0][2];
    if (eivals_out) {
      eivals_out[3 * i + 0] = eivals[0];
      eivals_out[3 * i + 1] = eivals[1];
      eivals_out[3 * i + 2] = eivals[2];
    }
  }
}

// Replays bin/filter_mrc/handlers.cpp:1751-1797 (in place); returns threshold.
float ref_saliency_cut(int64_t N, float *saliency, const float *mask,
                       float threshold_or_fraction, int is_fraction) {
  float hessian_score_threshold = threshold_or_fraction;
  if (is_fraction) {
    float hessian_score_fraction = threshold_or_fraction;
    size_t n_voxels = 0;
    for (int64_t i = 0; i < N; i++) {
      if (mask && (mask[i] == 0)) continue;
      n_voxels++;
    }
    vector<float> saliencies(n_voxels);
    size_t k = 0;
    for (int64_t i = 0; i < N; i++) {
      if (mask && (mask[i] == 0)) continue;
      saliencies[k] = saliency[i];
      k++;
    }
    sort(saliencies.rbegin(), saliencies.rend());
    k = floor(n_voxels * hessian_score_fraction);
    hessian_score_threshold = saliencies[k];
  }
  for (int64_t i = 0; i < N; i++)
    if (saliency[i] < hessian_score_threshold) saliency[i] = 0.0;
  return hessian_score_threshold;
}

// Radial decay table and unit displacement table exactly as TV3D builds them
// (feature.hpp:2419-2482 -> filter3d.hpp:546-601).  Returns hw.
// decay: (2hw+1)^3 floats [jz][jy][jx]; disp: (2hw+1)^3*3 floats (may be NULL).
int ref_tv_tables(float sigma, float cutoff_ratio, float *decay, float *disp,
                  int64_t capacity) {
  int hw = floor(sigma * cutoff_ratio);  // feature.hpp:1671
  int w = 2 * hw + 1;
  if (!decay) return hw;
  if (int64_t(w) * w * w > capacity) return -1;
  float sigmas[3] = {sigma, sigma, sigma};
  int halfwidth[3] = {hw, hw, hw};
  Filter3D<float, int> f =
      GenFilterGenGauss3D(sigmas, static_cast<float>(2.0), halfwidth);
  for (int iz = -hw; iz <= hw; iz++)
    for (int iy = -hw; iy <= hw; iy++)
      for (int ix = -hw; ix <= hw; ix++) {
        size_t o = (size_t(iz + hw) * w + (iy + hw)) * w + (ix + hw);
        decay[o] = f.aaafH[iz][iy][ix];
        if (disp) {
          // feature.hpp:2468-2482 (PrecalcDisplacement is private; same
          // expressions evaluated through the same types)
          float length = sqrt(ix * ix + iy * iy + iz * iz);
          if (length == 0) length = 1.0;
          disp[3 * o + 0] = ix / length;
          disp[3 * o + 1] = iy / length;
          disp[3 * o + 2] = iz / length;
        }
      }
  return hw;
}

// feature.hpp:1712-1901 (TVDenseStick).  tensor: dense N*6 (zero where
// mask_dst==0).
void ref_tv_dense_stick(int nx, int ny, int nz, const float *saliency,
                        const float *direction /*N*3*/, const float *mask_src,
                        const float *mask_dst, float sigma, int exponent,
                        float cutoff_ratio, int curves, int normalize,
                        float *tensor /*N*6*/) {
  int size[3] = {nx, ny, nz};
  size_t N = size_t(nx) * ny * nz;
  View3<const float> s(saliency, nx, ny, nz), ms(mask_src, nx, ny, nz),
      md(mask_dst, nx, ny, nz);
  View3<const Vec3> v(reinterpret_cast<const Vec3 *>(direction), nx, ny, nz);
  vector<float *> tp(N);
  for (size_t i = 0; i < N; i++)
    tp[i] = (mask_dst && mask_dst[i] == 0.0f) ? nullptr : tensor + 6 * i;
  memset(tensor, 0, N * 6 * sizeof(float));
  View3<float *> t(tp.data(), nx, ny, nz);
  TV3D<float, int, Vec3, float *> tv(sigma, exponent, cutoff_ratio);
  tv.TVDenseStick(size, s.p, v.p, t.p, ms.p, md.p, curves != 0, normalize != 0,
                  false, (ostream *)nullptr);
}

// Replays bin/filter_mrc/handlers.cpp:1870-1892: DiagonalizeFlatSym3
// (eigen3_simple.hpp:273) + ScoreTensorPlanar/Linear (feature.hpp:1593/1610).
// Voxels with mask==0 keep out[i].
void ref_tensor_score(int64_t N, const float *tensor /*N*6*/, const float *mask,
                      int eival_order, int score_kind, float *out) {
  auto order = order_from_int(eival_order);
#pragma omp parallel for
  for (int64_t i = 0; i < N; i++) {
    if (mask && mask[i] == 0.0f) continue;
    float diag[6];
    selfadjoint_eigen3::DiagonalizeFlatSym3(tensor + 6 * i, diag, order);
    float score;
    if (score_kind == 1)
      score = ScoreTensorLinear(diag);
    else
      score = ScoreTensorPlanar(diag);
    float peak_height = 1.0;
    score *= peak_height;
    out[i] = score;
  }
}

// Whole membrane path of HandleTV (handlers.cpp:1618-1892) without background
// subtraction: CalcHessian -> eigen/score -> cut -> TVDenseStick -> score.
// Any of the optional outputs may be NULL.  Returns the cut threshold used.
float ref_membrane(int nx, int ny, int nz, const float *src, const float *mask,
                   float sigma, float truncate_ratio, int eival_order,
                   float cut, int cut_is_fraction, float tv_sigma,
                   int tv_exponent, float tv_cutoff_ratio,
                   float *hess_saliency_out /*N, after cut*/,
                   float *direction_out /*N*3*/, float *tensor_out /*N*6*/,
                   float *out /*N*/) {
  size_t N = size_t(nx) * ny * nz;
  vector<float> grad(N * 3, 0.0f), hess(N * 6, 0.0f), sal(N, 0.0f),
      dir(N * 3, 0.0f);
  ref_calc_hessian(nx, ny, nz, src, mask, sigma, truncate_ratio, grad.data(),
                   hess.data());
  // aaaafDirection aliases aaaafGradient (handlers.cpp:1633): masked voxels keep
  // the (uninitialised there, zero here) gradient storage.
  dir = grad;
  ref_hessian_eigen_score(N, hess.data(), mask, eival_order, 0, sal.data(),
                          dir.data(), nullptr);
  float thr = ref_saliency_cut(N, sal.data(), mask, cut, cut_is_fraction);
  if (hess_saliency_out) memcpy(hess_saliency_out, sal.data(), N * sizeof(float));
  if (direction_out) memcpy(direction_out, dir.data(), 3 * N * sizeof(float));
  memcpy(out, sal.data(), N * sizeof(float));
  if (tv_sigma > 0.0f) {
    vector<float> tensor(N * 6, 0.0f);
    ref_tv_dense_stick(nx, ny, nz, sal.data(), dir.data(), mask, mask, tv_sigma,
                       tv_exponent, tv_cutoff_ratio, 0, 0, tensor.data());
    ref_tensor_score(N, tensor.data(), mask, eival_order, 0, out);
    if (tensor_out) memcpy(tensor_out, tensor.data(), 6 * N * sizeof(float));
  }
  return thr;
}

// lib/threshold/threshold.hpp:52-77, 117-169 applied element-wise as
// HandleThresholds does (handlers.cpp:1037-1080).
void ref_threshold2(int64_t N, const float *in, float *out, float a, float b,
                    float outA, float outB) {
  for (int64_t i = 0; i < N; i++) out[i] = Threshold2(in[i], a, b, outA, outB);
}
void ref_threshold4(int64_t N, const float *in, float *out, float a01, float b01,
                    float a10, float b10, float outA, float outB) {
  for (int64_t i = 0; i < N; i++)
    out[i] = Threshold4(in[i], a01, b01, a10, b10, outA, outB);
}
// handlers.cpp:1049-1053: single threshold (a == b)
void ref_threshold1(int64_t N, const float *in, float *out, float a, float outA,
                    float outB) {
  for (int64_t i = 0; i < N; i++) out[i] = (in[i] > a) ? outB : outA;
}
// lib/visfd/visfd_utils.hpp:685-790
float ref_average(int nx, int ny, int nz, const float *in, const float *w) {
  int size[3] = {nx, ny, nz};
  View3<const float> a(in, nx, ny, nz), ww(w, nx, ny, nz);
  return AverageArr(size, a.p, ww.p);
}
float ref_stddev(int nx, int ny, int nz, const float *in, const float *w) {
  int size[3] = {nx, ny, nz};
  View3<const float> a(in, nx, ny, nz), ww(w, nx, ny, nz);
  return StdDevArr(size, a.p, ww.p);
}

// lib/visfd/resample.hpp:53-166.  size_* = {nx, ny, nz} as int64 (converted to the
// reference's int); returns 1 if the reference throws.
int ref_bin3d(const int64_t size_src[3], const int64_t size_dst[3], const float *src, float *dst,
              const int *offset) {
  int ss[3] = {(int)size_src[0], (int)size_src[1], (int)size_src[2]};
  int ds[3] = {(int)size_dst[0], (int)size_dst[1], (int)size_dst[2]};
  View3<const float> a(src, ss[0], ss[1], ss[2]);
  View3<float> b(dst, ds[0], ds[1], ds[2]);
  try {
    BinArray3D(ss, ds, a.p, b.p, offset);
  } catch (const std::exception &) {
    return 1;
  }
  return 0;
}
int ref_unbin3d(const int64_t size_src[3], const int64_t size_dst[3], const float *src, float *dst,
                const int *offset) {
  int ss[3] = {(int)size_src[0], (int)size_src[1], (int)size_src[2]};
  int ds[3] = {(int)size_dst[0], (int)size_dst[1], (int)size_dst[2]};
  View3<const float> a(src, ss[0], ss[1], ss[2]);
  View3<float> b(dst, ds[0], ds[1], ds[2]);
  try {
    UnbinArray3D(ss, ds, a.p, b.p, offset);
  } catch (const std::exception &) {
    return 1;
  }
  return 0;
}

// The clustering step of HandleTV (bin/filter_mrc/handlers.cpp:1927-2034): the first eigenvector of every
// vote tensor becomes the voxel's direction (:1933-1950), then LabelConnected (lib/visfd/connect.hpp:171)
// with the arguments filter_mrc passes (:1963-1993: unsigned dot products, positive-definite tensors,
// connectivity 1, clusters sorted by size, maxima as seeds, no must-link constraints).  labels: cluster
// id from 1 in decreasing size, -1 where undefined.  Returns the number of clusters.  SURVEY 8f rank 1:
// the known answers for a future device implementation come from here.
int64_t ref_label_connected(int nx, int ny, int nz, const float *saliency, const float *mask, const float *tensor,
                            int eival_order, float threshold_saliency, float threshold_vector_saliency,
                            float threshold_vector_neighbor, float threshold_tensor_saliency,
                            float threshold_tensor_neighbor, int64_t *labels, float *direction_out) {
  int size[3] = {nx, ny, nz};
  const size_t N = size_t(nx) * ny * nz;
  vector<Vec3> dir(N);
  vector<float> tcopy(tensor, tensor + 6 * N);
  vector<float *> tptr(N);
  for (size_t i = 0; i < N; i++) {
    tptr[i] = tcopy.data() + 6 * i;
    float eivals[3], eivects[3][3];
    ConvertFlatSym2Evects3(tptr[i], eivals, eivects, order_from_int(eival_order));
    for (int d = 0; d < 3; d++) dir[i][d] = eivects[0][d];
  }
  vector<ptrdiff_t> lab(N);
  View3<const float> sal(saliency, nx, ny, nz), m(mask, nx, ny, nz);
  View3<Vec3> dtab(dir.data(), nx, ny, nz);
  View3<float *> ttab(tptr.data(), nx, ny, nz);
  View3<ptrdiff_t> ltab(lab.data(), nx, ny, nz);
  vector<array<float, 3> > centers;
  vector<float> sizes, saliencies;
  size_t n = LabelConnected(size, sal.p, ltab.p, m.p, threshold_saliency, dtab.p, threshold_vector_saliency,
                            threshold_vector_neighbor, false, ttab.p, threshold_tensor_saliency,
                            threshold_tensor_neighbor, true, 1, static_cast<ptrdiff_t>(-1), &centers, &sizes,
                            &saliencies, RegionSortCriteria::SORT_BY_SIZE, static_cast<float ***>(nullptr),
#ifndef DISABLE_STANDARDIZE_VECTOR_DIRECTION
                            dtab.p,
#endif
                            static_cast<const vector<vector<array<float, 3> > > *>(nullptr),
                            static_cast<const vector<vector<DirectionPairType> > *>(nullptr), true,
                            static_cast<ostream *>(nullptr));
  for (size_t i = 0; i < N; i++) labels[i] = (int64_t)lab[i];
  if (direction_out) memcpy(direction_out, dir.data(), 3 * N * sizeof(float));
  return (int64_t)n;
}

// lib/visfd/draw.hpp:90-237.  regions: n records of {int32 type (0 rect, 1 sphere), float p[6],
// float value}; rect p = xmin,xmax,ymin,ymax,zmin,zmax; sphere p = x0,y0,z0,r.
struct RegionRecord { int32_t type; float p[6]; float value; };
int ref_draw_regions(int nx, int ny, int nz, float *image, const float *mask, const RegionRecord *regions, int n,
                     int negative_means_subtract) {
  int size[3] = {nx, ny, nz};
  View3<float> img(image, nx, ny, nz);
  View3<const float> m(mask, nx, ny, nz);
  vector<SimpleRegion<float> > v((size_t)n);
  for (int i = 0; i < n; i++) {
    v[i].value = regions[i].value;
    if (regions[i].type == 1) {
      v[i].type = SimpleRegion<float>::SPHERE;
      v[i].data.sphere.x0 = regions[i].p[0];
      v[i].data.sphere.y0 = regions[i].p[1];
      v[i].data.sphere.z0 = regions[i].p[2];
      v[i].data.sphere.r = regions[i].p[3];
    } else {
      v[i].type = SimpleRegion<float>::RECT;
      v[i].data.rect.xmin = regions[i].p[0];
      v[i].data.rect.xmax = regions[i].p[1];
      v[i].data.rect.ymin = regions[i].p[2];
      v[i].data.rect.ymax = regions[i].p[3];
      v[i].data.rect.zmin = regions[i].p[4];
      v[i].data.rect.zmax = regions[i].p[5];
    }
  }
  DrawRegions(size, img.p, m.p, v, negative_means_subtract != 0);
  return 0;
}

// Blob list post-processing (lib/visfd/feature.hpp:521-616, :723-913, :926-969), in place on
// flat arrays; returns the new length.
namespace {
struct RefList {
  vector<array<float, 3> > crds;
  vector<float> diam, score;
  RefList(int64_t n, const float *c, const float *d, const float *s) : crds(n), diam(d, d + n), score(s, s + n) {
    for (int64_t i = 0; i < n; i++) crds[i] = {c[3 * i], c[3 * i + 1], c[3 * i + 2]};
  }
  int64_t store(float *c, float *d, float *s) {
    for (size_t i = 0; i < crds.size(); i++) {
      c[3 * i] = crds[i][0]; c[3 * i + 1] = crds[i][1]; c[3 * i + 2] = crds[i][2];
      d[i] = diam[i]; s[i] = score[i];
    }
    return (int64_t)crds.size();
  }
};
}
int64_t ref_blobs_sort(int64_t n, float *c, float *d, float *s, int criteria, int ascending) {
  RefList l(n, c, d, s);
  SortBlobs(l.crds, l.diam, l.score, (SortCriteria)criteria, ascending != 0);
  return l.store(c, d, s);
}
int64_t ref_blobs_discard_masked(int64_t n, float *c, float *d, float *s, const float *mask, int64_t nx, int64_t ny,
                                 int64_t nz) {
  RefList l(n, c, d, s);
  View3<const float> m(mask, (int)nx, (int)ny, (int)nz);
  DiscardMaskedBlobs(l.crds, l.diam, l.score, m.p);
  return l.store(c, d, s);
}
int64_t ref_blobs_discard_overlapping(int64_t n, float *c, float *d, float *s, float sep, float large, float small_,
                                      int criteria) {
  RefList l(n, c, d, s);
  DiscardOverlappingBlobs(l.crds, l.diam, l.score, sep, large, small_, (SortCriteria)criteria);
  return l.store(c, d, s);
}

// lib/mrc_simple: MrcSimple::Read(file name, rescale=false) and MrcSimple::Write(file name),
// the header handed over in the plain struct of include/visfd_mrc.h.
static void header_out(const MrcHeader &m, visfd_mrc_header *h) {
  memset(h, 0, sizeof *h);
  for (int d = 0; d < 3; d++) {
    h->nvoxels[d] = m.nvoxels[d]; h->nstart[d] = m.nstart[d]; h->mvoxels[d] = m.mvoxels[d];
    h->cellA[d] = m.cellA[d]; h->cellB[d] = m.cellB[d]; h->mapCRS[d] = m.mapCRS[d]; h->origin[d] = m.origin[d];
  }
  h->mode = m.mode; h->dmin = m.dmin; h->dmax = m.dmax; h->dmean = m.dmean; h->ispg = m.ispg; h->nsymbt = m.nsymbt;
  memcpy(h->extra_raw_data, m.extra_raw_data, sizeof h->extra_raw_data);
  memcpy(h->remaining_raw_data, m.remaining_raw_data, sizeof h->remaining_raw_data);
  h->use_signed_bytes = m.use_signed_bytes ? 1 : 0;
}
static void header_in(const visfd_mrc_header *h, MrcHeader &m) {
  for (int d = 0; d < 3; d++) {
    m.nvoxels[d] = h->nvoxels[d]; m.nstart[d] = h->nstart[d]; m.mvoxels[d] = h->mvoxels[d];
    m.cellA[d] = h->cellA[d]; m.cellB[d] = h->cellB[d]; m.mapCRS[d] = h->mapCRS[d]; m.origin[d] = h->origin[d];
  }
  m.mode = h->mode; m.dmin = h->dmin; m.dmax = h->dmax; m.dmean = h->dmean; m.ispg = h->ispg; m.nsymbt = h->nsymbt;
  memcpy(m.extra_raw_data, h->extra_raw_data, sizeof h->extra_raw_data);
  memcpy(m.remaining_raw_data, h->remaining_raw_data, sizeof h->remaining_raw_data);
  m.use_signed_bytes = h->use_signed_bytes != 0;
}
int ref_mrc_read(const char *path, visfd_mrc_header *h, float *voxels, int64_t capacity) {
  try {
    MrcSimple t;
    t.Read(string(path), false);
    header_out(t.header, h);
    const int64_t n = (int64_t)t.header.nvoxels[0] * t.header.nvoxels[1] * t.header.nvoxels[2];
    if (n > capacity) return 2;
    for (int iz = 0; iz < t.header.nvoxels[2]; iz++)
      for (int iy = 0; iy < t.header.nvoxels[1]; iy++)
        for (int ix = 0; ix < t.header.nvoxels[0]; ix++)
          voxels[((int64_t)iz * t.header.nvoxels[1] + iy) * t.header.nvoxels[0] + ix] = t.aaafI[iz][iy][ix];
  } catch (const std::exception &) {
    return 1;
  }
  return 0;
}
int ref_mrc_write(const char *path, visfd_mrc_header *h, const float *voxels) {
  try {
    MrcSimple t;
    t.Resize(h->nvoxels);
    header_in(h, t.header);
    for (int iz = 0; iz < h->nvoxels[2]; iz++)
      for (int iy = 0; iy < h->nvoxels[1]; iy++)
        for (int ix = 0; ix < h->nvoxels[0]; ix++)
          t.aaafI[iz][iy][ix] = voxels[((int64_t)iz * h->nvoxels[1] + iy) * h->nvoxels[0] + ix];
    t.Write(string(path));
    header_out(t.header, h);
  } catch (const std::exception &) {
    return 1;
  }
  return 0;
}

// lib/visfd/feature.hpp:56-427 (BlobDog).  Results are returned through
// caller-provided buffers of `capacity` entries; returns counts via n_min/n_max
// (clipped to capacity).  Order of the lists is thread-dependent in the
// reference; callers sort before comparing.
void ref_blob_dog(int nx, int ny, int nz, const float *src, const float *mask,
                  const float *sigmas, int n_sigmas, float delta,
                  float truncate_ratio, float minima_threshold,
                  float maxima_threshold, int use_threshold_ratios,
                  int64_t capacity, float *min_crds /*cap*3*/, float *min_sigma,
                  float *min_score, int64_t *n_min, float *max_crds,
                  float *max_sigma, float *max_score, int64_t *n_max) {
  int size[3] = {nx, ny, nz};
  View3<const float> s(src, nx, ny, nz), m(mask, nx, ny, nz);
  vector<float> sig(sigmas, sigmas + n_sigmas);
  vector<Vec3> minc, maxc;
  vector<float> mins, maxs, minsc, maxsc;
  BlobDog(size, s.p, m.p, sig, &minc, &maxc, &mins, &maxs, &minsc, &maxsc,
          (const float *)nullptr, delta, truncate_ratio, minima_threshold,
          maxima_threshold, use_threshold_ratios != 0, (ostream *)nullptr);
  *n_min = (int64_t)minc.size();
  *n_max = (int64_t)maxc.size();
  for (int64_t i = 0; i < min<int64_t>(capacity, *n_min); i++) {
    for (int d = 0; d < 3; d++) min_crds[3 * i + d] = minc[i][d];
    min_sigma[i] = mins[i];
    min_score[i] = minsc[i];
  }
  for (int64_t i = 0; i < min<int64_t>(capacity, *n_max); i++) {
    for (int d = 0; d < 3; d++) max_crds[3 * i + d] = maxc[i][d];
    max_sigma[i] = maxs[i];
    max_score[i] = maxsc[i];
  }
}

} // extern "C"
