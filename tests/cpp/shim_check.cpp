// tests/cpp/shim_check.cpp -- exercises the C++ host-side mirror (visfd_cuda_shim.hpp) the
// way filter_mrc would: pointer-table images (Alloc3D layout and a deliberately
// NON-contiguous one), the reference's call signatures, exceptions for bad input; every
// result is compared with the CPU oracle (oracle/libvisfd_oracle.so, test infrastructure).
// Exit code 0 = all checks passed.  Run by tests/test_gpu_shim.py on the GPU box.
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#include "../../visfd_b200/csrc/visfd_cuda_shim.hpp"

extern "C" {
float vo_apply_gauss(int64_t, int64_t, int64_t, const float *, float *, const float *, const float[3], const int[3], int);
void vo_apply_log(int64_t, int64_t, int64_t, const float *, float *, const float *, const float[3], float, float, float *, float *);
int vo_calc_hessian(int64_t, int64_t, int64_t, const float *, const float *, float, float, float *, float *, float *);
float vo_membrane(int64_t, int64_t, int64_t, const float *, const float *, float, float, int, float, int, float, int, float,
                  float *, float *, float *, float *);
void vo_tv_dense_stick(int64_t, int64_t, int64_t, const float *, const float *, const float *, const float *, float, int,
                       float, int, float *);
int vo_bin3d(const int64_t[3], const int64_t[3], const float *, float *, const int *);
int vo_unbin3d(const int64_t[3], const int64_t[3], const float *, float *, const int *);
int vo_draw_regions(int, int, int, float *, const float *, const visfd_region *, int, int);
}

// the shape of visfd::SimpleRegion<float> (lib/visfd/draw.hpp:41-87), as the caller would hand it over
struct Region {
  struct Rect { float xmin, xmax, ymin, ymax, zmin, zmax; };
  struct Sphere { float x0, y0, z0, r; };
  enum RegionType { RECT, SPHERE };
  RegionType type = RECT;
  union { Rect rect; Sphere sphere; } data;
  float value = 1;
};

// pointer tables over a flat buffer; `gap` > 0 makes rows non-contiguous
template <typename T>
struct Image3 {
  int nx, ny, nz;
  size_t pitch;
  std::vector<T> data;
  std::vector<T *> rows;
  std::vector<T **> planes;
  Image3(int nx_, int ny_, int nz_, int gap = 0) : nx(nx_), ny(ny_), nz(nz_), pitch(nx_ + gap) {
    data.assign(pitch * ny * nz, T());
    rows.resize((size_t)ny * nz);
    planes.resize(nz);
    for (int z = 0; z < nz; z++) {
      for (int y = 0; y < ny; y++) rows[(size_t)z * ny + y] = data.data() + ((size_t)z * ny + y) * pitch;
      planes[z] = rows.data() + (size_t)z * ny;
    }
  }
  T ***p() { return planes.data(); }
  std::vector<T> flat() const {
    std::vector<T> f((size_t)nx * ny * nz);
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        std::copy(rows[(size_t)z * ny + y], rows[(size_t)z * ny + y] + nx, f.begin() + ((size_t)z * ny + y) * nx);
    return f;
  }
};

static int failures = 0;
static void check(bool ok, const char *what) {
  std::printf("%s  %s\n", ok ? "PASS" : "FAIL", what);
  if (!ok) failures++;
}
static double max_rel(const float *a, const float *b, size_t n) {
  double scale = 0, worst = 0;
  for (size_t i = 0; i < n; i++) scale = std::max(scale, (double)std::fabs(b[i]));
  for (size_t i = 0; i < n; i++)
    worst = std::max(worst, std::fabs((double)a[i] - b[i]) / std::max((double)std::fabs(b[i]), 1e-3 * scale));
  return worst;
}

int main() {
  const int nx = 37, ny = 30, nz = 26;
  const int size[3] = {nx, ny, nz};
  const size_t N = (size_t)nx * ny * nz;
  std::mt19937 rng(7);
  std::normal_distribution<float> noise(0.f, 1.f);
  for (int gap = 0; gap <= 5; gap += 5) {  // contiguous (Alloc3D) and non-contiguous tables
    Image3<float> src(nx, ny, nz, gap), dst(nx, ny, nz, gap), mask(nx, ny, nz, gap);
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
          float r = std::sqrt((x - 18.f) * (x - 18.f) + (y - 15.f) * (y - 15.f) + (z - 13.f) * (z - 13.f));
          src.p()[z][y][x] = noise(rng) - 3.f * std::exp(-0.5f * (r - 9.f) * (r - 9.f) / 4.f);
          mask.p()[z][y][x] = (x < 4) ? 0.f : 1.f;
        }
    std::vector<float> fsrc = src.flat(), fmask = mask.flat(), want(N);
    char label[128];

    float sigma[3] = {1.5f, 1.5f, 1.5f};
    int hw[3] = {3, 3, 3};
    float A = visfd_cuda::ApplyGauss(size, src.p(), dst.p(), (float const *const *const *)nullptr, sigma, hw, true);
    float A0 = vo_apply_gauss(nx, ny, nz, fsrc.data(), want.data(), nullptr, sigma, hw, 1);
    std::snprintf(label, sizeof label, "ApplyGauss bit-exact (row gap %d)", gap);
    check(dst.flat() == want && A == A0, label);

    visfd_cuda::ApplyGauss(size, src.p(), dst.p(), mask.p(), 1.5f, 3, true);
    vo_apply_gauss(nx, ny, nz, fsrc.data(), want.data(), fmask.data(), sigma, hw, 1);
    std::snprintf(label, sizeof label, "ApplyGauss masked bit-exact (row gap %d)", gap);
    check(dst.flat() == want, label);

    float sg3[3] = {2.f, 2.f, 2.f};
    visfd_cuda::ApplyLog(size, src.p(), dst.p(), (float const *const *const *)nullptr, sg3, 0.02f, 2.6482f);
    vo_apply_log(nx, ny, nz, fsrc.data(), want.data(), nullptr, sg3, 0.02f, 2.6482f, nullptr, nullptr);
    std::snprintf(label, sizeof label, "ApplyLog bit-exact (row gap %d)", gap);
    check(dst.flat() == want, label);

    // CalcHessian into array<float,3>*** and a per-voxel float**** (nullptr where masked out)
    Image3<std::array<float, 3> > grad(nx, ny, nz, gap);
    std::vector<float> hbuf(N * 6, 0.f);
    Image3<float *> hess(nx, ny, nz);
    size_t k = 0;
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++)
          hess.p()[z][y][x] = (mask.p()[z][y][x] != 0.f) ? &hbuf[6 * k++] : nullptr;
    visfd_cuda::CalcHessian(size, src.p(), grad.p(), hess.p(), mask.p(), 1.2f, 2.6482f);
    std::vector<float> wg(N * 3, 0.f), wh(N * 6, 0.f);
    vo_calc_hessian(nx, ny, nz, fsrc.data(), fmask.data(), 1.2f, 2.6482f, wg.data(), wh.data(), nullptr);
    bool ok = true;
    for (int z = 0; z < nz && ok; z++)
      for (int y = 0; y < ny && ok; y++)
        for (int x = 0; x < nx; x++) {
          size_t i = ((size_t)z * ny + y) * nx + x;
          if (!hess.p()[z][y][x]) continue;
          if (std::memcmp(hess.p()[z][y][x], &wh[6 * i], 24) || std::memcmp(&grad.p()[z][y][x], &wg[3 * i], 12)) {
            ok = false;
            break;
          }
        }
    std::snprintf(label, sizeof label, "CalcHessian (compact float****) bit-exact (row gap %d)", gap);
    check(ok, label);

    // fused pipeline as HandleTV would call it, then the stand-alone TV3D class on its outputs
    visfd_membrane_params p = {1.2f, 2.6482f, VISFD_DECREASING_EIVALS, 0.1f, 1, 3.0f, 4, 1.41421354f};
    Image3<std::array<float, 3> > dir(nx, ny, nz, gap);
    float thr = visfd_cuda::MembranePipeline(size, src.p(), dst.p(), (float const *const *const *)nullptr, p, dir.p());
    std::vector<float> wsal(N), wdir(N * 3), wt(N * 6), wout(N);
    float thr0 = vo_membrane(nx, ny, nz, fsrc.data(), nullptr, p.sigma, p.truncate_ratio, 1, p.cut, 1, p.tv_sigma, 4,
                             p.tv_cutoff_ratio, wsal.data(), wdir.data(), wt.data(), wout.data());
    std::vector<float> got = dst.flat();
    std::snprintf(label, sizeof label, "MembranePipeline threshold identical, saliency within 1e-4 (row gap %d)", gap);
    check(thr == thr0 && max_rel(got.data(), wout.data(), N) <= 1e-4, label);

    Image3<float> sal(nx, ny, nz, gap);
    for (int z = 0; z < nz; z++)
      for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
          size_t i = ((size_t)z * ny + y) * nx + x;
          sal.p()[z][y][x] = wsal[i];
          dir.p()[z][y][x] = {wdir[3 * i], wdir[3 * i + 1], wdir[3 * i + 2]};
        }
    std::vector<float> tbuf(N * 6, 0.f);
    Image3<float *> ten(nx, ny, nz);
    for (size_t i = 0; i < N; i++) ten.data[i] = &tbuf[6 * i];
    visfd_cuda::TV3D tv(3.0f, 4, 1.41421354f);
    tv.TVDenseStick(size, sal.p(), dir.p(), ten.p(), nullptr, nullptr, false, false, false);
    double scale = 0, worst = 0;
    for (size_t i = 0; i < N * 6; i++) scale = std::max(scale, (double)std::fabs(wt[i]));
    for (size_t i = 0; i < N * 6; i++) worst = std::max(worst, std::fabs((double)tbuf[i] - wt[i]) / scale);
    std::snprintf(label, sizeof label, "TV3D::TVDenseStick tensor within 1e-5 of the volume scale (row gap %d)", gap);
    check(worst <= 1e-5, label);
  }

  // error behaviour: CalcHessian on a 2-voxel-thin image throws (feature.hpp:1260-1264)
  {
    const int small[3] = {5, 5, 2};
    Image3<float> s(5, 5, 2), d(5, 5, 2);
    Image3<std::array<float, 3> > g(5, 5, 2);
    std::vector<float> hb(50 * 6);
    Image3<float *> h(5, 5, 2);
    for (size_t i = 0; i < 50; i++) h.data[i] = &hb[6 * i];
    bool threw = false;
    try {
      visfd_cuda::CalcHessian(small, s.p(), g.p(), h.p(), nullptr, 1.0f, 2.5f);
    } catch (const std::exception &e) {
      threw = true;
    }
    check(threw, "CalcHessian throws on an image thinner than 3 voxels");
    threw = false;
    try {
      visfd_cuda::TV3D tv(1.0f, 4, 1.4f);
      tv.TVDenseStick(small, s.p(), g.p(), h.p(), nullptr, nullptr, false, true /* normalize */, false);
    } catch (const std::exception &e) {
      threw = true;
    }
    check(threw, "TVDenseStick(normalize=true) is rejected, not silently ignored");
  }
  // BinArray3D / UnbinArray3D (resample.hpp:53-166) through pointer tables, bit-exact
  for (int gap = 0; gap <= 3; gap += 3) {
    const int big[3] = {23, 17, 13}, small[3] = {11, 8, 6}, off[3] = {1, 0, 0};
    const int64_t big64[3] = {23, 17, 13}, small64[3] = {11, 8, 6};
    Image3<float> src(23, 17, 13, gap), dst(11, 8, 6, gap), back(23, 17, 13, gap);
    for (int z = 0; z < 13; z++)
      for (int y = 0; y < 17; y++)
        for (int x = 0; x < 23; x++) src.p()[z][y][x] = noise(rng);
    std::vector<float> fsrc = src.flat(), want(11 * 8 * 6), wantb(23 * 17 * 13);
    char label[128];
    visfd_cuda::BinArray3D(big, small, src.p(), dst.p(), off);
    vo_bin3d(big64, small64, fsrc.data(), want.data(), off);
    std::snprintf(label, sizeof label, "BinArray3D bit-exact (row gap %d)", gap);
    check(dst.flat() == want, label);
    visfd_cuda::UnbinArray3D(small, big, dst.p(), back.p());
    vo_unbin3d(small64, big64, want.data(), wantb.data(), nullptr);
    std::snprintf(label, sizeof label, "UnbinArray3D bit-exact (row gap %d)", gap);
    check(back.flat() == wantb, label);
    bool threw = false;
    const int bad[3] = {2, 0, 0};
    try {
      visfd_cuda::BinArray3D(big, small, src.p(), dst.p(), bad);
    } catch (const std::exception &e) {
      threw = true;
    }
    check(threw, "BinArray3D throws for an offset outside [0, bin size)");
  }
  // DrawRegions (draw.hpp:90-237): a box, a subtracted sphere and a small bright sphere, with and
  // without a mask, contiguous and gapped tables
  for (int gap = 0; gap <= 3; gap += 3) {
    const int size[3] = {31, 22, 17};
    std::vector<Region> regions(3);
    regions[0].type = Region::RECT;
    regions[0].data.rect = {2.0f, 25.4f, 1.0f, 19.0f, 0.0f, 30.0f};
    regions[0].value = 1.0f;
    regions[1].type = Region::SPHERE;
    regions[1].data.sphere = {14.0f, 10.0f, 8.0f, 6.5f};
    regions[1].value = -1.0f;
    regions[2].type = Region::SPHERE;
    regions[2].data.sphere = {15.2f, 10.7f, 8.4f, 2.6f};
    regions[2].value = 3.0f;
    std::vector<visfd_region> flat(3);
    flat[0] = {VISFD_REGION_RECT, {2.0f, 25.4f, 1.0f, 19.0f, 0.0f, 30.0f}, 1.0f};
    flat[1] = {VISFD_REGION_SPHERE, {14.0f, 10.0f, 8.0f, 6.5f, 0.0f, 0.0f}, -1.0f};
    flat[2] = {VISFD_REGION_SPHERE, {15.2f, 10.7f, 8.4f, 2.6f, 0.0f, 0.0f}, 3.0f};
    for (int with_mask = 0; with_mask <= 1; with_mask++) {
      Image3<float> img(31, 22, 17, gap), mask(31, 22, 17, gap);
      for (int z = 0; z < 17; z++)
        for (int y = 0; y < 22; y++)
          for (int x = 0; x < 31; x++) mask.p()[z][y][x] = noise(rng) > -0.3f ? 1.0f : 0.0f;
      std::vector<float> want = img.flat(), fmask = mask.flat();
      visfd_cuda::DrawRegions(size, img.p(), with_mask ? mask.p() : nullptr, regions, true);
      vo_draw_regions(31, 22, 17, want.data(), with_mask ? fmask.data() : nullptr, flat.data(), 3, 1);
      char label[128];
      std::snprintf(label, sizeof label, "DrawRegions bit-exact (row gap %d, mask %d)", gap, with_mask);
      check(img.flat() == want, label);
    }
  }
  std::printf("%s (%d failure%s)\n", failures ? "FAILED" : "OK", failures, failures == 1 ? "" : "s");
  return failures ? 1 : 0;
}
