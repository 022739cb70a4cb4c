// blobs_host.cpp -- blob list post-processing on the host (include/visfd_blobs.h): SortBlobs,
// DiscardMaskedBlobs, DiscardOverlappingBlobs and the score/diameter window, with the
// arithmetic types and evaluation order of the reference (Scalar = float; lib/visfd/feature.hpp,
// lib/visfd/visfd_utils.hpp:95-118).
#include "../../include/visfd_blobs.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <tuple>
#include <vector>

namespace {

struct List {
  std::vector<std::array<float, 3>> crds;
  std::vector<float> diam, score;
  List(int64_t n, const float *c, const float *d, const float *s) : crds((size_t)n), diam(d, d + n), score(s, s + n) {
    for (int64_t i = 0; i < n; i++) crds[(size_t)i] = {c[3 * i], c[3 * i + 1], c[3 * i + 2]};
  }
  int64_t store(float *c, float *d, float *s) const {
    for (size_t i = 0; i < crds.size(); i++) {
      c[3 * i] = crds[i][0]; c[3 * i + 1] = crds[i][1]; c[3 * i + 2] = crds[i][2];
      d[i] = diam[i];
      s[i] = score[i];
    }
    return (int64_t)crds.size();
  }
};

// SortBlobs(..., ascending_order, ignore_score_sign): feature.hpp:521-562
void sort_blobs(List &l, bool ascending, bool ignore_sign) {
  const size_t n = l.crds.size();
  if (n == 0) return;
  std::vector<std::tuple<float, size_t>> key(n);
  for (size_t i = 0; i < n; i++) key[i] = std::make_tuple(ignore_sign ? std::fabs(l.score[i]) : l.score[i], i);
  if (ascending) std::sort(key.begin(), key.end());
  else std::sort(key.rbegin(), key.rend());
  List old = l;
  for (size_t i = 0; i < n; i++) {
    const size_t j = std::get<1>(key[i]);
    l.crds[i] = old.crds[j];
    l.diam[i] = old.diam[j];
    l.score[i] = old.score[j];
  }
}

// SortBlobs(..., criteria, ascending_order): feature.hpp:572-616
bool sort_by_criteria(List &l, int criteria, bool ascending) {
  switch (criteria) {
    case VISFD_DO_NOT_SORT: return true;
    case VISFD_SORT_DECREASING: sort_blobs(l, ascending, false); return true;
    case VISFD_SORT_INCREASING: sort_blobs(l, !ascending, false); return true;
    case VISFD_SORT_DECREASING_MAGNITUDE: sort_blobs(l, ascending, true); return true;
    case VISFD_SORT_INCREASING_MAGNITUDE: sort_blobs(l, !ascending, true); return true;
    default: return false;
  }
}

inline float sqr(float x) { return x * x; }

// CalcSphereOverlap<float>: visfd_utils.hpp:95-118 (float operands, double constants)
float sphere_overlap(float rij, float Ri, float Rj) {
  if (Ri > Rj) std::swap(Ri, Rj);
  if (rij <= Ri) return (float)((4 * M_PI / 3) * Ri * Ri * Ri);
  const float xi = (float)(0.5 * (1.0 / rij) * (rij * rij + Ri * Ri - Rj * Rj));
  const float xj = (float)(0.5 * (1.0 / rij) * (rij * rij + Rj * Rj - Ri * Ri));
  return (float)((M_PI / 3) * (Ri * Ri * Ri * (2 - (xi / Ri) * (3 - sqr(xi / Ri))) +
                               Rj * Rj * Rj * (2 - (xj / Rj) * (3 - sqr(xj / Rj)))));
}

}  // namespace

extern "C" {

int64_t visfd_blobs_sort(int64_t n, float *crds, float *diameters, float *scores, int criteria, int ascending_order) {
  if (n < 0 || (n > 0 && (!crds || !diameters || !scores))) return -1;
  List l(n, crds, diameters, scores);
  if (!sort_by_criteria(l, criteria, ascending_order != 0)) return -1;
  return l.store(crds, diameters, scores);
}

int64_t visfd_blobs_filter(int64_t n, float *crds, float *diameters, float *scores, float score_lower, float score_upper,
                           float diameter_lower, float diameter_upper) {
  if (n < 0 || (n > 0 && (!crds || !diameters || !scores))) return -1;
  int64_t m = 0;
  for (int64_t i = 0; i < n; i++) {
    if (scores[i] >= score_lower && scores[i] <= score_upper && diameters[i] >= diameter_lower && diameters[i] <= diameter_upper) {
      for (int d = 0; d < 3; d++) crds[3 * m + d] = crds[3 * i + d];
      diameters[m] = diameters[i];
      scores[m] = scores[i];
      m++;
    }
  }
  return m;
}

int64_t visfd_blobs_discard_masked(int64_t n, float *crds, float *diameters, float *scores, const float *mask, int64_t nx,
                                   int64_t ny, int64_t nz) {
  if (n < 0 || (n > 0 && (!crds || !diameters || !scores))) return -1;
  int64_t m = 0;
  for (int64_t i = 0; i < n; i++) {
    const int ix = (int)std::floor(crds[3 * i] + 0.5), iy = (int)std::floor(crds[3 * i + 1] + 0.5),
              iz = (int)std::floor(crds[3 * i + 2] + 0.5);
    if (mask) {
      // the reference indexes the mask unchecked; a centre outside the image is an error here
      if (ix < 0 || iy < 0 || iz < 0 || ix >= nx || iy >= ny || iz >= nz) return -1;
      if (mask[((int64_t)iz * ny + iy) * nx + ix] == 0.0f) continue;
    }
    for (int d = 0; d < 3; d++) crds[3 * m + d] = crds[3 * i + d];
    diameters[m] = diameters[i];
    scores[m] = scores[i];
    m++;
  }
  return m;
}

int64_t visfd_blobs_discard_overlapping(int64_t n, float *crds, float *diameters, float *scores,
                                        float min_radial_separation_ratio, float max_volume_overlap_large,
                                        float max_volume_overlap_small, int criteria) {
  if (n < 0 || (n > 0 && (!crds || !diameters || !scores))) return -1;
  List l(n, crds, diameters, scores);
  if (!sort_by_criteria(l, criteria, false)) return -1;   // feature.hpp:739-745
  const int scale = 6;
  // bounding box of the blobs (feature.hpp:755-771; int bounds, float arithmetic, truncating stores)
  int bmin[3] = {0, 0, 0}, bmax[3] = {-1, -1, -1}, table[3];
  for (size_t i = 0; i < l.crds.size(); i++)
    for (int d = 0; d < 3; d++) {
      const float reff = std::ceil(l.diam[i] / 2);
      if ((l.crds[i][d] - reff < bmin[d]) || (bmin[d] > bmax[d])) bmin[d] = (int)(l.crds[i][d] - reff);
      if ((l.crds[i][d] + reff > bmax[d]) || (bmin[d] > bmax[d])) bmax[d] = (int)(l.crds[i][d] + reff);
    }
  for (int d = 0; d < 3; d++) table[d] = std::max(0, (1 + bmax[d] - bmin[d]) / scale);
  std::vector<std::vector<size_t>> occ((size_t)table[0] * table[1] * table[2]);
  auto cell = [&](int X, int Y, int Z) -> std::vector<size_t> & { return occ[((size_t)Z * table[1] + Y) * table[0] + X]; };
  List kept(0, nullptr, nullptr, nullptr);
  for (size_t i = 0; i < l.crds.size(); i++) {
    bool discard = false;
    const float reff_ = l.diam[i] / 2;
    const float Reff_ = reff_ / scale;
    const int Reff = (int)std::ceil(Reff_) + 1, Reffsq = Reff * Reff;
    const float ix = l.crds[i][0], iy = l.crds[i][1], iz = l.crds[i][2];
    const int Ix = (int)std::floor((ix - bmin[0]) / scale), Iy = (int)std::floor((iy - bmin[1]) / scale),
              Iz = (int)std::floor((iz - bmin[2]) / scale);
    auto inside = [&](int X, int Y, int Z) { return 0 <= X && X < table[0] && 0 <= Y && Y < table[1] && 0 <= Z && Z < table[2]; };
    for (int Jz = -Reff; Jz <= Reff && !discard; Jz++)
      for (int Jy = -Reff; Jy <= Reff && !discard; Jy++)
        for (int Jx = -Reff; Jx <= Reff && !discard; Jx++) {
          if (!inside(Ix + Jx, Iy + Jy, Iz + Jz)) continue;
          if (Jx * Jx + Jy * Jy + Jz * Jz > Reffsq) continue;
          for (size_t k : cell(Ix + Jx, Iy + Jy, Iz + Jz)) {
            const float kx = l.crds[k][0], ky = l.crds[k][1], kz = l.crds[k][2];
            const float rik = std::sqrt((ix - kx) * (ix - kx) + (iy - ky) * (iy - ky) + (iz - kz) * (iz - kz));
            const float ri = l.diam[i] / 2, rk = l.diam[k] / 2;
            const float vol_overlap = sphere_overlap(rik, ri, rk);
            if (rik < (ri + rk) * min_radial_separation_ratio) discard = true;
            const float vi = (float)((4 * M_PI / 3) * (ri * ri * ri)), vk = (float)((4 * M_PI / 3) * (rk * rk * rk));
            const float v_large = vk > vi ? vk : vi, v_small = vk > vi ? vi : vk;
            if ((vol_overlap / v_small > max_volume_overlap_small) || (vol_overlap / v_large > max_volume_overlap_large))
              discard = true;
          }
        }
    if (discard) continue;
    kept.crds.push_back(l.crds[i]);
    kept.diam.push_back(l.diam[i]);
    kept.score.push_back(l.score[i]);
    for (int Jz = -Reff; Jz <= Reff; Jz++)
      for (int Jy = -Reff; Jy <= Reff; Jy++)
        for (int Jx = -Reff; Jx <= Reff; Jx++) {
          if (!inside(Ix + Jx, Iy + Jy, Iz + Jz)) continue;
          if (Jx * Jx + Jy * Jy + Jz * Jz > Reffsq) continue;
          cell(Ix + Jx, Iy + Jy, Iz + Jz).push_back(i);
        }
  }
  return kept.store(crds, diameters, scores);
}

}  // extern "C"
