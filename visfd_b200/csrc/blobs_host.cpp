// blobs_host.cpp -- blob list post-processing on the host (include/visfd_blobs.h): SortBlobs,
// DiscardMaskedBlobs, DiscardOverlappingBlobs and the score/diameter window, with the
// arithmetic types and evaluation order of the reference (Scalar = float; lib/visfd/feature.hpp,
// lib/visfd/visfd_utils.hpp:95-118).
#include "../../include/visfd_blobs.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <tuple>
#include <unordered_map>
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


// ---- greedy non-max suppression (DiscardOverlappingBlobs, feature.hpp:723-913) ------------------------------------
// The reference walks the sorted list and keeps a blob unless an already kept one is too close or overlaps too much.
// It finds the kept blobs to compare with through a coarse occupancy table (cells of 6 voxels over the blobs' bounding
// box): every kept blob writes its number into all cells of a lattice ball around its own cell, and a new blob reads
// the cells of a lattice ball around its cell.  Which pairs get compared is therefore decided by whether two lattice
// balls, clipped to the table, share a cell -- a rule that is part of the result (a pair whose balls merely touch
// between lattice points is never compared).  Here the rule is evaluated directly: kept blobs sit in a hash grid keyed
// by their coarse cell (one entry per blob instead of one per covered cell), a newcomer scans the buckets within reach
// and asks for each kept blob whether the two footprints meet (an O(r^2) interval test), then applies the reference's
// pair criteria.  The decision is an OR over the kept blobs a newcomer can see, so it does not depend on the order in
// which they are visited.
struct Footprint {
  int cell[3];   // coarse cell of the centre
  int reach;     // radius of the lattice ball, in cells
};

class Suppressor {
  static constexpr int kCell = 6;   // voxels per coarse cell (feature.hpp:747)
  const List &blobs;
  const float sep_ratio, overlap_large, overlap_small;
  int lo[3] = {0, 0, 0}, dims[3] = {0, 0, 0};

  // the table's origin and size: bounding box of centre -+ ceil(radius), int bounds updated through float compares
  // and truncating stores exactly as feature.hpp:755-776 does
  void measure() {
    int hi[3] = {-1, -1, -1};
    for (size_t i = 0; i < blobs.crds.size(); i++) {
      const float pad = std::ceil(blobs.diam[i] / 2);
      for (int d = 0; d < 3; d++) {
        const bool empty_lo = lo[d] > hi[d];
        if (blobs.crds[i][d] - pad < lo[d] || empty_lo) lo[d] = (int)(blobs.crds[i][d] - pad);
        const bool empty_hi = lo[d] > hi[d];
        if (blobs.crds[i][d] + pad > hi[d] || empty_hi) hi[d] = (int)(blobs.crds[i][d] + pad);
      }
    }
    for (int d = 0; d < 3; d++) dims[d] = std::max(0, (1 + hi[d] - lo[d]) / kCell);
  }

  Footprint footprint(size_t i) const {
    Footprint f;
    const float coarse_radius = (blobs.diam[i] / 2) / kCell;
    f.reach = (int)std::ceil(coarse_radius) + 1;
    for (int d = 0; d < 3; d++) f.cell[d] = (int)std::floor((blobs.crds[i][d] - lo[d]) / kCell);
    return f;
  }

  // largest |dx| with dx^2 <= budget (budget >= 0)
  static int isqrt_floor(int budget) {
    int r = (int)std::sqrt((double)budget);
    while ((r + 1) * (r + 1) <= budget) r++;
    while (r * r > budget) r--;
    return r;
  }

  // do the two lattice balls share a cell inside the table?
  bool meet(const Footprint &a, const Footprint &b) const {
    const int a2 = a.reach * a.reach, b2 = b.reach * b.reach;
    const int z0 = std::max(0, std::max(a.cell[2] - a.reach, b.cell[2] - b.reach));
    const int z1 = std::min(dims[2] - 1, std::min(a.cell[2] + a.reach, b.cell[2] + b.reach));
    for (int z = z0; z <= z1; z++) {
      const int az = (z - a.cell[2]) * (z - a.cell[2]), bz = (z - b.cell[2]) * (z - b.cell[2]);
      const int y0 = std::max(0, std::max(a.cell[1] - a.reach, b.cell[1] - b.reach));
      const int y1 = std::min(dims[1] - 1, std::min(a.cell[1] + a.reach, b.cell[1] + b.reach));
      for (int y = y0; y <= y1; y++) {
        const int arem = a2 - az - (y - a.cell[1]) * (y - a.cell[1]);
        const int brem = b2 - bz - (y - b.cell[1]) * (y - b.cell[1]);
        if (arem < 0 || brem < 0) continue;
        const int ax = isqrt_floor(arem), bx = isqrt_floor(brem);
        const int x0 = std::max(0, std::max(a.cell[0] - ax, b.cell[0] - bx));
        const int x1 = std::min(dims[0] - 1, std::min(a.cell[0] + ax, b.cell[0] + bx));
        if (x0 <= x1) return true;
      }
    }
    return false;
  }

  // the reference's criteria for a newcomer i against a kept blob k (feature.hpp:820-862)
  bool suppresses(size_t k, size_t i) const {
    const float dx = blobs.crds[i][0] - blobs.crds[k][0], dy = blobs.crds[i][1] - blobs.crds[k][1],
                dz = blobs.crds[i][2] - blobs.crds[k][2];
    const float dist = std::sqrt(dx * dx + dy * dy + dz * dz);
    const float ri = blobs.diam[i] / 2, rk = blobs.diam[k] / 2;
    const float shared = sphere_overlap(dist, ri, rk);
    bool out = dist < (ri + rk) * sep_ratio;
    const float vol_i = (float)((4 * M_PI / 3) * (ri * ri * ri)), vol_k = (float)((4 * M_PI / 3) * (rk * rk * rk));
    const float bigger = std::max(vol_i, vol_k), smaller = std::min(vol_i, vol_k);
    if (shared / smaller > overlap_small || shared / bigger > overlap_large) out = true;
    return out;
  }

  static uint64_t bucket_key(int x, int y, int z) {
    return ((uint64_t)(uint32_t)(x + (1 << 20)) << 42) | ((uint64_t)(uint32_t)(y + (1 << 20)) << 21) |
           (uint64_t)(uint32_t)(z + (1 << 20));
  }

 public:
  Suppressor(const List &l, float sep, float ov_large, float ov_small)
      : blobs(l), sep_ratio(sep), overlap_large(ov_large), overlap_small(ov_small) {
    measure();
  }

  void run(List &kept) const {
    std::unordered_map<uint64_t, std::vector<size_t>> grid;   // coarse cell of the centre -> kept blobs
    std::vector<Footprint> prints(blobs.crds.size());
    int widest = 0;                                           // largest reach among the kept blobs
    for (size_t i = 0; i < blobs.crds.size(); i++) {
      const Footprint me = prints[i] = footprint(i);
      bool drop = false;
      const int span = me.reach + widest;
      for (int z = me.cell[2] - span; z <= me.cell[2] + span && !drop; z++)
        for (int y = me.cell[1] - span; y <= me.cell[1] + span && !drop; y++)
          for (int x = me.cell[0] - span; x <= me.cell[0] + span && !drop; x++) {
            const auto hit = grid.find(bucket_key(x, y, z));
            if (hit == grid.end()) continue;
            for (size_t k : hit->second)
              if (meet(me, prints[k]) && suppresses(k, i)) { drop = true; break; }
          }
      if (drop) continue;
      kept.crds.push_back(blobs.crds[i]);
      kept.diam.push_back(blobs.diam[i]);
      kept.score.push_back(blobs.score[i]);
      grid[bucket_key(me.cell[0], me.cell[1], me.cell[2])].push_back(i);
      widest = std::max(widest, me.reach);
    }
  }
};

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
  const Suppressor nms(l, min_radial_separation_ratio, max_volume_overlap_large, max_volume_overlap_small);
  List kept(0, nullptr, nullptr, nullptr);
  nms.run(kept);
  return kept.store(crds, diameters, scores);
}

}  // extern "C"
