// blob.cu -- scale-space blob detection, BlobDog (lib/visfd/feature.hpp:56-427):
// a ring of three LoG-filtered volumes (ApplyLog per scale, gauss.cu) and, for every
// interior scale, a strict 80-neighbour extremum test in (x,y,z,scale) that appends
// candidates (x,y,z,score) to a device list through a warp-aggregated atomic cursor.
// Scan traffic: the three volumes are read once from HBM (12 B/voxel); the 27-point
// neighbourhoods come out of L1/L2.
//
// The reference's running score filter (:267-303) depends on OpenMP thread order; only
// its deterministic consequences are kept: in absolute mode a candidate must beat the
// threshold strictly, in ratio mode candidates are pre-filtered on the device against
// the best score of the PREVIOUS scales (never stricter than the final filter, :362-417)
// and the final filter runs on the host once the global best scores are known.
#include "common.cuh"
#include "kernels.cuh"
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>

namespace visfd_cuda {

struct BlobCand {
  float x, y, z, score;
};

struct BlobScanArgs {
  const float *prev, *cur, *next, *mask;
  int nx, ny, nz;             // nz: planes of the slab
  int z_offset, nz_global;    // the slab's first plane in the image, the image's planes
  int own_z0;                 // first receiver plane (slab-local); the grid spans the receiver planes
  float min_thr, max_thr;     // device pre-filter: minima need score < min_thr, maxima > max_thr
  BlobCand *mins, *maxs;
  unsigned long long *counters;  // [0] minima, [1] maxima
  unsigned long long capacity;
};

__device__ __forceinline__ void append(BlobCand *list, unsigned long long *counter,
                                       unsigned long long capacity, bool pred, const BlobCand &c) {
  unsigned m = __ballot_sync(0xffffffffu, pred);
  if (!m) return;
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  const int leader = __ffs(m) - 1;
  if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (pred) {
    unsigned long long slot = base + __popc(m & ((1u << lane) - 1u));
    if (slot < capacity) list[slot] = c;
  }
}

__global__ void __launch_bounds__(256) blob_scan_kernel(BlobScanArgs a) {
  const int ix = blockIdx.x * 64 + (threadIdx.x & 63);
  const int iy = blockIdx.y * 4 + (threadIdx.x >> 6);
  const int iz = a.own_z0 + blockIdx.z;
  const int gz = a.z_offset + iz;
  bool is_min = false, is_max = false;
  float e = 0.0f;
  // every neighbour must be inside the image (feature.hpp:245-258); the slab's halo guarantees that a
  // neighbour inside the image is inside the slab
  if (ix >= 1 && ix < a.nx - 1 && iy >= 1 && iy < a.ny - 1 && gz >= 1 && gz < a.nz_global - 1 && iz >= 1 &&
      iz < a.nz - 1) {
    const size_t sy = a.nx, sz = (size_t)a.nx * a.ny;
    const size_t c = (size_t)iz * sz + (size_t)iy * sy + ix;
    e = __ldg(a.cur + c);
    is_min = e < 0.0f;   // minima need score < 0, maxima score > 0 (:270-271, :289-290)
    is_max = e > 0.0f;
    if (a.mask && __ldg(a.mask + c) == 0.0f) is_min = is_max = false;
    const float *vol[3] = {a.cur, a.prev, a.next};
    for (int r = 0; r < 3 && (is_min || is_max); r++) {
      const float *v = vol[r];
      for (int jz = -1; jz <= 1 && (is_min || is_max); jz++)
        for (int jy = -1; jy <= 1; jy++) {
          const size_t row = c + jz * (ptrdiff_t)sz + jy * (ptrdiff_t)sy;
#pragma unroll
          for (int jx = -1; jx <= 1; jx++) {
            if (r == 0 && jx == 0 && jy == 0 && jz == 0) continue;
            float nb = __ldg(v + row + jx);
            if (nb <= e) is_min = false;
            if (nb >= e) is_max = false;
            if (r == 0 && a.mask && __ldg(a.mask + row + jx) == 0.0f) is_min = is_max = false;
          }
        }
    }
  }
  BlobCand cand{(float)ix, (float)iy, (float)gz, e};
  append(a.mins, a.counters + 0, a.capacity, is_min && e < a.min_thr, cand);
  append(a.maxs, a.counters + 1, a.capacity, is_max && e > a.max_thr, cand);
}

static void sort_raster(std::vector<BlobCand> &v) {
  std::sort(v.begin(), v.end(), [](const BlobCand &p, const BlobCand &q) {
    if (p.z != q.z) return p.z < q.z;
    if (p.y != q.y) return p.y < q.y;
    return p.x < q.x;
  });
}

// feature.hpp:362-417: the final score filter, once the best scores of the WHOLE image are known
void blob_final_filter(BlobList &minima, BlobList &maxima, float minima_threshold, float maxima_threshold,
                       int use_threshold_ratios, float gmin, float gmax) {
  const float INF = std::numeric_limits<float>::infinity();
  bool filt = (minima_threshold != INF) || (maxima_threshold != -INF);
  bool drop_min = false, drop_max = false;
  if (use_threshold_ratios) {
    // An infinite RATIO admits (almost) nothing in the reference's running filter
    // (+-inf times the running best score); the deterministic reading is an empty list.
    if (maxima_threshold == -INF) drop_max = true;
    if (minima_threshold == INF) drop_min = true;
    if (filt) {
      minima_threshold *= gmin;
      maxima_threshold *= gmax;
    }
  }
  auto keep = [&](BlobList &l, bool drop, bool is_min) {
    BlobList out;
    if (!drop)
      for (size_t i = 0; i < l.score.size(); i++)
        if (!filt || (is_min ? l.score[i] <= minima_threshold : l.score[i] >= maxima_threshold)) {
          out.crds.insert(out.crds.end(), {l.crds[3 * i], l.crds[3 * i + 1], l.crds[3 * i + 2]});
          out.sigma.push_back(l.sigma[i]);
          out.score.push_back(l.score[i]);
        }
    l = std::move(out);
  };
  keep(minima, drop_min, true);
  keep(maxima, drop_max, false);
}

// Slab form: the slab holds image planes [z_offset, z_offset + nz); candidates are produced for the
// slab-local planes [own_z0, own_z1) with z reported as the IMAGE plane.  best[0] / best[1] receive the
// best minimum / maximum score found (1 / -1 if none, feature.hpp:122-123).  finalize = false leaves out
// the final filter, which needs the best scores of the whole image (blob_final_filter).
void blob_dog_device(visfd_ctx *ctx, i64 nx, i64 ny, i64 nz, i64 z_offset, i64 nz_global, i64 own_z0, i64 own_z1,
                     const float *src, const float *mask,
                     const float *sigmas, int n_sigmas, float delta, float truncate_ratio,
                     float minima_threshold, float maxima_threshold, int use_threshold_ratios,
                     BlobList &minima, BlobList &maxima, bool finalize, float best[2]) {
  VREQUIRE(nx > 0 && ny > 0 && nz > 0, "empty volume");
  VREQUIRE(z_offset >= 0 && z_offset + nz <= nz_global && 0 <= own_z0 && own_z0 <= own_z1 && own_z1 <= nz,
           "slab / receiver planes outside the image");
  VREQUIRE(nx < 2147483647LL / 2 && ny <= 4 * 65535LL && nz <= 65535 && nz_global < 2147483647LL,
           "volume too large for the blob scan launch");
  const i64 N = nx * ny * nz;
  const float INF = std::numeric_limits<float>::infinity();
  Scratch<float> ring[3];
  for (int k = 0; k < 3; k++) ring[k].reset(ctx, N);
  unsigned long long capacity = (unsigned long long)std::max<i64>(1 << 16, N / 32);
  Scratch<BlobCand> dmins(ctx, capacity), dmaxs(ctx, capacity);
  Scratch<unsigned long long> counters(ctx, 2);

  std::vector<BlobCand> mins, maxs;
  std::vector<float> min_sig, max_sig;
  float gmin = 1.0f, gmax = -1.0f;  // feature.hpp:122-123 ("impossible" initial values)

  for (int ir = 0; ir < n_sigmas; ir++) {
    float s3[3] = {sigmas[ir], sigmas[ir], sigmas[ir]};
    float sa[3], sb[3], scale;
    int hw[3];
    log_params(s3, delta, truncate_ratio, sa, sb, hw, &scale);
    // an interior slab end needs hw planes for the LoG of the receiver planes' neighbours, + 1 for those
    VREQUIRE((z_offset == 0 || own_z0 >= hw[2] + 1) && (z_offset + nz == nz_global || nz - own_z1 >= hw[2] + 1),
             "slab lacks the halo (LoG half-width + 1 planes) around the receiver planes");
    dog_device(ctx, nx, ny, nz, z_offset, nz_global, src, ring[ir % 3].get(), mask, sa, sb, hw, scale, nullptr, nullptr);
    if (ir < 2 || own_z1 == own_z0) continue;
    // device pre-filter thresholds for this scale
    float min_thr = minima_threshold, max_thr = maxima_threshold;
    if (use_threshold_ratios) {
      // final cut will be ratio * (global best); the best of the previous scales gives a
      // bound that can only be looser (for ratios >= 0)
      min_thr = (gmin < 0.0f && minima_threshold >= 0.0f && minima_threshold != INF) ? minima_threshold * gmin : INF;
      max_thr = (gmax > 0.0f && maxima_threshold >= 0.0f && maxima_threshold != -INF) ? maxima_threshold * gmax : -INF;
      // (<= becomes < on the device: widen by one ulp so that equality survives to the host filter)
      if (min_thr != INF) min_thr = std::nextafter(min_thr, INF);
      if (max_thr != -INF) max_thr = std::nextafter(max_thr, -INF);
    }
    for (int attempt = 0; attempt < 2; attempt++) {
      BlobScanArgs a;
      a.prev = ring[(ir - 2) % 3].get();
      a.cur = ring[(ir - 1) % 3].get();
      a.next = ring[ir % 3].get();
      a.mask = mask;
      a.nx = (int)nx; a.ny = (int)ny; a.nz = (int)nz;
      a.z_offset = (int)z_offset; a.nz_global = (int)nz_global; a.own_z0 = (int)own_z0;
      a.min_thr = min_thr; a.max_thr = max_thr;
      a.mins = dmins.get(); a.maxs = dmaxs.get();
      a.counters = counters.get();
      a.capacity = capacity;
      unsigned long long h[2];
      {
        StageTimer t(ctx, "blob_scan");
        VCK(cudaMemsetAsync(counters.get(), 0, 2 * sizeof(unsigned long long), ctx->stream));
        dim3 grid(div_up(nx, 64), div_up(ny, 4), (unsigned)(own_z1 - own_z0));
        blob_scan_kernel<<<grid, 256, 0, ctx->stream>>>(a);
        VCK(cudaGetLastError());
        ctx->count_launch();
      }
      VCK(cudaMemcpyAsync(h, counters.get(), sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
      VCK(cudaStreamSynchronize(ctx->stream));
      if (h[0] > capacity || h[1] > capacity) {
        VREQUIRE(attempt == 0, "blob candidate list overflow");
        capacity = std::max(h[0], h[1]);
        dmins.reset(ctx, capacity);
        dmaxs.reset(ctx, capacity);
        continue;
      }
      std::vector<BlobCand> lm(h[0]), lx(h[1]);
      if (h[0]) VCK(cudaMemcpyAsync(lm.data(), dmins.get(), h[0] * sizeof(BlobCand), cudaMemcpyDeviceToHost, ctx->stream));
      if (h[1]) VCK(cudaMemcpyAsync(lx.data(), dmaxs.get(), h[1] * sizeof(BlobCand), cudaMemcpyDeviceToHost, ctx->stream));
      VCK(cudaStreamSynchronize(ctx->stream));
      sort_raster(lm);
      sort_raster(lx);
      for (auto &c : lm) { mins.push_back(c); min_sig.push_back(sigmas[ir - 1]); gmin = std::min(gmin, c.score); }
      for (auto &c : lx) { maxs.push_back(c); max_sig.push_back(sigmas[ir - 1]); gmax = std::max(gmax, c.score); }
      break;
    }
  }

  minima = BlobList();
  maxima = BlobList();
  for (size_t i = 0; i < mins.size(); i++) {
    minima.crds.insert(minima.crds.end(), {mins[i].x, mins[i].y, mins[i].z});
    minima.sigma.push_back(min_sig[i]);
    minima.score.push_back(mins[i].score);
  }
  for (size_t i = 0; i < maxs.size(); i++) {
    maxima.crds.insert(maxima.crds.end(), {maxs[i].x, maxs[i].y, maxs[i].z});
    maxima.sigma.push_back(max_sig[i]);
    maxima.score.push_back(maxs[i].score);
  }
  if (best) { best[0] = gmin; best[1] = gmax; }
  if (finalize) blob_final_filter(minima, maxima, minima_threshold, maxima_threshold, use_threshold_ratios, gmin, gmax);
}

}  // namespace visfd_cuda
