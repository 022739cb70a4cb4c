// multi.cu -- the membrane pipeline on several GPUs of one node behind ONE C call: a C++ host such
// as filter_mrc hands over its host arrays and gets the result back, with no MPI, no Python and no process per GPU.
//
//   visfd_cuda_membrane_multi(ndev, devices, nx, ny, nz, src_host, mask_host, params, out_host, &threshold)
//
// One worker thread per device.  The volume is cut into Z-slabs exactly as visfd_b200/slab.py does for the
// process-per-GPU driver (rank boundaries on multiples of 8 planes, slabs widened by a halo of RAW SOURCE planes,
// halo = tv_halfwidth + 1 + gauss_halfwidth, slab start rounded down to a multiple of 8), so every stage is
// slab-local and the result is bit-identical to the one-GPU result.  Because the source lives in host memory, each
// device uploads its slab -- halo included -- straight from the caller's array over its own PCIe link: there is no
// device-to-device halo traffic at all.  The only exchange is the radix select of the `-tv-best` cut: three rounds of
// 2048-bin histograms, summed by the host between two thread barriers (the cut must be a GLOBAL order statistic,
// bin/filter_mrc/handlers.cpp:1766-1782).  Each device delivers its planes of the result into the caller's array
// chunk by chunk behind its voting kernels (visfd_cuda_vote_slab_host).
//
// The same device may be listed several times (two slabs on one GPU): that is how the one-GPU test box exercises
// the slab logic.
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

#include "common.cuh"
#include "kernels.cuh"

using namespace visfd_cuda;

namespace {

class Barrier {   // C++17: no std::barrier
  std::mutex m;
  std::condition_variable cv;
  const int n;
  int waiting = 0;
  uint64_t generation = 0;
 public:
  explicit Barrier(int count) : n(count) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const uint64_t g = generation;
    if (++waiting == n) {
      waiting = 0;
      generation++;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != g; });
    }
  }
};

struct SlabPlan {
  int64_t own0, own1;     // global planes this worker delivers
  int64_t slab0, slab1;   // global planes it holds
  int64_t vote0, vote1;   // global voter planes
};

// visfd_b200/slab.py: partition() + make_plan()
std::vector<SlabPlan> make_plans(int64_t nz, int world, int gauss_hw, int tv_hw) {
  const int64_t unit = (nz / 8 >= world) ? 8 : 1;
  const int64_t base = (nz / unit) / world, rem = (nz / unit) % world;
  const int64_t halo = tv_hw > 0 ? tv_hw + 1 + gauss_hw : 1 + gauss_hw;
  std::vector<SlabPlan> plans((size_t)world);
  int64_t z = 0;
  for (int r = 0; r < world; r++) {
    const int64_t n = (base + (r < rem ? 1 : 0)) * unit;
    SlabPlan &p = plans[(size_t)r];
    p.own0 = z;
    p.own1 = (r == world - 1) ? nz : z + n;
    z += n;
    p.slab0 = std::max<int64_t>(0, p.own0 - halo) / 8 * 8;
    p.slab1 = std::min(nz, p.own1 + halo);
    p.vote0 = std::max<int64_t>(0, p.own0 - std::max(tv_hw, 0));
    p.vote1 = std::min(nz, p.own1 + std::max(tv_hw, 0));
  }
  return plans;
}

// contexts are kept between calls: one per (slot, device)
std::mutex g_ctx_mutex;
std::vector<std::pair<int, visfd_ctx *> > g_ctx;   // index = slot

visfd_ctx *context_for(int slot, int device) {
  std::lock_guard<std::mutex> lk(g_ctx_mutex);
  if ((int)g_ctx.size() <= slot) g_ctx.resize((size_t)slot + 1, {-1, nullptr});
  auto &e = g_ctx[(size_t)slot];
  if (e.second && e.first != device) {
    visfd_cuda_destroy(e.second);
    e.second = nullptr;
  }
  if (!e.second) {
    visfd_ctx *c = nullptr;
    if (visfd_cuda_init(device, &c) != 0) throw Error(std::string("visfd_cuda_membrane_multi: ") + visfd_cuda_last_error());
    e = {device, c};
  }
  return e.second;
}

}  // namespace

extern "C" int visfd_cuda_membrane_multi(int ndev, const int *devices, int64_t nx, int64_t ny, int64_t nz,
                                         const float *src_host, const float *mask_host,
                                         const visfd_membrane_params *p, float *out_host, float *threshold_out,
                                         double *device_ms /* ndev values or NULL */) {
  try {
    VREQUIRE(ndev >= 1 && devices, "visfd_cuda_membrane_multi: need at least one device");
    VREQUIRE(nx > 0 && ny > 0 && nz > 0 && src_host && p && out_host, "visfd_cuda_membrane_multi: bad arguments");
    VREQUIRE(!is_device_pointer(src_host) && !is_device_pointer(out_host) && !is_device_pointer(mask_host),
             "visfd_cuda_membrane_multi takes HOST arrays (use the slab entry points for device-resident data)");
    const int gauss_hw = (int)floor(p->sigma * p->truncate_ratio);
    const int tv_hw = p->tv_sigma > 0.0f ? tv_halfwidth(p->tv_sigma, p->tv_cutoff_ratio) : 0;
    const int world = (int)std::min<int64_t>(ndev, std::max<int64_t>(1, nz / 8));   // at least 8 planes per worker
    const std::vector<SlabPlan> plans = make_plans(nz, world, gauss_hw, tv_hw);
    const size_t plane = (size_t)nx * (size_t)ny;

    std::vector<visfd_ctx *> ctxs((size_t)world);
    for (int r = 0; r < world; r++) ctxs[(size_t)r] = context_for(r, devices[r]);

    Barrier barrier(world);
    std::vector<std::string> errors((size_t)world);
    std::vector<std::vector<uint64_t> > hists((size_t)world, std::vector<uint64_t>(2048));
    // state of the distributed radix select, written by worker 0 between barriers
    uint32_t prefix = 0;
    int prefix_bits = 0;
    uint64_t rank_k = 0;
    float threshold = p->cut;
    std::atomic<bool> failed{false};
    std::mutex fail_mutex;
    std::vector<double> ms((size_t)world, 0.0);

    auto worker = [&](int r) {
      visfd_ctx *ctx = ctxs[(size_t)r];
      const SlabPlan &pl = plans[(size_t)r];
      const int64_t nzl = pl.slab1 - pl.slab0;
      auto fail = [&](const std::string &m) {
        std::lock_guard<std::mutex> lk(fail_mutex);
        errors[(size_t)r] = m;
        failed = true;
      };
      // stages run under try/catch individually: a worker that failed still meets the others at every barrier
      Scratch<float> src, mask, smoothed, saliency, out;
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      try {
        VCK(cudaSetDevice(ctx->device));
        VCK(cudaEventCreate(&e0));
        VCK(cudaEventCreate(&e1));
        VCK(cudaEventRecord(e0, ctx->stream));
        src.reset(ctx, (size_t)nzl * plane);
        smoothed.reset(ctx, (size_t)nzl * plane);
        saliency.reset(ctx, (size_t)nzl * plane);
        VCK(cudaMemcpyAsync(src.get(), src_host + (size_t)pl.slab0 * plane, (size_t)nzl * plane * sizeof(float),
                            cudaMemcpyHostToDevice, ctx->stream));
        if (mask_host) {
          mask.reset(ctx, (size_t)nzl * plane);
          VCK(cudaMemcpyAsync(mask.get(), mask_host + (size_t)pl.slab0 * plane, (size_t)nzl * plane * sizeof(float),
                              cudaMemcpyHostToDevice, ctx->stream));
        }
        if (visfd_cuda_ridge_saliency_slab(ctx, nx, ny, nzl, pl.slab0, nz, src.get(), mask.get(), p->sigma,
                                           p->truncate_ratio, p->eival_order, VISFD_SCORE_PLANAR, smoothed.get(),
                                           saliency.get()) != 0)
          throw Error(visfd_cuda_last_error());
        src.free();   // the raw slab is not needed any more
      } catch (const std::exception &ex) {
        fail(ex.what());
      }
      // ---- global cut: radix select over the workers' own planes ----
      if (p->cut_is_fraction) {
        bool first = true;
        for (;;) {
          barrier.wait();                    // prefix / prefix_bits of this round are final
          if (failed || prefix_bits >= 32) break;
          try {
            VCK(cudaSetDevice(ctx->device));
            select_hist_device(ctx, (i64)((pl.own1 - pl.own0) * (int64_t)plane),
                               saliency.get() + (size_t)(pl.own0 - pl.slab0) * plane,
                               mask_host ? mask.get() + (size_t)(pl.own0 - pl.slab0) * plane : nullptr, prefix,
                               prefix_bits, hists[(size_t)r].data());
            VCK(cudaStreamSynchronize(ctx->stream));
          } catch (const std::exception &ex) {
            fail(ex.what());
          }
          barrier.wait();                    // all histograms are in
          if (r == 0 && !failed) {
            std::vector<uint64_t> sum(2048, 0);
            for (int q = 0; q < world; q++)
              for (int b = 0; b < 2048; b++) sum[(size_t)b] += hists[(size_t)q][(size_t)b];
            if (first) {
              uint64_t total = 0;
              for (uint64_t v : sum) total += v;
              if (total == 0) {
                fail("saliency cut: no un-masked voxels");
              } else {
                // i = floor(n_voxels * fraction) in float, as handlers.cpp:1779-1782; clamped to the last element
                const uint64_t k = (uint64_t)std::max(0.0f, floorf((float)total * p->cut));
                rank_k = std::min(k, total - 1);
              }
            }
            if (!failed) select_step_host(sum.data(), &prefix, &prefix_bits, &rank_k);
          }
          first = false;
        }
        if (r == 0 && !failed) threshold = key_to_float(prefix);
        barrier.wait();                      // threshold is final
      }
      // ---- voting on the own planes, delivered into the caller's array ----
      if (!failed) {
        try {
          VCK(cudaSetDevice(ctx->device));
          out.reset(ctx, (size_t)(pl.own1 - pl.own0) * plane);
          if (visfd_cuda_vote_slab_host(ctx, nx, ny, nzl, pl.slab0, nz, pl.own0 - pl.slab0, pl.own1 - pl.slab0,
                                        pl.vote0 - pl.slab0, pl.vote1 - pl.slab0, saliency.get(), smoothed.get(),
                                        mask_host ? mask.get() : nullptr, threshold, p, out.get(), nullptr,
                                        out_host + (size_t)pl.own0 * plane) != 0)
            throw Error(visfd_cuda_last_error());
          VCK(cudaEventRecord(e1, ctx->stream));
          VCK(cudaEventSynchronize(e1));
          float t = 0;
          VCK(cudaEventElapsedTime(&t, e0, e1));
          ms[(size_t)r] = t;
        } catch (const std::exception &ex) {
          fail(ex.what());
        }
      }
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
    };

    std::vector<std::thread> threads;
    for (int r = 1; r < world; r++) threads.emplace_back(worker, r);
    worker(0);
    for (auto &t : threads) t.join();
    for (int r = 0; r < world; r++)
      if (!errors[(size_t)r].empty())
        throw Error("visfd_cuda_membrane_multi, slab " + std::to_string(r) + " on device " +
                    std::to_string(devices[r]) + ": " + errors[(size_t)r]);
    if (threshold_out) *threshold_out = threshold;
    if (device_ms)
      for (int r = 0; r < ndev; r++) device_ms[r] = r < world ? ms[(size_t)r] : 0.0;
    return 0;
  } catch (const std::exception &ex) {
    set_last_error(ex.what());
    return 1;
  }
}
