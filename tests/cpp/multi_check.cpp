// tests/cpp/multi_check.cpp -- drives visfd_cuda_membrane_multi the way a C++ host (filter_mrc) would: plain host
// arrays, a list of CUDA devices, one call.  The result must equal visfd_cuda_membrane on one GPU BIT FOR BIT
// (thresholds included) for every split.  Usage: multi_check [ndev]   (default: all visible devices; with one
// visible device the same device is listed several times, which exercises the slab logic all the same).
// Exit code 0 = all checks passed.  Run by tests/test_gpu_multi.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <random>
#include <vector>
#include "../../include/visfd_cuda.h"


static std::vector<float> volume(int nx, int ny, int nz, unsigned seed) {
  std::mt19937 g(seed);
  std::normal_distribution<float> noise(0.0f, 1.0f);
  std::vector<float> v((size_t)nx * ny * nz);
  const float cx = 0.5f * nx, cy = 0.5f * ny, cz = 0.5f * nz, R = 0.3f * std::min(nx, std::min(ny, nz));
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        const float r = std::sqrt((x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz));
        v[((size_t)z * ny + y) * nx + x] = noise(g) - 3.0f * std::exp(-0.5f * (r - R) * (r - R) / 4.0f);
      }
  return v;
}

int main(int argc, char **argv) {
  const int visible = visfd_cuda_device_count();
  if (visible < 1) { std::printf("no CUDA device\n"); return 2; }
  const int want = argc > 1 ? std::atoi(argv[1]) : 0;
  int failures = 0;
  visfd_ctx *ctx = nullptr;
  if (visfd_cuda_init(0, &ctx) != 0) { std::printf("init: %s\n", visfd_cuda_last_error()); return 2; }
  struct Case { int nx, ny, nz; float sigma, tv_sigma, cut; int fraction; bool mask; };
  const Case cases[] = {{48, 40, 72, 1.5f, 4.3f, 0.08f, 1, false},    // 9 units of 8 planes
                        {40, 36, 61, 1.2f, 3.1f, 0.10f, 1, true},     // ragged tail, mask
                        {64, 48, 96, 2.0f, 6.0f, 0.05f, 1, false},    // halo wider than a slab at 8 workers
                        {32, 32, 40, 1.0f, 0.0f, 0.20f, 1, false},    // no voting: saliency after the cut
                        {40, 40, 48, 1.5f, 4.0f, 0.002f, 0, false}};  // absolute threshold
  for (const Case &c : cases) {
    const size_t N = (size_t)c.nx * c.ny * c.nz;
    std::vector<float> src = volume(c.nx, c.ny, c.nz, 11u + (unsigned)c.nz), mask, one(N), multi(N);
    if (c.mask) {
      mask.assign(N, 1.0f);
      for (int z = 0; z < c.nz; z++)
        for (int y = 0; y < c.ny; y++)
          for (int x = 0; x < c.nx; x++)
            if (x < 3 || y > c.ny - 5 || (z % 17) == 0) mask[((size_t)z * c.ny + y) * c.nx + x] = 0.0f;
    }
    visfd_membrane_params p;
    p.sigma = c.sigma; p.truncate_ratio = 2.6482f; p.eival_order = VISFD_DECREASING_EIVALS;
    p.cut = c.cut; p.cut_is_fraction = c.fraction;
    p.tv_sigma = c.tv_sigma; p.tv_exponent = 4; p.tv_cutoff_ratio = 1.41421354f;
    float thr1 = 0;
    if (visfd_cuda_membrane(ctx, c.nx, c.ny, c.nz, src.data(), c.mask ? mask.data() : nullptr, &p, one.data(), nullptr,
                            nullptr, nullptr, &thr1) != 0) {
      std::printf("one GPU: %s\n", visfd_cuda_last_error());
      return 2;
    }
    for (int ndev : {1, 2, 3, 4, 8}) {
      if (want && ndev != want && ndev != 1) continue;
      std::vector<int> devs((size_t)ndev);
      for (int r = 0; r < ndev; r++) devs[(size_t)r] = r % visible;
      std::fill(multi.begin(), multi.end(), -7.0f);
      float thrn = 0;
      std::vector<double> ms((size_t)ndev);
      if (visfd_cuda_membrane_multi(ndev, devs.data(), c.nx, c.ny, c.nz, src.data(), c.mask ? mask.data() : nullptr, &p,
                                    multi.data(), &thrn, ms.data()) != 0) {
        std::printf("FAIL %dx%dx%d ndev %d: %s\n", c.nx, c.ny, c.nz, ndev, visfd_cuda_last_error());
        failures++;
        continue;
      }
      size_t diff = 0;
      for (size_t i = 0; i < N; i++) diff += std::memcmp(&one[i], &multi[i], sizeof(float)) != 0;
      const bool ok = diff == 0 && std::memcmp(&thr1, &thrn, sizeof(float)) == 0;
      std::printf("%s %dx%dx%d%s tv %.1f: %d worker(s) on %d device(s): %zu voxels differ, threshold %.9g vs %.9g\n",
                  ok ? "ok  " : "FAIL", c.nx, c.ny, c.nz, c.mask ? " masked" : "", c.tv_sigma, ndev, std::min(ndev, visible), diff,
                  thrn, thr1);
      failures += !ok;
    }
  }
  // error path: a device that does not exist
  {
    std::vector<float> src = volume(16, 16, 16, 3), out(16 * 16 * 16);
    visfd_membrane_params p = {1.0f, 2.6482f, 1, 0.1f, 1, 2.0f, 4, 1.41421354f};
    const int bad[2] = {0, 1000};
    if (visfd_cuda_membrane_multi(2, bad, 16, 16, 16, src.data(), nullptr, &p, out.data(), nullptr, nullptr) == 0) {
      std::printf("FAIL: device 1000 accepted\n");
      failures++;
    } else {
      std::printf("ok   bad device refused: %s\n", visfd_cuda_last_error());
    }
  }
  visfd_cuda_destroy(ctx);
  std::printf("%s (%d failures)\n", failures ? "FAILED" : "OK", failures);
  return failures ? 1 : 0;
}
