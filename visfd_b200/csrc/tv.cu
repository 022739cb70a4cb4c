// tv.cu -- TV3D dense stick voting (lib/visfd/feature.hpp:1712-2037, receiver loop
// TVReceiveStickVotes :2218-2384) as a sparse-voter / dense-receiver gather.
//
// The reference visits all (2hw+1)^3 offsets of every receiver and skips the ~95 % of
// them whose voter has zero saliency.  Here the voters (saliency != 0 after the cut,
// mask != 0) are first compacted into a list ordered by 8x8x8 BRICK (count -> exclusive
// scan -> fill: three streaming passes over the saliency volume), 32 B per voter:
//     A = {x, y, z, saliency * mask_weight / table_total}   B = {nx, ny, nz, 0}
// The gather kernel runs one CTA per 8x8x8 receiver tile.  The tile's neighbourhood is
// a set of <= (2R+2)(2R+1) brick ROWS (R = ceil(hw/8)), each a contiguous range of the
// voter list trimmed to the bricks that can reach the tile; the ranges are streamed
// through a double-buffered shared-memory ring with cp.async (LDGSTS).  Each of the 8
// warps owns a 4x4x4 receiver patch (32 lanes x 2 z-adjacent receivers per lane, 12
// register accumulators); lanes first test 32 voters in parallel against the patch
// (ballot), then the warp walks the surviving voters, broadcasting each one from shared
// memory.  Per (receiver, voter) pair that is ~33 FP32-pipe instructions (35 FLOP by
// the SURVEY's count), with no table lookup: the radial decay exp(-r^2/sigma^2) is one
// MUFU.EX2 and 1/r one MUFU.RSQ.  The reference's decay TABLE is only needed for two
// things, both computed on the host with the reference's own float expressions
// (lib/visfd/filter3d.hpp:546-601): the normalisation constant (sum over the cube) and
// which lattice points on the shell r^2 == hw^2 survive the truncation threshold.
//
// Epilogue (fused, accumulators still in registers): optional store of the 6-component
// tensor (-save-progress) and DiagonalizeFlatSym3 + ScoreTensorPlanar/Linear
// (bin/filter_mrc/handlers.cpp:1870-1892) in double.
#include "common.cuh"
#include "kernels.cuh"
#include "eigen3.cuh"
#include <cmath>
#include <algorithm>

namespace visfd_cuda {

constexpr int BR = 8;            // brick edge
constexpr int BR3 = BR * BR * BR;
constexpr int TV_CHUNK = 512;    // voters per shared-memory stage
constexpr int TV_THREADS = 256;
constexpr int TV_MAX_REACH = 7;  // bricks; hw <= 56
constexpr int TV_MAX_ROWS = (2 * TV_MAX_REACH + 2) * (2 * TV_MAX_REACH + 1);
constexpr int TV_MAX_SHELL = 512;

int tv_halfwidth(float sigma, float cutoff_ratio) {
  return (int)floor(sigma * cutoff_ratio);  // feature.hpp:1669-1675
}

// ---------------------------------------------------------------------------------
// host: what we need from the reference's decay table
// ---------------------------------------------------------------------------------
struct DecayInfo {
  float total;                      // sum of the un-normalised table (float, raster order)
  std::vector<uint32_t> shell_keep; // packed |dx| | |dy|<<8 | |dz|<<16 of surviving shell points
};

// GenFilterGenGauss3D(sigma, m=2, hw): lib/visfd/filter3d.hpp:546-601
static DecayInfo decay_info(float sigma, int hw) {
  DecayInfo info;
  float thr = 1.0f;
  {
    float h = (sigma > 0) ? expf(-powf(hw / sigma, 2.0f)) : 1.0f;
    if (h < thr) thr = h;
  }
  float total = 0;
  const int hw2 = hw * hw;
  for (int iz = -hw; iz <= hw; iz++)
    for (int iy = -hw; iy <= hw; iy++)
      for (int ix = -hw; ix <= hw; ix++) {
        float x = (!((sigma == 0.0f) && (ix == 0))) ? ix / sigma : 0.0f;
        float y = (!((sigma == 0.0f) && (iy == 0))) ? iy / sigma : 0.0f;
        float z = (!((sigma == 0.0f) && (iz == 0))) ? iz / sigma : 0.0f;
        float r = sqrtf(x * x + y * y + z * z);
        float h = (r > 0) ? expf(-powf(r, 2.0f)) : 1.0f;
        if (fabsf(h) < thr) h = 0.0f;
        total += h;
        int r2 = ix * ix + iy * iy + iz * iz;
        if (r2 == hw2 && h != 0.0f && ix >= 0 && iy >= 0 && iz >= 0)
          info.shell_keep.push_back((uint32_t)ix | ((uint32_t)iy << 8) | ((uint32_t)iz << 16));
        // Lattice points strictly inside (outside) the shell are kept (dropped) by a
        // margin of exp(1/sigma^2) - 1 >> float epsilon, so r^2 < hw^2 decides them.
      }
  info.total = total;
  return info;
}

// ---------------------------------------------------------------------------------
// voter list construction
// ---------------------------------------------------------------------------------
struct VoterSrc {
  const float *sal;
  const float *mask_src;
  float thr;
  int nx, ny;
  i64 nz;  // slab planes
  int nbx, nby, nbz;
};

__device__ __forceinline__ bool is_voter(const VoterSrc &v, int x, int y, i64 z, float &w) {
  if (x >= v.nx || y >= v.ny || z >= v.nz) return false;
  i64 i = (z * v.ny + y) * (i64)v.nx + x;
  float s = __ldg(v.sal + i);
  if (!(s >= v.thr) || s == 0.0f) return false;  // cut: handlers.cpp:1792; skip: feature.hpp:2268
  if (v.mask_src) {
    float m = __ldg(v.mask_src + i);
    if (m == 0.0f) return false;                 // feature.hpp:2259-2260
    s *= m;                                      // a weight multiplies the decay, :2261-2265
  }
  w = s;
  return true;
}

// one CTA (512 threads) per brick
__global__ void __launch_bounds__(BR3) voter_count_kernel(VoterSrc v, uint32_t *__restrict__ counts) {
  const int b = blockIdx.x;
  const int bx = b % v.nbx, by = (b / v.nbx) % v.nby, bz = b / (v.nbx * v.nby);
  const int t = threadIdx.x;
  float w;
  bool p = is_voter(v, bx * BR + (t & 7), by * BR + ((t >> 3) & 7), (i64)bz * BR + (t >> 6), w);
  int c = __syncthreads_count(p);
  if (t == 0) counts[b] = (uint32_t)c;
}

// Exclusive scan of n uint32 counters into off[0..n] (off[n] = total): per-block scan,
// scan of the block sums by one CTA, then offset add.
constexpr int SCAN_T = 1024, SCAN_I = 4, SCAN_B = SCAN_T * SCAN_I;

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *sh, uint32_t &total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t x = v;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[w] = x;
  __syncthreads();
  if (w == 0) {
    uint32_t s = sh[lane];
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    sh[lane] = s;
  }
  __syncthreads();
  uint32_t base = w ? sh[w - 1] : 0;
  total = sh[31];
  __syncthreads();
  return base + x - v;
}

__global__ void __launch_bounds__(SCAN_T)
scan_local_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t *__restrict__ sums, i64 n) {
  __shared__ uint32_t sh[32];
  i64 base = (i64)blockIdx.x * SCAN_B + (i64)threadIdx.x * SCAN_I;
  uint32_t v[SCAN_I], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_I; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    s += v[k];
  }
  uint32_t tot;
  uint32_t ex = block_excl_scan(s, sh, tot);
#pragma unroll
  for (int k = 0; k < SCAN_I; k++) {
    if (base + k < n) out[base + k] = ex;
    ex += v[k];
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_T) scan_sums_kernel(uint32_t *__restrict__ sums, i64 m, uint32_t *__restrict__ total_out) {
  __shared__ uint32_t sh[32];
  uint32_t carry = 0;
  for (i64 base = 0; base < m; base += SCAN_T) {
    i64 i = base + threadIdx.x;
    uint32_t v = (i < m) ? sums[i] : 0u;
    uint32_t tot;
    uint32_t ex = block_excl_scan(v, sh, tot);
    if (i < m) sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_T)
scan_add_kernel(uint32_t *__restrict__ out, const uint32_t *__restrict__ sums, i64 n, const uint32_t *__restrict__ total) {
  i64 base = (i64)blockIdx.x * SCAN_B + (i64)threadIdx.x * SCAN_I;
  uint32_t add = sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_I; k++)
    if (base + k < n) out[base + k] += add;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;
}

struct DirSrc {
  const float *direction;  // N*3, or NULL
  const float *smoothed;   // used when direction == NULL
  float ridge_sigma;
  int order;
  i64 z_offset, nz_global;
};

// Declared in ridge.cu's translation unit as static device code; restated here because
// the fill kernel needs the same stencil (kept tiny on purpose).
__device__ __forceinline__ void voter_direction(const DirSrc &d, int nx, int ny, int x, int y, i64 z, float n[3]) {
  i64 i = (z * ny + y) * (i64)nx + x;
  if (d.direction) {
    n[0] = __ldg(d.direction + 3 * i + 0);
    n[1] = __ldg(d.direction + 3 * i + 1);
    n[2] = __ldg(d.direction + 3 * i + 2);
    return;
  }
  // finite-difference Hessian of the smoothed image with the centre clamped at the
  // GLOBAL border (visfd_utils.hpp:530-616), times sigma^2 (feature.hpp:1331-1333)
  int cx = x, cy = y;
  i64 zg = d.z_offset + z;
  if (cx == 0) cx++; else if (cx == nx - 1) cx--;
  if (cy == 0) cy++; else if (cy == ny - 1) cy--;
  if (zg == 0) zg++; else if (zg == d.nz_global - 1) zg--;
  const i64 sy = nx, sz = (i64)nx * ny;
  const float *p = d.smoothed + ((zg - d.z_offset) * ny + cy) * (i64)nx + cx;
#define F(a, b, c) __ldg(p + (a) + (b) * sy + (c) * sz)
  float s2 = __fmul_rn(d.ridge_sigma, d.ridge_sigma);
  float c = F(0, 0, 0), c2 = __fmul_rn(2.0f, c);
  float hxx = __fmul_rn(__fsub_rn(__fadd_rn(F(1, 0, 0), F(-1, 0, 0)), c2), s2);
  float hyy = __fmul_rn(__fsub_rn(__fadd_rn(F(0, 1, 0), F(0, -1, 0)), c2), s2);
  float hzz = __fmul_rn(__fsub_rn(__fadd_rn(F(0, 0, 1), F(0, 0, -1)), c2), s2);
  float xy = __fsub_rn(__fsub_rn(__fadd_rn(F(1, 1, 0), F(-1, -1, 0)), F(1, -1, 0)), F(-1, 1, 0));
  float yz = __fsub_rn(__fsub_rn(__fadd_rn(F(0, 1, 1), F(0, -1, -1)), F(0, 1, -1)), F(0, -1, 1));
  float xz = __fsub_rn(__fsub_rn(__fadd_rn(F(1, 0, 1), F(-1, 0, -1)), F(-1, 0, 1)), F(1, 0, -1));
#undef F
  Sym3d m = {hxx, hyy, hzz, __fmul_rn(__fmul_rn(0.25f, xy), s2), __fmul_rn(__fmul_rn(0.25f, yz), s2),
             __fmul_rn(__fmul_rn(0.25f, xz), s2)};
  double ev[3], e0[3];
  sym3_eigen_first(m, d.order, ev, e0);
  n[0] = (float)e0[0];
  n[1] = (float)e0[1];
  n[2] = (float)e0[2];
}

__global__ void __launch_bounds__(BR3)
voter_fill_kernel(VoterSrc v, DirSrc d, const uint32_t *__restrict__ off, float inv_total,
                  float4 *__restrict__ va, float4 *__restrict__ vb) {
  __shared__ uint32_t wsum[BR3 / 32];
  const int b = blockIdx.x;
  const uint32_t o0 = off[b], o1 = off[b + 1];
  if (o0 == o1) return;  // uniform per CTA
  const int bx = b % v.nbx, by = (b / v.nbx) % v.nby, bz = b / (v.nbx * v.nby);
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int x = bx * BR + (t & 7), y = by * BR + ((t >> 3) & 7);
  const i64 z = (i64)bz * BR + (t >> 6);
  float wt = 0.0f;
  bool p = is_voter(v, x, y, z, wt);
  unsigned bal = __ballot_sync(0xffffffffu, p);
  if (lane == 0) wsum[w] = __popc(bal);
  __syncthreads();
  if (!p) return;
  uint32_t rank = __popc(bal & ((1u << lane) - 1u));
  for (int k = 0; k < w; k++) rank += wsum[k];
  float n[3];
  voter_direction(d, v.nx, v.ny, x, y, z, n);
  va[o0 + rank] = make_float4((float)x, (float)y, (float)z, wt * inv_total);
  vb[o0 + rank] = make_float4(n[0], n[1], n[2], 0.0f);
}

// ---------------------------------------------------------------------------------
// gather
// ---------------------------------------------------------------------------------
struct GatherArgs {
  const float4 *va, *vb;
  const uint32_t *off;
  const uint32_t *shell;  // device copy of DecayInfo::shell_keep
  int n_shell;
  int nx, ny;
  i64 nz;                 // slab planes (voter bricks cover [0,nz))
  int nbx, nby, nbz;
  i64 own_z0, own_z1;     // receiver planes (slab-local)
  int ntx, nty;           // receiver tiles in x,y
  int hw;
  float hw2;              // (float) hw*hw
  float neg_c;            // -log2(e)/sigma^2
  float half_exp;         // exponent/2, generic path
  const float *mask_dst;  // slab-indexed, or NULL
  float *tensor;          // own-planes-indexed * 6, or NULL
  float *score;           // own-planes-indexed, or NULL
  int order, score_kind;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float fast_rsqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __noinline__ float shell_weight(const uint32_t *shell, int n, float dx, float dy, float dz) {
  uint32_t key = (uint32_t)fabsf(dx) | ((uint32_t)fabsf(dy) << 8) | ((uint32_t)fabsf(dz) << 16);
  for (int k = 0; k < n; k++)
    if (shell[k] == key) return 1.0f;
  return 0.0f;
}

// EXPO: 2 or 4 = the reference's special cases (feature.hpp:2328-2339); 0 = pow()
template <int EXPO, bool CURVES>
__device__ __forceinline__ void vote_pair(float dx, float dy, float dz, float dxy2, float sxy,
                                          const float4 &a, const float4 &n, const GatherArgs &g,
                                          const uint32_t *shell, float T[6]) {
  float r2 = fmaf(dz, dz, dxy2);
  float sd = fmaf(dz, n.z, sxy);
  float inv = fast_rsqrt(fmaxf(r2, 0.25f));  // r2 == 0: sd == 0, the vote is sal*decay*n n^T
  float inv2 = inv * inv;
  float e = fast_ex2(r2 * g.neg_c);
  float sin2 = sd * sd * inv2;               // (r_hat . n)^2
  float ang2 = CURVES ? sin2 : 1.0f - sin2;  // feature.hpp:2318-2326
  float ang;
  if (EXPO == 2) ang = ang2;
  else if (EXPO == 4) ang = ang2 * ang2;
  else ang = __powf(fmaxf(ang2, 0.0f), g.half_exp);
  float w = a.w * e * ang;
  w = (r2 < g.hw2) ? w : 0.0f;
  if (r2 == g.hw2) w = a.w * e * ang * shell_weight(shell, g.n_shell, dx, dy, dz);
  float t = (sd + sd) * inv2;
  float vx, vy, vz;  // rotated normal: 2 s r_hat - n (surfaces) / n - 2 s r_hat (curves)
  if (CURVES) {
    vx = fmaf(-t, dx, n.x); vy = fmaf(-t, dy, n.y); vz = fmaf(-t, dz, n.z);
  } else {
    vx = fmaf(t, dx, -n.x); vy = fmaf(t, dy, -n.y); vz = fmaf(t, dz, -n.z);
  }
  float wx = w * vx, wy = w * vy, wz = w * vz;
  T[0] = fmaf(wx, vx, T[0]);
  T[1] = fmaf(wy, vy, T[1]);
  T[2] = fmaf(wz, vz, T[2]);
  T[3] = fmaf(wx, vy, T[3]);
  T[4] = fmaf(wy, vz, T[4]);
  T[5] = fmaf(wx, vz, T[5]);
}

__device__ __forceinline__ int axis_gap(int lo_a, int hi_a, int lo_b, int hi_b) {
  // distance between integer intervals [lo_a,hi_a] and [lo_b,hi_b]
  return max(0, max(lo_b - hi_a, lo_a - hi_b));
}

template <int EXPO, bool CURVES>
__global__ void __launch_bounds__(TV_THREADS, 2) tv_gather_kernel(GatherArgs g) {
  __shared__ __align__(16) float4 s_a[2][TV_CHUNK];
  __shared__ __align__(16) float4 s_b[2][TV_CHUNK];
  __shared__ uint32_t s_row_start[TV_MAX_ROWS];
  __shared__ uint32_t s_row_prefix[TV_MAX_ROWS + 1];
  __shared__ uint32_t s_shell[TV_MAX_SHELL];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int tx = tile % g.ntx, ty = (tile / g.ntx) % g.nty, tz = tile / (g.ntx * g.nty);
  const int X0 = tx * BR, Y0 = ty * BR;
  const i64 Z0 = g.own_z0 + (i64)tz * BR;
  const int Z1 = (int)min((i64)(Z0 + BR - 1), g.own_z1 - 1);  // last receiver plane of the tile

  // ---- neighbourhood rows --------------------------------------------------------
  const int bz_lo = (int)max((i64)0, (Z0 - g.hw) >> 3);
  const int bz_hi = (int)min((i64)g.nbz - 1, (i64)(Z1 + g.hw) >> 3);
  const int by_lo = max(0, (Y0 - g.hw) >> 3);
  const int by_hi = min(g.nby - 1, (Y0 + BR - 1 + g.hw) >> 3);
  const int nry = by_hi - by_lo + 1;
  const int nrows = max(0, (bz_hi - bz_lo + 1)) * nry;
  for (int r = tid; r < nrows; r += TV_THREADS) {
    int bz = bz_lo + r / nry, by = by_lo + r % nry;
    int dz = axis_gap((int)Z0, Z1, bz * BR, bz * BR + BR - 1);
    int dy = axis_gap(Y0, Y0 + BR - 1, by * BR, by * BR + BR - 1);
    int rem = g.hw * g.hw - dz * dz - dy * dy;
    uint32_t start = 0, len = 0;
    if (rem >= 0) {
      int d = (int)floorf(sqrtf((float)rem));
      while ((d + 1) * (d + 1) <= rem) d++;
      while (d * d > rem) d--;
      int bx_lo = max(0, (X0 - d) >> 3), bx_hi = min(g.nbx - 1, (X0 + BR - 1 + d) >> 3);
      i64 rb = ((i64)bz * g.nby + by) * g.nbx;
      start = g.off[rb + bx_lo];
      len = g.off[rb + bx_hi + 1] - start;
    }
    s_row_start[r] = start;
    s_row_prefix[r + 1] = len;
  }
  for (int k = tid; k < g.n_shell; k += TV_THREADS) s_shell[k] = g.shell[k];
  __syncthreads();
  if (tid == 0) {
    uint32_t acc = 0;
    s_row_prefix[0] = 0;
    for (int r = 0; r < nrows; r++) {
      acc += s_row_prefix[r + 1];
      s_row_prefix[r + 1] = acc;
    }
  }
  __syncthreads();
  const uint32_t total = s_row_prefix[nrows];
  const int nchunks = (int)((total + TV_CHUNK - 1) / TV_CHUNK);

  // ---- receivers of this lane ------------------------------------------------------
  const int px = X0 + (warp & 1) * 4, py = Y0 + ((warp >> 1) & 1) * 4;
  const int pz = (int)Z0 + (warp >> 2) * 4;
  const int ix = px + (lane & 3), iy = py + ((lane >> 2) & 3), iz = pz + (lane >> 4) * 2;
  const float rx = (float)ix, ry = (float)iy, rz0 = (float)iz, rz1 = (float)(iz + 1);
  const float pcx = px + 1.5f, pcy = py + 1.5f, pcz = pz + 1.5f;
  float T0[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, T1[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

  // ---- stream the voter ranges ---------------------------------------------------------
  int cur_row = 0;  // per-thread cursor into the row table (flat indices only grow)
  auto issue = [&](int c) {
    const int buf = c & 1;
    const uint32_t base = (uint32_t)c * TV_CHUNK;
#pragma unroll
    for (int k = 0; k < TV_CHUNK / TV_THREADS; k++) {
      uint32_t e = base + tid + k * TV_THREADS;
      if (e < total) {
        while (e >= s_row_prefix[cur_row + 1]) cur_row++;
        uint32_t gi = s_row_start[cur_row] + (e - s_row_prefix[cur_row]);
        cp_async16(&s_a[buf][e - base], g.va + gi);
        cp_async16(&s_b[buf][e - base], g.vb + gi);
      }
    }
    cp_async_commit();
  };
  if (nchunks > 0) issue(0);
  for (int c = 0; c < nchunks; c++) {
    if (c + 1 < nchunks) issue(c + 1); else cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int buf = c & 1;
    const int cn = (int)min((uint32_t)TV_CHUNK, total - (uint32_t)c * TV_CHUNK);
    for (int base = 0; base < cn; base += 32) {
      bool pass = false;
      if (base + lane < cn) {
        float4 a = s_a[buf][base + lane];
        float ddx = fmaxf(fabsf(a.x - pcx) - 1.5f, 0.0f);
        float ddy = fmaxf(fabsf(a.y - pcy) - 1.5f, 0.0f);
        float ddz = fmaxf(fabsf(a.z - pcz) - 1.5f, 0.0f);
        pass = fmaf(ddx, ddx, fmaf(ddy, ddy, ddz * ddz)) <= g.hw2;
      }
      unsigned m = __ballot_sync(0xffffffffu, pass);
      while (m) {
        int j = base + __ffs(m) - 1;
        m &= m - 1;
        const float4 a = s_a[buf][j];
        const float4 n = s_b[buf][j];
        float dx = rx - a.x, dy = ry - a.y;
        float dxy2 = fmaf(dx, dx, dy * dy);
        float sxy = fmaf(dx, n.x, dy * n.y);
        vote_pair<EXPO, CURVES>(dx, dy, rz0 - a.z, dxy2, sxy, a, n, g, s_shell, T0);
        vote_pair<EXPO, CURVES>(dx, dy, rz1 - a.z, dxy2, sxy, a, n, g, s_shell, T1);
      }
    }
    __syncthreads();
  }

  // ---- epilogue ---------------------------------------------------------------------
  if (ix < g.nx && iy < g.ny) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const i64 z = iz + r;
      if (z >= g.own_z1) break;
      const float *T = r ? T1 : T0;
      const i64 slab_i = (z * g.ny + iy) * (i64)g.nx + ix;
      const i64 out_i = ((z - g.own_z0) * g.ny + iy) * (i64)g.nx + ix;
      const bool masked = g.mask_dst && __ldg(g.mask_dst + slab_i) == 0.0f;  // feature.hpp:2002-2003
      if (g.tensor) {
        float *o = g.tensor + 6 * out_i;
#pragma unroll
        for (int k = 0; k < 6; k++) o[k] = masked ? 0.0f : T[k];
      }
      if (g.score) {
        float sc = 0.0f;
        if (!masked) {
          Sym3d m = {T[0], T[1], T[2], T[3], T[4], T[5]};
          double ev[3];
          sym3_eigenvalues(m, g.order, ev);
          sc = score_from_eivals(ev, g.score_kind, 1);
        }
        g.score[out_i] = sc;
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------
void tv_device(visfd_ctx *ctx, i64 nx, i64 ny, i64 nz_local, i64 z_offset, i64 nz_global,
               i64 own_z0, i64 own_z1, const float *saliency, float thr, const float *direction,
               const float *smoothed, float ridge_sigma, int eival_order, int score_kind,
               const float *mask_src, const float *mask_dst, const TVParams &p, float *tensor,
               float *score) {
  VREQUIRE(nx > 0 && ny > 0 && nz_local > 0, "empty volume");
  VREQUIRE(own_z0 >= 0 && own_z1 <= nz_local && own_z0 <= own_z1, "receiver planes outside the slab");
  VREQUIRE(p.sigma > 0.0f, "tensor-voting sigma must be positive");
  VREQUIRE(direction || smoothed, "tensor voting needs voter directions");
  const int hw = tv_halfwidth(p.sigma, p.cutoff_ratio);
  VREQUIRE(hw >= 0 && hw <= TV_MAX_REACH * BR, "tensor-voting radius too large (max 56 voxels)");
  VREQUIRE(nx <= (1 << 23) && ny <= (1 << 23) && nz_local <= (1 << 23), "slab too large");
  if (own_z1 == own_z0) return;
  DecayInfo info = decay_info(p.sigma, hw);
  VREQUIRE((int)info.shell_keep.size() <= TV_MAX_SHELL, "too many lattice points on the support shell");

  const int nbx = (int)div_up(nx, BR), nby = (int)div_up(ny, BR), nbz = (int)div_up(nz_local, BR);
  const i64 n_bricks = (i64)nbx * nby * nbz;
  VREQUIRE(n_bricks < 2147483647LL, "too many bricks for one launch");
  VoterSrc vs{saliency, mask_src, thr, (int)nx, (int)ny, nz_local, nbx, nby, nbz};

  Scratch<uint32_t> counts(ctx, n_bricks), off(ctx, n_bricks + 1);
  const i64 n_scan_blocks = (n_bricks + SCAN_B - 1) / SCAN_B;
  Scratch<uint32_t> sums(ctx, n_scan_blocks + 1);
  uint32_t n_voters = 0;
  Scratch<float4> va, vb;
  {
    StageTimer t(ctx, "compact");
    voter_count_kernel<<<(unsigned)n_bricks, BR3, 0, ctx->stream>>>(vs, counts.get());
    VCK(cudaGetLastError());
    scan_local_kernel<<<(unsigned)n_scan_blocks, SCAN_T, 0, ctx->stream>>>(counts.get(), off.get(), sums.get(), n_bricks);
    VCK(cudaGetLastError());
    scan_sums_kernel<<<1, SCAN_T, 0, ctx->stream>>>(sums.get(), n_scan_blocks, sums.get() + n_scan_blocks);
    VCK(cudaGetLastError());
    scan_add_kernel<<<(unsigned)n_scan_blocks, SCAN_T, 0, ctx->stream>>>(off.get(), sums.get(), n_bricks,
                                                                        sums.get() + n_scan_blocks);
    VCK(cudaGetLastError());
    ctx->count_launch(4);
    VCK(cudaMemcpyAsync(&n_voters, sums.get() + n_scan_blocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    VCK(cudaStreamSynchronize(ctx->stream));
    // (a 32-bit voter count: 4.29e9 voters would need 137 GB of voter records anyway)
    va.reset(ctx, std::max<size_t>(n_voters, 1));
    vb.reset(ctx, std::max<size_t>(n_voters, 1));
    if (n_voters > 0) {
      DirSrc ds{direction, smoothed, ridge_sigma, eival_order, z_offset, nz_global};
      voter_fill_kernel<<<(unsigned)n_bricks, BR3, 0, ctx->stream>>>(vs, ds, off.get(), 1.0f / info.total,
                                                                     va.get(), vb.get());
      VCK(cudaGetLastError());
      ctx->count_launch();
    }
  }
  ctx->last_voters = n_voters;

  Scratch<uint32_t> shell(ctx, std::max<size_t>(info.shell_keep.size(), 1));
  if (!info.shell_keep.empty())
    VCK(cudaMemcpyAsync(shell.get(), info.shell_keep.data(), info.shell_keep.size() * sizeof(uint32_t),
                        cudaMemcpyHostToDevice, ctx->stream));

  GatherArgs g;
  g.va = va.get(); g.vb = vb.get(); g.off = off.get();
  g.shell = shell.get(); g.n_shell = (int)info.shell_keep.size();
  g.nx = (int)nx; g.ny = (int)ny; g.nz = nz_local;
  g.nbx = nbx; g.nby = nby; g.nbz = nbz;
  g.own_z0 = own_z0; g.own_z1 = own_z1;
  g.ntx = nbx; g.nty = nby;
  g.hw = hw; g.hw2 = (float)(hw * hw);
  g.neg_c = (float)(-1.4426950408889634 / ((double)p.sigma * (double)p.sigma));
  g.half_exp = 0.5f * (float)p.exponent;
  g.mask_dst = mask_dst; g.tensor = tensor; g.score = score;
  g.order = eival_order; g.score_kind = score_kind;
  const i64 ntz = div_up(own_z1 - own_z0, BR);
  const i64 n_tiles = (i64)g.ntx * g.nty * ntz;
  VREQUIRE(n_tiles < 2147483647LL, "too many receiver tiles for one launch");
  {
    StageTimer t(ctx, "tv");
    const unsigned grid = (unsigned)n_tiles;
    if (p.curves) {
      if (p.exponent == 2) tv_gather_kernel<2, true><<<grid, TV_THREADS, 0, ctx->stream>>>(g);
      else if (p.exponent == 4) tv_gather_kernel<4, true><<<grid, TV_THREADS, 0, ctx->stream>>>(g);
      else tv_gather_kernel<0, true><<<grid, TV_THREADS, 0, ctx->stream>>>(g);
    } else {
      if (p.exponent == 2) tv_gather_kernel<2, false><<<grid, TV_THREADS, 0, ctx->stream>>>(g);
      else if (p.exponent == 4) tv_gather_kernel<4, false><<<grid, TV_THREADS, 0, ctx->stream>>>(g);
      else tv_gather_kernel<0, false><<<grid, TV_THREADS, 0, ctx->stream>>>(g);
    }
    VCK(cudaGetLastError());
    ctx->count_launch();
  }
  // the Scratch buffers are returned to the pool on scope exit; the pool is
  // stream-ordered, so the kernels above keep exclusive use until they finish.
  VCK(cudaStreamSynchronize(ctx->stream));
}

}  // namespace visfd_cuda
