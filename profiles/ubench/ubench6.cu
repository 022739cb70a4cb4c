// vote loop variants, round 2: dispatch-slot diet.
//   V0 : tv.cu table kernel as is   (FMUL2 on {t0.x,t1.x} pairs => MOVs; IADD3 for the table address)
//   V1 : qn and w as scalar FMULs straight from the LDS.64 results (no MOVs)
//   V2 : V1 + table address = the bit pattern of a DENORMAL r2 accumulator (no IADD3)
//   V3 : V2 + scalar cross terms
//   V4 : V0 + denormal address
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
struct __align__(16) VoterRec { float4 a, b, c; };
__device__ __forceinline__ float2 bc(float x) { return make_float2(x, x); }
__device__ __forceinline__ float sfma(float a, float b, float c) { float d; asm("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float smul(float a, float b) { float d; asm("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
template <bool SCALAR_CROSS>
__device__ __forceinline__ void accum(float2 wx, float2 wy, float2 wz, float2 T[6]) {
  T[0] = __ffma2_rn(wx, wx, T[0]);
  if (SCALAR_CROSS) {
    T[3].x = sfma(wx.x, wy.x, T[3].x); T[3].y = sfma(wx.y, wy.y, T[3].y);
    T[5].x = sfma(wx.x, wz.x, T[5].x); T[5].y = sfma(wx.y, wz.y, T[5].y);
  } else {
    T[3] = __ffma2_rn(wx, wy, T[3]); T[5] = __ffma2_rn(wx, wz, T[5]);
  }
  T[1] = __ffma2_rn(wy, wy, T[1]);
  if (SCALAR_CROSS) { T[4].x = sfma(wy.x, wz.x, T[4].x); T[4].y = sfma(wy.y, wz.y, T[4].y); }
  else T[4] = __ffma2_rn(wy, wz, T[4]);
  T[2] = __ffma2_rn(wz, wz, T[2]);
}
constexpr int NV = 64;

// V: bit0 scalar qn/w, bit1 denormal address, bit2 scalar cross
template <int V>
__device__ __forceinline__ void pair_lut(float rx, float ry, float rxy2m, float dxy, float2 fz, const float4 &ea, const float4 &eb,
                                         const float4 &ec, unsigned tab, float2 T[6]) {
  const float2 rz = __fadd2_rn(fz, bc(ea.z));
  const float2 r2m = __ffma2_rn(rz, rz, bc(rxy2m));
  const float2 d = __ffma2_rn(rz, bc(eb.z), bc(dxy));
  float2 t0, t1;
  if (V & 8) {   // no table: fake values derived without memory
    t0 = make_float2(1e-11f, -1e30f); t1 = make_float2(2e-11f, -2e30f);
    if (__float_as_uint(r2m.x) == 0x7fffffffu) t0.x = r2m.y;
  } else if (V & 2) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(t0.x), "=f"(t0.y) : "r"(__float_as_uint(r2m.x)));
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(t1.x), "=f"(t1.y) : "r"(__float_as_uint(r2m.y)));
  } else {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(t0.x), "=f"(t0.y) : "r"(__float_as_uint(r2m.x) + tab));
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(t1.x), "=f"(t1.y) : "r"(__float_as_uint(r2m.y) + tab));
  }
  float2 qn, w;
  if (V & 1) { qn.x = smul(d.x, t0.y); qn.y = smul(d.y, t1.y); }
  else qn = __fmul2_rn(d, make_float2(t0.y, t1.y));
  const float2 ang2 = __ffma2_rn(qn, d, bc(ea.w));
  if (V & 1) { w.x = smul(t0.x, ang2.x); w.y = smul(t1.x, ang2.y); }
  else w = __fmul2_rn(make_float2(t0.x, t1.x), ang2);
  const float2 hx = __ffma2_rn(qn, bc(rx), bc(ec.x));
  const float2 hy = __ffma2_rn(qn, bc(ry), bc(ec.y));
  const float2 hz = __ffma2_rn(qn, rz, bc(ec.z));
  const float2 wx = __fmul2_rn(w, hx), wy = __fmul2_rn(w, hy), wz = __fmul2_rn(w, hz);
  if (V & 32) {
    T[0] = __ffma2_rn(wx, wx, T[0]); T[3] = __ffma2_rn(wy, wy, T[3]); T[5] = __ffma2_rn(wz, wz, T[5]);
    T[1] = __ffma2_rn(wx, wx, T[1]); T[4] = __ffma2_rn(wy, wy, T[4]); T[2] = __ffma2_rn(wz, wz, T[2]);
  } else
  accum<(V & 4) != 0>(wx, wy, wz, T);
}
template <int V>
__device__ __forceinline__ void vote_lut(const VoterRec *q, float fx, float fy, float2 fz01, float2 fz23, float base, unsigned tab, float2 T[12]) {
  float4 ea, eb, ec;
  if (V & 16) {   // same record every time, perturbed in registers so that nothing is hoisted
    ea = make_float4(fx * 3.0f, fy * 2.0f, fz01.y * 5.0f, 1e10f); eb = make_float4(fx, fy, 1e12f, 0.f); ec = make_float4(0.1f, 0.2f, fx, 0.f);
    asm volatile("" : "+f"(ea.x), "+f"(ea.y), "+f"(ea.z), "+f"(eb.x));
  } else { ea = q->a; eb = q->b; ec = q->c; }
  const float rx = fx + ea.x, ry = fy + ea.y;
  const float rxy2m = fmaf(ry, ry, fmaf(rx, rx, base));
  const float dxy = fmaf(ry, eb.y, rx * eb.x);
  pair_lut<V>(rx, ry, rxy2m, dxy, fz01, ea, eb, ec, tab, T);
  pair_lut<V>(rx, ry, rxy2m, dxy, fz23, ea, eb, ec, tab, T + 6);
}

struct Args {
  const VoterRec *rec_n, *rec_d;   // normal-scale / denormal-address records  [1024][NV]
  const float2 *tab_n, *tab_d;     // {E, -1/r2} / {E', -tau/r2}
  int n_r2, reps;
  float *out;
};
constexpr float RHO = 4.235164736271502e-22f;   // 2^-71

template <int V, int UNROLL, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k(Args g) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // table first (address baked into the denormal accumulator), rings after
  unsigned char *tabp = smem;
  VoterRec *ring = reinterpret_cast<VoterRec *>(smem + (size_t)g.n_r2 * 128) + warp * NV;
  const int gw = blockIdx.x * WARPS + warp;
  const VoterRec *src = ((V & 2) ? g.rec_d : g.rec_n) + (size_t)(gw % 1024) * NV;
  for (int i = lane; i < NV; i += 32) ring[i] = src[i];
  const float2 *tsrc = (V & 2) ? g.tab_d : g.tab_n;
  for (int i = threadIdx.x; i < g.n_r2 * 16; i += 32 * WARPS) reinterpret_cast<float2 *>(tabp)[i] = tsrc[i / 16];
  __syncthreads();
  const unsigned tab_lane = (unsigned)__cvta_generic_to_shared(tabp) + 8u * (lane & 15);
  const float magic = 65536.0f;
  const unsigned tab = tab_lane - __float_as_uint(magic);
  const float sc = (V & 2) ? RHO : 1.0f;
  const float base = (V & 2) ? __uint_as_float(tab_lane) : magic;
  const float fx = (lane & 3) * sc, fy = ((lane >> 2) & 3) * sc;
  const int half = lane >> 4;
  const float2 fz01 = make_float2(0 * sc, 1 * sc), fz23 = make_float2(2 * sc, 3 * sc);
  float2 T[12];
  for (int i = 0; i < 12; i++) T[i] = make_float2(0.f, 0.f);
  for (int it = 0; it < g.reps; it++) {
    const VoterRec *q = ring + half, *end = q + NV;
#pragma unroll 1
    for (; q < end; q += 2 * UNROLL) {
#pragma unroll
      for (int u = 0; u < UNROLL; u++) vote_lut<V>(q + 2 * u, fx, fy, fz01, fz23, base, tab, T);
    }
  }
  for (int kk = 0; kk < 12; kk++) {
    T[kk].x += __shfl_xor_sync(0xffffffffu, T[kk].x, 16);
    T[kk].y += __shfl_xor_sync(0xffffffffu, T[kk].y, 16);
  }
  if (half == 0 && gw < 1024) {
    const int col = lane & 15;
    for (int z = 0; z < 4; z++)
      for (int c = 0; c < 6; c++) {
        const float2 v = T[(z >> 1) * 6 + c];
        g.out[((size_t)gw * 64 + z * 16 + col) * 6 + c] = (z & 1) ? v.y : v.x;
      }
  }
}

static Args base_args;
static float *d_out;
static std::vector<float> ref_out;

template <int V, int UNROLL, int WARPS>
void run(const char *name, int n_r2) {
  Args g = base_args;
  g.n_r2 = n_r2; g.out = d_out;
  const size_t smem = WARPS * NV * sizeof(VoterRec) + (size_t)n_r2 * 128;
  cudaFuncSetAttribute(k<V, UNROLL, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = 148;
  cudaMemset(d_out, 0, 1024 * 64 * 6 * 4);
  g.reps = 4;
  k<V, UNROLL, WARPS><<<grid, 32 * WARPS, smem>>>(g);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  g.reps = 200;
  cudaEventRecord(e0);
  k<V, UNROLL, WARPS><<<grid, 32 * WARPS, smem>>>(g);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  g.reps = 1;
  k<V, UNROLL, WARPS><<<grid, 32 * WARPS, smem>>>(g);
  cudaDeviceSynchronize();
  std::vector<float> h(1024 * 64 * 6);
  const int nw = std::min(1024, grid * WARPS);
  cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxv = 0;
  if (ref_out.empty()) ref_out = h;
  for (size_t i = 0; i < (size_t)nw * 64 * 6; i += 6) {
    double fro = 0, err = 0;
    for (int c = 0; c < 6; c++) { fro += (double)ref_out[i + c] * ref_out[i + c]; err += ((double)h[i + c] - ref_out[i + c]) * ((double)h[i + c] - ref_out[i + c]); }
    maxerr = std::max(maxerr, sqrt(err) / (sqrt(fro) + 1e-30)); maxv = std::max(maxv, sqrt(fro));
  }
  const double evals64_per_smsp = (double)WARPS / 4 * 200 * (NV / 2) * 2;
  const double cyc = ms * 1e-3 * 1.965e9;
  printf("%-40s warps/SM %2d unroll %d smem %6zu: %7.3f ms %6.1f cyc per 64 evals   rel.err vs first %.2e (max|T| %.3g) %s\n", name,
         WARPS, UNROLL, smem, ms, cyc / evals64_per_smsp, maxerr, maxv, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  const float sigma = 14.1986f;
  const int total_w = 1024;
  std::vector<VoterRec> rn((size_t)total_w * NV), rd((size_t)total_w * NV);
  srand(1);
  auto rnd = []() { return rand() / (float)RAND_MAX; };
  const float BETA = ldexpf(1.0f, 40), BETA2 = ldexpf(1.0f, -3), KK = ldexpf(1.0f, 38), TAU = ldexpf(1.0f, 100);
  for (auto i = 0u; i < rn.size(); i++) {
    float x, y, z;
    do { x = floorf((rnd() * 2 - 1) * 23); y = floorf((rnd() * 2 - 1) * 23); z = floorf((rnd() * 2 - 1) * 23); }
    while (x * x + y * y + z * z > 22.5f * 22.5f);
    x += 1; y += 2; z += 1;
    float nx = rnd() - 0.5f, ny = rnd() - 0.5f, nz = rnd() - 0.5f, nn = sqrtf(nx * nx + ny * ny + nz * nz);
    nx /= nn; ny /= nn; nz /= nn;
    const float w = (0.2f + rnd()) * 1e-3f, w4 = 4 * w;
    const float lam = (float)pow((double)w4, 1.0 / 6.0);
    rn[i].a = make_float4(-x, -y, -z, lam * lam); rn[i].b = make_float4(lam * nx, lam * ny, lam * nz, 0); rn[i].c = make_float4(lam * nx / 2, lam * ny / 2, lam * nz / 2, 0);
    rd[i].a = make_float4(-x * RHO, -y * RHO, -z * RHO, lam * lam * KK);
    rd[i].b = make_float4(lam * nx * BETA, lam * ny * BETA, lam * nz * BETA, 0);
    rd[i].c = make_float4(lam * nx * BETA2, lam * ny * BETA2, lam * nz * BETA2, 0);
  }
  VoterRec *d_rn, *d_rd;
  cudaMalloc(&d_rn, rn.size() * sizeof(VoterRec)); cudaMalloc(&d_rd, rd.size() * sizeof(VoterRec));
  cudaMemcpy(d_rn, rn.data(), rn.size() * sizeof(VoterRec), cudaMemcpyHostToDevice);
  cudaMemcpy(d_rd, rd.data(), rd.size() * sizeof(VoterRec), cudaMemcpyHostToDevice);
  const int NT = 800;
  std::vector<float2> tn(NT), td(NT);
  for (int r2 = 0; r2 < NT; r2++) {
    const float E = (r2 <= 400) ? (float)exp(-0.5 * r2 / ((double)sigma * sigma)) : 0.0f;
    tn[r2] = make_float2(E, r2 ? -1.0f / r2 : 0.0f);
    td[r2] = make_float2(E * ldexpf(1.0f, -36), r2 ? -TAU / r2 : 0.0f);
  }
  float2 *d_tn, *d_td; cudaMalloc(&d_tn, NT * sizeof(float2)); cudaMalloc(&d_td, NT * sizeof(float2));
  cudaMemcpy(d_tn, tn.data(), NT * sizeof(float2), cudaMemcpyHostToDevice);
  cudaMemcpy(d_td, td.data(), NT * sizeof(float2), cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, 1024 * 64 * 6 * 4);
  base_args.rec_n = d_rn; base_args.rec_d = d_rd; base_args.tab_n = d_tn; base_args.tab_d = d_td;
  if (only < 0 || only == 0) run<0, 4, 16>("V0 tv.cu table loop", NT);
  if (only < 0 || only == 1) run<1, 4, 16>("V1 scalar qn/w", NT);
  if (only < 0 || only == 2) run<2, 4, 16>("V4 denormal address, packed qn/w", NT);
  if (only < 0 || only == 3) run<3, 4, 16>("V2 scalar qn/w + denormal address", NT);
  if (only < 0 || only == 4) run<7, 4, 16>("V3 V2 + scalar cross", NT);
  if (only < 0 || only == 5) run<5, 4, 16>("V1 + scalar cross", NT);
  if (only < 0 || only == 6) run<3, 2, 16>("V2 unroll 2", NT);
  if (only < 0 || only == 7) run<3, 8, 16>("V2 unroll 8", NT);
  if (only < 0 || only == 8) run<3, 4, 12>("V2 12 warps", NT);
  if (only < 0 || only == 9) run<3, 4, 20>("V2 20 warps", NT);
  if (only < 0 || only == 10) run<7, 4, 20>("V3 20 warps", NT);
  if (only < 0 || only == 11) run<3 + 8, 4, 16>("V2 no table LDS", NT);
  if (only < 0 || only == 12) run<3 + 16, 4, 16>("V2 no voter LDS", NT);
  if (only < 0 || only == 13) run<3 + 24, 4, 16>("V2 no LDS at all", NT);
  if (only < 0 || only == 14) run<3 + 32, 4, 16>("V2 squares only", NT);
  if (only < 0 || only == 15) run<3 + 56, 4, 16>("V2 no LDS, squares only", NT);
  if (only < 0 || only == 16) run<3 + 32, 8, 16>("V2 squares only unroll 8", NT);
  if (only < 0 || only == 17) run<3, 8, 12>("V2 unroll 8 12 warps", NT);
  if (only < 0 || only == 18) run<3, 8, 8>("V2 unroll 8 8 warps", NT);
  return 0;
}
