// mrc_io.cpp -- bulk, 64-bit-indexed MRC/REC reader and writer with the file semantics of
// the reference's lib/mrc_simple (see include/visfd_mrc.h).  Host code only.
#include "../../include/visfd_mrc.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(const std::string &m) {
  g_err = m;
  return 1;
}

struct File {
  FILE *f = nullptr;
  File(const char *path, const char *mode) { f = std::fopen(path, mode); }
  ~File() {
    if (f) std::fclose(f);
  }
};

constexpr size_t HEADER_BYTES = 1024;
constexpr int32_t IMOD_STAMP = 1146047817;

// mrc_header.cpp:12-149
void parse_header(const unsigned char *raw, visfd_mrc_header *h) {
  auto i32 = [&](int word) { int32_t v; std::memcpy(&v, raw + 4 * word, 4); return v; };
  auto f32 = [&](int word) { float v; std::memcpy(&v, raw + 4 * word, 4); return v; };
  for (int d = 0; d < 3; d++) h->nvoxels[d] = i32(d);
  h->mode = i32(3);
  if (h->mode == 0 && i32(38) == IMOD_STAMP) h->use_signed_bytes = i32(39) & 1;
  for (int d = 0; d < 3; d++) h->nstart[d] = i32(4 + d);
  for (int d = 0; d < 3; d++) h->mvoxels[d] = i32(7 + d);
  for (int d = 0; d < 3; d++) h->cellA[d] = f32(10 + d);
  for (int d = 0; d < 3; d++) h->cellB[d] = f32(13 + d);
  for (int d = 0; d < 3; d++) h->mapCRS[d] = i32(16 + d);
  h->dmin = f32(19);
  h->dmax = f32(20);
  h->dmean = f32(21);
  h->ispg = i32(22);
  h->nsymbt = i32(23);
  std::memcpy(h->extra_raw_data, raw + 4 * 24, sizeof h->extra_raw_data);
  for (int d = 0; d < 3; d++) h->origin[d] = f32(49 + d);
  std::memcpy(h->remaining_raw_data, raw + 4 * 52, sizeof h->remaining_raw_data);
}

// mrc_header.cpp:163-232
void format_header(const visfd_mrc_header *h, int32_t mode, unsigned char *raw) {
  auto put = [&](int word, const void *v) { std::memcpy(raw + 4 * word, v, 4); };
  for (int d = 0; d < 3; d++) put(d, &h->nvoxels[d]);
  put(3, &mode);
  for (int d = 0; d < 3; d++) put(4 + d, &h->nstart[d]);
  for (int d = 0; d < 3; d++) put(7 + d, &h->mvoxels[d]);
  for (int d = 0; d < 3; d++) put(10 + d, &h->cellA[d]);
  for (int d = 0; d < 3; d++) put(13 + d, &h->cellB[d]);
  for (int d = 0; d < 3; d++) put(16 + d, &h->mapCRS[d]);
  put(19, &h->dmin);
  put(20, &h->dmax);
  put(21, &h->dmean);
  put(22, &h->ispg);
  put(23, &h->nsymbt);
  std::memcpy(raw + 4 * 24, h->extra_raw_data, sizeof h->extra_raw_data);
  for (int d = 0; d < 3; d++) put(49 + d, &h->origin[d]);
  std::memcpy(raw + 4 * 52, h->remaining_raw_data, sizeof h->remaining_raw_data);
}

// PermuteCArray (mrc_simple.cpp:72-84): target[i] = old[p[i]], through an `int` temporary
// whatever the element type is -- float entries are truncated on the way, as there.
template <typename T>
void permute(T *a, const int32_t p[3]) {
  int32_t tmp[3];
  for (int i = 0; i < 3; i++) tmp[i] = (int32_t)a[i];
  for (int i = 0; i < 3; i++) a[i] = (T)tmp[p[i]];
}

bool ends_with_rec(const char *path) {
  const size_t n = std::strlen(path);
  return n > 4 && std::strcmp(path + n - 4, ".rec") == 0;
}

size_t mode_bytes(int32_t mode) {
  switch (mode) {
    case 0: return 1;
    case 1: case 6: return 2;
    case 2: return 4;
    default: return 0;
  }
}

// header part of MrcSimple::Read; axis_order[] is filled when the file is not x-fastest
int read_header(FILE *f, const char *path, visfd_mrc_header *h, bool *permuted, int32_t axis_order[3]) {
  unsigned char raw[HEADER_BYTES];
  if (std::fread(raw, 1, HEADER_BYTES, f) != HEADER_BYTES) return fail(std::string("Error: \"") + path + "\" is shorter than an MRC header.\n");
  visfd_mrc_header_init(h);
  if (ends_with_rec(path)) h->use_signed_bytes = 0;  // mrc_simple.cpp:186-192 (the header may still override it)
  parse_header(raw, h);
  *permuted = !(h->mapCRS[0] == 1 && h->mapCRS[1] == 2 && h->mapCRS[2] == 3);
  if (*permuted) {
    for (int d = 0; d < 3; d++) {
      axis_order[d] = h->mapCRS[d] - 1;
      if (axis_order[d] < 0 || axis_order[d] > 2) return fail("Error: invalid axis order (mapCRS) in the MRC header.\n");
    }
    if (axis_order[0] == axis_order[1] || axis_order[0] == axis_order[2] || axis_order[1] == axis_order[2])
      return fail("Error: invalid axis order (mapCRS) in the MRC header.\n");
    h->mapCRS[0] = 1; h->mapCRS[1] = 2; h->mapCRS[2] = 3;
    permute(h->nvoxels, axis_order);
    permute(h->mvoxels, axis_order);
    permute(h->origin, axis_order);
    permute(h->cellA, axis_order);
  }
  for (int d = 0; d < 3; d++) h->mvoxels[d] = h->nvoxels[d];   // mrc_simple.cpp:158-160
  for (int d = 0; d < 3; d++)
    if (h->nvoxels[d] < 0) return fail("Error: negative image size in the MRC header.\n");
  return 0;
}

template <typename T>
void convert(const unsigned char *in, float *out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    T v;
    std::memcpy(&v, in + i * sizeof(T), sizeof(T));
    out[i] = static_cast<float>(v);
  }
}

}  // namespace

extern "C" {

const char *visfd_mrc_last_error(void) { return g_err.c_str(); }

void visfd_mrc_header_init(visfd_mrc_header *h) {
  std::memset(h, 0, sizeof *h);
  h->mapCRS[0] = 1; h->mapCRS[1] = 2; h->mapCRS[2] = 3;
  h->dmin = 0.0f;
  h->dmax = -1.0f;
  h->mode = -1;
  h->cellB[0] = h->cellB[1] = h->cellB[2] = 90.0f;
  h->use_signed_bytes = 1;
}

int visfd_mrc_read_header(const char *path, visfd_mrc_header *h) {
  if (!path || !h) return fail("NULL argument");
  File file(path, "rb");
  if (!file.f) return fail(std::string("Error: Unable to open \"") + path + "\" for reading.\n");
  bool permuted;
  int32_t order[3];
  return read_header(file.f, path, h, &permuted, order);
}

int visfd_mrc_read(const char *path, visfd_mrc_header *h, float *voxels, int64_t capacity) {
  if (!path || !h || !voxels) return fail("NULL argument");
  File file(path, "rb");
  if (!file.f) return fail(std::string("Error: Unable to open \"") + path + "\" for reading.\n");
  bool permuted;
  int32_t order[3] = {0, 1, 2};
  if (int rc = read_header(file.f, path, h, &permuted, order)) return rc;
  const size_t bpv = mode_bytes(h->mode);
  if (bpv == 0) return fail("UNSUPPORTED MODE in MRC file (unsupported MRC format)");
  const int64_t nx = h->nvoxels[0], ny = h->nvoxels[1], nz = h->nvoxels[2];
  const int64_t n = nx * ny * nz;
  if (n > capacity) return fail("the voxel buffer is smaller than the image in the file");
  // file extents (fastest first): nvoxels[inv[k]], where inv is the inverse of axis_order
  // (mrc_simple.cpp:214-226)
  int32_t inv[3] = {0, 1, 2};
  if (permuted)
    for (int i = 0; i < 3; i++) inv[order[i]] = i;
  const int64_t NX = h->nvoxels[inv[0]], NY = h->nvoxels[inv[1]];
  const int64_t stride[3] = {1, nx, nx * ny};   // of x, y, z in `voxels`
  // one file row (NX entries) at a time through a conversion buffer of ~4 MB of rows
  const int64_t rows_total = (NX > 0) ? n / NX : 0;
  const int64_t rows_per_chunk = std::max<int64_t>(1, (int64_t)(4 << 20) / std::max<int64_t>(1, NX * (int64_t)bpv));
  std::vector<unsigned char> raw((size_t)(rows_per_chunk * NX) * bpv);
  std::vector<float> rowf(permuted ? (size_t)NX : 0);
  // The reference zero-fills the image before reading (Alloc, mrc_simple.cpp:55-66).  That
  // only shows for a cyclic axis order (mapCRS 2,3,1 / 3,1,2), where its index mapping
  // (mrc_simple.cpp:235-245) is not one-to-one: some voxels are never written, others are
  // written twice or fall outside the array (undefined behaviour there, skipped here).
  if (permuted) std::memset(voxels, 0, (size_t)n * sizeof(float));
  for (int64_t r0 = 0; r0 < rows_total; r0 += rows_per_chunk) {
    const int64_t nr = std::min(rows_per_chunk, rows_total - r0);
    const size_t want = (size_t)(nr * NX) * bpv;
    // the reference ignores short reads (the stream just fails and entries keep their
    // garbage); a truncated file is an error here
    if (std::fread(raw.data(), 1, want, file.f) != want) return fail(std::string("Error: \"") + path + "\" ends before the last voxel.\n");
    for (int64_t r = 0; r < nr; r++) {
      const unsigned char *in = raw.data() + (size_t)(r * NX) * bpv;
      const int64_t iY = (r0 + r) % NY, iZ = (r0 + r) / NY;
      float *out = permuted ? rowf.data() : voxels + (r0 + r) * NX;
      switch (h->mode) {
        case 0:
          if (h->use_signed_bytes) convert<int8_t>(in, out, (size_t)NX); else convert<uint8_t>(in, out, (size_t)NX);
          break;
        case 1: convert<int16_t>(in, out, (size_t)NX); break;
        case 6: convert<uint16_t>(in, out, (size_t)NX); break;
        default: std::memcpy(out, in, (size_t)NX * 4); break;
      }
      if (permuted) {
        // file index (iX, iY, iZ) -> image index ixyz[inv[.]] (mrc_simple.cpp:235-245)
        const int64_t fi[3] = {0, iY, iZ};
        // image coordinate c = x,y,z takes file index fi[inv[c]]; the one with inv[c] == 0 runs with iX
        int64_t base = 0, step = 0;
        for (int c = 0; c < 3; c++) {
          if (inv[c] == 0) step = stride[c]; else base += fi[inv[c]] * stride[c];
        }
        for (int64_t iX = 0; iX < NX; iX++) {
          const int64_t at = base + iX * step;
          if (at < n) voxels[at] = rowf[(size_t)iX];
        }
      }
    }
  }
  return 0;
}

int visfd_mrc_write(const char *path, visfd_mrc_header *h, const float *voxels) {
  if (!path || !h || !voxels) return fail("NULL argument");
  const int64_t n = (int64_t)h->nvoxels[0] * h->nvoxels[1] * h->nvoxels[2];
  // FindMinMaxMean (mrc_simple.cpp:396-426)
  double total = 0.0, dmin = 0.0, dmax = -1.0;
  for (int64_t i = 0; i < n; i++) {
    const float v = voxels[i];
    total += v;
    if (dmin > dmax) {
      dmin = v;
      dmax = v;
    } else {
      if (v > dmax) dmax = v;
      if (v < dmin) dmin = v;
    }
  }
  h->dmin = (float)dmin;
  h->dmax = (float)dmax;
  h->dmean = (float)(total / (double)(long long)n);
  File file(path, "wb");
  if (!file.f) return fail(std::string("Error: Unable to open \"") + path + "\" for writing.\n");
  unsigned char raw[HEADER_BYTES];
  format_header(h, 2, raw);
  if (std::fwrite(raw, 1, HEADER_BYTES, file.f) != HEADER_BYTES) return fail("Error: write failed.\n");
  const size_t chunk = (size_t)64 << 20;
  for (size_t off = 0; off < (size_t)n * 4; off += chunk) {
    const size_t len = std::min(chunk, (size_t)n * 4 - off);
    if (std::fwrite(reinterpret_cast<const unsigned char *>(voxels) + off, 1, len, file.f) != len) return fail("Error: write failed.\n");
  }
  return 0;
}

}  // extern "C"
