/* visfd_mrc.h -- MRC/REC file I/O for the membrane path's input and output volumes
 * (SURVEY.md 8f rank 2): the C ABI that replaces lib/mrc_simple (class MrcSimple,
 * mrc_simple.hpp:48-202; MrcHeader, mrc_header.hpp:20-150) for filter_mrc's reads and writes
 * (bin/filter_mrc/filter_mrc.cpp:96-101 tomo_in.Read, :786-791 tomo_out.Write).
 *
 * Same file semantics as the reference, which are its own (not the MRC2014 standard's):
 * 1024-byte header taken in native byte order with no MAP/machine-stamp check; voxel data
 * start at byte 1024 whatever `nsymbt` says; modes 0 (8-bit), 1 (int16), 2 (float32) and
 * 6 (uint16) are read, anything else is an error; the writer always emits mode 2 and
 * recomputes dmin/dmax/dmean.  What differs is how: the voxels are moved with bulk reads
 * and writes and indexed with 64 bits, where the reference issues one stream call per voxel
 * (13 Mvoxel/s) and multiplies the three extents in `int` (no volume above 2^31 voxels).
 * Host code only; nothing here needs a GPU.
 */
#ifndef VISFD_MRC_H
#define VISFD_MRC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* The fields of MrcHeader (mrc_header.hpp:22-102) in file order; 4-byte word index of each
 * in the comment. */
typedef struct visfd_mrc_header {
  int32_t nvoxels[3];            /*  0- 2  columns, rows, sections (x, y, z)                 */
  int32_t mode;                  /*  3                                                       */
  int32_t nstart[3];             /*  4- 6                                                    */
  int32_t mvoxels[3];            /*  7- 9  set equal to nvoxels by Read (mrc_simple.cpp:160) */
  float cellA[3];                /* 10-12  box size in Angstroms; voxel width = cellA/nvoxels */
  float cellB[3];                /* 13-15                                                    */
  int32_t mapCRS[3];             /* 16-18                                                    */
  float dmin, dmax, dmean;       /* 19-21                                                    */
  int32_t ispg, nsymbt;          /* 22-23                                                    */
  char extra_raw_data[100];      /* 24-48  copied verbatim                                   */
  float origin[3];               /* 49-51                                                    */
  char remaining_raw_data[816];  /* 52-255 copied verbatim                                   */
  int32_t use_signed_bytes;      /* not in the file: how mode-0 bytes are read               */
} visfd_mrc_header;

/* MrcHeader's defaults (mrc_header.hpp:108-141); the raw areas are zeroed. */
void visfd_mrc_header_init(visfd_mrc_header *h);

/* MrcSimple::Read(file name, rescale = false) up to the voxel data: the header as the
 * reference leaves it after reading -- mode-0 signedness from the IMOD stamp
 * (mrc_header.cpp:68-74) or unsigned for a name ending in ".rec" (mrc_simple.cpp:186-192),
 * mvoxels = nvoxels, and for a file that is not stored x-fastest (mapCRS != 1,2,3) the
 * extents, mvoxels, origin and cellA permuted into x,y,z order and mapCRS reset
 * (mrc_simple.cpp:111-152; origin and cellA pass through an int there and so do they here).
 * Returns 0, or non-zero with a message in visfd_mrc_last_error(). */
int visfd_mrc_read_header(const char *path, visfd_mrc_header *h);

/* The whole of MrcSimple::Read: header as above and the voxels converted to float32 in
 * [z][y][x] order (x fastest) into `voxels`, which must hold `capacity` >= nx*ny*nz floats. */
int visfd_mrc_read(const char *path, visfd_mrc_header *h, float *voxels, int64_t capacity);

/* MrcSimple::Write (mrc_simple.cpp:356-392): dmin/dmax/dmean recomputed over all voxels
 * (FindMinMaxMean, :396-426: running min/max in double, sum in double in raster order) and
 * stored back into *h, header written with mode 2, then the float32 voxels. */
int visfd_mrc_write(const char *path, visfd_mrc_header *h, const float *voxels);

const char *visfd_mrc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
