/* visfd_blobs.h -- post-processing of blob lists: what filter_mrc does
 * with the candidates BlobDog returns before it writes them out
 * (bin/filter_mrc/handlers.cpp:428-640 HandleBlobsNonmaxSuppression, :876-881 SortBlobs).
 * Lists are small (10^2..10^6 entries) and the algorithms sequential by definition (a
 * priority order decides who survives), so this is host code; it needs no GPU.
 * A list is three parallel arrays: crds[3*i..3*i+2] = x,y,z (voxels), diameter[i], score[i].
 * Every function works IN PLACE and returns the new length (>= 0) or -1 on bad arguments.
 */
#ifndef VISFD_BLOBS_H
#define VISFD_BLOBS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* SortCriteria, lib/visfd/visfd_utils.hpp:49-55 */
#define VISFD_DO_NOT_SORT 0
#define VISFD_SORT_DECREASING 1
#define VISFD_SORT_INCREASING 2
#define VISFD_SORT_DECREASING_MAGNITUDE 3
#define VISFD_SORT_INCREASING_MAGNITUDE 4

/* SortBlobs(crds, diameters, scores, criteria, ascending_order): lib/visfd/feature.hpp:521-616
 * (std::sort of (score or |score|, index) tuples, so ties keep index order). */
int64_t visfd_blobs_sort(int64_t n, float *crds, float *diameters, float *scores, int criteria,
                         int ascending_order);

/* The score / diameter window of HandleBlobsNonmaxSuppression (handlers.cpp:505-520): keep
 * blobs with lower <= value <= upper for both. */
int64_t visfd_blobs_filter(int64_t n, float *crds, float *diameters, float *scores,
                           float score_lower, float score_upper, float diameter_lower,
                           float diameter_upper);

/* DiscardMaskedBlobs: lib/visfd/feature.hpp:926-969 -- drop blobs whose centre, rounded to the
 * nearest voxel, has mask == 0.  mask: [nz][ny][nx]. */
int64_t visfd_blobs_discard_masked(int64_t n, float *crds, float *diameters, float *scores,
                                   const float *mask, int64_t nx, int64_t ny, int64_t nz);

/* DiscardOverlappingBlobs: lib/visfd/feature.hpp:723-913 -- greedy non-maximum suppression in
 * the order given by `criteria` (filter_mrc passes VISFD_SORT_DECREASING_MAGNITUDE): a blob is
 * dropped if an already accepted blob lies closer than (r_i + r_k) * min_radial_separation_ratio
 * or overlaps more than the given fractions of the smaller / larger sphere's volume
 * (INFINITY disables a criterion).  Candidates are found through the reference's coarse
 * occupancy table (cells of `scale` = 6 voxels), which is part of the result. */
int64_t visfd_blobs_discard_overlapping(int64_t n, float *crds, float *diameters, float *scores,
                                        float min_radial_separation_ratio,
                                        float max_volume_overlap_large,
                                        float max_volume_overlap_small, int criteria);

#ifdef __cplusplus
}
#endif
#endif
