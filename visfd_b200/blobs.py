"""ctypes plumbing over include/visfd_blobs.h (host-side blob list post-processing).  The same
class binds the reference's own functions through oracle/_ref (`ref_blobs_*`, test
infrastructure).  A list is an (n, 5) float32 array of rows x, y, z, diameter, score (voxels)."""
import ctypes as C

import numpy as np

DO_NOT_SORT, SORT_DECREASING, SORT_INCREASING, SORT_DECREASING_MAGNITUDE, SORT_INCREASING_MAGNITUDE = range(5)
_f, _i, _i64 = C.c_float, C.c_int, C.c_int64


class BlobLists:
    def __init__(self, lib, prefix="visfd_blobs_"):
        self.lib, self.px = lib, prefix

    def _call(self, name, blobs, *args):
        b = np.ascontiguousarray(blobs, np.float32).reshape(-1, 5)
        crds = np.ascontiguousarray(b[:, :3])
        diam = np.ascontiguousarray(b[:, 3])
        score = np.ascontiguousarray(b[:, 4])
        fn = getattr(self.lib, self.px + name)
        fn.restype = _i64
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        n = fn(_i64(len(b)), p(crds), p(diam), p(score), *args)
        if n < 0:
            raise ValueError(f"{name}: bad arguments")
        return np.concatenate([crds[:n], diam[:n, None], score[:n, None]], axis=1)

    def sort(self, blobs, criteria, ascending=True):
        """SortBlobs (lib/visfd/feature.hpp:521-616)"""
        return self._call("sort", blobs, _i(criteria), _i(int(ascending)))

    def filter(self, blobs, score_lower=-np.inf, score_upper=np.inf, diameter_lower=-np.inf, diameter_upper=np.inf):
        """the score / diameter window of bin/filter_mrc/handlers.cpp:505-520"""
        return self._call("filter", blobs, _f(score_lower), _f(score_upper), _f(diameter_lower), _f(diameter_upper))

    def discard_masked(self, blobs, mask):
        """DiscardMaskedBlobs (lib/visfd/feature.hpp:926-969)"""
        m = np.ascontiguousarray(mask, np.float32)
        nz, ny, nx = m.shape
        return self._call("discard_masked", blobs, m.ctypes.data_as(C.c_void_p), _i64(nx), _i64(ny), _i64(nz))

    def discard_overlapping(self, blobs, min_radial_separation_ratio, max_volume_overlap_large=np.inf,
                            max_volume_overlap_small=np.inf, criteria=SORT_DECREASING_MAGNITUDE):
        """DiscardOverlappingBlobs (lib/visfd/feature.hpp:723-913)"""
        return self._call("discard_overlapping", blobs, _f(min_radial_separation_ratio), _f(max_volume_overlap_large),
                          _f(max_volume_overlap_small), _i(criteria))


def open_library():
    from .capi import load_library
    return BlobLists(load_library(), "visfd_blobs_")
