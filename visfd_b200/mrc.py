"""ctypes plumbing over include/visfd_mrc.h (host-only MRC/REC file I/O with the file
semantics of the reference's lib/mrc_simple).  The same functions bind the reference's own
MrcSimple through oracle/_ref (`ref_mrc_read` / `ref_mrc_write`, test infrastructure) --
`MrcIO(lib, prefix)` below serves both."""
import ctypes as C

import numpy as np


class MrcHeader(C.Structure):
    """visfd_mrc_header (include/visfd_mrc.h); fields in file order."""
    _fields_ = [("nvoxels", C.c_int32 * 3), ("mode", C.c_int32), ("nstart", C.c_int32 * 3),
                ("mvoxels", C.c_int32 * 3), ("cellA", C.c_float * 3), ("cellB", C.c_float * 3),
                ("mapCRS", C.c_int32 * 3), ("dmin", C.c_float), ("dmax", C.c_float), ("dmean", C.c_float),
                ("ispg", C.c_int32), ("nsymbt", C.c_int32), ("extra_raw_data", C.c_char * 100),
                ("origin", C.c_float * 3), ("remaining_raw_data", C.c_char * 816),
                ("use_signed_bytes", C.c_int32)]

    def as_bytes(self):
        return bytes(memoryview(self))


class MrcError(RuntimeError):
    pass


class MrcIO:
    def __init__(self, lib, prefix="visfd_mrc_"):
        self.lib, self.px = lib, prefix
        self.err = getattr(lib, "visfd_mrc_last_error", None)
        if self.err is not None:
            self.err.restype = C.c_char_p

    def _ck(self, rc, what):
        if rc != 0:
            raise MrcError(self.err().decode() if self.err is not None else f"{what} failed ({rc})")

    def read_header(self, path):
        h = MrcHeader()
        self._ck(getattr(self.lib, self.px + "read_header")(str(path).encode(), C.byref(h)), "read_header")
        return h

    def read(self, path, capacity=None):
        """-> (header, float32 array [nz][ny][nx]) as MrcSimple::Read(path, rescale=False) leaves them"""
        if capacity is None:
            import os
            # >= the voxel count for every mode; a missing file is reported by the library
            capacity = max(1, os.path.getsize(path)) if os.path.exists(path) else 1
        buf = np.empty(capacity, np.float32)
        h = MrcHeader()
        self._ck(getattr(self.lib, self.px + "read")(str(path).encode(), C.byref(h),
                                                      buf.ctypes.data_as(C.c_void_p), C.c_int64(capacity)), "read")
        nx, ny, nz = h.nvoxels
        return h, buf[:nx * ny * nz].reshape(nz, ny, nx).copy()

    def write(self, path, header, voxels):
        """MrcSimple::Write: recomputes dmin/dmax/dmean into `header`, writes mode 2"""
        v = np.ascontiguousarray(voxels, np.float32)
        assert tuple(v.shape) == (header.nvoxels[2], header.nvoxels[1], header.nvoxels[0])
        self._ck(getattr(self.lib, self.px + "write")(str(path).encode(), C.byref(header),
                                                       v.ctypes.data_as(C.c_void_p)), "write")


def open_library():
    """MRC I/O of the product library (visfd_b200/libvisfd_cuda.so; needs no GPU)."""
    from .capi import load_library
    return MrcIO(load_library(), "visfd_mrc_")
