"""ctypes bindings for the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

  * ``Oracle("port")``      -> oracle/libvisfd_oracle.so  (our C-style restatement,
                               oracle/visfd_oracle.cpp, symbols ``vo_*``)
  * ``Oracle("reference")`` -> oracle/_ref/libvisfd_ref.so (the unmodified reference
                               headers behind oracle/ref_shim.cpp, symbols ``ref_*``)

Both expose the same numpy-level methods so tests can run either against the
CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` arm may import this module; the product never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libvisfd_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libvisfd_ref.so")

_f = C.c_float
_i = C.c_int
_i64 = C.c_int64
_p = C.c_void_p


def build(kind="port"):
    """(Re)build a checker library with oracle/Makefile."""
    target = "oracle" if kind == "port" else "ref"
    subprocess.check_call(["make", "-s", "-C", HERE, target])


def have(kind):
    return os.path.exists(PORT_SO if kind == "port" else REF_SO)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)



REGION_DTYPE = np.dtype([("type", np.int32), ("p", np.float32, 6), ("value", np.float32)])


def pack_regions(regions):
    """List of ("rect", xmin, xmax, ymin, ymax, zmin, zmax, value) / ("sphere", x0, y0, z0, r, value)
    -> the 32-byte records both libraries take (mirrors visfd::SimpleRegion<float>, draw.hpp:41-87)."""
    rec = np.zeros(len(regions), REGION_DTYPE)
    for i, r in enumerate(regions):
        if r[0] == "sphere":
            rec[i] = (1, tuple(r[1:5]) + (0.0, 0.0), r[5])
        elif r[0] == "rect":
            rec[i] = (0, tuple(r[1:7]), r[7])
        else:
            raise ValueError("unknown region kind %r" % (r[0],))
    return rec

class Oracle:
    def __init__(self, kind="port"):
        assert kind in ("port", "reference")
        self.kind = kind
        path = PORT_SO if kind == "port" else REF_SO
        if not os.path.exists(path):
            build(kind)
        self.lib = C.CDLL(path)
        self.px = "vo_" if kind == "port" else "ref_"
        # the reference shim uses int sizes (like the reference), the port int64
        self.size_t = _i64 if kind == "port" else _i

    def _fn(self, name, restype=None):
        f = getattr(self.lib, self.px + name)
        f.restype = restype
        return f

    def _dims(self, shape):
        nz, ny, nx = shape
        t = self.size_t
        return t(nx), t(ny), t(nz)

    # ---- taps ------------------------------------------------------------
    def gen_gauss1d(self, sigma, hw):
        t = np.zeros(2 * hw + 1, np.float32)
        self._fn("gen_gauss1d")(_f(sigma), _i(hw), _ptr(t))
        return t

    # ---- Gaussian family ---------------------------------------------------
    def apply_gauss(self, src, sigma, hw, mask=None, normalize=True):
        src = _f32(src)
        mask = _f32(mask)
        dst = np.empty_like(src)
        sg = (_f * 3)(*([sigma] * 3 if np.isscalar(sigma) else sigma))
        h = (_i * 3)(*([hw] * 3 if np.isscalar(hw) else hw))
        A = self._fn("apply_gauss", _f)(*self._dims(src.shape), _ptr(src), _ptr(dst),
                                         _ptr(mask), sg, h, _i(int(normalize)))
        return dst, A

    def apply_dog(self, src, sigma_a, sigma_b, hw, mask=None):
        src = _f32(src)
        mask = _f32(mask)
        dst = np.empty_like(src)
        sa = (_f * 3)(*([sigma_a] * 3 if np.isscalar(sigma_a) else sigma_a))
        sb = (_f * 3)(*([sigma_b] * 3 if np.isscalar(sigma_b) else sigma_b))
        h = (_i * 3)(*([hw] * 3 if np.isscalar(hw) else hw))
        A, B = _f(), _f()
        self._fn("apply_dog")(*self._dims(src.shape), _ptr(src), _ptr(dst), _ptr(mask),
                              sa, sb, h, C.byref(A), C.byref(B))
        return dst, A.value, B.value

    def apply_log(self, src, sigma, delta=0.02, truncate_ratio=2.5, mask=None):
        src = _f32(src)
        mask = _f32(mask)
        dst = np.empty_like(src)
        sg = (_f * 3)(*([sigma] * 3 if np.isscalar(sigma) else sigma))
        A, B = _f(), _f()
        self._fn("apply_log")(*self._dims(src.shape), _ptr(src), _ptr(dst), _ptr(mask),
                              sg, _f(delta), _f(truncate_ratio), C.byref(A), C.byref(B))
        return dst, A.value, B.value

    # ---- Hessian / eigen -----------------------------------------------------
    def calc_hessian(self, src, sigma, truncate_ratio, mask=None):
        src = _f32(src)
        mask = _f32(mask)
        N = src.size
        grad = np.zeros((N, 3), np.float32)
        hess = np.zeros((N, 6), np.float32)
        if self.kind == "port":
            rc = self._fn("calc_hessian", _i)(*self._dims(src.shape), _ptr(src), _ptr(mask),
                                              _f(sigma), _f(truncate_ratio), _ptr(grad),
                                              _ptr(hess), None)
        else:
            rc = self._fn("calc_hessian", _i)(*self._dims(src.shape), _ptr(src), _ptr(mask),
                                              _f(sigma), _f(truncate_ratio), _ptr(grad),
                                              _ptr(hess))
        if rc != 0:
            raise ValueError("CalcHessian requires an image at least 3 voxels wide")
        return grad.reshape(src.shape + (3,)), hess.reshape(src.shape + (6,))

    def hessian_eigen_score(self, hess, order=1, score_kind=0, mask=None):
        hess = _f32(hess)
        shape = hess.shape[:-1]
        N = int(np.prod(shape))
        mask = _f32(mask)
        sal = np.zeros(N, np.float32)
        dire = np.zeros((N, 3), np.float32)
        ev = np.zeros((N, 3), np.float32)
        self._fn("hessian_eigen_score")(_i64(N), _ptr(hess), _ptr(mask), _i(order),
                                        _i(score_kind), _ptr(sal), _ptr(dire), _ptr(ev))
        return sal.reshape(shape), dire.reshape(shape + (3,)), ev.reshape(shape + (3,))

    def saliency_cut(self, sal, cut, is_fraction, mask=None):
        out = np.array(sal, dtype=np.float32, copy=True, order="C")
        mask = _f32(mask)
        thr = self._fn("saliency_cut", _f)(_i64(out.size), _ptr(out), _ptr(mask), _f(cut),
                                           _i(int(is_fraction)))
        return out, thr

    # ---- tensor voting ---------------------------------------------------------
    def tv_tables(self, sigma, cutoff_ratio):
        if self.kind == "port":
            hw = self._fn("tv_halfwidth", _i)(_f(sigma), _f(cutoff_ratio))
            w = 2 * hw + 1
            decay = np.zeros((w, w, w), np.float32)
            disp = np.zeros((w, w, w, 3), np.float32)
            self._fn("tv_tables")(_f(sigma), _i(hw), _ptr(decay), _ptr(disp))
        else:
            hw = self._fn("tv_tables", _i)(_f(sigma), _f(cutoff_ratio), None, None, _i64(0))
            w = 2 * hw + 1
            decay = np.zeros((w, w, w), np.float32)
            disp = np.zeros((w, w, w, 3), np.float32)
            self._fn("tv_tables", _i)(_f(sigma), _f(cutoff_ratio), _ptr(decay), _ptr(disp),
                                      _i64(decay.size))
        return hw, decay, disp

    def tv_dense_stick(self, sal, direction, sigma, exponent, cutoff_ratio, mask_src=None,
                       mask_dst=None, curves=False):
        sal = _f32(sal)
        direction = _f32(direction)
        mask_src = _f32(mask_src)
        mask_dst = _f32(mask_dst)
        tensor = np.zeros(sal.shape + (6,), np.float32)
        if self.kind == "port":
            self._fn("tv_dense_stick")(*self._dims(sal.shape), _ptr(sal), _ptr(direction),
                                       _ptr(mask_src), _ptr(mask_dst), _f(sigma),
                                       _i(exponent), _f(cutoff_ratio), _i(int(curves)),
                                       _ptr(tensor))
        else:
            self._fn("tv_dense_stick")(*self._dims(sal.shape), _ptr(sal), _ptr(direction),
                                       _ptr(mask_src), _ptr(mask_dst), _f(sigma),
                                       _i(exponent), _f(cutoff_ratio), _i(int(curves)), _i(0),
                                       _ptr(tensor))
        return tensor

    def tensor_score(self, tensor, order=1, score_kind=0, mask=None, out=None):
        tensor = _f32(tensor)
        shape = tensor.shape[:-1]
        mask = _f32(mask)
        res = np.zeros(shape, np.float32) if out is None else np.array(out, np.float32, order="C")
        self._fn("tensor_score")(_i64(res.size), _ptr(tensor), _ptr(mask), _i(order),
                                 _i(score_kind), _ptr(res))
        return res

    def membrane(self, src, sigma, truncate_ratio, order, cut, cut_is_fraction, tv_sigma,
                 tv_exponent, tv_cutoff_ratio, mask=None, want_tensor=True, background_sigma=0.0, normalize=True):
        src = _f32(src)
        mask = _f32(mask)
        sal = np.zeros(src.shape, np.float32)
        dire = np.zeros(src.shape + (3,), np.float32)
        tensor = np.zeros(src.shape + (6,), np.float32) if want_tensor else None
        out = np.zeros(src.shape, np.float32)
        if background_sigma > 0.0:      # `-membrane-background` (handlers.cpp:1577-1592); restatement only
            if self.kind != "port":
                raise RuntimeError("membrane with background subtraction: only the restatement has this entry")
            thr = self._fn("membrane_background", _f)(*self._dims(src.shape), _ptr(src), _ptr(mask), _f(sigma),
                                                      _f(truncate_ratio), _i(order), _f(cut), _i(int(cut_is_fraction)),
                                                      _f(tv_sigma), _i(tv_exponent), _f(tv_cutoff_ratio),
                                                      _f(background_sigma), _i(int(normalize)), _ptr(sal), _ptr(dire),
                                                      _ptr(tensor), _ptr(out))
            return dict(threshold=thr, hess_saliency=sal, direction=dire, tensor=tensor, out=out)
        thr = self._fn("membrane", _f)(*self._dims(src.shape), _ptr(src), _ptr(mask), _f(sigma),
                                       _f(truncate_ratio), _i(order), _f(cut),
                                       _i(int(cut_is_fraction)), _f(tv_sigma), _i(tv_exponent),
                                       _f(tv_cutoff_ratio), _ptr(sal), _ptr(dire), _ptr(tensor),
                                       _ptr(out))
        return dict(threshold=thr, hess_saliency=sal, direction=dire, tensor=tensor, out=out)

    def label_connected(self, saliency, tensor, threshold_saliency, angle_deg=15.0, order=1, mask=None,
                        want_direction=False, thresholds=None):
        """The clustering step of HandleTV (handlers.cpp:1927-2034 -> LabelConnected, connect.hpp:171), both
        libraries: -> (labels int64, -1 undefined; number of clusters).  angle_deg as -connect-angle (settings.cpp:3075-3086)."""
        sal = _f32(saliency)
        ten = _f32(tensor)
        mask = _f32(mask)
        labels = np.zeros(sal.shape, np.int64)
        c = float(np.float32(np.cos(angle_deg * np.pi / 180.0)))
        tvs, tvn, tts, ttn = thresholds if thresholds is not None else (c, c, c, c)
        nz, ny, nx = sal.shape
        dire = np.zeros(sal.shape + (3,), np.float32) if want_direction else None
        n = self._fn("label_connected", _i64)(*self._dims(sal.shape), _ptr(sal), _ptr(mask), _ptr(ten), _i(order),
                                             _f(threshold_saliency), _f(tvs), _f(tvn), _f(tts), _f(ttn),
                                             labels.ctypes.data_as(C.c_void_p), _ptr(dire))
        if want_direction:
            return labels, int(n), dire
        return labels, int(n)

    def surface_points(self, saliency, direction, labels=None, mask=None, select_cluster=1, voxel_width=(1.0, 1.0, 1.0),
                       curve_ds=0.2, find_ridge=True, max_distance=1.3, capacity=None):
        """The oriented point cloud HandleTV writes with -normals-file (handlers.cpp:2039-2309): rows
        {x, y, z, nx, ny, nz} in raster order of the source voxel (restatement only: the reference code is inline in
        HandleTV; pinned on the stock binary's PLY files)."""
        if self.kind != "port":
            raise RuntimeError("surface_points: only the restatement has this entry")
        sal, dire, lab, mask = _f32(saliency), _f32(direction), _f32(labels), _f32(mask)
        cap = int(capacity if capacity is not None else sal.size)
        rows = np.zeros((cap, 6), np.float32)
        vw = (C.c_float * 3)(*[float(v) for v in voxel_width])
        n = self._fn("surface_points", _i64)(*self._dims(sal.shape), _ptr(sal), _ptr(dire), _ptr(lab), _ptr(mask),
                                            _i(select_cluster), vw, _f(curve_ds), _i(int(find_ridge)), _f(max_distance),
                                            _ptr(rows), _i64(cap))
        return rows[:min(int(n), cap)], int(n)

    # ---- thresholds ----------------------------------------------------------
    def threshold1(self, a, thr, outA=0.0, outB=1.0):
        a = _f32(a)
        out = np.empty_like(a)
        self._fn("threshold1")(_i64(a.size), _ptr(a), _ptr(out), _f(thr), _f(outA), _f(outB))
        return out

    def threshold2(self, a, t_a, t_b, outA=0.0, outB=1.0):
        a = _f32(a)
        out = np.empty_like(a)
        self._fn("threshold2")(_i64(a.size), _ptr(a), _ptr(out), _f(t_a), _f(t_b), _f(outA),
                               _f(outB))
        return out

    def threshold4(self, a, a01, b01, a10, b10, outA=0.0, outB=1.0):
        a = _f32(a)
        out = np.empty_like(a)
        self._fn("threshold4")(_i64(a.size), _ptr(a), _ptr(out), _f(a01), _f(b01), _f(a10),
                               _f(b10), _f(outA), _f(outB))
        return out

    def average(self, a, w=None):
        a = _f32(a)
        w = _f32(w)
        if self.kind == "port":
            return self._fn("average", _f)(_i64(a.size), _ptr(a), _ptr(w))
        return self._fn("average", _f)(*self._dims(a.shape), _ptr(a), _ptr(w))

    def stddev(self, a, w=None):
        a = _f32(a)
        w = _f32(w)
        if self.kind == "port":
            return self._fn("stddev", _f)(_i64(a.size), _ptr(a), _ptr(w))
        return self._fn("stddev", _f)(*self._dims(a.shape), _ptr(a), _ptr(w))

    # ---- blobs ---------------------------------------------------------------
    def _resample(self, name, a, dst_shape, offset):
        a = _f32(a)
        out = np.empty(tuple(dst_shape), np.float32)
        ss = (_i64 * 3)(a.shape[2], a.shape[1], a.shape[0])
        ds = (_i64 * 3)(dst_shape[2], dst_shape[1], dst_shape[0])
        off = None if offset is None else (_i * 3)(*[int(v) for v in offset])
        rc = self._fn(name, _i)(ss, ds, _ptr(a), _ptr(out), off)
        if rc:
            raise ValueError("the reference rejects these binning arguments")
        return out

    def bin3d(self, a, bin_size=None, dst_shape=None, offset=None):
        """BinArray3D, lib/visfd/resample.hpp:53-104"""
        if dst_shape is None:
            dst_shape = tuple(int(n) // int(bin_size) for n in a.shape)
        return self._resample("bin3d", a, dst_shape, offset)

    def unbin3d(self, a, dst_shape, offset=None):
        """UnbinArray3D, lib/visfd/resample.hpp:106-166"""
        return self._resample("unbin3d", a, dst_shape, offset)

    def draw_regions(self, image, regions, mask=None, negative_means_subtract=False):
        """DrawRegions, lib/visfd/draw.hpp:90-237.  regions: ("rect", xmin, xmax, ymin, ymax, zmin,
        zmax, value) or ("sphere", x0, y0, z0, r, value) in voxels.  Returns the painted copy."""
        img = np.array(image, np.float32, order="C", copy=True)
        mask = _f32(mask)
        rec = pack_regions(regions)
        nz, ny, nx = img.shape
        self._fn("draw_regions", _i)(_i(nx), _i(ny), _i(nz), _ptr(img), _ptr(mask), _ptr(rec), _i(len(rec)),
                                     _i(int(negative_means_subtract)))
        return img

    def blob_dog(self, src, sigmas, delta=0.02, truncate_ratio=2.5, mask=None,
                 minima_threshold=np.inf, maxima_threshold=-np.inf, use_threshold_ratios=True,
                 capacity=1 << 20):
        src = _f32(src)
        mask = _f32(mask)
        sg = _f32(np.asarray(sigmas))
        mc = np.zeros((capacity, 3), np.float32)
        ms = np.zeros(capacity, np.float32)
        msc = np.zeros(capacity, np.float32)
        xc = np.zeros((capacity, 3), np.float32)
        xs = np.zeros(capacity, np.float32)
        xsc = np.zeros(capacity, np.float32)
        nmin, nmax = _i64(), _i64()
        self._fn("blob_dog")(*self._dims(src.shape), _ptr(src), _ptr(mask), _ptr(sg),
                             _i(len(sg)), _f(delta), _f(truncate_ratio), _f(minima_threshold),
                             _f(maxima_threshold), _i(int(use_threshold_ratios)),
                             _i64(capacity), _ptr(mc), _ptr(ms), _ptr(msc), C.byref(nmin),
                             _ptr(xc), _ptr(xs), _ptr(xsc), C.byref(nmax))
        a, b = min(nmin.value, capacity), min(nmax.value, capacity)

        def pack(c, s, sc, n):
            t = np.concatenate([c[:n], s[:n, None], sc[:n, None]], axis=1)
            # canonical order: (sigma, z, y, x)
            idx = np.lexsort((t[:, 0], t[:, 1], t[:, 2], t[:, 3]))
            return t[idx]
        return pack(mc, ms, msc, a), pack(xc, xs, xsc, b)
