"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and
exports every symbol include/visfd_cuda.h declares; host-only helpers agree with the
oracle; a context cannot be created without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    """every function declared in include/*.h (visfd_cuda.h: the GPU path; visfd_mrc.h: MRC file I/O)"""
    names = set()
    for header in sorted(os.listdir(os.path.join(ROOT, "include"))):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(visfd_(?:cuda|mrc|blobs)_\w+)\s*\(", text))
    return sorted(names)


def test_header_symbols_exported():
    import visfd_b200
    lib = visfd_b200.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/*.h but not exported"
    assert "visfd_mrc_read" in names and "visfd_cuda_bin3d" in names and "visfd_blobs_discard_overlapping" in names


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "visfd_cuda.h"\n#include "visfd_mrc.h"\n#include "visfd_blobs.h"\n'
                   'int main(void){ visfd_membrane_params p; visfd_mrc_header h; (void)p; (void)h; return 0; }\n')
    import subprocess
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_host_helpers_match_oracle(oracle, golden):
    import visfd_b200 as vb
    for k, (s, hw) in enumerate(golden["taps_cases"]):
        assert np.array_equal(vb.gen_gauss1d(float(s), int(hw)), golden[f"taps_{k}"])
    lib = oracle.lib
    lib.vo_gauss_halfwidth.restype = ctypes.c_int
    lib.vo_tv_halfwidth.restype = ctypes.c_int
    for s in (0.3, 0.8269, 2.0, 2.99991, 3.2, 4.0, 6.4, 7.13, 8.0, 12.8):
        assert vb.gauss_halfwidth(s, -1.0, 0.03) == lib.vo_gauss_halfwidth(
            ctypes.c_float(s), ctypes.c_float(-1.0), ctypes.c_float(0.03))
        assert vb.gauss_halfwidth(s, 2.5, 0.03) == lib.vo_gauss_halfwidth(
            ctypes.c_float(s), ctypes.c_float(2.5), ctypes.c_float(0.03))
        r = float(np.float32(np.sqrt(2.0)))
        assert vb.tv_halfwidth(s * 4.733, r) == lib.vo_tv_halfwidth(ctypes.c_float(s * 4.733), ctypes.c_float(r))
    # sigma -> half-width table of config C2 and the C4 vote radius
    assert [vb.gauss_halfwidth(s) for s in (2, 3.2, 4, 6.4, 8, 12.8)] == [5, 8, 10, 16, 21, 33]
    assert vb.tv_halfwidth(14.1986, float(np.float32(np.sqrt(2.0)))) == 20


def test_select_step_is_a_radix_select(oracle):
    """host half of the cut: narrowing over 2048-bin histograms reproduces the sort-based
    threshold (the histograms themselves are built here with numpy)"""
    import visfd_b200 as vb
    lib = vb.load_library()
    rng = np.random.default_rng(0)
    for n, frac in ((1000, 0.05), (4097, 0.5), (50000, 0.013)):
        v = (rng.standard_normal(n) ** 2 * 1e6).astype(np.float32)
        v[::7] = 0
        v[:20] *= -1
        bits = v.view(np.uint32)
        keys = np.where(bits & 0x80000000, ~bits, bits | 0x80000000).astype(np.uint32)
        want_cut, want_thr = oracle.saliency_cut(v, frac, True)
        rank = int(np.floor(np.float32(n) * np.float32(frac)))
        prefix, pbits = ctypes.c_uint32(0), ctypes.c_int(0)
        r = ctypes.c_uint64(rank)
        while pbits.value < 32:
            nb = 11 if 32 - pbits.value >= 11 else 32 - pbits.value
            sel = keys if pbits.value == 0 else keys[(keys >> np.uint32(32 - pbits.value)) == prefix.value]
            b = (sel >> np.uint32(32 - pbits.value - nb)) & np.uint32((1 << nb) - 1)
            hist = np.zeros(2048, np.uint64)
            hist[:1 << nb] = np.bincount(b, minlength=1 << nb)
            assert lib.visfd_cuda_select_step(hist.ctypes.data_as(ctypes.c_void_p), ctypes.byref(prefix),
                                              ctypes.byref(pbits), ctypes.byref(r)) == 0
        assert np.float32(lib.visfd_cuda_key_to_float(prefix)) == np.float32(want_thr)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import visfd_b200
    with pytest.raises(visfd_b200.VisfdCudaError):
        visfd_b200.Context(0)
