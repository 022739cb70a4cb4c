"""The reference's own command-line program (bin/filter_mrc, host C++ unchanged) rebuilt with its hot path redirected
to libvisfd_cuda.so (integration/build_filter_mrc_cuda.py), driven through the reference's own tests for the path:
tests/test_membrane_detection.sh:8-16 (BASELINE config 1, both passes) and tests/test_blob_detection.sh:21.
Expected outputs: what the STOCK binary produced when the fixtures were made (tests/golden/make_golden.py)."""
import os
import subprocess

import numpy as np
import pytest

from util import rel_err, TOL_SALIENCY
from visfd_b200 import mrc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "integration", "_build", "filter_mrc_cuda")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        pytest.skip("integration/_build/filter_mrc_cuda is missing: run `python integration/build_filter_mrc_cuda.py` "
                    "where the reference tree is available (__graft_entry__.build() does); like oracle/_ref it is built "
                    "in the dev container and shipped to the GPU box")
    return EXE


def write_rec(io, path, vol):
    h = mrc.MrcHeader()
    nz, ny, nx = vol.shape
    h.nvoxels[:] = (nx, ny, nz)
    h.mvoxels[:] = (nx, ny, nz)
    h.mode = 2
    h.cellA[:] = (float(nx), float(ny), float(nz))
    h.cellB[:] = (90.0, 90.0, 90.0)
    h.mapCRS[:] = (1, 2, 3)
    io.write(path, h, vol)


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode == 0, "%s\n%s" % (" ".join(cmd), p.stderr[-3000:])
    return p.stderr


def test_membrane_detection_sh_through_the_cuda_cli(exe, golden, tmp_path):
    io = mrc.open_library()
    write_rec(io, tmp_path / "test_image_membrane.rec", golden["c1_in_raw"])
    base = ["-w", "19.2", "-in", "test_image_membrane.rec", "-membrane", "minima", "55", "-tv", "4",
            "-tv-angle-exponent", "4", "-bin", "2"]
    # ---- pass 1 (test_membrane_detection.sh:8): fused GPU pipeline, tensors saved for pass 2 ----
    log = run([exe] + base + ["-out", "pass1.rec", "-save-progress", "test_image_membrane"], tmp_path)
    assert "membrane pipeline on the GPU" in log
    _, out1 = io.read(tmp_path / "pass1.rec")
    want = golden["c1_out"]
    assert out1.shape == want.shape
    assert rel_err(out1, want) <= TOL_SALIENCY
    assert int(np.count_nonzero(out1)) == int(np.count_nonzero(want)) == 419
    assert np.unravel_index(np.argmax(out1), out1.shape) == (2, 3, 2)
    assert all((tmp_path / ("test_image_membrane_tensor_%d.rec" % d)).exists() for d in range(6))
    # ---- pass 2 (:9): tensors loaded back, clustering by visfd_cuda::LabelConnected, normals file ----
    log = run([exe] + base + ["-out", "pass2.rec", "-load-progress", "test_image_membrane", "-connect", "1e+09",
                              "-connect-angle", "30", "-normals-file", "normals.ply", "-select-cluster", "1"], tmp_path)
    assert "Number of clusters found: 1" in log
    _, out2 = io.read(tmp_path / "pass2.rec")
    cli = golden["c1_connect_labels"]
    assert np.array_equal(out2, cli.astype(np.float32))
    assert int((out2 == 1).sum()) == 69
    assert "surface point cloud on the GPU" in log
    ply = (tmp_path / "normals.ply").read_text().splitlines()
    assert "element vertex 58" in ply
    # ... and they are the stock binary's 58 vertices (tests/golden/surface_points.npz), to the six digits of the file
    rows = np.array([[float(v) for v in l.split()] for l in ply[ply.index("end_header") + 1:]], np.float64)
    want_ply = np.load(os.path.join(ROOT, "tests", "golden", "surface_points.npz"))["c1_ply"].astype(np.float64)
    assert rows.shape == want_ply.shape
    den = np.maximum(np.abs(want_ply), 1e-3 * np.abs(want_ply).max(axis=0))
    assert (np.abs(rows - want_ply) / den).max() <= 1e-4
    # ---- both passes in one run: pipeline, tensors and clustering all on the GPU ----
    log = run([exe] + base + ["-out", "both.rec", "-connect", "1e+09", "-connect-angle", "30"], tmp_path)
    assert "membrane pipeline on the GPU" in log and "Number of clusters found: 1" in log
    _, out3 = io.read(tmp_path / "both.rec")
    assert np.array_equal(out3, cli.astype(np.float32))


def test_blob_detection_sh_through_the_cuda_cli(exe, golden, tmp_path):
    io = mrc.open_library()
    write_rec(io, tmp_path / "test_blob_detect.rec", golden["blobfix_img"])
    write_rec(io, tmp_path / "test_blob_detect_mask.rec", golden["blobfix_mask"])
    # test_blob_detection.sh:21
    run([exe, "-w", "19.6", "-mask", "test_blob_detect_mask.rec", "-in", "test_blob_detect.rec", "-blob", "minima",
         "test_blobs.txt", "160.0", "280.0", "1.01"], tmp_path)
    got = np.loadtxt(tmp_path / "test_blobs.txt", ndmin=2)
    want = golden["blobfix_cli"]
    assert got.shape == want.shape == (11, 5)
    assert np.array_equal(got, want), "blob list differs from the stock binary's"
    # :25-29 non-max suppression (host code of the reference, reading the list above): exactly 2 blobs
    run([exe, "-w", "19.6", "-mask", "test_blob_detect_mask.rec", "-in", "test_blob_detect.rec", "-discard-blobs",
         "test_blobs.txt", "nms.txt", "-blob-separation", "1.1", "-minima-threshold", "-90"], tmp_path)
    nms = np.loadtxt(tmp_path / "nms.txt", ndmin=2)
    assert np.array_equal(nms, golden["blobnms_cli"])


def test_gauss_dog_through_the_cuda_cli(exe, golden, oracle, tmp_path):
    """-gauss and -dog (HandleGauss / HandleDog, handlers.cpp:218-357) through the CLI == the restated reference."""
    io = mrc.open_library()
    vol = golden["blob_vol"]
    write_rec(io, tmp_path / "in.rec", vol)
    run([exe, "-w", "1", "-in", "in.rec", "-out", "g.rec", "-gauss", "2.0"], tmp_path)
    _, g = io.read(tmp_path / "g.rec")
    ratio = float(np.float32(np.sqrt(np.float32(-2) * np.log(np.float32(0.03)))))   # -truncate-threshold default
    hw = max(1, int(np.floor(np.float32(2.0) * np.float32(ratio))))
    want, _ = oracle.apply_gauss(vol, 2.0, hw)
    assert np.array_equal(g, want)
    run([exe, "-w", "1", "-in", "in.rec", "-out", "d.rec", "-dog", "1.5", "3.0"], tmp_path)
    _, d = io.read(tmp_path / "d.rec")
    hwa = max(1, int(np.floor(np.float32(1.5) * np.float32(ratio))))
    hwb = max(1, int(np.floor(np.float32(3.0) * np.float32(ratio))))
    ga, _ = oracle.apply_gauss(vol, 1.5, hwa)
    gb, _ = oracle.apply_gauss(vol, 3.0, hwb)
    assert np.array_equal(d, ga - gb)


def test_membrane_background_through_the_cuda_cli(exe, tmp_path):
    """`-membrane-background` no longer leaves the fused GPU path (handlers.cpp:1577-1592): the stock binary's image"""
    sp = np.load(os.path.join(ROOT, "tests", "golden", "surface_points.npz"))
    io = mrc.open_library()
    write_rec(io, tmp_path / "vol.rec", sp["s_vol"])
    log = run([exe, "-w", "1", "-in", "vol.rec", "-out", "bg.rec", "-membrane", "minima", "3.4641", "-tv", "2.5",
               "-tv-angle-exponent", "4", "-bin", "1", "-membrane-background", "6"], tmp_path)
    assert "membrane pipeline on the GPU" in log
    _, out = io.read(tmp_path / "bg.rec")
    want = sp["s_background6_out"]
    assert np.abs(out - want).max() <= 2e-4 * np.abs(want).max()


def test_cli_filters_a_volume_beyond_2_31_voxels(exe, oracle, tmp_path):
    """with Alloc3D's products in size_t and MrcSimple's named-file Read / Write going through
    include/visfd_mrc.h (both edits of integration/build_filter_mrc_cuda.py), the reference's own program reads,
    filters (`-gauss 2`, on the GPU) and writes a 1296 x 1296 x 1280 volume = 2.15e9 voxels > 2^31.  Checked against
    the oracle on crops (bit for bit), including the planes that hold flat index 2^31 and the global z border."""
    import shutil
    import psutil
    nz, ny, nx = 1280, 1296, 1296
    n = nz * ny * nx
    assert n > 2 ** 31
    if shutil.disk_usage(tmp_path).free < 3 * 4 * n or psutil.virtual_memory().available < 6 * 4 * n:
        pytest.skip("needs ~26 GB of scratch disk and ~52 GB of free host memory")
    rng = np.random.default_rng(11)
    tile = rng.standard_normal((16, ny, nx)).astype(np.float32)
    hdr = mrc.MrcHeader()
    hdr.nvoxels[:] = (nx, ny, nz)
    hdr.mvoxels[:] = (nx, ny, nz)
    hdr.mode = 2
    hdr.cellA[:] = (float(nx), float(ny), float(nz))
    hdr.cellB[:] = (90.0, 90.0, 90.0)
    hdr.mapCRS[:] = (1, 2, 3)

    def plane_block(z0, z1):
        return np.stack([tile[z % 16] + np.float32(0.01 * z) for z in range(z0, z1)])

    src = tmp_path / "big.rec"
    with open(src, "wb") as f:
        f.write(hdr.as_bytes()[:1024])     # (the struct carries one more field than the file header)
        for z0 in range(0, nz, 16):
            plane_block(z0, min(z0 + 16, nz)).tofile(f)
    run([exe, "-w", "1", "-in", "big.rec", "-out", "big_gauss.rec", "-gauss", "2"], tmp_path)
    out = np.memmap(tmp_path / "big_gauss.rec", dtype=np.float32, mode="r", offset=1024, shape=(nz, ny, nx))
    assert os.path.getsize(tmp_path / "big_gauss.rec") == 1024 + 4 * n
    hw = 5   # floor(2 * 2.6482)
    z_of_2_31 = (2 ** 31) // (ny * nx)
    for (za, zb, ya, yb, xa, xb) in ((600, 640, 100, 164, 1200, 1296),           # interior in z, the x border
                                     (z_of_2_31 - 24, nz, 0, 48, 0, 64),          # flat index 2^31, the z and y borders
                                     (0, 24, ny - 40, ny, 640, 700)):             # the first planes
        # margin of hw towards the interior, the volume's own border otherwise (the oracle renormalises there too)
        lo = [max(za - hw, 0), max(ya - hw, 0), max(xa - hw, 0)]
        hi = [min(zb + hw, nz), min(yb + hw, ny), min(xb + hw, nx)]
        crop = plane_block(lo[0], hi[0])[:, lo[1]:hi[1], lo[2]:hi[2]]
        want, _ = oracle.apply_gauss(np.ascontiguousarray(crop), 2.0, hw)
        sl = (slice(za - lo[0], zb - lo[0]), slice(ya - lo[1], yb - lo[1]), slice(xa - lo[2], xb - lo[2]))
        got = np.asarray(out[za:zb, ya:yb, xa:xb])
        assert np.array_equal(got, want[sl]), (za, ya, xa, float(np.abs(got - want[sl]).max()))
    del out
