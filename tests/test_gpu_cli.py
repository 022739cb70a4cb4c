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
