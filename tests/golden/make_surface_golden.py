"""Golden vectors for the oriented point cloud (`-normals-file`, bin/filter_mrc/handlers.cpp:2039-2309), made by the
STOCK filter_mrc binary (oracle/_ref/filter_mrc, compiled from /root/reference by oracle/Makefile):
  c1_ply     the 58 vertices of the reference's own test, tests/test_membrane_detection.sh:9 (BASELINE config 1)
  s_*        a 40x44x48 synthetic tomogram (visfd_b200.synth, seed 3): `-membrane minima 3.4641 -tv 2.5
             -tv-angle-exponent 4 -bin 1 -connect T -connect-angle 30 -select-cluster 1 -normals-file`:
             the volume, the cluster image the binary wrote, T and the PLY rows;
  s_background6_out  the same volume through `-membrane ... -membrane-background 6` (no clustering): the output image.
Run in the development container (needs /root/reference); the output travels as tests/golden/surface_points.npz."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from visfd_b200 import mrc, synth  # noqa: E402

REFERENCE = os.environ.get("VISFD_REFERENCE", "/root/reference")
FM = os.path.join(ROOT, "oracle", "_ref", "filter_mrc")


def read_ply(path):
    lines = open(path).read().splitlines()
    k = lines.index("end_header")
    n = int([l for l in lines[:k] if l.startswith("element vertex")][0].split()[2])
    rows = np.array([[float(v) for v in l.split()] for l in lines[k + 1:]], np.float32).reshape(-1, 6)
    assert len(rows) == n
    return rows


def main():
    io = mrc.open_library()
    out = {}
    with tempfile.TemporaryDirectory() as td:
        fixture = os.path.join(REFERENCE, "tests", "test_image_membrane.rec")
        subprocess.run([FM, "-w", "19.2", "-in", fixture, "-out", os.path.join(td, "c1.rec"), "-membrane", "minima", "55",
                        "-tv", "4", "-tv-angle-exponent", "4", "-bin", "2", "-connect", "1e+09", "-connect-angle", "30",
                        "-select-cluster", "1", "-normals-file", os.path.join(td, "c1.ply")], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        out["c1_ply"] = read_ply(os.path.join(td, "c1.ply"))
        vol = synth.tomogram((40, 44, 48), seed=3)
        h = mrc.MrcHeader()
        nz, ny, nx = vol.shape
        h.nvoxels[:] = (nx, ny, nz)
        h.mvoxels[:] = (nx, ny, nz)
        h.mode = 2
        h.cellA[:] = (float(nx), float(ny), float(nz))
        h.cellB[:] = (90.0, 90.0, 90.0)
        h.mapCRS[:] = (1, 2, 3)
        io.write(os.path.join(td, "vol.rec"), h, vol)
        base = [FM, "-w", "1", "-in", os.path.join(td, "vol.rec"), "-membrane", "minima", "3.4641", "-tv", "2.5",
                "-tv-angle-exponent", "4", "-bin", "1"]
        subprocess.run(base + ["-out", os.path.join(td, "p1.rec")], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        _, o1 = io.read(os.path.join(td, "p1.rec"))
        thr = float(np.float32(np.percentile(o1[o1 > 0], 75)))
        subprocess.run(base + ["-out", os.path.join(td, "p2.rec"), "-connect", repr(thr), "-connect-angle", "30",
                               "-select-cluster", "1", "-normals-file", os.path.join(td, "s.ply")], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        # `-membrane-background 6` (handlers.cpp:1577-1592): both scores times (source - blurred source)
        subprocess.run(base + ["-out", os.path.join(td, "pb.rec"), "-membrane-background", "6"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _, out["s_background6_out"] = io.read(os.path.join(td, "pb.rec"))
        _, lab = io.read(os.path.join(td, "p2.rec"))
        out["s_vol"] = vol
        out["s_labels"] = lab.astype(np.uint8)
        assert np.array_equal(out["s_labels"].astype(np.float32), lab)
        out["s_threshold"] = np.float32(thr)
        out["s_ply"] = read_ply(os.path.join(td, "s.ply"))
    print("C1: %d vertices; synthetic: %d vertices, %d voxels in cluster 1, threshold %.9g" %
          (len(out["c1_ply"]), len(out["s_ply"]), int((out["s_labels"] == 1).sum()), thr))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "surface_points.npz"), **out)


if __name__ == "__main__":
    main()
