#!/usr/bin/env python
"""Build the reference's OWN command-line program, bin/filter_mrc, with its membrane / blob hot path
redirected to libvisfd_cuda.so -- the proof that the library is a drop-in behind visfd's API.

What it does (nothing from the reference tree is committed here, and the tree is never written to):
  1. copies bin/filter_mrc/handlers.cpp, feature_variants.hpp, lib/visfd/alloc3d.hpp and lib/mrc_simple/mrc_simple.cpp
     from $VISFD_REFERENCE (default /root/reference) into integration/_build/src/ and applies the edits below, each anchored on a
     short, unique piece of the original text (the script fails loudly if an anchor is missing or
     ambiguous, e.g. after an upstream change);
  2. compiles the patched copies together with the UNMODIFIED settings.cpp, filter_mrc.cpp,
     handlers_unsupported.cpp and mrc_header.cpp (read in place) with the reference's
     own flags (setup_gcc.sh: -O3 -DNDEBUG -fopenmp) plus -DVISFD_USE_CUDA;
  3. links libvisfd_cuda.so (rpath relative to the binary) -> integration/_build/filter_mrc_cuda.

The edits are what INTEGRATION.md calls "the patch": every one is guarded by VISFD_USE_CUDA, so the
patched sources still build the stock program without the define.
  * handlers.cpp: include the shim after `using namespace visfd;`
  * HandleGauss / HandleDog / HandleLoGDoG: visfd:: -> visfd_cuda:: (same argument lists)
  * HandleBinning: BinArray3D -> visfd_cuda::BinArray3D
  * HandleTV: lines 1618-1892 (CalcHessian, eigen loop, cut, TVDenseStick, score loop) replaced by ONE call,
    visfd_cuda::MembranePipeline (with `-membrane-background` too), unless the run loads saved tensors or detects
    edges or curves; the clustering call becomes visfd_cuda::LabelConnected; the loop that collects the oriented
    point cloud of -normals-file (2050-2299) becomes visfd_cuda::SurfacePointCloud
  * HandleLabelConnected: visfd_cuda::LabelConnected
  * feature_variants.hpp (BlobDogNM): BlobDogD -> visfd_cuda::BlobDogD
  * lib/visfd/alloc3d.hpp: the byte counts and row offsets of Alloc3D in size_t instead of int (:33-35, :60)
  * lib/mrc_simple/mrc_simple.cpp: MrcSimple::Read(file name) / Write(file name) through include/visfd_mrc.h (bulk I/O,
    64-bit sizes) -- with the two edits above the same binary opens, filters and writes volumes of 2^31 voxels and more

tests/test_gpu_cli.py runs the reference's own two tests for the path through this binary.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("VISFD_REFERENCE", "/root/reference")
FM = os.path.join(REF, "bin", "filter_mrc")
OUT = os.path.join(HERE, "_build")
SRC = os.path.join(OUT, "src")
CXX = "/usr/bin/g++"


def edit(text, anchor, replacement, count=1, after=None, what=""):
    """Replace `count` occurrences of `anchor` (the first ones following `after`, if given)."""
    start = 0
    if after is not None:
        assert text.count(after) >= 1, "anchor missing: %r (%s)" % (after, what)
        start = text.index(after)
    head, tail = text[:start], text[start:]
    assert tail.count(anchor) >= count, "anchor missing: %r (%s)" % (anchor, what)
    return head + tail.replace(anchor, replacement, count)


SHIM_INCLUDE = """using namespace visfd;
#ifdef VISFD_USE_CUDA
#define VISFD_CUDA_ERR_BASE visfd::VisfdErr   /* shim errors ARE VisfdErr (lib/visfd/err_visfd.hpp) */
#include <visfd_cuda_shim.hpp>
#define VISFD_NS visfd_cuda
#else
#define VISFD_NS visfd
#endif
"""

FUSED = """#ifdef VISFD_USE_CUDA
  if (visfd_fused) {
    visfd_membrane_params p;
    p.sigma           = sigma;
    p.truncate_ratio  = filter_truncate_ratio;
    p.eival_order     = (eival_order == selfadjoint_eigen3::DECREASING_EIVALS) ? VISFD_DECREASING_EIVALS
                                                                                : VISFD_INCREASING_EIVALS;
    p.cut             = settings.hessian_score_threshold;
    p.cut_is_fraction = settings.hessian_score_threshold_is_a_fraction;
    p.tv_sigma        = settings.tv_sigma;
    p.tv_exponent     = settings.tv_exponent;
    p.tv_cutoff_ratio = settings.tv_truncate_ratio;
    const bool want_tensor = (settings.tv_sigma > 0.0) &&
      ((settings.save_intermediate_fname_base != "") || settings.cluster_connected_voxels ||
       (settings.out_normals_fname != ""));
    cerr << "-- membrane pipeline on the GPU (libvisfd_cuda) --" << endl;
    visfd_cuda::MembranePipeline(image_size, tomo_in.aaafI, tomo_out.aaafI, mask.aaafI, p,
                                 aaaafGradient,  // = aaaafDirection below: eivects[0] of the Hessian
                                 want_tensor ? hessian_tensor.aaaafI : nullptr,
                                 subtract_background ? settings.width_b[0] : 0.0f,   // -membrane-background
                                 settings.normalize_near_boundaries);
  }
  else
#endif
  CalcHessian(image_size,"""


def patch_handlers(t):
    t = edit(t, "using namespace visfd;\n", SHIM_INCLUDE, what="shim include")
    t = edit(t, "  A = ApplyGauss(tomo_in.header.nvoxels,", "  A = VISFD_NS::ApplyGauss(tomo_in.header.nvoxels,", what="HandleGauss")
    t = edit(t, "  ApplyDog(tomo_in.header.nvoxels,", "  VISFD_NS::ApplyDog(tomo_in.header.nvoxels,", what="HandleDog")
    t = edit(t, "  ApplyLog(tomo_in.header.nvoxels,", "  VISFD_NS::ApplyLog(tomo_in.header.nvoxels,", what="HandleLoGDoG")
    # ---- HandleTV ----
    tv = "HandleTV(const Settings &settings,"
    t = edit(t, "  bool subtract_background = (settings.width_b[0] > 0.0);\n  if (subtract_background) {\n",
             "  bool subtract_background = (settings.width_b[0] > 0.0);\n  bool visfd_fused = false;\n"
             "#ifdef VISFD_USE_CUDA\n"
             "  visfd_fused = (settings.load_intermediate_fname_base == \"\") &&\n"
             "                (settings.filter_type == Settings::SURFACE_RIDGE);\n"
             "#endif\n"
             "  if (subtract_background && (! visfd_fused)) {   // fused: the peak height is applied on the GPU\n",
             after=tv, what="flag")
    t = edit(t, "  CalcHessian(image_size,", FUSED, after=tv, what="fused call")
    t = edit(t, "  for(int iz=0; iz < image_size[2]; iz++)\n    for(int iy=0; iy < image_size[1]; iy++)\n"
                "      for(int ix=0; ix < image_size[0]; ix++)\n        tomo_out.aaafI[iz][iy][ix] = 0.0;\n",
             "  if (! visfd_fused)\n  for(int iz=0; iz < image_size[2]; iz++)\n    for(int iy=0; iy < image_size[1]; iy++)\n"
             "      for(int ix=0; ix < image_size[0]; ix++)\n        tomo_out.aaafI[iz][iy][ix] = 0.0;\n",
             after="Diagonalizing the Hessians", what="zero fill")
    t = edit(t, "  for(int iz=0; iz < image_size[2]; iz++) {\n    #pragma omp parallel for collapse(2)",
             "  if (! visfd_fused)\n  for(int iz=0; iz < image_size[2]; iz++) {\n    #pragma omp parallel for collapse(2)",
             after="Diagonalizing the Hessians", what="eigen loop")
    t = edit(t, "  { // Use thresholding to reduce the number of voxels that we have to consider",
             "  if (! visfd_fused)\n  { // Use thresholding to reduce the number of voxels that we have to consider",
             after="Diagonalizing the Hessians", what="cut")
    t = edit(t, "  if (settings.tv_sigma > 0.0) {", "  if ((settings.tv_sigma > 0.0) && (! visfd_fused)) {",
             after="float ****aaaafVoteTensor = hessian_tensor.aaaafI;", what="voting")
    t = edit(t, "      LabelConnected(image_size, //image size", "      VISFD_NS::LabelConnected(image_size, //image size",
             after=tv, what="clustering in HandleTV")
    # the point cloud of -normals-file: one call instead of the loop over the voxels (handlers.cpp:2050-2299)
    t = edit(t, "    for (int iz = 0; iz < image_size[2]; ++iz) {\n      for (int iy = 0; iy < image_size[1]; ++iy) {\n"
                "        for (int ix = 0; ix < image_size[0]; ++ix) {\n          if (mask.aaafI && (mask.aaafI[iz][iy][ix] == 0.0))",
             "#ifdef VISFD_USE_CUDA\n"
             "    cerr << \"-- surface point cloud on the GPU (libvisfd_cuda) --\" << endl;\n"
             "    visfd_cuda::SurfacePointCloud(image_size, aaafSaliency, aaaafDirection, aaafVoxel2Cluster, mask.aaafI,\n"
             "                                  settings.select_cluster, voxel_width, settings.surface_normal_curve_ds,\n"
             "                                  settings.surface_find_ridge, settings.max_distance_to_feature, crds, norms);\n"
             "    if (false)\n"
             "#endif\n"
             "    for (int iz = 0; iz < image_size[2]; ++iz) {\n      for (int iy = 0; iy < image_size[1]; ++iy) {\n"
             "        for (int ix = 0; ix < image_size[0]; ++ix) {\n          if (mask.aaafI && (mask.aaafI[iz][iy][ix] == 0.0))",
             after="aaafVoxel2Cluster = tomo_out.aaafI; //tomo_out was filled by LabelConnected()", what="point cloud")
    t = edit(t, "    LabelConnected(tomo_in.header.nvoxels, //image size",
             "    VISFD_NS::LabelConnected(tomo_in.header.nvoxels, //image size", what="HandleLabelConnected")
    t = edit(t, "  BinArray3D(tomo_in.header.nvoxels,", "  VISFD_NS::BinArray3D(tomo_in.header.nvoxels,", what="HandleBinning")
    t = edit(t, "    BinArray3D(mask.header.nvoxels,", "    VISFD_NS::BinArray3D(mask.header.nvoxels,", what="HandleBinning mask")
    return t


MRC_BRIDGE = """#include "mrc_simple.hpp"
#ifdef VISFD_USE_CUDA
// lib/mrc_simple moves the voxels one stream call at a time (13 Mvoxel/s); with VISFD_USE_CUDA the named-file
// Read / Write go through include/visfd_mrc.h (libvisfd_cuda.so: same file semantics, bulk I/O, 64-bit sizes)
#include <visfd_mrc.h>
static void visfd_header_to_c(const MrcHeader &a, visfd_mrc_header &b) {
  visfd_mrc_header_init(&b);
  for (int d = 0; d < 3; d++) {
    b.nvoxels[d] = a.nvoxels[d]; b.nstart[d] = a.nstart[d]; b.mvoxels[d] = a.mvoxels[d];
    b.cellA[d] = a.cellA[d]; b.cellB[d] = a.cellB[d]; b.mapCRS[d] = a.mapCRS[d]; b.origin[d] = a.origin[d];
  }
  b.mode = a.mode; b.dmin = a.dmin; b.dmax = a.dmax; b.dmean = a.dmean; b.ispg = a.ispg; b.nsymbt = a.nsymbt;
  memcpy(b.extra_raw_data, a.extra_raw_data, sizeof(b.extra_raw_data));
  memcpy(b.remaining_raw_data, a.remaining_raw_data, sizeof(b.remaining_raw_data));
  b.use_signed_bytes = a.use_signed_bytes ? 1 : 0;
}
static void visfd_header_from_c(const visfd_mrc_header &b, MrcHeader &a) {
  for (int d = 0; d < 3; d++) {
    a.nvoxels[d] = b.nvoxels[d]; a.nstart[d] = b.nstart[d]; a.mvoxels[d] = b.mvoxels[d];
    a.cellA[d] = b.cellA[d]; a.cellB[d] = b.cellB[d]; a.mapCRS[d] = b.mapCRS[d]; a.origin[d] = b.origin[d];
  }
  a.mode = b.mode; a.dmin = b.dmin; a.dmax = b.dmax; a.dmean = b.dmean; a.ispg = b.ispg; a.nsymbt = b.nsymbt;
  memcpy(a.extra_raw_data, b.extra_raw_data, sizeof(b.extra_raw_data));
  memcpy(a.remaining_raw_data, b.remaining_raw_data, sizeof(b.remaining_raw_data));
  a.use_signed_bytes = b.use_signed_bytes != 0;
}
#endif
"""

MRC_READ = """                     float ***aaafMask) {
#ifdef VISFD_USE_CUDA
  {
    visfd_mrc_header h;
    if (visfd_mrc_read_header(in_file_name.c_str(), &h))
      throw MrcfileErr(string("Error: ") + visfd_mrc_last_error() + "\\n");
    visfd_header_from_c(h, header);
    Dealloc();
    Alloc();
    const long long n = (long long)header.nvoxels[0] * header.nvoxels[1] * header.nvoxels[2];
    if (visfd_mrc_read(in_file_name.c_str(), &h, &(aaafI[0][0][0]), n))
      throw MrcfileErr(string("Error: ") + visfd_mrc_last_error() + "\\n");
    visfd_header_from_c(h, header);
    if (rescale)
      Rescale01(aaafMask);
    return;
  }
#endif
  Int len_in_file_name = in_file_name.size();"""

MRC_WRITE = """void MrcSimple::Write(string out_file_name) {
#ifdef VISFD_USE_CUDA
  {
    visfd_mrc_header h;
    visfd_header_to_c(header, h);
    if (visfd_mrc_write(out_file_name.c_str(), &h, &(aaafI[0][0][0])))
      throw MrcfileErr(string("Error: ") + visfd_mrc_last_error() + "\\n");
    header.dmin = h.dmin; header.dmax = h.dmax; header.dmean = h.dmean;   // FindMinMaxMean's side effect
    return;
  }
#endif
"""


def patch_mrc_simple(t):
    t = edit(t, '#include "mrc_simple.hpp"\n', MRC_BRIDGE, what="mrc bridge")
    t = edit(t, "                     float ***aaafMask) {\n  Int len_in_file_name = in_file_name.size();", MRC_READ,
             after="void MrcSimple::Read(string in_file_name,", what="MrcSimple::Read(file name)")
    t = edit(t, "void MrcSimple::Write(string out_file_name) {\n", MRC_WRITE, what="MrcSimple::Write(file name)")
    return t


def patch_alloc3d(t):
    """Alloc3D multiplies the three extents in `Integer` (int): 64-bit products instead (alloc3d.hpp:33-35, :60)."""
    t = edit(t, "sizeof(Entry*)  * (size[2]*size[1]) +", "sizeof(Entry*)  * ((size_t)size[2]*(size_t)size[1]) +", what="row table bytes")
    t = edit(t, "sizeof(Entry)   * (size[2]*size[1]*size[0])];",
             "sizeof(Entry)   * ((size_t)size[2]*(size_t)size[1]*(size_t)size[0])];", what="voxel bytes")
    t = edit(t, "sizeof(Entry)*(iz*size[0]*size[1]+\n                                                             iy*size[0]));",
             "sizeof(Entry)*((size_t)iz*(size_t)size[0]*(size_t)size[1]+\n"
             "                                                             (size_t)iy*(size_t)size[0]));", what="row offset")
    return t



def patch_feature_variants(t):
    return edit(t, "  BlobDogD(image_size,", "  VISFD_NS::BlobDogD(image_size,", what="BlobDogNM")


def main():
    if not os.path.isdir(FM):
        print("reference tree not found at %s: nothing built (the GPU box uses the prebuilt binary)" % REF)
        return 0
    os.makedirs(SRC, exist_ok=True)
    ms = os.path.join(REF, "lib", "mrc_simple")
    for folder, name, fn in ((FM, "handlers.cpp", patch_handlers), (FM, "feature_variants.hpp", patch_feature_variants),
                             (ms, "mrc_simple.cpp", patch_mrc_simple),
                             (os.path.join(REF, "lib", "visfd"), "alloc3d.hpp", patch_alloc3d)):
        with open(os.path.join(folder, name)) as f:
            text = f.read()
        with open(os.path.join(SRC, name), "w") as f:
            f.write(fn(text))
    inc = ["-I" + SRC, "-I" + FM] + ["-I" + os.path.join(REF, "lib", d) for d in
                                    ("visfd", "threshold", "mrc_simple", "random_gen", "eigen_simple")]
    inc += ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "visfd_b200", "csrc")]
    srcs = [os.path.join(SRC, "handlers.cpp")] + [os.path.join(FM, s) for s in
                                                  ("settings.cpp", "handlers_unsupported.cpp", "filter_mrc.cpp")]
    srcs += [os.path.join(SRC, "mrc_simple.cpp"), os.path.join(ms, "mrc_header.cpp")]
    exe = os.path.join(OUT, "filter_mrc_cuda")
    cmd = [CXX, "-std=c++17", "-O3", "-DNDEBUG", "-fopenmp", "-DVISFD_USE_CUDA"] + inc + srcs + [
        "-L" + os.path.join(ROOT, "visfd_b200"), "-lvisfd_cuda", "-Wl,-rpath,$ORIGIN/../../visfd_b200", "-o", exe, "-lm"]
    subprocess.check_call(cmd)
    print("built", exe)
    return 0


if __name__ == "__main__":
    sys.exit(main())
