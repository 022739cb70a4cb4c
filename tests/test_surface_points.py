"""The oriented point cloud of `-normals-file` (bin/filter_mrc/handlers.cpp:2039-2309).
CPU: the restatement (oracle/visfd_oracle.cpp::vo_surface_points) against the PLY files the STOCK filter_mrc binary
wrote (tests/golden/surface_points.npz, made by tests/golden/make_surface_golden.py) -- the reference's own test
(58 vertices) and a synthetic tomogram (3457).  GPU: visfd_cuda_surface_points against the restatement on the same
inputs, and the whole chain from the raw volume against the stock binary's file."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SQ2 = float(np.float32(np.sqrt(2.0)))


@pytest.fixture(scope="module")
def sp_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "surface_points.npz"))


def ply_err(rows, want):
    """column-wise relative error with a floor of 1e-3 of the column's range (the file holds 6 significant digits)"""
    rows, want = np.asarray(rows, np.float64), np.asarray(want, np.float64)
    assert rows.shape == want.shape, (rows.shape, want.shape)
    den = np.maximum(np.abs(want), 1e-3 * np.abs(want).max(axis=0))
    return float((np.abs(rows - want) / den).max())


def c1_inputs(oracle, golden):
    sigma, ratio, tv_sigma, expo, cutoff, frac = [float(v) for v in golden["c1_params"]]
    m = oracle.membrane(golden["c1_in_binned"], sigma, ratio, 1, frac, True, tv_sigma, int(expo), cutoff, want_tensor=True)
    _, n, dire = oracle.label_connected(m["out"], m["tensor"], 1e9, 30.0, order=1, want_direction=True)
    assert n == 1
    return m["out"], dire, golden["c1_connect_labels"]


def synthetic_params():
    sigma = np.float32(np.float32(3.4641) / np.sqrt(3.0))                # settings.cpp: thickness / sqrt(3), -w 1
    tv_sigma = np.float32(np.float32(2.5) * sigma)
    ratio = np.float32(np.sqrt(np.float32(-2) * np.log(np.float32(0.03))))
    return float(sigma), float(ratio), float(tv_sigma)


def test_restatement_reproduces_the_stock_binary_c1(oracle, golden, sp_golden):
    sal, dire, labels = c1_inputs(oracle, golden)
    vw = float(np.float32(19.2) * np.float32(2))
    rows, n = oracle.surface_points(sal, dire, labels=labels, select_cluster=1, voxel_width=(vw, vw, vw))
    assert n == 58
    assert ply_err(rows, sp_golden["c1_ply"]) <= 2e-5


def test_restatement_reproduces_the_stock_binary_synthetic(oracle, sp_golden):
    sigma, ratio, tv_sigma = synthetic_params()
    m = oracle.membrane(sp_golden["s_vol"], sigma, ratio, 1, 0.05, True, tv_sigma, 4, SQ2, want_tensor=True)
    thr = float(sp_golden["s_threshold"])
    lab, n, dire = oracle.label_connected(m["out"], m["tensor"], thr, 30.0, order=1, want_direction=True)
    labels = sp_golden["s_labels"].astype(np.float32)
    assert np.array_equal(lab == 1, labels == 1)
    rows, n = oracle.surface_points(m["out"], dire, labels=labels, select_cluster=1)
    assert n == len(sp_golden["s_ply"]) == 3457
    assert ply_err(rows, sp_golden["s_ply"]) <= 2e-5


def test_restatement_variants(oracle, golden):
    """no curve step, no ridge step, a mask, every voxel (no clustering), capacity"""
    sal, dire, labels = c1_inputs(oracle, golden)
    n_sel = int((labels == 1).sum())
    rows, n = oracle.surface_points(sal, dire, labels=labels, curve_ds=0.0, find_ridge=False)
    assert n == n_sel
    idx = np.argwhere(labels == 1)[:, ::-1].astype(np.float32)          # raster order, (x, y, z) in voxels
    assert np.array_equal(rows[:, :3], idx)
    mask = np.ones_like(sal)
    mask[:, :, :4] = 0
    rows_m, n_m = oracle.surface_points(sal, dire, labels=labels, mask=mask, curve_ds=0.0, find_ridge=False)
    assert n_m == int(((labels == 1) & (mask != 0)).sum())
    rows_a, n_a = oracle.surface_points(sal, dire, labels=None, voxel_width=(2.0, 3.0, 4.0))
    assert n_a == sal.size and np.array_equal(rows_a[:, 3:], dire.reshape(-1, 3))
    assert np.array_equal(rows_a[1, :3], [2.0, 0.0, 0.0])
    rows_c, n_c = oracle.surface_points(sal, dire, labels=labels, capacity=10)
    assert n_c == 58 and len(rows_c) == 10


@pytest.mark.gpu
def test_gpu_matches_the_restatement(ctx, oracle, golden, sp_golden):
    vw = float(np.float32(19.2) * np.float32(2))
    sal, dire, labels = c1_inputs(oracle, golden)
    for kw in (dict(voxel_width=(vw, vw, vw)), dict(curve_ds=0.0), dict(find_ridge=False), dict(max_distance=0.0),
               dict(curve_ds=0.35, max_distance=0.6)):
        want, n_want = oracle.surface_points(sal, dire, labels=labels, **kw)
        got, n_got = ctx.surface_points(sal, dire, labels=labels, **kw)
        assert n_got == n_want, kw
        assert ply_err(got, want) <= 2e-5, kw
    # the synthetic case on the stock binary's inputs (restated bit for bit by the oracle), host and device arrays
    sigma, ratio, tv_sigma = synthetic_params()
    m = oracle.membrane(sp_golden["s_vol"], sigma, ratio, 1, 0.05, True, tv_sigma, 4, SQ2, want_tensor=True)
    _, _, dire = oracle.label_connected(m["out"], m["tensor"], float(sp_golden["s_threshold"]), 30.0, order=1,
                                        want_direction=True)
    labels = sp_golden["s_labels"].astype(np.float32)
    got, n_got = ctx.surface_points(m["out"], dire, labels=labels)
    assert n_got == 3457
    assert ply_err(got, sp_golden["s_ply"]) <= 2e-5
    import torch
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (m["out"], dire, labels)]
    got_d, n_d = ctx.surface_points(dev[0], dev[1], labels=dev[2])
    assert n_d == n_got and np.array_equal(got_d, got)
    mask = np.ones_like(labels)
    mask[:, :, 20:24] = 0
    want, n_want = oracle.surface_points(m["out"], dire, labels=labels, mask=mask)
    got, n_got = ctx.surface_points(m["out"], dire, labels=labels, mask=mask)
    assert n_got == n_want and ply_err(got, want) <= 2e-5
    got, n_got = ctx.surface_points(None, dire, labels=None, voxel_width=(1.5, 1.0, 2.0), capacity=1000)
    assert n_got == labels.size and len(got) == 1000
    assert np.array_equal(got[:, 3:], dire.reshape(-1, 3)[:1000])


@pytest.mark.gpu
def test_gpu_chain_from_the_raw_volume(ctx, sp_golden):
    """volume -> GPU membrane pipeline -> GPU LabelConnected -> GPU point cloud = the stock binary's PLY file"""
    sigma, ratio, tv_sigma = synthetic_params()
    res = ctx.membrane(sp_golden["s_vol"], sigma, ratio, 1, 0.05, True, tv_sigma, 4, SQ2, want_tensor=True)
    thr = float(sp_golden["s_threshold"])
    lc = ctx.label_connected(res["out"], res["tensor"], thr, 30.0, want_direction=True)
    labels = np.where(lc["labels"] < 0, 0, lc["labels"]).astype(np.float32)
    assert np.array_equal(labels == 1, sp_golden["s_labels"] == 1)
    got, n = ctx.surface_points(res["out"], lc["direction"], labels=labels)
    want = sp_golden["s_ply"]
    # the inputs carry the 1e-5 of the voting kernel, so a handful of discrete decisions (the sample next to the mean,
    # the 1.3-voxel cut) may differ: compare the bulk
    assert abs(n - len(want)) <= 0.01 * len(want)
    if n == len(want):
        den = np.maximum(np.abs(want), 1e-3 * np.abs(want).max(axis=0))
        bad = (np.abs(got.astype(np.float64) - want) / den).max(axis=1) > 1e-3
        assert bad.mean() <= 0.01, bad.mean()


# ---- `-membrane-background` (handlers.cpp:1577-1592): lives here because the stock binary's answer for it was
# made on the same synthetic volume ------------------------------------------------------------------------------
def test_restatement_background_subtraction_is_the_stock_binary(oracle, sp_golden):
    sigma, ratio, tv_sigma = synthetic_params()
    m = oracle.membrane(sp_golden["s_vol"], sigma, ratio, 1, 0.05, True, tv_sigma, 4, SQ2, background_sigma=6.0)
    assert np.array_equal(m["out"], sp_golden["s_background6_out"])


@pytest.mark.gpu
def test_gpu_background_subtraction(ctx, oracle, sp_golden):
    from util import vote_score_err, rel_err, TOL_SALIENCY
    sigma, ratio, tv_sigma = synthetic_params()
    vol = sp_golden["s_vol"]
    want = oracle.membrane(vol, sigma, ratio, 1, 0.05, True, tv_sigma, 4, SQ2, background_sigma=6.0, want_tensor=True)
    got = ctx.membrane(vol, sigma, ratio, 1, 0.05, True, tv_sigma, 4, SQ2, background_sigma=6.0, want_saliency=True)
    assert got["threshold"] == want["threshold"] or abs(got["threshold"] - want["threshold"]) <= 1e-5 * abs(want["threshold"])
    assert rel_err(got["hess_saliency"], want["hess_saliency"]) <= TOL_SALIENCY
    # the post-vote score times the peak height: judged like the score itself, against the tensor's trace times |peak|
    peak = np.abs(sp_golden["s_background6_out"]).max()
    assert np.abs(got["out"] - sp_golden["s_background6_out"]).max() <= 2e-4 * peak
    # no voting, masked, device arrays
    mask = np.ones_like(vol)
    mask[:, :10, :] = 0
    want = oracle.membrane(vol, sigma, ratio, 1, 0.05, True, 0.0, 4, SQ2, mask=mask, background_sigma=4.0)
    import torch
    got = ctx.membrane(torch.from_numpy(vol).cuda(), sigma, ratio, 1, 0.05, True, 0.0, 4, SQ2,
                       mask=torch.from_numpy(mask).cuda(), background_sigma=4.0)
    assert rel_err(got["out"].cpu().numpy(), want["out"]) <= TOL_SALIENCY
    assert np.all(got["out"].cpu().numpy()[mask == 0] == 0)
