"""Multi-rank plumbing of the Z-slab driver (visfd_b200/slab.py) on CPU: world_size 2 and 3
over gloo.  The compute backend is a stand-in built on the CPU oracle (allowed here: this
is test infrastructure), so what is under test is the slab plan, the point-to-point halo
exchange (including halos that reach past the nearest neighbour) and the distributed
radix select; the result must equal the single-process oracle bit for bit."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from visfd_b200.slab import partition, make_plan  # noqa: E402

SQ2 = float(np.float32(np.sqrt(2.0)))


def test_partition_and_plan():
    assert partition(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert partition(1024, 8)[3] == (384, 512)
    # C4 on 8 GPUs: gauss hw 7, tv hw 20 -> halo 28
    plans = [make_plan(1024, 8, r, 7, 20) for r in range(8)]
    assert plans[0].slab == (0, 156) and plans[0].vote == (0, 148) and plans[0].own_local == (0, 128)
    assert plans[3].slab == (352, 512 + 28) and plans[3].vote == (364, 532) and plans[3].vote_local == (12, 180)   # 384 - 28 = 356 -> 352: brick aligned
    assert plans[7].slab == (864, 1024)
    for r, p in enumerate(plans):
        # every send has its matching receive
        for (dst, a, b) in p.sends:
            assert (r, a, b) in plans[dst].recvs
        got = sorted([(a, b) for (_, a, b) in p.recvs] + [p.own])
        assert got[0][0] == p.slab[0] and got[-1][1] == p.slab[1]
        assert all(got[i][1] == got[i + 1][0] for i in range(len(got) - 1))
    # halo wider than a neighbour's whole slab: data comes from two ranks away
    p0 = make_plan(15, 3, 0, 2, 3)
    assert p0.slab == (0, 11) and {q for (q, _, _) in p0.recvs} == {1, 2}
    # no voting: only the Gaussian + stencil halo
    assert make_plan(100, 2, 1, 7, 0).halo == 8


class OracleBackend:
    """CPU stand-in for visfd_b200.Context built on the oracle (tests only)."""

    def __init__(self):
        from oracle.pyoracle import Oracle
        import visfd_b200
        self.o = Oracle("port")
        self.lib = visfd_b200.load_library()   # host-only helpers work without a GPU
        self.vb = visfd_b200

    def tv_halfwidth(self, s, r):
        return self.vb.tv_halfwidth(s, r)

    def ridge_saliency_slab(self, src, z_offset, nz_global, sigma, ratio, order=1, score_kind=0, mask=None,
                            smoothed=None, saliency=None):
        import torch
        a = src.numpy()
        hw = int(np.floor(np.float32(sigma) * np.float32(ratio)))
        sm, _ = self.o.apply_gauss(a, sigma, hw)
        _, h = self.o.calc_hessian(a, sigma, ratio)
        sal, dire, _ = self.o.hessian_eigen_score(h, order=order, score_kind=score_kind)
        smoothed.copy_(torch.from_numpy(sm))
        saliency.copy_(torch.from_numpy(sal))
        self.dire = dire
        return smoothed, saliency

    def select_hist(self, sal, prefix, bits, mask=None):
        v = np.ascontiguousarray(sal.numpy()).reshape(-1)
        b = v.view(np.uint32)
        keys = np.where(b & 0x80000000, ~b, b | 0x80000000).astype(np.uint32)
        nb = 11 if 32 - bits >= 11 else 32 - bits
        sel = keys if bits == 0 else keys[(keys >> np.uint32(32 - bits)) == prefix]
        idx = (sel >> np.uint32(32 - bits - nb)) & np.uint32((1 << nb) - 1)
        hist = np.zeros(2048, np.uint64)
        hist[:1 << nb] = np.bincount(idx, minlength=1 << nb)
        return hist

    def select_step(self, hist, prefix, bits, rank):
        import ctypes as C
        p, b, r = C.c_uint32(prefix), C.c_int(bits), C.c_uint64(rank)
        h = np.ascontiguousarray(hist, np.uint64)
        assert self.lib.visfd_cuda_select_step(h.ctypes.data_as(C.c_void_p), C.byref(p), C.byref(b), C.byref(r)) == 0
        return p.value, b.value, r.value

    def key_to_float(self, key):
        import ctypes as C
        return self.lib.visfd_cuda_key_to_float(C.c_uint32(key))

    def vote_slab(self, saliency, smoothed, z_offset, nz_global, own, vote, thr, p, mask=None, want_tensor=False,
                  out=None):
        import torch
        sal = saliency.numpy()[vote[0]:vote[1]].copy()
        sal[sal < np.float32(thr)] = 0
        dire = self.dire[vote[0]:vote[1]]
        t = self.o.tv_dense_stick(sal, dire, p.tv_sigma, p.tv_exponent, p.tv_cutoff_ratio)
        sc = self.o.tensor_score(t, order=p.eival_order, score_kind=0)
        a, b = own[0] - vote[0], own[1] - vote[0]
        return torch.from_numpy(sc[a:b].copy()), torch.from_numpy(t[a:b].copy())


    def blob_dog_slab(self, src, z_offset, nz_global, own, sigmas, delta, truncate_ratio, mask=None, capacity=0,
                      minima_threshold=np.inf, maxima_threshold=-np.inf, use_threshold_ratios=True):
        """the oracle on the slab, unfiltered; candidates of the own planes only, z as image plane.  A slab
        end inside the image is `image border` to the oracle, which only discards candidates on the slab's
        first / last plane -- never own planes, thanks to the halo."""
        a = src.numpy()
        kw = dict(minima_threshold=minima_threshold, maxima_threshold=maxima_threshold, use_threshold_ratios=False)
        if use_threshold_ratios:   # ratio thresholds are applied by blob_finalize
            kw = dict(minima_threshold=0.0, maxima_threshold=0.0, use_threshold_ratios=False)
        mn, mx = self.o.blob_dog(a, sigmas, delta, truncate_ratio, **kw)
        out = []
        for rows in (mn, mx):
            rows = rows[(rows[:, 2] >= own[0]) & (rows[:, 2] < own[1])].copy()
            rows[:, 2] += z_offset
            out.append(rows)
        best = (min([1.0] + out[0][:, 4].tolist()), max([-1.0] + out[1][:, 4].tolist()))
        return out[0], out[1], best

    def blob_finalize(self, mins, maxs, best, **kw):
        return self.vb.capi.blob_finalize(mins, maxs, best, lib=self.lib, **kw)


def _blob_worker(rank, world, port, shape, seed, sigmas, kw, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from visfd_b200 import synth
        from visfd_b200.slab import SlabBlobs
        pipe = SlabBlobs(OracleBackend(), shape, sigmas, 0.02, 2.6482, rank=rank, world=world, dist=dist, device="cpu")
        z0, z1 = pipe.plan.own
        own = torch.from_numpy(synth.tomogram(shape, seed=seed, z0=z0, z1=z1, n_shells=0, blobs=12, blob_sigma=(1.0, 2.5)))
        ret[rank] = pipe.run(own, **kw)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_blobs_match_single_process(world):
    """SlabBlobs over gloo: halo exchange, per-slab scan, all-gather of the lists, all-reduce of the best
    scores, final ratio filter == BlobDog on the whole volume, rows and order included."""
    import torch.multiprocessing as mp
    from oracle.pyoracle import Oracle
    from visfd_b200 import synth
    shape, seed = (30, 18, 20), 9
    sigmas = 1.0 * 1.3 ** np.arange(5)
    vol = synth.tomogram(shape, seed=seed, n_shells=0, blobs=12, blob_sigma=(1.0, 2.5))
    for kw in (dict(minima_threshold=0.4, maxima_threshold=0.4, use_threshold_ratios=True),
               dict(minima_threshold=0.0, maxima_threshold=-np.inf, use_threshold_ratios=False)):
        port = 29700 + world + os.getpid() % 200
        ret = mp.Manager().dict()
        mp.spawn(_blob_worker, args=(world, port, shape, seed, sigmas, kw, ret), nprocs=world, join=True)
        want = Oracle("port").blob_dog(vol, sigmas, 0.02, 2.6482, **kw)
        assert len(want[0]) > 0
        for r in range(world):                      # every rank ends up with the whole list
            assert np.array_equal(ret[r][0], want[0]) and np.array_equal(ret[r][1], want[1])


def _worker(rank, world, port, shape, seed, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from visfd_b200 import synth, MembraneParams
        from visfd_b200.slab import SlabMembrane
        p = MembraneParams(1.0, 2.6482, 1, 0.12, 1, 2.4, 4, SQ2)
        be = OracleBackend()
        pipe = SlabMembrane(be, shape, p, rank=rank, world=world, dist=dist, device="cpu")
        z0, z1 = pipe.plan.own
        own = torch.from_numpy(synth.tomogram(shape, seed=seed, z0=z0, z1=z1))
        out, tensor = pipe.run(own, want_tensor=True)
        ret[rank] = (z0, z1, out.numpy().copy(), float(pipe.threshold))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (24, 14, 16)), (3, (15, 12, 14))])
def test_slab_pipeline_matches_single_process(world, shape):
    import torch.multiprocessing as mp
    from oracle.pyoracle import Oracle
    from visfd_b200 import synth
    seed = 5
    port = 29600 + world + os.getpid() % 200
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, shape, seed, ret), nprocs=world, join=True)
    vol = synth.tomogram(shape, seed=seed)
    want = Oracle("port").membrane(vol, 1.0, 2.6482, 1, 0.12, True, 2.4, 4, SQ2, want_tensor=False)
    got = np.zeros(shape, np.float32)
    for r in range(world):
        z0, z1, o, thr = ret[r]
        got[z0:z1] = o
        assert np.float32(thr) == np.float32(want["threshold"])       # the cut is global
    assert np.array_equal(got, want["out"])


def test_plans_cover_every_plane_once():
    """host-only invariants of the z-slab plans for many (planes, ranks, radii): the own ranges tile
    [0, nz) in order, rank boundaries and slab starts are multiples of 8 whenever there are 8 planes per
    rank, a slab holds its own planes plus the halo (clipped at the image), the voter range lies inside
    the slab, and every receive has the matching send on the other rank."""
    rng = np.random.default_rng(11)
    for _ in range(300):
        world = int(rng.integers(1, 9))
        nz = int(rng.integers(world, 700))
        ghw, thw = int(rng.integers(0, 12)), int(rng.integers(0, 30))
        plans = [make_plan(nz, world, r, ghw, thw) for r in range(world)]
        halo = (thw + 1 + ghw) if thw > 0 else (1 + ghw)
        z = 0
        for r, p in enumerate(plans):
            assert p.own[0] == z and p.own[1] >= p.own[0]
            z = p.own[1]
            assert p.halo == halo
            assert p.slab[0] <= max(0, p.own[0] - halo) and p.slab[1] == min(nz, p.own[1] + halo)
            assert p.slab[0] >= max(0, p.own[0] - halo - 7)
            if nz // 8 >= world:
                assert p.own[0] % 8 == 0 and p.slab[0] % 8 == 0
            assert p.slab[0] <= p.vote[0] <= p.own[0] and p.own[1] <= p.vote[1] <= p.slab[1]
            got = sorted([(a, b) for (_, a, b) in p.recvs] + ([p.own] if p.own[1] > p.own[0] else []))
            if got:
                assert got[0][0] == p.slab[0] or p.own[1] == p.own[0]
                assert all(got[i][1] == got[i + 1][0] for i in range(len(got) - 1))
            for (dst, a, b) in p.sends:
                assert (r, a, b) in plans[dst].recvs
            for (src, a, b) in p.recvs:
                assert (r, a, b) in plans[src].sends
        assert z == nz
