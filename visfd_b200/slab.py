"""Z-slab multi-GPU driver of the membrane pipeline (one process per GPU).

Partition: rank r owns global planes [z0_r, z1_r).  It works on a SLAB = its own planes
widened by a halo of RAW SOURCE planes

    halo = tv_halfwidth + 1 + gauss_halfwidth        (clipped at the volume borders)

fetched ONCE from the ranks that own them (point-to-point send/recv: NCCL over NVLink on
the GPUs, gloo in the CPU tests).  With that halo every stage is slab-local:

    Gaussian (needs gauss_halfwidth planes)  ->  finite-difference Hessian (1 plane)
    ->  ridge saliency  ->  [global cut]  ->  voting (voters within tv_halfwidth planes)

and the only collective in the data path is the all-reduce of the 2048-bin radix-select
histograms (3 rounds) that makes the `-tv-best` cut a GLOBAL order statistic, exactly as
the reference's sort over all voxels (bin/filter_mrc/handlers.cpp:1751-1797).

The compute backend is injectable: on a GPU it is visfd_b200.Context (the C ABI); the
CPU tests (gloo, world_size 2) plug in a stand-in so that this file's plumbing -- the
plan, the halo exchange, the distributed select -- is tested without a GPU.
"""
from dataclasses import dataclass
import os

import numpy as np


def partition(nz, world):
    """Contiguous, near-equal plane ranges [z0, z1) per rank.  When there are at least 8 planes per rank
    the boundaries are multiples of 8 (the last rank takes the ragged tail): the voting kernel's 4x4x4
    receiver patches and the voter bricks (both 4^3; 8 planes = one layer of 8x8x4 receiver tiles twice over, a margin kept from the first voter-list kernels) then fall where they fall in the undivided volume, which makes
    the result independent of the number of ranks bit for bit."""
    unit = 8 if nz // 8 >= world else 1
    base, rem = divmod(nz // unit, world)
    out, z = [], 0
    for r in range(world):
        n = (base + (1 if r < rem else 0)) * unit
        out.append((z, nz if r == world - 1 else z + n))
        z += n
    return out


@dataclass
class SlabPlan:
    rank: int
    world: int
    nz: int
    own: tuple          # global [z0, z1)
    slab: tuple         # global [lo, hi) = own widened by halo, clipped, lo rounded down to a multiple of 8
    vote: tuple         # global voter range = own widened by tv_halfwidth, clipped
    halo: int
    recvs: list         # [(src_rank, global_lo, global_hi)]
    sends: list         # [(dst_rank, global_lo, global_hi)]

    @property
    def own_local(self):
        return (self.own[0] - self.slab[0], self.own[1] - self.slab[0])

    @property
    def vote_local(self):
        return (self.vote[0] - self.slab[0], self.vote[1] - self.slab[0])


def make_plan(nz, world, rank, gauss_hw, tv_hw):
    parts = partition(nz, world)
    halo = (tv_hw + 1 + gauss_hw) if tv_hw > 0 else (1 + gauss_hw)

    def widened(r, h):
        # The slab starts on a multiple of 8 planes: the voting stage orders its voters by 4^3 brick,
        # and with the slab's bricks coinciding with the whole volume's every receiver meets its
        # voters in the same order -- the multi-GPU result is then BIT-identical to the one-GPU one
        # (float sums depend on the order), at the price of up to 7 extra halo planes.
        return (max(0, parts[r][0] - h) // 8 * 8, min(nz, parts[r][1] + h))

    def reach(r, h):
        return (max(0, parts[r][0] - h), min(nz, parts[r][1] + h))

    def overlap(a, b):
        lo, hi = max(a[0], b[0]), min(a[1], b[1])
        return (lo, hi) if lo < hi else None

    slab = widened(rank, halo)
    recvs, sends = [], []
    for q in range(world):
        if q == rank:
            continue
        o = overlap(slab, parts[q])
        if o:
            recvs.append((q, o[0], o[1]))
        o = overlap(widened(q, halo), parts[rank])
        if o:
            sends.append((q, o[0], o[1]))
    return SlabPlan(rank, world, nz, parts[rank], slab, reach(rank, max(tv_hw, 0)), halo, recvs, sends)


def exchange_halo(plan, own_planes, slab_buf, dist=None):
    """Fill slab_buf (planes plan.slab) from this rank's own planes and its neighbours'.
    own_planes: tensor [own planes][ny][nx]; slab_buf: tensor [slab planes][ny][nx]."""
    lo = plan.slab[0]
    o0, o1 = plan.own
    slab_buf[o0 - lo:o1 - lo].copy_(own_planes)
    if plan.world == 1:
        return
    ops = []
    for (dst, a, b) in plan.sends:
        ops.append(dist.P2POp(dist.isend, own_planes[a - o0:b - o0], dst))
    for (src, a, b) in plan.recvs:
        ops.append(dist.P2POp(dist.irecv, slab_buf[a - lo:b - lo], src))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def distributed_cut_threshold(backend, saliency_own, fraction, mask_own=None, dist=None, world=1, device=None):
    """k-th largest un-masked saliency over ALL ranks, k = floor(n * fraction) with the
    reference's float product (handlers.cpp:1779-1782): radix select over all-reduced
    histograms.  backend needs select_hist / select_step / key_to_float."""
    import torch
    prefix, bits, rank_k, first = 0, 0, 0, True
    while bits < 32:
        hist = backend.select_hist(saliency_own, prefix, bits, mask=mask_own)
        if world > 1:
            # uint64 counters travel as int64 (counts < 2^63)
            t = torch.from_numpy(hist.astype(np.int64))
            if device is not None:
                t = t.to(device)
            dist.all_reduce(t)
            hist = t.cpu().numpy().astype(np.uint64)
        if first:
            total = int(hist.sum())
            if total == 0:
                raise RuntimeError("saliency cut: no un-masked voxels")
            k = int(np.floor(np.float32(total) * np.float32(fraction)))
            rank_k = min(max(k, 0), total - 1)
            first = False
        prefix, bits, rank_k = backend.select_step(hist, prefix, bits, rank_k)
    return backend.key_to_float(prefix)


def distributed_mean_stddev(backend, own, weights_own=None, dist=None, world=1, device=None):
    """AverageArr / StdDevArr (lib/visfd/visfd_utils.hpp:685-790) of a volume spread over the ranks: two
    all-reduces of (sum, weight) pairs in double.  Needed by `-cl` (SelectIntensityRangeGauss,
    lib/threshold/threshold.hpp:248-258), whose mean and standard deviation are those of the whole image."""
    import torch

    def reduced(center, squared):
        s = backend.moment_sums(own, weights_own, center, squared)
        if world > 1:
            t = torch.tensor(s, dtype=torch.float64, device=device)
            dist.all_reduce(t)
            s = (float(t[0].item()), float(t[1].item()))
        return s

    s0, s1 = reduced(0.0, False)
    mean = float(np.float32(s0 / s1))
    q0, q1 = reduced(mean, True)
    return mean, float(np.float32(np.sqrt(q0 / q1)))


class _Trace:
    """VISFD_SLAB_TRACE=1: wall-clock per phase of SlabMembrane.run (with device syncs), to stderr."""

    def __init__(self, device):
        import time
        import torch
        self.time, self.torch, self.device = time, torch, device
        self.marks = []
        if device is not None:
            torch.cuda.synchronize(device)
        self.t = time.perf_counter()

    def mark(self, name):
        if self.device is not None:
            self.torch.cuda.synchronize(self.device)
        t = self.time.perf_counter()
        self.marks.append((name, (t - self.t) * 1e3))
        self.t = t

    def report(self, rank, stage_ms):
        import sys
        print("slab trace rank %d: " % rank + ", ".join("%s %.1f ms" % m for m in self.marks) +
              " | kernels: " + ", ".join("%s %.1f" % kv for kv in stage_ms.items()), file=sys.stderr)


class SlabMembrane:
    """HandleTV's pipeline (handlers.cpp:1618-1892) on this rank's slab."""

    def __init__(self, backend, shape, params, rank=0, world=1, dist=None, device=None):
        import torch
        self.backend, self.params, self.dist, self.device = backend, params, dist, device
        self.nz, self.ny, self.nx = shape
        self.gauss_hw = int(np.floor(np.float32(params.sigma) * np.float32(params.truncate_ratio)))
        self.tv_hw = backend.tv_halfwidth(params.tv_sigma, params.tv_cutoff_ratio) if params.tv_sigma > 0 else 0
        self.plan = make_plan(self.nz, world, rank, self.gauss_hw, self.tv_hw)
        n_slab = self.plan.slab[1] - self.plan.slab[0]
        kw = dict(dtype=torch.float32, device=device)
        self.slab_src = None if world == 1 else torch.empty((n_slab, self.ny, self.nx), **kw)
        self.smoothed = torch.empty((n_slab, self.ny, self.nx), **kw)
        self.saliency = torch.empty((n_slab, self.ny, self.nx), **kw)
        self.threshold = None
        self.stage_ms = {}

    def run(self, own_src, out=None, want_tensor=False, out_host=None):
        """own_src: this rank's planes (device tensor).  Returns (out, tensor) for them.
        out_host (optional, host array/tensor of the own planes): also receives the result,
        copied chunk by chunk behind the voting kernels."""
        p, plan, be = self.params, self.plan, self.backend
        if hasattr(be, "reset_stage_ms"):
            be.reset_stage_ms()
        self.stage_ms = {}
        trace = _Trace(self.device) if os.environ.get("VISFD_SLAB_TRACE") == "1" else None
        if plan.world == 1:
            src = own_src                      # the slab IS the volume: no copy
        else:
            if self.slab_src is None:
                import torch
                self.slab_src = torch.empty_like(self.smoothed)
            exchange_halo(plan, own_src, self.slab_src, self.dist)
            src = self.slab_src
        if trace:
            trace.mark("halo")
        be.ridge_saliency_slab(src, plan.slab[0], self.nz, p.sigma, p.truncate_ratio,
                               order=p.eival_order, smoothed=self.smoothed, saliency=self.saliency)
        if trace:
            trace.mark("gauss+ridge")
        o0, o1 = plan.own_local
        if p.cut_is_fraction:
            thr = distributed_cut_threshold(be, self.saliency[o0:o1], p.cut, dist=self.dist, world=plan.world,
                                            device=self.device)
        else:
            thr = p.cut
        self.threshold = thr
        if trace:
            trace.mark("cut")
        res = be.vote_slab(self.saliency, self.smoothed, plan.slab[0], self.nz, plan.own_local, plan.vote_local,
                           thr, p, want_tensor=want_tensor, out=out, **({"out_host": out_host} if out_host is not None else {}))
        self._grab("gauss", "ridge", "select", "compact", "tv")
        if trace:
            trace.mark("vote")
            trace.report(plan.rank, self.stage_ms)
        return res

    def _grab(self, *names):
        sm = getattr(self.backend, "stage_ms", None)
        if sm is None:
            return
        for n in names:
            v = sm(n)
            if v >= 0:
                self.stage_ms[n] = self.stage_ms.get(n, 0.0) + v


# ---- scale-space blob detection over z-slabs ------------------------------------------
def log_halfwidth(sigma, delta, truncate_ratio):
    """Shared half-width of the two Gaussians of ApplyLog (lib/visfd/filter3d.hpp:1455-1466), with the
    library's arithmetic (log_params, gauss.cu): the wider sigma in double, rounded to float, times the
    float ratio."""
    s_b = np.float32(np.float64(np.float32(sigma)) * (1.0 + 0.5 * np.float64(np.float32(delta))))
    return int(np.floor(np.float32(truncate_ratio) * s_b))


class SlabBlobs:
    """BlobDog (lib/visfd/feature.hpp:56-427) on this rank's z-slab.  Halo = largest LoG half-width + 1
    planes of raw source (fetched like the membrane pipeline's); every rank scans its own planes; the
    candidate lists are all-gathered and concatenated per scale in rank (= z) order, which is the
    reference's (scale, z, y, x) order; the best scores are all-reduced (min / max) before the final
    ratio filter (feature.hpp:341-344, :369-372)."""

    def __init__(self, backend, shape, sigmas, delta=0.02, truncate_ratio=2.5, rank=0, world=1, dist=None,
                 device=None):
        self.backend, self.dist, self.device = backend, dist, device
        self.nz, self.ny, self.nx = shape
        self.sigmas = np.ascontiguousarray(np.asarray(sigmas), np.float32)
        self.delta, self.truncate_ratio = delta, truncate_ratio
        hw = max(log_halfwidth(s, delta, truncate_ratio) for s in self.sigmas)
        # make_plan's halo is tv_hw + 1 + gauss_hw: (hw, tv 0) gives hw + 1
        self.plan = make_plan(self.nz, world, rank, hw, 0)
        self.slab_src = None

    def run(self, own_src, minima_threshold=np.inf, maxima_threshold=-np.inf, use_threshold_ratios=False,
            capacity=1 << 20):
        import torch
        plan, be = self.plan, self.backend
        if plan.world == 1:
            src = own_src
        else:
            if self.slab_src is None:
                self.slab_src = torch.empty((plan.slab[1] - plan.slab[0], self.ny, self.nx), dtype=torch.float32,
                                            device=self.device)
            exchange_halo(plan, own_src, self.slab_src, self.dist)
            src = self.slab_src
        kw = dict(minima_threshold=minima_threshold, maxima_threshold=maxima_threshold,
                  use_threshold_ratios=use_threshold_ratios)
        mins, maxs, best = be.blob_dog_slab(src, plan.slab[0], self.nz, plan.own_local, self.sigmas, self.delta,
                                            self.truncate_ratio, capacity=capacity, **kw)
        if plan.world > 1:
            parts = [None] * plan.world
            self.dist.all_gather_object(parts, (mins, maxs))
            t = torch.tensor([best[0], -best[1]], dtype=torch.float32, device=self.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
            best = (float(t[0].item()), -float(t[1].item()))
            mins = _concat_by_scale([p[0] for p in parts], self.sigmas)
            maxs = _concat_by_scale([p[1] for p in parts], self.sigmas)
        return be.blob_finalize(mins, maxs, best, **kw)


def _concat_by_scale(lists, sigmas):
    """Per scale (in the order of `sigmas`), the ranks' rows in rank order."""
    out = []
    for s in sigmas:
        for rows in lists:
            rows = np.asarray(rows, np.float32).reshape(-1, 5)
            out.append(rows[rows[:, 3] == np.float32(s)])
    return np.concatenate(out, axis=0) if out else np.zeros((0, 5), np.float32)
