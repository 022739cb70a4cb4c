#!/usr/bin/env python
"""The Z-slab membrane pipeline over NCCL against the same volume on ONE GPU, bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29517 tests/multi_gpu_check.py

Every rank builds the same synthetic volume, runs `SlabMembrane` (raw-source halo exchange by NCCL
send/recv, all-reduced radix-select histograms, slab voting) on its own planes with the C4 parameter
set (halo 28 planes), and rank 0 compares the gathered result with `visfd_cuda_membrane` on the
whole volume.  Repeated a few times, with device work still in flight on torch's stream when the
pipeline is entered, to catch stream-ordering mistakes.  Run by tests/test_gpu_multi.py when the box
has more than one GPU."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import visfd_b200
    from visfd_b200 import synth
    from visfd_b200.capi import MembraneParams
    from visfd_b200.slab import SlabMembrane, partition

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    shapes = [(max(256, 40 * world), 96, 136), (max(250, 40 * world + 2), 90, 131), (max(203, 40 * world + 3), 64, 77)]
    sigma = float(np.float32(np.float32(5.196) / np.sqrt(3.0)))
    tv_sigma = float(np.float32(np.float32(4.733) * np.float32(sigma)))
    ratio = float(np.sqrt(-2.0 * np.log(0.03)))
    sq2 = float(np.float32(np.sqrt(2.0)))
    params = MembraneParams(sigma, ratio, visfd_b200.DECREASING_EIVALS, 0.05, 1, tv_sigma, 4, sq2)
    ctx = visfd_b200.Context(local)
    ok = True
    for trial, shape in enumerate(shapes):
        vol = synth.tomogram(shape, seed=40 + trial, n_shells=2)
        o0, o1 = partition(shape[0], world)[rank]
        pipe = SlabMembrane(ctx, shape, params, rank=rank, world=world, dist=dist, device=dev)
        # the upload and a scaling round trip are still queued on torch's stream when run() starts
        own = torch.from_numpy(vol[o0:o1]).to(dev, non_blocking=True)
        own = (own * 2.0) * 0.5
        res, _ = pipe.run(own)
        parts = [None] * world
        dist.all_gather_object(parts, res.cpu().numpy())
        thr = [None] * world
        dist.all_gather_object(thr, float(pipe.threshold))
        if rank == 0:
            whole = ctx.membrane(torch.from_numpy(vol).to(dev), sigma, ratio, visfd_b200.DECREASING_EIVALS, 0.05, True,
                                 tv_sigma, 4, sq2)
            want = whole["out"].cpu().numpy()
            got = np.concatenate(parts, axis=0)
            same_thr = all(np.float32(t) == np.float32(whole["threshold"]) for t in thr)
            err = float(np.abs(got.astype(np.float64) - want).max() / want.max())
            same = np.array_equal(got, want) and same_thr
            print(f"  thresholds equal: {same_thr}; max |diff| / max = {err:.3g}", flush=True)
            print(f"trial {trial}: world {world}, shape {shape}, threshold {whole['threshold']:.6g}, "
                  f"max {want.max():.6g}: {'identical' if same else 'DIFFERENT'}", flush=True)
            if not same:
                bad = np.argwhere(got != want)
                print("  first differing voxels (z,y,x):", bad[:5].tolist(), "count", len(bad), flush=True)
            ok = ok and same
    # ---- scale-space blob detection over the same ranks (all-gathered lists, all-reduced best scores) ----
    from visfd_b200.slab import SlabBlobs
    bshape = (max(192, 48 * world), 80, 96)
    sigmas = (1.5 * 1.25 ** np.arange(6)).astype(np.float32)
    bvol = synth.tomogram(bshape, seed=50, n_shells=0, blobs=60, blob_sigma=(1.5, 4.0))
    o0, o1 = partition(bshape[0], world)[rank]
    for kw in (dict(minima_threshold=0.3, maxima_threshold=0.3, use_threshold_ratios=True),
               dict(minima_threshold=0.0, maxima_threshold=-np.inf, use_threshold_ratios=False)):
        blobs = SlabBlobs(ctx, bshape, sigmas, 0.02, ratio, rank=rank, world=world, dist=dist, device=dev)
        own = torch.from_numpy(bvol[o0:o1]).to(dev, non_blocking=True)
        got = blobs.run(own, **kw)
        if rank == 0:
            want = ctx.blob_dog(torch.from_numpy(bvol).to(dev), sigmas, 0.02, ratio, **kw)
            same = np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
            print(f"blobs ({'ratio' if kw['use_threshold_ratios'] else 'absolute'} thresholds): world {world}, shape {bshape}, "
                  f"{len(want[0])} minima, {len(want[1])} maxima: {'identical' if same else 'DIFFERENT'}", flush=True)
            ok = ok and same and len(want[0]) > 0
    # ---- mean / standard deviation of the whole image from per-rank sums (-cl) ----------------------------
    from visfd_b200.slab import distributed_mean_stddev
    own = torch.from_numpy(bvol[o0:o1]).to(dev)
    mean, std = distributed_mean_stddev(ctx, own, None, dist=dist, world=world, device=dev)
    if rank == 0:
        m1, s1 = ctx.mean_stddev(torch.from_numpy(bvol).to(dev))
        same = abs(mean - m1) <= 1e-6 * abs(m1) + 1e-9 and abs(std - s1) <= 1e-6 * s1
        print(f"mean/stddev over {world} ranks: {mean:.7g} {std:.7g} (one GPU {m1:.7g} {s1:.7g}): "
              f"{'equal' if same else 'DIFFERENT'}", flush=True)
        ok = ok and same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK " + ("OK" if ok else "FAILED"), flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
