"""Synthetic float32 tomograms: N(0,1) noise plus dark membranes
(spherical shells, depth -3, Gaussian profile of width 2 voxels) and optional dark
blobs.  Every z-plane is generated from its own seed (seed, global z), so a Z-slab
generated on one rank is bit-identical to the same planes of the whole volume
generated elsewhere -- on the host with numpy or on the GPU with torch.
"""
import numpy as np


def _shell_centres(shape, n_shells):
    nz, ny, nx = shape
    r0 = 0.3 * min(shape)
    if n_shells <= 0:
        return []
    out = [((nz - 1) / 2.0, (ny - 1) / 2.0, (nx - 1) / 2.0, r0)]
    rng = np.random.default_rng(12345)
    for _ in range(n_shells - 1):
        out.append((rng.uniform(0, nz), rng.uniform(0, ny), rng.uniform(0, nx), r0 * rng.uniform(0.4, 0.9)))
    return out


def tomogram(shape, seed=0, z0=0, z1=None, n_shells=1, blobs=0, noise=1.0, blob_sigma=(3.0, 6.0)):
    """Planes [z0, z1) of the synthetic tomogram of the given full shape (numpy, host)."""
    nz, ny, nx = shape
    z1 = nz if z1 is None else z1
    out = np.empty((z1 - z0, ny, nx), np.float32)
    y = np.arange(ny, dtype=np.float32)[:, None]
    x = np.arange(nx, dtype=np.float32)[None, :]
    shells = _shell_centres(shape, n_shells)
    brng = np.random.default_rng(seed + 7919)
    blob_list = [(brng.uniform(0, nz), brng.uniform(0, ny), brng.uniform(0, nx), brng.uniform(*blob_sigma))
                 for _ in range(blobs)]
    for z in range(z0, z1):
        rng = np.random.default_rng([seed, z])
        a = rng.standard_normal((ny, nx), dtype=np.float32) * np.float32(noise)
        for (cz, cy, cx, R) in shells:
            r = np.sqrt((x - cx) ** 2 + (y - cy) ** 2 + np.float32((z - cz) ** 2))
            a -= 3.0 * np.exp(-0.5 * ((r - R) / 2.0) ** 2)
        for (bz, by, bx, bs) in blob_list:
            if abs(z - bz) < 5 * bs:
                r2 = (x - bx) ** 2 + (y - by) ** 2 + np.float32((z - bz) ** 2)
                a -= 4.0 * np.exp(-0.5 * r2 / (bs * bs))
        out[z - z0] = a
    return out


def tomogram_torch(shape, device, seed=0, z0=0, z1=None, n_shells=1, out=None):
    """Same construction on the GPU with torch (different noise stream than numpy, but
    the same plane-seeded consistency between slabs)."""
    import torch
    nz, ny, nx = shape
    z1 = nz if z1 is None else z1
    if out is None:
        out = torch.empty((z1 - z0, ny, nx), dtype=torch.float32, device=device)
    y = torch.arange(ny, dtype=torch.float32, device=device)[:, None]
    x = torch.arange(nx, dtype=torch.float32, device=device)[None, :]
    shells = _shell_centres(shape, n_shells)
    g = torch.Generator(device=device)
    for z in range(z0, z1):
        g.manual_seed(seed * 1000003 + z)
        a = out[z - z0]
        a.normal_(generator=g)
        for (cz, cy, cx, R) in shells:
            r = torch.sqrt((x - cx) ** 2 + (y - cy) ** 2 + float((z - cz) ** 2))
            a -= 3.0 * torch.exp(-0.5 * ((r - R) / 2.0) ** 2)
    return out
