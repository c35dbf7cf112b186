"""Synthetic building-shaped voxel clouds (the benchmark workload).

"B470" is the building SURVEY.md section 8(d) / BASELINE.md define: a 542 x 542 x 68
voxel shell (21.68 m x 21.68 m x 2.72 m at 4 cm voxels = 470 m^2 floor) made of a
floor, a ceiling, eight walls perpendicular to y and eight perpendicular to x,
concatenated in that order with the duplicates at the intersections kept
(1,177,224 input rows, 1,155,656 unique voxels) and shuffled with numpy seed 0.
"""
import numpy as np


def building_coords(nx=542, ny=542, nz=68, n_walls=8, seed=0, batch_index=0, shuffle=True):
    """int64 [N,4] rows (x, y, z, batch) of a hollow building shell."""
    xs, ys, zs = np.arange(nx), np.arange(ny), np.arange(nz)
    parts = []
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    for z in (0, nz - 1):  # floor, ceiling
        parts.append(np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, z)], 1))
    wx, wz = np.meshgrid(xs, zs, indexing="ij")
    for y in np.linspace(0, ny - 1, n_walls).astype(int):  # walls perpendicular to y
        parts.append(np.stack([wx.ravel(), np.full(wx.size, y), wz.ravel()], 1))
    wy, wz = np.meshgrid(ys, zs, indexing="ij")
    for x in np.linspace(0, nx - 1, n_walls).astype(int):  # walls perpendicular to x
        parts.append(np.stack([np.full(wy.size, x), wy.ravel(), wz.ravel()], 1))
    c = np.concatenate(parts, 0).astype(np.int64)
    if shuffle:
        rs = np.random.RandomState(seed)
        c = c[rs.permutation(c.shape[0])]
    return np.concatenate([c, np.full((c.shape[0], 1), batch_index, np.int64)], 1)


def b470(seed=0):
    """The BASELINE.json building: coords int64 [1177224,4], feats f32 [1177224,9]."""
    import torch

    coords = building_coords(seed=seed)
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(coords.shape[0], 9, generator=g)
    return coords, feats.numpy()


def jittered_building(seed, jitter=0.15):
    """Config-4 style buildings: B470 generator with X, Y jittered by +-15 %."""
    rs = np.random.RandomState(1000 + seed)
    nx = int(round(542 * (1 + rs.uniform(-jitter, jitter))))
    ny = int(round(542 * (1 + rs.uniform(-jitter, jitter))))
    return building_coords(nx=nx, ny=ny, seed=seed)


def small_building(nx=40, ny=36, nz=12, n_walls=3, seed=0, batch_index=0):
    """A miniature of the same shape for parity tests that a CPU checker finishes in seconds."""
    return building_coords(nx, ny, nz, n_walls, seed, batch_index)
