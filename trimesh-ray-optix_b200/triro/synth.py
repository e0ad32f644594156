"""Synthetic meshes and ray sets of the BASELINE.json configurations.

trimesh is not available offline, so the icosphere is generated here with the same
construction `trimesh.creation.icosphere` uses (a unit icosahedron subdivided 4-to-1 with the
new vertices pushed to the sphere); the reference demos use it at test/test.py:15 and
README.md:31.  Everything is deterministic; meshes are numpy arrays (float32 vertices,
int32 faces), ray sets are torch tensors so they can be generated on the GPU.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = [
    "icosphere",
    "heightfield",
    "triangle_soup",
    "cube",
    "readme_rays",
    "pinhole_rays",
    "gen_rays",
    "random_rays",
]


def icosphere(subdivisions: int = 3, radius: float = 1.0):
    """Unit icosahedron subdivided `subdivisions` times -> 20*4**s faces, outward CCW winding."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array(
        [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
         [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array(
        [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
         [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
         [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(subdivisions):
        nv = len(v)
        # unique edges -> midpoint vertex ids
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        e.sort(axis=1)
        key = e[:, 0] * nv + e[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        mid = (v[uniq // nv] + v[uniq % nv]) * 0.5
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid], axis=0)
        nf = len(f)
        m01 = nv + inv[:nf]
        m12 = nv + inv[nf:2 * nf]
        m20 = nv + inv[2 * nf:]
        f = np.concatenate([
            np.stack([f[:, 0], m01, m20], axis=1),
            np.stack([m01, f[:, 1], m12], axis=1),
            np.stack([m20, m12, f[:, 2]], axis=1),
            np.stack([m01, m12, m20], axis=1)], axis=0)
    return (v * radius).astype(np.float32), f.astype(np.int32)


def heightfield(nx: int = 2048, ny: int = 1024, amplitude: float = 0.1):
    """Closed-form terrain over [-1,1]^2 on an nx x ny cell grid -> 2*nx*ny triangles (config 3)."""
    xs = np.linspace(-1.0, 1.0, nx + 1, dtype=np.float64)
    ys = np.linspace(-1.0, 1.0, ny + 1, dtype=np.float64)
    x, y = np.meshgrid(xs, ys, indexing="xy")        # [ny+1, nx+1]
    z = np.zeros_like(x)
    for k in range(4):                                # four octaves, no RNG
        fr = 2.0 ** k * math.pi
        z += (0.5 ** k) * (np.sin(fr * x + 0.3 * k) * np.cos(fr * y * 1.3 - 0.7 * k))
    z *= amplitude / 1.875
    v = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    i = np.arange(ny, dtype=np.int64)[:, None] * (nx + 1) + np.arange(nx, dtype=np.int64)[None, :]
    a, b, c, d = i, i + 1, i + (nx + 1), i + (nx + 2)
    f = np.concatenate([np.stack([a, b, d], axis=-1).reshape(-1, 3), np.stack([a, d, c], axis=-1).reshape(-1, 3)], axis=0)
    return v, f.astype(np.int32)


def triangle_soup(n: int = 1_000_000, sigma: float = 0.003, seed: int = 7):
    """n unconnected triangles: centres U([-1,1]^3), vertex offsets N(0, sigma^2) (config 4)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1.0, 1.0, size=(n, 1, 3))
    v = (c + rng.normal(0.0, sigma, size=(n, 3, 3))).reshape(-1, 3).astype(np.float32)
    f = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
    return v, f


def cube(half: float = 0.5):
    """Closed axis-aligned cube, outward CCW winding, 12 triangles (contains_points cases)."""
    h = half
    v = np.array([[-h, -h, -h], [h, -h, -h], [h, h, -h], [-h, h, -h], [-h, -h, h], [h, -h, h], [h, h, h], [-h, h, h]],
                 dtype=np.float32)
    f = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [2, 3, 7], [2, 7, 6],
                  [1, 2, 6], [1, 6, 5], [3, 0, 4], [3, 4, 7]], dtype=np.int32)
    return v, f


def readme_rays(n: int = 800, device="cpu"):
    """The README quick-start grid (README.md:35-39): d = (x, y, -1) unnormalised, origin (0,0,3)
    as a stride-0 broadcast.  Returns (origins [n,n,3] broadcast view, directions [n,n,3])."""
    import torch

    y, x = torch.meshgrid([torch.linspace(1, -1, n), torch.linspace(-1, 1, n)], indexing="ij")
    z = -torch.ones_like(x)
    d = torch.stack([x, y, z], dim=-1).to(device)
    o = torch.tensor([0.0, 0.0, 3.0], device=device).broadcast_to(d.shape)
    return o, d


def pinhole_rays(w: int = 3840, h: int = 2160, device="cpu", origin=(0.0, 0.0, 3.0)):
    """Pinhole camera rays after test/performance_test.py:10-20 (gen_rays) with identity rotation:
    d = normalize(x-(w-1)/2, y-(h-1)/2, -f), f = w*25/36.  Returns (origins broadcast, directions)."""
    import torch

    f = w * 25.0 / 36.0
    y, x = torch.meshgrid([torch.linspace(0, h - 1, h, device=device), torch.linspace(0, w - 1, w, device=device)],
                          indexing="ij")
    x = x - (w - 1) / 2
    y = y - (h - 1) / 2
    z = -torch.ones_like(x) * f
    d = torch.stack([x, y, z], dim=-1)
    d = d / torch.norm(d, dim=-1, keepdim=True)
    o = torch.tensor(origin, dtype=torch.float32, device=device).broadcast_to(d.shape)
    return o, d.contiguous()


def gen_rays(cam_mat, w: int, h: int, f: float, device="cpu"):
    """The reference's gen_rays (test/performance_test.py:10-20) verbatim in behaviour: [h, w, 3] directions."""
    import torch

    y, x = torch.meshgrid([torch.linspace(0, h - 1, h), torch.linspace(0, w - 1, w)], indexing="ij")
    x = x - (w - 1) / 2
    y = y - (h - 1) / 2
    z = -torch.ones_like(x) * f
    dirs = torch.stack([x, y, z], dim=-1).to(device)
    dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
    return dirs @ torch.transpose(torch.as_tensor(cam_mat, dtype=torch.float32, device=device), 0, 1)


def random_rays(n: int, seed: int = 1234, device="cpu", zlo: float = 0.3, zhi: float = 0.5, box: bool = False):
    """Incoherent rays (configs 3-5): origins U([-1,1]^2 x [zlo,zhi]) (or U([-1,1]^3) with box=True),
    directions normalised N(0, I)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    o = torch.rand((n, 3), generator=g, device=device, dtype=torch.float32) * 2.0 - 1.0
    if not box:
        o[:, 2] = (o[:, 2] + 1.0) * 0.5 * (zhi - zlo) + zlo
    d = torch.randn((n, 3), generator=g, device=device, dtype=torch.float32)
    d = d / torch.norm(d, dim=-1, keepdim=True)
    return o, d
