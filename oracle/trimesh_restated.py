"""numpy restatement of trimesh's pure-Python ray path (`trimesh.ray.ray_triangle`) — TEST / BASELINE
INFRASTRUCTURE, never imported by the product.

trimesh is the library the reference mirrors (README.md:3 of the reference) and benchmarks against
(`mesh.ray.intersects_location(..., multiple_hits=False)`, test/performance_test.py:75, with the PyEmbree backend:
1.27 Mrays/s on an i5-13490F per the reference README).  Neither trimesh nor embree can be installed offline, so this
file restates — FROM MEMORY of trimesh 4.x, flagged as such — the algorithm of its dependency-free backend:

    candidates (r-tree broad phase)  ->  ray/plane intersection  ->  barycentric containment test with a
    1e-12-ish tolerance  ->  keep hits in front of the origin  ->  per ray: all hits or the nearest one

The r-tree broad phase is replaced by testing every (ray, triangle) pair in chunks (no rtree offline), which is
what trimesh degenerates to for a small mesh.  It serves as the "trimesh-class" CPU figure for BASELINE config 1
(320 / 1 280 triangles) next to the OpenMP BVH port used everywhere else.
"""
from __future__ import annotations

import numpy as np

TOL_ZERO = 1e-12      # trimesh.constants.tol.zero


def ray_triangle_id(vertices, faces, ray_origins, ray_directions, multiple_hits=True, chunk=4096):
    """-> (index_tri, index_ray, locations), float64, like trimesh.ray.ray_triangle.ray_triangle_id."""
    v = np.asarray(vertices, dtype=np.float64)
    tri = v[np.asarray(faces)]                                     # [T,3,3]
    o_all = np.asarray(ray_origins, dtype=np.float64).reshape(-1, 3)
    d_all = np.asarray(ray_directions, dtype=np.float64).reshape(-1, 3)
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    e1, e2 = b - a, c - a
    normal = np.cross(e1, e2)                                      # plane normals (unnormalised)
    d00 = (e1 * e1).sum(1); d01 = (e1 * e2).sum(1); d11 = (e2 * e2).sum(1)
    denom = d00 * d11 - d01 * d01
    out_tri, out_ray, out_loc = [], [], []
    for s in range(0, len(o_all), chunk):
        o, d = o_all[s:s + chunk], d_all[s:s + chunk]              # [R,3]
        # plane/line intersection: t = dot(a - o, n) / dot(d, n)
        dn = d @ normal.T                                          # [R,T]
        num = np.einsum("tk,tk->t", a, normal)[None, :] - o @ normal.T
        with np.errstate(divide="ignore", invalid="ignore"):
            t = num / dn
        valid = np.abs(dn) > TOL_ZERO
        p = o[:, None, :] + t[:, :, None] * d[:, None, :]          # [R,T,3] candidate locations
        # barycentric coordinates of p in each triangle (trimesh.triangles.points_to_barycentric, 'cross'-free form)
        w = p - a[None]
        d20 = np.einsum("rtk,tk->rt", w, e1); d21 = np.einsum("rtk,tk->rt", w, e2)
        with np.errstate(divide="ignore", invalid="ignore"):
            bv = (d11[None] * d20 - d01[None] * d21) / denom[None]
            bw = (d00[None] * d21 - d01[None] * d20) / denom[None]
        bu = 1.0 - bv - bw
        inside = (bu > -TOL_ZERO) & (bv > -TOL_ZERO) & (bw > -TOL_ZERO)
        hit = valid & inside & (t > TOL_ZERO) & np.isfinite(t)     # in front of the origin
        if multiple_hits:
            r_idx, t_idx = np.nonzero(hit)
        else:
            tt = np.where(hit, t, np.inf)
            t_idx = tt.argmin(axis=1)
            r_idx = np.nonzero(np.isfinite(tt[np.arange(len(o)), t_idx]))[0]
            t_idx = t_idx[r_idx]
        out_tri.append(t_idx); out_ray.append(r_idx + s); out_loc.append(p[r_idx, t_idx])
    return (np.concatenate(out_tri).astype(np.int64), np.concatenate(out_ray).astype(np.int64),
            np.concatenate(out_loc).reshape(-1, 3))


def intersects_location(vertices, faces, ray_origins, ray_directions, multiple_hits=True):
    """trimesh.ray.intersects_location: (locations, index_ray, index_tri)."""
    t, r, l = ray_triangle_id(vertices, faces, ray_origins, ray_directions, multiple_hits)
    return l, r, t
