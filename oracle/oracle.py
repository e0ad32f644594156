"""CPU oracle for the ray/mesh hot path — TEST INFRASTRUCTURE, never imported by the product.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference legs) may
import this module.  PARITY STATUS: "parity unpinned" beyond the reference's hand-derivable
known answers K1-K7 (the reference's arithmetic is inside the closed OptiX runtime and its
tests hold no assertions; see raymesh_oracle.c and DESIGN.md).

Two layers:
  * `query(...)`: numpy front end of oracle/raymesh_oracle.c (TRUTH = binary64 with a grazing
    classifier, MIRROR = binary32 with the product's exact operation sequence);
  * `OracleIntersector`: restatement in numpy of the reference's host logic around the trace
    calls — strided ray fetch (triro/backend/shaders.cu:27-63), stream compaction
    (triro/ray/ray_optix.py:142-144), clamp + cumsum + per-ray packing of all hits
    (triro/backend/ray.cpp:333-342, shaders.cu:226-246), intersects_id (ray_optix.py:191-223),
    contains_points (ray_optix.py:231-279).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TRUTH, MIRROR = 0, 1
GZ_EDGE_CLOSEST, GZ_EDGE_ANY, GZ_TIE, GZ_TNEAR, GZ_EDGEON = 1, 2, 4, 8, 16
TMAX = 1.0e7          # shaders.cu:86
MAX_ANYHIT_SIZE = 8   # LaunchParams.h:8
MAX_LIST = 64


def build(force: bool = False) -> str:
    """Compile liboracle.so next to this file (gcc + OpenMP)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "raymesh_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O2", "-fopenmp", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-o", so, src, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_bvh_build.restype = C.c_void_p
        _LIB.oracle_bvh_build.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        _LIB.oracle_bvh_free.argtypes = [C.c_void_p]
        _LIB.oracle_query.restype = C.c_int
        _LIB.oracle_query.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int64, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 8 + \
                                     [C.c_int] + [C.c_void_p] * 4
        _LIB.oracle_num_threads.restype = C.c_int
        _LIB.oracle_set_num_threads.argtypes = [C.c_int]
    return _LIB


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of later query() calls (bench.py: all host cores, whatever OMP_NUM_THREADS a launcher set)."""
    lib().oracle_set_num_threads(int(n))
    return num_threads()


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleMesh:
    """Mesh + optional binned-SAH BVH2 used only to cull candidates (same answers as brute force)."""

    def __init__(self, vertices, faces, use_bvh: bool | None = None):
        self.vertices = np.ascontiguousarray(np.asarray(vertices, dtype=np.float32).reshape(-1, 3))
        self.faces = np.ascontiguousarray(np.asarray(faces, dtype=np.int32).reshape(-1, 3))
        if use_bvh is None:
            use_bvh = len(self.faces) > 64
        self._bvh = None
        if use_bvh:
            self._bvh = lib().oracle_bvh_build(_p(self.vertices), len(self.vertices), _p(self.faces), len(self.faces))

    def __del__(self):
        if getattr(self, "_bvh", None):
            lib().oracle_bvh_free(self._bvh)
            self._bvh = None


def query(mesh: OracleMesh, origins, directions, mode: int = TRUTH, closest_only: bool = False, list_cap: int = 0,
          want=("hit", "front", "tri", "loc", "uv", "t", "count", "flags"), brute: bool = False):
    """Runs the oracle over flat [n,3] float32 rays; returns a dict of numpy arrays."""
    o = np.ascontiguousarray(np.asarray(origins, dtype=np.float32).reshape(-1, 3))
    d = np.ascontiguousarray(np.asarray(directions, dtype=np.float32).reshape(-1, 3))
    assert o.shape == d.shape
    n = len(o)
    out = {}
    def mk(name, shape, dt):
        if name in want:
            out[name] = np.zeros(shape, dtype=dt)
            return out[name]
        return None
    hit = mk("hit", (n,), np.uint8); front = mk("front", (n,), np.uint8); tri = mk("tri", (n,), np.int32)
    loc = mk("loc", (n, 3), np.float32); uv = mk("uv", (n, 2), np.float32); t = mk("t", (n,), np.float64)
    count = mk("count", (n,), np.int32); flags = mk("flags", (n,), np.uint32)
    ln = lt = ltt = ll = None
    if list_cap > 0:
        assert list_cap <= MAX_LIST
        ln = np.zeros((n,), np.int32); lt = np.full((n, list_cap), -1, np.int32)
        ltt = np.full((n, list_cap), np.inf, np.float64); ll = np.zeros((n, list_cap, 3), np.float32)
        out.update(list_n=ln, list_tri=lt, list_t=ltt, list_loc=ll)
    rc = lib().oracle_query(_p(mesh.vertices), len(mesh.vertices), _p(mesh.faces), len(mesh.faces),
                            None if brute else mesh._bvh, int(mode), int(bool(closest_only)), n, _p(o), _p(d), TMAX,
                            _p(hit), _p(front), _p(tri), _p(loc), _p(uv), _p(t), _p(count), _p(flags),
                            int(list_cap), _p(ln), _p(lt), _p(ltt), _p(ll))
    if rc != 0:
        raise RuntimeError(f"oracle_query failed: {rc}")
    return out


def fetch_rays(origins, directions):
    """Strided ray fetch of the reference (shaders.cu:27-63 with ray.cpp:151-159,177-179):
    the batch shape is that of `origins`; both tensors are read through their own strides.
    numpy/torch views already encode strides, so the restatement is: flatten in row-major order."""
    o = np.asarray(origins)
    d = np.asarray(directions)
    assert o.shape[-1] == 3 and d.shape == o.shape, "origins and directions must both be [*b, 3]"
    return o.shape[:-1], np.ascontiguousarray(o.reshape(-1, 3), dtype=np.float32), np.ascontiguousarray(
        d.reshape(-1, 3), dtype=np.float32)


class OracleIntersector:
    """numpy restatement of triro.ray.ray_optix.RayMeshIntersector (reference ray_optix.py:18-279)."""

    def __init__(self, vertices, faces, mode: int = TRUTH, use_bvh: bool | None = None):
        self.mesh = OracleMesh(vertices, faces, use_bvh)
        self.mode = mode
        v = self.mesh.vertices
        # ray_optix.py:43-46 — min/max over all vertices
        self.mesh_aabb = (v.min(axis=0), v.max(axis=0)) if len(v) else (np.zeros(3, np.float32), np.zeros(3, np.float32))

    # -- ops.py:84-192 ------------------------------------------------------------
    def intersects_any(self, origins, directions):
        b, o, d = fetch_rays(origins, directions)
        return query(self.mesh, o, d, self.mode, want=("count",))["count"].reshape(b) > 0

    def intersects_first(self, origins, directions):
        b, o, d = fetch_rays(origins, directions)
        return query(self.mesh, o, d, self.mode, closest_only=True, want=("tri",))["tri"].reshape(b)

    def intersects_count(self, origins, directions):
        b, o, d = fetch_rays(origins, directions)
        return query(self.mesh, o, d, self.mode, want=("count",))["count"].reshape(b)

    def closest_raw(self, origins, directions):
        b, o, d = fetch_rays(origins, directions)
        r = query(self.mesh, o, d, self.mode, want=("hit", "front", "tri", "loc", "uv", "t", "flags", "count"))
        return b, r

    def intersects_closest(self, origins, directions, stream_compaction: bool = False):
        b, r = self.closest_raw(origins, directions)
        hit = r["hit"].astype(bool).reshape(b)
        front = r["front"].astype(bool).reshape(b)
        tri = r["tri"].reshape(b)
        loc = r["loc"].reshape(*b, 3)
        uv = r["uv"].reshape(*b, 2)
        if stream_compaction:
            # ray_optix.py:142-144
            ray_idx = np.arange(hit.size, dtype=np.int32)[hit.reshape(-1)]
            return hit, front[hit], ray_idx, tri[hit], loc[hit], uv[hit]
        return hit, front, tri, loc, uv

    def intersects_location(self, origins, directions, sort_by_t: bool = True):
        """ray.cpp:324-378: per ray min(count, 8) entries, rays ascending.  The reference's order
        within a ray is BVH traversal order (unspecified); the oracle lists by ascending t."""
        b, o, d = fetch_rays(origins, directions)
        r = query(self.mesh, o, d, self.mode, list_cap=MAX_LIST, want=("count",))
        n = len(o)
        cnt = np.minimum(r["count"], MAX_ANYHIT_SIZE)          # ray.cpp:334-335
        off = np.concatenate([[0], np.cumsum(cnt)])             # ray.cpp:336-342
        nh = int(off[-1])
        loc = np.zeros((nh, 3), np.float32); ri = np.zeros((nh,), np.int32); ti = np.zeros((nh,), np.int32)
        for i in np.nonzero(cnt)[0]:
            k = cnt[i]
            loc[off[i]:off[i] + k] = r["list_loc"][i, :k]
            ti[off[i]:off[i] + k] = r["list_tri"][i, :k]
            ri[off[i]:off[i] + k] = i
        return loc, ri, ti, r["count"], r

    def intersects_id(self, origins, directions, return_locations=False, multiple_hits=True):
        if multiple_hits:                                        # ray_optix.py:207-214
            loc, ri, ti, _, _ = self.intersects_location(origins, directions)
            return (ti, ri, loc) if return_locations else (ti, ri)
        hit, _, ri, ti, loc, _ = self.intersects_closest(origins, directions, stream_compaction=True)
        return (ti, ri, loc) if return_locations else (ti, ri)   # ray_optix.py:215-223

    # -- ray_optix.py:231-279 -----------------------------------------------------
    DEFAULT_DIRECTION = (0.4395064455, 0.617598629942, 0.652231566745)

    def contains_core(self, points, direction):
        """Deterministic core: (inside_aabb, count(+dir), count(-dir)) for [n,3] points."""
        p = np.ascontiguousarray(np.asarray(points, dtype=np.float32).reshape(-1, 3))
        d = np.tile(np.asarray(direction, dtype=np.float32).reshape(1, 3), (len(p), 1))
        lo, hi = self.mesh_aabb
        inside = ~((~(p > lo)).any(axis=1) | (~(p < hi)).any(axis=1))       # :238-241
        cp = query(self.mesh, p, d, self.mode, want=("count",))["count"]    # :254-260, all points
        cm = query(self.mesh, p, -d, self.mode, want=("count",))["count"]
        return inside, cp, cm

    def contains_points(self, points, check_direction=None):
        import torch  # the reference draws the retry direction from torch's CPU RNG (:273)

        p = np.asarray(points, dtype=np.float32)
        contains = np.zeros(p.shape[:-1], dtype=bool)                        # :236
        direction = self.DEFAULT_DIRECTION if check_direction is None else np.asarray(check_direction, np.float32)
        lo, hi = self.mesh_aabb
        inside_pre = ~((~(p > lo)).any(axis=1) | (~(p < hi)).any(axis=1))
        if not inside_pre.any():                                             # :243-244
            return contains
        inside, cp, cm = self.contains_core(p, direction)
        agree = (cp % 2 == 1) & (cm % 2 == 1)                                # :263-264
        contain = inside & agree                                             # :265 ((inside & agree & mod2[0]) == 1)
        broken = ~agree & ((cp == 0) | (cm == 0))                            # :267
        if not broken.any():                                                 # :269-270
            return contain
        if check_direction is None:                                          # :272-277
            new_direction = (torch.rand(3) - 0.5).numpy()
            contains = contain.copy()
            contains[broken] = self.contains_points(p[broken], new_direction)
        return contains                                                      # :279 (all False when a direction was given)
