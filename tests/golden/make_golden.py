"""Generates tests/golden/host_logic.npz by EXECUTING THE REFERENCE'S OWN PYTHON
(/root/reference/triro/ray/ray_optix.py, unmodified) in this container.

The reference's native backend needs the OptiX SDK and a GPU, neither of which exists here, so
its module `triro.backend.ops` is replaced by a stub that answers the five trace calls with the
oracle's binary32 mirror on CPU tensors, and Tensor.cuda() is patched to the identity.  What the
fixture therefore pins is the reference's HOST LOGIC — tuple orders, dtypes, stream compaction
(ray_optix.py:142-144), intersects_id (:191-223) and the contains_points decision procedure with
its retry / all-False quirks (:231-279) — on fixed seeded inputs.  The trace results themselves
come from the oracle (see oracle/raymesh_oracle.c for why those are "parity unpinned").

Run from the repo root (only where /root/reference exists):  python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200"))
REF = "/root/reference/triro/ray/ray_optix.py"

from oracle import oracle  # noqa: E402
from triro import synth  # noqa: E402  (mesh generators only; no product code path is used)


def load_reference_class(mesh_holder):
    """Import the reference's ray_optix.py with stubbed dependencies."""
    hops = types.ModuleType("triro.backend.ops")

    def _oi(accel):
        return accel.oracle

    def t(a, dtype=None):
        x = torch.from_numpy(np.ascontiguousarray(a))
        return x if dtype is None else x.to(dtype)

    def intersects_any(accel, o, d):
        return t(_oi(accel).intersects_any(o.numpy(), d.numpy()))

    def intersects_first(accel, o, d):
        return t(_oi(accel).intersects_first(o.numpy(), d.numpy()))

    def intersects_closest(accel, o, d):
        hit, front, tri, loc, uv = _oi(accel).intersects_closest(o.numpy(), d.numpy())
        return t(hit), t(front), t(tri), t(loc), t(uv)

    def intersects_count(accel, o, d):
        return t(_oi(accel).intersects_count(o.numpy(), d.numpy()))

    def intersects_location(accel, o, d):
        loc, ri, ti, _, _ = _oi(accel).intersects_location(o.numpy(), d.numpy())
        return t(loc), t(ri), t(ti)

    class _Module:
        class OptixAccelStructureWrapperCPP:
            def buildAccelStructure(self, v, f):
                self.oracle = oracle.OracleIntersector(v.numpy(), f.numpy(), mode=oracle.MIRROR)

            def freeAccelStructure(self):
                pass

    hops.get_module = lambda: _Module
    for fn in (intersects_any, intersects_first, intersects_closest, intersects_count, intersects_location):
        setattr(hops, fn.__name__, fn)

    # the reference passes its Python wrapper to hops.*; give the stub access to the oracle through it
    def unwrap(fn):
        return lambda accel, o, d: fn(accel._inner, o, d)

    for name in ("intersects_any", "intersects_first", "intersects_closest", "intersects_count", "intersects_location"):
        setattr(hops, name, unwrap(getattr(hops, name)))

    pkg = types.ModuleType("triro"); pkg.__path__ = []
    backend = types.ModuleType("triro.backend"); backend.__path__ = []
    backend.ops = hops
    sys.modules.update({"triro_ref_stub": pkg})
    saved = {k: sys.modules.get(k) for k in ("triro", "triro.backend", "triro.backend.ops", "trimesh", "jaxtyping")}
    sys.modules["triro"] = pkg
    sys.modules["triro.backend"] = backend
    sys.modules["triro.backend.ops"] = hops
    sys.modules["trimesh"] = types.ModuleType("trimesh")
    if "jaxtyping" not in sys.modules:
        try:
            import jaxtyping  # noqa: F401
        except Exception:
            jt = types.ModuleType("jaxtyping")

            class _Ann:
                def __class_getitem__(cls, item):
                    return cls

            jt.Float32 = jt.Int32 = jt.Bool = _Ann
            sys.modules["jaxtyping"] = jt
    spec = importlib.util.spec_from_file_location("reference_ray_optix", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    return mod.RayMeshIntersector


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self          # no GPU here: .cuda() is the identity
    RefRMI = load_reference_class(None)
    out = {}

    def rays(n, seed, scale):
        o, d = synth.random_rays(n, seed=seed, box=True)
        return (o * scale).contiguous(), d.contiguous()

    # ---- closest + compaction + id on an icosphere (reference ray_optix.py:117-146, :191-223)
    v, f = synth.icosphere(2)
    r = RefRMI(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
    o, d = rays(3000, 21, 2.0)
    o3, d3 = o.reshape(10, 300, 3), d.reshape(10, 300, 3)
    dense = r.intersects_closest(o3, d3)
    comp = r.intersects_closest(o3, d3, stream_compaction=True)
    out.update(ico_v=v, ico_f=f, ico_o=o3.numpy(), ico_d=d3.numpy())
    for i, name in enumerate(("hit", "front", "tri", "loc", "uv")):
        out[f"ico_dense_{name}"] = dense[i].numpy()
    for i, name in enumerate(("hit", "front", "ray", "tri", "loc", "uv")):
        out[f"ico_comp_{name}"] = comp[i].numpy()
    a = r.intersects_id(o3, d3, return_locations=True, multiple_hits=False)
    out.update(ico_id1_tri=a[0].numpy(), ico_id1_ray=a[1].numpy(), ico_id1_loc=a[2].numpy())
    b = r.intersects_id(o3, d3, return_locations=True, multiple_hits=True)
    out.update(ico_idm_tri=b[0].numpy(), ico_idm_ray=b[1].numpy(), ico_idm_loc=b[2].numpy())
    out["ico_any"] = r.intersects_any(o3, d3).numpy()
    out["ico_first"] = r.intersects_first(o3, d3).numpy()
    out["ico_count"] = r.intersects_count(o3, d3).numpy()
    out["ico_aabb_lo"] = r.mesh_aabb[0].numpy(); out["ico_aabb_hi"] = r.mesh_aabb[1].numpy()

    # ---- contains_points truth table on the unit cube (reference ray_optix.py:231-279)
    cv, cf = synth.cube(0.5)
    rc = RefRMI(vertices=torch.from_numpy(cv), faces=torch.from_numpy(cf))
    g = torch.Generator().manual_seed(5)
    pts = (torch.rand((400, 3), generator=g) * 2 - 1) * 0.8           # inside and outside the cube
    out.update(cube_v=cv, cube_f=cf, cube_pts=pts.numpy())
    torch.manual_seed(1234)
    out["cube_default"] = rc.contains_points(pts).numpy()
    inside_only = pts[(pts.abs() < 0.45).all(dim=1)]
    out["cube_inside_pts"] = inside_only.numpy()
    xdir = torch.tensor([1.0, 0.0, 0.0])
    out["cube_inside_xdir"] = rc.contains_points(inside_only, xdir).numpy()
    out["cube_mixed_xdir"] = rc.contains_points(pts, xdir).numpy()            # quirk: all False
    far = torch.tensor([[5.0, 5.0, 5.0], [0.5, 0.5, 0.5], [-3.0, 0.0, 0.0]])
    out["cube_far_pts"] = far.numpy()
    out["cube_far"] = rc.contains_points(far).numpy()
    # sphere: K7 and a seeded cloud
    sv, sf = synth.icosphere(3)
    rs = RefRMI(vertices=torch.from_numpy(sv), faces=torch.from_numpy(sf))
    sp = (torch.rand((600, 3), generator=g) * 2 - 1) * 1.05
    sp[0] = torch.tensor([0.0, 0.0, 0.999])
    out.update(sph_v=sv, sph_f=sf, sph_pts=sp.numpy())
    torch.manual_seed(99)
    out["sph_default"] = rs.contains_points(sp).numpy()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_logic.npz"), **out)
    print("wrote host_logic.npz with", len(out), "arrays;",
          "cube_default inside:", int(out["cube_default"].sum()), "sph_default inside:", int(out["sph_default"].sum()),
          "ico hits:", int(out["ico_dense_hit"].sum()))


if __name__ == "__main__":
    main()
