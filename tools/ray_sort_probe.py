"""Upper bound of ray re-ordering: times the trace kernels on incoherent batches whose rays were pre-sorted (for free)
by Morton code of the origin and/or direction octant.  Measured: at most x1.10 (octant-major), so no sort pass."""
import os, sys
sys.path.insert(0, "trimesh-ray-optix_b200"); sys.path.insert(0, ".")
import torch
from triro import synth
from triro.ray.ray_optix import RayMeshIntersector
dev = torch.device("cuda:0")
def part1by2(x):
    x = x & 0x3ff
    x = (x | (x << 16)) & 0x30000ff
    x = (x | (x << 8)) & 0x300f00f
    x = (x | (x << 4)) & 0x30c30c3
    x = (x | (x << 2)) & 0x9249249
    return x
def keys(o, d, lo, hi, obits=10):
    q = ((o - lo) / (hi - lo)).clamp(0, 0.999999)
    qi = (q * (1 << obits)).long()
    m = part1by2(qi[:, 0]) | (part1by2(qi[:, 1]) << 1) | (part1by2(qi[:, 2]) << 2)
    octant = ((d[:, 0] < 0).long()) | ((d[:, 1] < 0).long() << 1) | ((d[:, 2] < 0).long() << 2)
    return m, octant
def timeit(fn, reps=5):
    ts = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    return min(ts)
for cfg in ("hf4m", "soup1m", "hf16m"):
    if cfg == "hf4m":
        v, f = synth.heightfield(2048, 1024); o, d = synth.random_rays(20_000_000, seed=1234, device=dev)
    elif cfg == "hf16m":
        v, f = synth.heightfield(4096, 2048); o, d = synth.random_rays(20_000_000, seed=1234, device=dev)
    else:
        v, f = synth.triangle_soup(1_000_000); o, d = synth.random_rays(10_000_000, seed=9, device=dev, box=True)
    r = RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
    lo, hi = o.min(0).values, o.max(0).values
    n = len(o)
    base = {m: timeit(lambda: getattr(r, "intersects_" + m)(o, d)) for m in ("closest", "any", "count")}
    print(cfg, "unsorted:", {k: f"{v:.3f} ms {n / v / 1e3:.0f} Mrays/s" for k, v in base.items()})
    m, octant = keys(o, d, lo, hi)
    for name, key in (("origin30", m), ("octant|origin30", (octant << 30) | m), ("origin15", m >> 15), ("octant|origin15", (octant << 15) | (m >> 15)), ("origin30|octant", (m << 3) | octant)):
        perm = torch.argsort(key)
        os_, ds_ = o[perm].contiguous(), d[perm].contiguous()
        res = {mm: timeit(lambda: getattr(r, "intersects_" + mm)(os_, ds_)) for mm in ("closest", "any", "count")}
        print("  sorted by", name, {k: f"{v:.3f} ms x{base[k] / v:.2f}" for k, v in res.items()})
    del r
