"""Short driver for timing / ncu captures: builds one BASELINE config and runs the trace kernels a few times.
usage: python tools/profile_closest.py [config2|soup1m|hf4m|readme] [reps] [modes]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT)
import torch
from triro import synth
from triro.ray.ray_optix import RayMeshIntersector
from triro.backend import ops as hops

cfg = sys.argv[1] if len(sys.argv) > 1 else "config2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["closest", "any", "count", "first"]
dev = torch.device("cuda:0")
if cfg == "config2":
    v, f = synth.icosphere(7); o, d = synth.pinhole_rays(3840, 2160, device=dev)
elif cfg == "readme":
    v, f = synth.icosphere(3); o, d = synth.readme_rays(800, device=dev)
elif cfg == "soup1m":
    v, f = synth.triangle_soup(1_000_000); o, d = synth.random_rays(10_000_000, seed=9, device=dev, box=True)
elif cfg == "hf4m":
    v, f = synth.heightfield(2048, 1024); o, d = synth.random_rays(20_000_000, seed=1234, device=dev)
else:
    raise SystemExit(cfg)
r = RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
h = r.as_wrapper.header
print(cfg, "tris", h["n_tris"], "nodes", h["n_nodes"], "depth", h["depth"], "blob MB", h["used_bytes"] / 1e6,
      "thr", os.environ.get("TRIRO_REFILL_THRESHOLD"))
n = o.numel() // 3
fns = dict(closest=r.intersects_closest, any=r.intersects_any, count=r.intersects_count, first=r.intersects_first)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in modes:
    st = hops.trace_stats(r.as_wrapper, o, d, name if name != "first" else "closest")
    bpr = 80 * st["nodes_per_ray"] + 48 * st["tris_per_ray"] + (12 if o.stride(0) == 0 or cfg in ("config2", "readme") else 24) + dict(closest=26, any=1, count=4, first=4)[name]
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); res = fns[name](o, d); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    best, med = min(ts), statistics.median(ts)
    print(f"{name:8s} min {best:8.3f} ms  med {med:8.3f} ms  {n / best / 1e3:9.1f} Mrays/s  nodes/ray {st['nodes_per_ray']:.2f} tris/ray {st['tris_per_ray']:.2f}"
          f"  B/ray {bpr:.0f}  roofline frac {bpr * n / (best * 1e-3) / 6543.1e9:.3f}")
