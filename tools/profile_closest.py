"""Short driver for ncu captures: builds one BASELINE config and runs the closest-hit kernel a few times.
usage: python tools/profile_closest.py [config2|soup1m|hf4m] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT)
import torch
from triro import synth
from triro.ray.ray_optix import RayMeshIntersector
from triro.backend import ops as hops

cfg = sys.argv[1] if len(sys.argv) > 1 else "config2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
if cfg == "config2":
    v, f = synth.icosphere(7); o, d = synth.pinhole_rays(3840, 2160, device=dev)
elif cfg == "soup1m":
    v, f = synth.triangle_soup(1_000_000); o, d = synth.random_rays(10_000_000, seed=9, device=dev, box=True)
elif cfg == "hf4m":
    v, f = synth.heightfield(2048, 1024); o, d = synth.random_rays(20_000_000, seed=1234, device=dev)
else:
    raise SystemExit(cfg)
r = RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
print(cfg, r.as_wrapper.header)
for mode in ("closest", "any", "count"):
    print(mode, hops.trace_stats(r.as_wrapper, o, d, mode))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    ev0.record(); res = r.intersects_closest(o, d); ev1.record(); torch.cuda.synchronize()
    print("closest ms", ev0.elapsed_time(ev1), "Mrays/s", o.numel() / 3 / ev0.elapsed_time(ev1) / 1e3)
for name, fn in (("any", r.intersects_any), ("count", r.intersects_count), ("first", r.intersects_first)):
    for i in range(2):
        ev0.record(); res = fn(o, d); ev1.record(); torch.cuda.synchronize()
    print(name, "ms", ev0.elapsed_time(ev1), "Mrays/s", o.numel() / 3 / ev0.elapsed_time(ev1) / 1e3)
