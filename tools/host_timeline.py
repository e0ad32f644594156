"""Per-chunk timeline of the host-buffer path (rt_host_trace_closest) on BASELINE config 2.
Needs a library built with -DRT_HOST_TIMELINE:  TRIRO_NVCC_EXTRA=-DRT_HOST_TIMELINE python trimesh-ray-optix_b200/triro/backend/build.py --force; python tools/host_timeline.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT)
import torch
from triro import synth
from triro.backend import ops as hops
from triro.ray.ray_optix import RayMeshIntersector


v, f = synth.icosphere(7)
r = RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
o, d = synth.pinhole_rays(3840, 2160, device="cpu")
o1 = o.reshape(-1, 3)[:1].contiguous().pin_memory()
d = d.contiguous().pin_memory()
out = hops.host_closest(r.as_wrapper, o1, d)
n = d.numel() // 3
wb = torch.empty(1, dtype=torch.uint8)
for rep in range(6):
    if rep == 5:
        os.environ["TRIRO_HOST_TIMELINE"] = "1"
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = hops.host_closest(r.as_wrapper, o1, d, out=out)
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {dt * 1e3:.3f} ms  {n / dt / 1e6:.0f} Mrays/s", flush=True)
