"""PCIe probe for the e2e path: pinned H2D / D2H alone and overlapped, plus rt_host_trace_closest chunk sweep."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT)
import torch
dev = torch.device("cuda:0")
n = 8294400
h_in = torch.empty(n * 12, dtype=torch.uint8).pin_memory(); d_in = torch.empty(n * 12, dtype=torch.uint8, device=dev)
h_out = torch.empty(n * 26, dtype=torch.uint8).pin_memory(); d_out = torch.empty(n * 26, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D {h_in.numel()/a/1e9:.1f} GB/s ({a*1e3:.2f} ms)  D2H {h_out.numel()/b/1e9:.1f} GB/s ({b*1e3:.2f} ms)  overlapped {c*1e3:.2f} ms -> {n/c/1e6:.0f} Mrays/s bound")
from triro import synth
from triro.ray.ray_optix import RayMeshIntersector
from triro.backend import ops as hops
v, f = synth.icosphere(7)
r = RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
o, d = synth.pinhole_rays(3840, 2160, device="cpu")
dh = d.reshape(-1, 3).contiguous().pin_memory(); oh = torch.tensor([0.0, 0.0, 3.0]).pin_memory()
out = hops.host_closest(r.as_wrapper, oh, dh)
for chunk in os.environ.get("CHUNKS", "default").split(","):
    if chunk != "default": os.environ["TRIRO_HOST_CHUNK"] = chunk
    for slots in os.environ.get("SLOTS", "default").split(","):
        if slots != "default": os.environ["TRIRO_HOST_SLOTS"] = slots
        e = t(lambda: hops.host_closest(r.as_wrapper, oh, dh, out=out), reps=8)
        print(f"chunk {chunk} slots {slots}: {e*1e3:.2f} ms  {n/e/1e6:.0f} Mrays/s")
