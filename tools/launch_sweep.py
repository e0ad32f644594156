"""Time per intersects_closest call against launch size (camera rays on the 327 680-triangle icosphere, or
`soup` / `heightfield`), back-to-back calls timed with the wall clock like the reference benchmark."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT)
import torch
from triro import synth
from triro.ray.ray_optix import RayMeshIntersector

cam_mat = torch.tensor([[5.6272650e-01, 2.7091104e-01, 7.8099048e-01], [8.2602328e-01, -1.4769979e-01, -5.4393965e-01],
                        [3.2007132e-02, -9.5120555e-01, 3.0689341e-01]]).cuda()
scene = sys.argv[1] if len(sys.argv) > 1 else "icosphere"
if scene == "icosphere":
    v, f = synth.icosphere(7); dist = 3.0
elif scene == "tiny":
    v, f = synth.icosphere(0); dist = 3.0
elif scene == "soup":
    v, f = synth.triangle_soup(1_000_000, seed=3); dist = 3.0
else:
    v, f = synth.heightfield(1448); dist = 3.0
r = RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
origin = (cam_mat[:, 2] * dist)
print(f"scene {scene}: {len(f)} triangles")
for rw, rh in ((320, 180), (640, 360), (1280, 720), (1920, 1080), (3840, 2160)):
    rf = int(rw * 25 / 36)
    d = synth.gen_rays(cam_mat, rw, rh, rf, device="cuda")
    if scene != "icosphere":
        d = -d
    o = origin.broadcast_to(d.shape)
    line = f"{rw}x{rh} ({rw * rh / 1e6:.2f} Mrays): "
    for rep in range(2):
        iters = max(20, min(2000, int(2e8 / (rw * rh))))
        for _ in range(10):
            r.intersects_closest(o, d)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(iters):
            res = r.intersects_closest(o, d)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / iters
        line += f" {dt * 1e6:8.1f} us {rw * rh / dt / 1e6:7.0f} Mrays/s |"
    print(line + f" hit {float(res[0].float().mean()):.3f}")
