"""The reference's own benchmark loop (test/performance_test.py:29-60) on a synthetic scene: 640x360 pinhole rays
(its cam_mat / focal), N back-to-back intersects_closest calls timed with the wall clock exactly like the reference
(no explicit synchronisation inside the loop; one at the end).  The reference's scene file is not available offline,
so the mesh is the 327 680-triangle icosphere placed in front of the camera."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT)
import torch
from triro import synth
from triro.ray.ray_optix import RayMeshIntersector

cam_mat = torch.tensor([[5.6272650e-01, 2.7091104e-01, 7.8099048e-01], [8.2602328e-01, -1.4769979e-01, -5.4393965e-01],
                        [3.2007132e-02, -9.5120555e-01, 3.0689341e-01]]).cuda()
rw = 640; rh = int(rw * 9 / 16); rf = int(rw * 25 / 36)
GPU_ITER = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
v, f = synth.icosphere(7)
cam_origin = (cam_mat[:, 2] * 3.0).cpu()
ray_dirs = synth.gen_rays(cam_mat, rw, rh, rf, device="cuda")
ray_origins = cam_origin.cuda().broadcast_to(ray_dirs.shape)
r = RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
for _ in range(100):
    result = r.intersects_closest(ray_origins, ray_dirs)
torch.cuda.synchronize()
t0 = time.time()
for i in range(GPU_ITER):
    result = r.intersects_closest(ray_origins, ray_dirs)
torch.cuda.synchronize()
dt = time.time() - t0
n = rw * rh
print(f"GPU time: {dt:.3f} s / {GPU_ITER} iters -> {dt / GPU_ITER * 1e6:.1f} us per call, {n * GPU_ITER / dt / 1e6:.0f} Mrays/s "
      f"(hit fraction {float(result[0].float().mean()):.3f}; reference README: 83.6 us per call = 2755 Mrays/s on an RTX 3090 with RT cores)")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); result = r.intersects_closest(ray_origins, ray_dirs); e1.record(); torch.cuda.synchronize()
print(f"device time of one call: {e0.elapsed_time(e1) * 1e3:.1f} us")
