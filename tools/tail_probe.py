"""Timing probe for launch-tail experiments: the reference's 640x360 loop (us per call, no sync inside the loop), one cold
640x360 launch, and the big scenes under the AUTO schedule.  Select an experiment library with TRIRO_B200_LIB.

usage: python tools/tail_probe.py [tag]   -> gpurun_out/<tag>_tail.json
"""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT)
import torch
from triro import synth
from triro.ray.ray_optix import RayMeshIntersector

tag = sys.argv[1] if len(sys.argv) > 1 else "probe"
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
out = {"lib": os.environ.get("TRIRO_B200_LIB", "default")}


def timed(fn, reps=7, warm=2):
    ts = []
    for i in range(warm + reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    return min(ts), statistics.median(ts)


def loop_us(fn, calls=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(calls):
            fn()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / calls * 1e6)
    return best


v, f = synth.icosphere(7)
r = RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
o, d = synth.pinhole_rays(640, 360, device=dev)
out["loop640_us"] = loop_us(lambda: r.intersects_closest(o, d))
out["cold640_ms"] = timed(lambda: r.intersects_closest(o, d))
for w, h in ((160, 90), (1280, 720)):
    o2, d2 = synth.pinhole_rays(w, h, device=dev)
    out[f"loop{w}_us"] = loop_us(lambda: r.intersects_closest(o2, d2))
o, d = synth.readme_rays(800, device=dev)
out["readme800_loop_us"] = loop_us(lambda: r.intersects_closest(o, d))
o, d = synth.pinhole_rays(3840, 2160, device=dev)
out["config2_ms"] = timed(lambda: r.intersects_closest(o, d))
out["config2_count_ms"] = timed(lambda: r.intersects_count(o, d))
v, f = synth.icosphere(8)
r = RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
out["ico8_ms"] = timed(lambda: r.intersects_closest(o, d))
v, f = synth.heightfield(2048, 1024)
r = RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
o, d = synth.random_rays(20_000_000, seed=1234, device=dev)
out["hf4m_closest_ms"] = timed(lambda: r.intersects_closest(o, d), reps=4)
out["hf4m_any_ms"] = timed(lambda: r.intersects_any(o, d), reps=4)
v, f = synth.triangle_soup(1_000_000)
r = RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
o, d = synth.random_rays(10_000_000, seed=9, device=dev, box=True)
out["soup_closest_ms"] = timed(lambda: r.intersects_closest(o, d), reps=4)
out["soup_count_ms"] = timed(lambda: r.intersects_count(o, d), reps=4)
out["soup_location_ms"] = timed(lambda: r.intersects_location(o, d), reps=3)
pts = (torch.rand((4_000_000, 3), device=dev) - 0.5) * 2.0
out["soup_contains_ms"] = timed(lambda: r.contains_points(pts), reps=3)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_tail.json"), "w"), indent=1)
print(json.dumps(out))
