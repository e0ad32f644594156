"""Builder leaf size (TRIRO_LEAF_TRIS=1..3, read once per process): Mrays/s, nodes / triangles per ray, blob size and
build time for the three benchmark scenes.  usage: TRIRO_LEAF_TRIS=2 python tools/leaf_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT)
import torch
from triro import synth
from triro.ray.ray_optix import RayMeshIntersector
from triro.backend import ops as hops

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for cfg in ("config2", "soup1m", "hf4m"):
    if cfg == "config2":
        v, f = synth.icosphere(7); o, d = synth.pinhole_rays(3840, 2160, device=dev)
    elif cfg == "soup1m":
        v, f = synth.triangle_soup(1_000_000); o, d = synth.random_rays(10_000_000, seed=9, device=dev, box=True)
    else:
        v, f = synth.heightfield(2048, 1024); o, d = synth.random_rays(20_000_000, seed=1234, device=dev)
    r = RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
    acc = hops.AccelStructure(); vd, fd = torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)
    bms = []
    for _ in range(3):
        acc.build(vd, fd, timing=True); bms.append(acc.build_ms)
    h = r.as_wrapper.header
    n = o.numel() // 3
    line = f"leaf {os.environ.get('TRIRO_LEAF_TRIS', 'default')} {cfg:8s} nodes {h['n_nodes']} blob {h['used_bytes'] / 1e6:.1f} MB build {min(bms[1:]):.3f} ms |"
    for name, fn in (("closest", r.intersects_closest), ("count", r.intersects_count)):
        st = hops.trace_stats(r.as_wrapper, o, d, name)
        ts = []
        for i in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(o, d); e1.record(); torch.cuda.synchronize()
            if i >= 2: ts.append(e0.elapsed_time(e1))
        line += f" {name} {n / min(ts) / 1e3:8.1f} Mrays/s (nodes {st['nodes_per_ray']:.2f} tris {st['tris_per_ray']:.2f}) |"
    print(line, flush=True)
    del r, acc
