"""BVH build time (device time of rt_bvh_build, CUDA events) for the BASELINE mesh sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from triro import synth
from triro.backend import ops as hops
dev = torch.device("cuda:0")
cases = [("icosphere2", lambda: synth.icosphere(2)), ("icosphere7", lambda: synth.icosphere(7)),
         ("soup1m", lambda: synth.triangle_soup(1_000_000)), ("heightfield4m", lambda: synth.heightfield(2048, 1024)),
         ("heightfield16m", lambda: synth.heightfield(4096, 2048))]
only = sys.argv[1].split(",") if len(sys.argv) > 1 else None
for name, gen in cases:
    if only and name not in only: continue
    v, f = gen()
    vd, fd = torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)
    acc = hops.AccelStructure()
    ts = []
    for i in range(6):
        acc.build(vd, fd, timing=True); ts.append(acc.build_ms)
    h = acc.header
    print(f"{name:15s} tris {len(f):9d}  build min {min(ts[1:]):8.3f} ms  med {sorted(ts[1:])[2]:8.3f} ms  nodes {h['n_nodes']} depth {h['depth']} blob {h['used_bytes']/1e6:.1f} MB  ({len(f)/min(ts[1:])/1e3:.1f} Mtris/s)")
    acc.free(); del vd, fd
