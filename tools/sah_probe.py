"""How much of the traversal cost is tree quality?  Builds the BVH8 blob twice on the CPU (tests/hostsim: the product's
own collapse / quantisation code) - once from the product's Morton/Karras topology, once from a top-down binned-SAH
binary tree - and counts wide nodes and triangles fetched per ray for the benchmark ray sets.  No GPU needed."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from triro import synth
from hostsim import hostsim

def stats(blob, o, d, mode):
    r = hostsim.trace(blob, mode, o, d)["stats"]
    return r["nodes"] / r["rays"], r["tris"] / r["rays"]

cases = []
v, f = synth.icosphere(7); o, d = synth.pinhole_rays(960, 540, device="cpu")
cases.append(("config2 icosphere 327k, camera", v, f, np.broadcast_to(o.numpy(), d.shape).reshape(-1, 3).copy(), d.numpy().reshape(-1, 3)))
v, f = synth.triangle_soup(1_000_000); o, d = synth.random_rays(200_000, seed=9, device="cpu", box=True)
cases.append(("config4 soup 1M, random", v, f, o.numpy(), d.numpy()))
v, f = synth.heightfield(1024, 512); o, d = synth.random_rays(400_000, seed=1234, device="cpu")
cases.append(("heightfield 1M, random", v, f, o.numpy(), d.numpy()))
for name, v, f, o, d in cases:
    t0 = time.time(); lb = hostsim.build_blob(v, f); t1 = time.time(); sb = hostsim.build_blob_sah(v, f); t2 = time.time()
    assert hostsim.check_blob(sb) is not None
    for mode in ("closest", "count"):
        a = stats(lb, o, d, mode); b = stats(sb, o, d, mode)
        print(f"{name:34s} {mode:8s} LBVH nodes/ray {a[0]:6.2f} tris/ray {a[1]:6.2f} | SAH {b[0]:6.2f} {b[1]:6.2f} | "
              f"bytes ratio {(80 * b[0] + 48 * b[1]) / (80 * a[0] + 48 * a[1]):.2f}  (build {t1 - t0:.1f}s / {t2 - t1:.1f}s)")
