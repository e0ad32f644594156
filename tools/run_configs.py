"""Runs the five BASELINE.json configurations on ONE GPU (config 5 as its per-GPU slice), checks
size-independent properties of the results at full size, and writes profiles/<tag>_configs.{json,md}.

usage: python tools/run_configs.py [tag] [comma-separated configs]
"""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch
from triro import synth
from triro.ray.ray_optix import RayMeshIntersector
from triro.backend import ops as hops

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
only = sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "2", "3", "4", "5"]
dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
rows = []


def timed(fn, reps=5, warm=2):
    ts = []
    for i in range(warm + reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    return min(ts), statistics.median(ts), out


def build(v, f):
    t0 = time.perf_counter()
    r = RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))
    torch.cuda.synchronize()
    acc = hops.AccelStructure()
    vd, fd = r.mesh_vertices, r.mesh_faces
    bt = []
    for _ in range(4):
        acc.build(vd, fd, timing=True); bt.append(acc.build_ms)
    acc.free()
    return r, min(bt[1:])


def record(cfg, query, n, ms, med, stats_mode, r, o, d, b_in, b_out, build_ms, note=""):
    st = hops.trace_stats(r.as_wrapper, o, d, stats_mode) if stats_mode else None
    bpr = b_in + b_out + (80 * st["nodes_per_ray"] + 48 * st["tris_per_ray"] if st else 0)
    h = r.as_wrapper.header
    row = dict(config=cfg, query=query, tris=h["n_tris"], rays=n, ms=ms, ms_median=med, mrays_s=n / ms / 1e3,
               nodes_per_ray=st["nodes_per_ray"] if st else None, tris_per_ray=st["tris_per_ray"] if st else None,
               bytes_per_ray=bpr, roofline_mrays_s=PEAK * 1e9 / bpr / 1e6, roofline_frac=bpr * n / (ms * 1e-3) / (PEAK * 1e9),
               build_ms=build_ms, blob_mb=h["used_bytes"] / 1e6, bvh_nodes=h["n_nodes"], bvh_depth=h["depth"], note=note)
    rows.append(row)
    print(json.dumps(row))


if "1" in only:   # README quick-start
    for sub in (2, 3):
        v, f = synth.icosphere(sub)
        r, bms = build(v, f)
        o, d = synth.readme_rays(800, device=dev)
        ms, med, out = timed(lambda: r.intersects_closest(o, d, stream_compaction=True))
        hit = out[0]
        assert torch.equal(out[2].long(), torch.nonzero(hit.reshape(-1)).reshape(-1))              # ray_idx = positions of hits
        assert bool(out[1].all()) and abs(float(hit.float().mean()) - 0.098) < 0.004              # all front faces; disc area
        record(f"1 (icosphere subdiv {sub})", "intersects_closest(stream_compaction=True)", 640000, ms, med, "closest", r, o, d, 12,
               1 + 29 * float(hit.float().mean()), bms, "includes scan + host sync + scatter")

if "2" in only:
    v, f = synth.icosphere(7)
    r, bms = build(v, f)
    o, d = synth.pinhole_rays(3840, 2160, device=dev)
    n = 3840 * 2160
    ms, med, out = timed(lambda: r.intersects_closest(o, d))
    hit, front, tri, loc, uv = out
    # properties at full size: hit locations lie on the faceted unit sphere, all front faces, uv reconstructs loc
    nrm = loc[hit].norm(dim=1)
    assert float(nrm.min()) > 0.9995 and float(nrm.max()) < 1.0 + 1e-6 and bool(front[hit].all())
    tv = r.mesh_vertices[r.mesh_faces[tri[hit].long()].long()]
    u = uv[hit]
    rec = u[:, :1] * tv[:, 0] + u[:, 1:] * tv[:, 1] + (1 - u[:, :1] - u[:, 1:]) * tv[:, 2]
    assert float((rec - loc[hit]).abs().max()) < 2e-6
    assert torch.equal(r.intersects_first(o, d), tri) and torch.equal(r.intersects_any(o, d), hit)
    record("2", "intersects_closest", n, ms, med, "closest", r, o, d, 12, 26, bms)
    ms, med, _ = timed(lambda: r.intersects_closest(o, d, stream_compaction=True))
    record("2", "intersects_closest(stream_compaction=True)", n, ms, med, "closest", r, o, d, 12, 26 + 1 + 29 * 0.3367 * 2, bms,
           "dense trace + scan + sync + scatter")

if "3" in only:
    v, f = synth.heightfield(2048, 1024)
    r, bms = build(v, f)
    n = 100_000_000
    o = torch.empty((n, 3), device=dev); d = torch.empty((n, 3), device=dev)
    for i in range(10):
        oc, dc = synth.random_rays(10_000_000, seed=1234 + i, device=dev)
        o[i * 10_000_000:(i + 1) * 10_000_000] = oc; d[i * 10_000_000:(i + 1) * 10_000_000] = dc
    ms, med, anyh = timed(lambda: r.intersects_any(o, d), reps=3, warm=1)
    record("3", "intersects_any", n, ms, med, "any", r, o, d, 24, 1, bms)
    ms, med, cnt = timed(lambda: r.intersects_count(o, d), reps=3, warm=1)
    record("3", "intersects_count", n, ms, med, "count", r, o, d, 24, 4, bms)
    assert torch.equal(cnt > 0, anyh)                                   # any == (count > 0)
    # a heightfield is a function graph: rays starting above it going upwards never hit, downward rays inside the
    # footprint cross it an odd number of times
    up = d[:, 2] > 0.2
    assert not bool(anyh[up].any())
    del o, d

if "4" in only:
    v, f = synth.triangle_soup(1_000_000)
    r, bms = build(v, f)
    n = 10_000_000
    g = torch.Generator(device=dev); g.manual_seed(8)
    pts = torch.rand((n, 3), generator=g, device=dev) * 2 - 1
    torch.manual_seed(0)
    ms, med, inside = timed(lambda: r.contains_points(pts), reps=3, warm=1)
    rows.append(dict(config="4", query="contains_points (default direction)", tris=1_000_000, rays=n, ms=ms, ms_median=med,
                     mrays_s=n / ms / 1e3, note=f"2 traversals per point fused; inside fraction {float(inside.float().mean()):.2e}"))
    print(json.dumps(rows[-1]))
    xdir = torch.tensor([1.0, 0.0, 0.0], device=dev)
    ms, med, inside_x = timed(lambda: r.contains_points(pts, xdir), reps=3, warm=1)
    rows.append(dict(config="4", query="contains_points (+x)", tris=1_000_000, rays=n, ms=ms, ms_median=med, mrays_s=n / ms / 1e3,
                     note="explicit direction (reference quirk: all False when any point is 'broken')"))
    print(json.dumps(rows[-1]))
    # +x parity check against two count traces
    dirs = xdir.broadcast_to(pts.shape)
    cp, cm = r.intersects_count(pts, dirs), r.intersects_count(pts, -dirs)
    contain, broken, flags = hops.contains_parity(r.as_wrapper, pts, [1, 0, 0], *r._aabb_host)
    lo, hi = r.mesh_aabb
    inside_aabb = ((pts > lo) & (pts < hi)).all(dim=1)
    assert torch.equal(contain, inside_aabb & (cp % 2 == 1) & (cm % 2 == 1))
    assert torch.equal(broken, ~((cp % 2 == 1) & (cm % 2 == 1)) & ((cp == 0) | (cm == 0)))
    o, d = synth.random_rays(n, seed=9, device=dev, box=True)
    ms, med, out = timed(lambda: r.intersects_location(o, d), reps=3, warm=1)
    loc, ri, ti = out
    cnt = r.intersects_count(o, d)
    assert loc.shape[0] == int(torch.clamp(cnt, max=8).sum()) and bool((ri[1:] >= ri[:-1]).all())
    over8 = float((cnt > 8).float().mean())
    record("4", "intersects_location (all hits)", n, ms, med, "count", r, o, d, 24, 4 + 20 * float(cnt.clamp(max=8).float().mean()), bms,
           f"single traversal + scan + sync + scatter; mean hits/ray {float(cnt.float().mean()):.2f}, rays with > 8 hits {over8:.2e}")
    ms, med, _ = timed(lambda: r.intersects_closest(o, d), reps=3, warm=1)
    record("4", "intersects_closest (1M-triangle soup, incoherent)", n, ms, med, "closest", r, o, d, 24, 26, bms)
    del o, d, pts

if "5" in only:   # per-GPU slice of config 5: 16.8 M triangles, 125 M rays
    v, f = synth.heightfield(4096, 2048)
    r, bms = build(v, f)
    n = 125_000_000
    o = torch.empty((n, 3), device=dev); d = torch.empty((n, 3), device=dev)
    for i in range(5):
        oc, dc = synth.random_rays(25_000_000, seed=100 + i, device=dev)
        o[i * 25_000_000:(i + 1) * 25_000_000] = oc; d[i * 25_000_000:(i + 1) * 25_000_000] = dc
    ms, med, out = timed(lambda: r.intersects_closest(o, d), reps=3, warm=1)
    hit, front, tri, loc, uv = out
    assert bool((tri[~hit] == -1).all()) and bool((tri[hit] >= 0).all())
    zs = loc[hit][:, 2]
    assert float(zs.abs().max()) <= 0.1 + 1e-5                              # hits lie on the terrain (|z| <= amplitude)
    record("5 (per-GPU slice)", "intersects_closest", n, ms, med, "closest", r, o, d, 24, 26, bms,
           "16.8 M-triangle heightfield, 125 M random rays = one rank of the 8-GPU / 1 B-ray job; blob 1.0 GB > L2")
    # 1M-triangle MESH (connected, coherent camera): icosphere subdiv 8 = 1.31 M triangles
    del o, d
    v, f = synth.icosphere(8)
    r, bms = build(v, f)
    o, d = synth.pinhole_rays(3840, 2160, device=dev)
    ms, med, out = timed(lambda: r.intersects_closest(o, d))
    record("1M-triangle mesh (icosphere subdiv 8, 1.31 M tris)", "intersects_closest", 3840 * 2160, ms, med, "closest", r, o, d, 12, 26, bms,
           "north_star target: >= 70 % of the memory roofline on a 1M-triangle mesh")

OUT = os.path.join(ROOT, "gpurun_out")     # merged back from the GPU box; copy into profiles/ afterwards
os.makedirs(OUT, exist_ok=True)
with open(os.path.join(OUT, f"{tag}_configs.json"), "w") as fh:
    json.dump(dict(peak_hbm_gbs=PEAK, rows=rows), fh, indent=1)
with open(os.path.join(OUT, f"{tag}_configs.md"), "w") as fh:
    fh.write(f"| config | query | tris | rays | ms | Mrays/s | nodes/ray | tris/ray | B/ray | roofline Mrays/s | frac (of measured {PEAK:.0f} GB/s) | build ms | note |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for x in rows:
        g = lambda k, fmt="{:.2f}": (fmt.format(x[k]) if x.get(k) is not None else "–")
        fh.write(f"| {x['config']} | {x['query']} | {x.get('tris','–')} | {x['rays']} | {g('ms','{:.3f}')} | {g('mrays_s','{:.0f}')} | {g('nodes_per_ray')} | {g('tris_per_ray')} | {g('bytes_per_ray','{:.0f}')} | {g('roofline_mrays_s','{:.0f}')} | {g('roofline_frac','{:.3f}')} | {g('build_ms','{:.3f}')} | {x.get('note','')} |\n")
print("written gpurun_out/%s_configs.{json,md}" % tag)
