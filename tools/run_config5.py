"""BASELINE config 5: 16.8 M-triangle scene, 1 B rays sharded over N GPUs (125 M rays per GPU),
BVH built on rank 0 and NCCL-broadcast, closest hit with stream compaction, hit gather.

launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_config5.py [rays_per_gpu]
Prints one JSON line on rank 0 (also written to gpurun_out/config5_N<N>.json).
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from triro import synth
from triro.distributed import ShardedRayMeshIntersector, all_counts, gather_fixed

rays_per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000_000
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

v, f = synth.heightfield(4096, 2048)                 # every rank holds the mesh description; only rank 0 builds
vt, ft = torch.from_numpy(v), torch.from_numpy(f)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
sh = ShardedRayMeshIntersector.build(vt, ft, src=0)
torch.cuda.synchronize(); dist.barrier()
build_bcast_ms = (time.perf_counter() - t0) * 1e3
# broadcast alone (blob already resident): time a second broadcast of the same bytes
blob = sh.local.as_wrapper._inner.used()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dist.broadcast(blob, src=0); e1.record(); torch.cuda.synchronize()
bcast_ms = e0.elapsed_time(e1)

o = torch.empty((rays_per_gpu, 3), device=dev); d = torch.empty((rays_per_gpu, 3), device=dev)
chunk = 25_000_000
for i in range(0, rays_per_gpu, chunk):
    m = min(chunk, rays_per_gpu - i)
    oc, dc = synth.random_rays(m, seed=100 + rank * 16 + i // chunk, device=dev)
    o[i:i + m] = oc; d[i:i + m] = dc
    del oc, dc
r = sh.local
times = []
for it in range(4):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hit, front, ray_idx, tri_idx, loc, uv = r.intersects_closest(o, d, stream_compaction=True)
    e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = torch.tensor([min(times[1:])], dtype=torch.float64, device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# hit gather: per-rank totals -> global ray numbering -> gather of the compacted (ray, tri) pairs to every rank
base = rank * rays_per_gpu
gather_times = []
for it in range(3):                      # the first pass pays NCCL's lazy all-gather set-up and the allocator
    g_ray = g_tri = None
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    counts = all_counts(ray_idx.shape[0], dev)
    g_ray = gather_fixed(ray_idx.long() + base, counts)
    g_tri = gather_fixed(tri_idx, counts)
    torch.cuda.synchronize(); dist.barrier()
    gather_times.append((time.perf_counter() - t0) * 1e3)
gather_ms = min(gather_times[1:])
assert g_ray.shape[0] == sum(counts) and bool((g_ray[1:] > g_ray[:-1]).all())      # ascending global ray order
del g_ray, g_tri
# whole call, full 6-tuple: (a) trace + compaction + NCCL all-gather of every array to every rank,
# (b) fused: every rank's scatter kernel packs its hits straight into rank 0's tensors over NVLink
def whole(fn, reps=3):
    ts = []
    for it in range(reps):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter(); res = fn(); torch.cuda.synchronize(); dist.barrier()
        ts.append((time.perf_counter() - t0) * 1e3)
    t = torch.tensor([min(ts[1:])], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), res
t_nccl, res_nccl = whole(lambda: (lambda r6: [gather_fixed(r6[0], [rays_per_gpu] * world)] +
                                  [gather_fixed(x if i != 1 else x.long() + base, counts) for i, x in enumerate(r6[1:])])(
                                      r.intersects_closest(o, d, stream_compaction=True)))
n_check = int(res_nccl[2].shape[0]); del res_nccl
from triro.backend import ops as hops
from triro.distributed import PeerPacked
packed = PeerPacked(1 << (sum(counts) - 1).bit_length(), world * rays_per_gpu, dev)
def fused(scatter_to_peer):
    hit, front, tri, loc, uv = hops.intersects_closest(r.as_wrapper, o, d)
    ws, total = hops.compact_scan(hit)
    rb = 8 if world * rays_per_gpu > 2**31 - 1 else 4
    if not scatter_to_peer:       # pack locally, then bulk peer-to-peer copies into the root's tensors
        mine = dict(front=torch.empty(total, dtype=torch.uint8, device=dev), ray=torch.empty(total, dtype=torch.int64 if rb == 8 else torch.int32, device=dev),
                    tri=torch.empty(total, dtype=torch.int32, device=dev), loc=torch.empty(3 * total, dtype=torch.float32, device=dev),
                    uv=torch.empty(2 * total, dtype=torch.float32, device=dev))
        hops.compact_scatter_at(hit, ws, front, tri, loc, uv, base, rb, mine["front"].data_ptr(), mine["ray"].data_ptr(),
                                mine["tri"].data_ptr(), mine["loc"].data_ptr(), mine["uv"].data_ptr())
    cs = all_counts(total, dev)
    row0 = sum(cs[:rank])
    if scatter_to_peer:
        hops.compact_scatter_at(hit, ws, front, tri, loc, uv, base, rb, *packed.addresses(0, row0, rb))
    else:
        for name, t in mine.items():
            packed.peer_rows(0, name, row0, total, rb).copy_(t)
    packed.peer_hit_mask(0, base, base + rays_per_gpu).copy_(hit.view(torch.uint8))
    return sum(cs), rb
t_fused_scatter, _ = whole(lambda: fused(True))
t_fused, (h_fused, rb) = whole(lambda: fused(False))
if rank == 0:
    v = packed.local_views(h_fused, rb, (world * rays_per_gpu,))
    assert h_fused == n_check and bool((v["ray"][1:] > v["ray"][:-1]).all()) and int(v["hit"].sum()) == h_fused
assert int(hit.sum()) == ray_idx.shape[0] and bool((tri_idx >= 0).all())
if rank == 0:
    h = r.as_wrapper.header
    line = dict(config="5", n_gpus=world, tris=h["n_tris"], rays_total=world * rays_per_gpu, rays_per_gpu=rays_per_gpu,
                query="intersects_closest(stream_compaction=True)", ms=float(ms[0]),
                mrays_s=world * rays_per_gpu / float(ms[0]) / 1e3, build_plus_broadcast_ms=build_bcast_ms,
                broadcast_ms=bcast_ms, broadcast_gb_s=blob.numel() / bcast_ms / 1e6, blob_mb=blob.numel() / 1e6,
                hits_total=sum(counts), gather_ray_tri_ms=gather_ms, gather_first_call_ms=gather_times[0], whole_call_nccl_allgather_ms=t_nccl, whole_call_peer_copies_to_root_ms=t_fused, whole_call_peer_scatter_to_root_ms=t_fused_scatter, gathered_bytes=sum(counts) * 12)
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(line, open(os.path.join(ROOT, "gpurun_out", f"config5_N{world}.json"), "w"), indent=1)
dist.destroy_process_group()
