"""Strong-scaling probe (torchrun, N GPUs): ONE 66 M-ray batch on the 16.8 M-triangle heightfield split N ways, dense
5-tuple assembled on rank 0 by peer copies, for several chunk counts of the trace / copy pipeline."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from triro import synth
from triro.distributed import ShardedRayMeshIntersector, PeerOutputs
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
v, f = synth.heightfield(4096, 2048)
sh = ShardedRayMeshIntersector.build(torch.from_numpy(v), torch.from_numpy(f), src=0)
n = 8 * 3840 * 2160
o = torch.empty((n, 3), device=dev); d = torch.empty((n, 3), device=dev)
for i in range(0, n, 25_000_000):
    m = min(25_000_000, n - i)
    oc, dc = synth.random_rays(m, seed=7 * 16 + i // 25_000_000, device=dev)
    o[i:i + m] = oc; d[i:i + m] = dc
outs = PeerOutputs(n, dev)
res = {}
for chunks in (1, 2, 4, 8, 16):
    for kw in (dict(kernel_stores=False, chunks=chunks),):
        ts = []
        for it in range(5):
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            sh.intersects_closest_to_root(o, d, root=0, outputs=outs, **kw)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        t = torch.tensor([min(ts[1:])], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[f"chunks{chunks}"] = float(t)
ts = []
for it in range(4):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    sh.intersects_closest_to_root(o, d, root=0, outputs=outs, kernel_stores=True); torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
t = torch.tensor([min(ts[1:])], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
res["kernel_stores"] = float(t)
if rank == 0:
    print(json.dumps({"world": world, "rays": n, "ms": res}))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"world": world, "rays": n, "ms": res}, open(os.path.join(ROOT, "gpurun_out", f"strong_probe_n{world}.json"), "w"), indent=1)
dist.destroy_process_group()
