"""Fused trace + gather (peer stores into the root's tensors from inside k_trace) against trace + NCCL all-gather,
on BASELINE config 2 split over N ranks (each rank traces 1/N of ONE 4K frame: strong scaling of a single call).
launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/peer_gather.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "trimesh-ray-optix_b200")); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from triro import synth
from triro.distributed import ShardedRayMeshIntersector, PeerOutputs

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
v, f = synth.icosphere(7)
sh = ShardedRayMeshIntersector.build(torch.from_numpy(v), torch.from_numpy(f), src=0)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
o, d = synth.pinhole_rays(3840, 2160 * scale, device=dev)
n = d.numel() // 3

def timed(fn, reps=8):
    ts = []
    for i in range(reps + 3):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); dist.barrier()
        if i >= 3: ts.append((time.perf_counter() - t0) * 1e3)
    t = torch.tensor([min(ts)], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), r

t_local, res_local = timed(lambda: sh.intersects_closest(o, d, gather=False))
t_gather, res_gather = timed(lambda: sh.intersects_closest(o, d, gather=True))
outs = PeerOutputs(n, dev)
t_peer, res_peer = timed(lambda: sh.intersects_closest_to_root(o, d, root=0, outputs=outs, kernel_stores=True))
t_copy, res_copy = timed(lambda: sh.intersects_closest_to_root(o, d, root=0, outputs=outs, kernel_stores=False))
ok = True
if rank == 0:          # res_peer and res_copy alias the same symmetric buffer: the copy route ran last
    for a, b in zip(res_gather, res_copy):
        ok = ok and bool(torch.equal(a, b))
res_peer = sh.intersects_closest_to_root(o, d, root=0, outputs=outs, kernel_stores=True)      # collective: every rank calls it
if rank == 0:
    for a, b in zip(res_gather, res_peer):
        ok = ok and bool(torch.equal(a, b))
    line = dict(n_gpus=world, rays=n, trace_only_sharded_ms=t_local, trace_plus_nccl_allgather_ms=t_gather,
                trace_with_peer_stores_to_root_ms=t_peer, trace_then_peer_copies_to_root_ms=t_copy, identical=ok,
                mrays_s_peer=n / t_peer / 1e3, mrays_s_allgather=n / t_gather / 1e3)
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(line, open(os.path.join(ROOT, "gpurun_out", f"peer_gather_N{world}.json"), "w"), indent=1)
dist.destroy_process_group()
