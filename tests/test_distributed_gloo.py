"""world_size-2 gloo test of the ray-sharding layer (triro/distributed.py) on CPU tensors.

The per-rank tracer is an oracle-backed stand-in with the RayMeshIntersector method surface, so
what is tested is the host logic: slice bounds, global ray-index rebasing, padded all_gather of
fixed and variable-length results, blob broadcast.  Property: the gathered N-rank result equals
the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle
from triro import synth


class OracleBackedLocal:
    """torch-CPU stand-in for a per-rank RayMeshIntersector."""

    def __init__(self, v, f):
        self.oi = oracle.OracleIntersector(v, f, mode=oracle.MIRROR)

    @staticmethod
    def _t(x):
        return torch.from_numpy(np.ascontiguousarray(x))

    def intersects_any(self, o, d):
        return self._t(self.oi.intersects_any(o.numpy(), d.numpy()))

    def intersects_first(self, o, d):
        return self._t(self.oi.intersects_first(o.numpy(), d.numpy()))

    def intersects_count(self, o, d):
        return self._t(self.oi.intersects_count(o.numpy(), d.numpy()))

    def intersects_closest(self, o, d, stream_compaction=False):
        return tuple(self._t(x) for x in self.oi.intersects_closest(o.numpy(), d.numpy(), stream_compaction))

    def intersects_location(self, o, d):
        loc, ri, ti, _, _ = self.oi.intersects_location(o.numpy(), d.numpy())
        return self._t(loc), self._t(ri), self._t(ti)

    def contains_points(self, p, check_direction=None):
        return self._t(self.oi.contains_points(p.numpy(), check_direction))

    def contains_parity(self, points, direction, active=None, out=None, stop_when_broken=False):
        """Stand-in for RayMeshIntersector.contains_parity (the fused kernel): same outputs, masked in-place update."""
        inside, cp, cm = self.oi.contains_core(points.numpy(), direction)
        agree = (cp % 2 == 1) & (cm % 2 == 1)
        contain, broken = self._t(inside & agree), self._t(~agree & ((cp == 0) | (cm == 0)))
        if active is None:
            return contain, broken, torch.tensor([int(inside.any()), int(broken.any())], dtype=torch.int32)
        act = active.clone()
        flags = torch.tensor([int(inside[act.numpy()].any()), int(broken[act].any())], dtype=torch.int32)
        out[0][act] = contain[act]
        out[1][act] = broken[act]
        return out[0], out[1], flags


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from triro.distributed import ShardedRayMeshIntersector, broadcast_blob, shard_bounds

        v, f = synth.icosphere(2)
        local = OracleBackedLocal(v, f)
        sh = ShardedRayMeshIntersector(local)
        o, d = synth.random_rays(1001, seed=4, box=True)          # odd count: uneven slices
        o = (o * 1.6).reshape(7, 143, 3); d = d.reshape(7, 143, 3)
        res = {}
        res["any"] = sh.intersects_any(o, d)
        res["first"] = sh.intersects_first(o, d)
        res["count"] = sh.intersects_count(o, d)
        res["closest"] = sh.intersects_closest(o, d)
        res["compact"] = sh.intersects_closest(o, d, stream_compaction=True)
        res["location"] = sh.intersects_location(o, d)
        res["id"] = sh.intersects_id(o, d, return_locations=True, multiple_hits=False)
        pts = (torch.rand((501, 3), generator=torch.Generator().manual_seed(2)) * 2 - 1) * 0.9
        res["contains"] = sh.contains_points(pts, torch.tensor([0.3, 0.5, 0.8]))
        # default direction on an OPEN mesh (one face removed): some points are 'broken' -> the retry branch; the
        # decisions are taken over all ranks and the retry direction comes from rank 0, so N ranks == 1 process
        torch.manual_seed(1234 + rank)           # ranks deliberately disagree about their own RNG
        open_local = OracleBackedLocal(v, f[1:])
        sh_open = ShardedRayMeshIntersector(open_local)
        if rank == 0:
            torch.manual_seed(77)
        res["contains_retry"] = sh_open.contains_points(pts * 1.2)
        # explicit direction with broken points on rank 1's slice only: the reference answers all False everywhere
        res["contains_quirk"] = sh_open.contains_points(pts * 1.2, torch.tensor([0.3, 0.5, 0.8]))
        part, (lo, hi) = sh.intersects_first(o, d, gather=False)
        assert (lo, hi) == shard_bounds(1001, world, rank) and part.shape == (hi - lo,)
        blob = torch.arange(1000, dtype=torch.uint8) if rank == 0 else None
        got = broadcast_blob(blob, src=0, device="cpu")
        assert torch.equal(got, torch.arange(1000, dtype=torch.uint8))
        if rank == 0:
            q.put({k: [x.numpy() for x in (val if isinstance(val, tuple) else (val,))] for k, val in res.items()})
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    v, f = synth.icosphere(2)
    single = OracleBackedLocal(v, f)
    o, d = synth.random_rays(1001, seed=4, box=True)
    o = (o * 1.6).reshape(7, 143, 3); d = d.reshape(7, 143, 3)
    exp = {
        "any": (single.intersects_any(o, d),), "first": (single.intersects_first(o, d),),
        "count": (single.intersects_count(o, d),), "closest": single.intersects_closest(o, d),
        "compact": single.intersects_closest(o, d, stream_compaction=True),
        "location": single.intersects_location(o, d),
    }
    hit, _, ray, tri, loc, _ = exp["compact"]
    exp["id"] = (tri, ray, loc)
    pts = (torch.rand((501, 3), generator=torch.Generator().manual_seed(2)) * 2 - 1) * 0.9
    exp["contains"] = (single.contains_points(pts, torch.tensor([0.3, 0.5, 0.8])),)
    single_open = OracleBackedLocal(v, f[1:])
    torch.manual_seed(77)
    exp["contains_retry"] = (single_open.contains_points(pts * 1.2),)
    assert single_open.oi.contains_core((pts * 1.2).numpy(), single_open.oi.DEFAULT_DIRECTION)[1].size == 501
    exp["contains_quirk"] = (single_open.contains_points(pts * 1.2, torch.tensor([0.3, 0.5, 0.8])),)
    for k, vals in exp.items():
        assert len(vals) == len(got[k]), k
        for a, b in zip(vals, got[k]):
            assert a.shape == b.shape and np.array_equal(a.numpy(), b), f"{k}: sharded result differs"
