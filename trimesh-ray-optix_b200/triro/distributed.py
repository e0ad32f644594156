"""Ray-slice data parallelism over several GPUs (one process per GPU, torch.distributed).

The reference is single-GPU (one global OptiX context, triro/backend/base.cpp:33-41); this
layer is what BASELINE.json's north_star adds: the mesh/BVH is built once on `src` and
broadcast as one flat blob over NCCL (NVLink 5 / NVSwitch), every rank traces a contiguous
slice of the flattened ray index space, and results are either kept sharded or gathered —
variable-length all-hit / compacted results included, with ray indices rebased to the global
numbering so that the concatenation equals the single-GPU answer.

There is no compute/collective fusion here on purpose: the path has no exchange step between
kernels (rays are independent), the only collectives are one broadcast per build and an
optional gather of results.

The collective plumbing is backend-agnostic (works with gloo on CPU tensors), which is how the
host logic is tested without GPUs (tests/test_distributed_gloo.py).
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def bind_to_device_cpus(device_index: int) -> List[int] | None:
    """Pin the calling thread (and the threads / pinned host allocations it makes from now on) to the CPU cores
    that are local to CUDA device `device_index` (same NUMA node / PCIe root), via NVML's ideal CPU affinity.
    With one process per GPU on a two-socket host this keeps each rank's pinned buffers and copy submission off the
    inter-socket link, which is what the host-buffer path (rt_host_trace_closest) is bound by at 8 ranks.
    Returns the CPU list, or None when NVML / the mapping is unavailable (nothing changed)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        handle = None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device_index]) if vis and vis.split(",")[device_index].isdigit() else device_index
            handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n rays owned by `rank`: ceil(n / world) rays per rank."""
    per = (n + world - 1) // world if world > 0 else n
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def slice_rays(origins: torch.Tensor, directions: torch.Tensor, lo: int, hi: int):
    """Slice the FLATTENED batch index space [lo, hi) of [*b, 3] tensors without copying when the
    batch is one-dimensional; strided / broadcast multi-dimensional inputs are flattened by index
    (reshape copies only what it must)."""
    o = origins.reshape(-1, 3) if origins.dim() != 2 else origins
    d = directions.reshape(-1, 3) if directions.dim() != 2 else directions
    return o[lo:hi], d[lo:hi]


def broadcast_blob(blob: torch.Tensor | None, src: int = 0, group=None, device=None) -> torch.Tensor:
    """Broadcast the used prefix of a BVH blob from `src` to every rank (ncclBroadcast)."""
    rank = dist.get_rank(group)
    size = torch.zeros(1, dtype=torch.int64, device=device if device is not None else (blob.device if blob is not None else "cpu"))
    if rank == src:
        size[0] = blob.numel()
    dist.broadcast(size, src=src, group=group)
    n = int(size.item())
    if rank != src:
        blob = torch.empty(n, dtype=torch.uint8, device=size.device)
    dist.broadcast(blob, src=src, group=group)
    return blob


def gather_fixed(x: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """Concatenate per-rank tensors whose leading dimension is counts[rank] (ranks in order).
    One padded all-gather into a flat buffer (every rank contributes max(counts) rows; the padding is
    uninitialised and dropped), then one compaction copy - none when all ranks hold the same count."""
    world = dist.get_world_size(group)
    m = max(counts) if len(counts) else 0
    if m == 0:
        return x[:0]
    as_bool = x.dtype == torch.bool
    src = x.to(torch.uint8) if as_bool else x
    if src.shape[0] == m:
        pad = src.contiguous()
    else:
        pad = torch.empty((m, *src.shape[1:]), dtype=src.dtype, device=src.device)
        pad[: src.shape[0]] = src
    flat = torch.empty((world * m, *src.shape[1:]), dtype=src.dtype, device=src.device)
    try:
        dist.all_gather_into_tensor(flat, pad, group=group)
    except (RuntimeError, NotImplementedError, AttributeError):      # backend without the flat variant
        bufs = list(flat.split(m, dim=0))
        dist.all_gather(bufs, pad, group=group)
    if all(c == m for c in counts):
        out = flat
    else:
        out = torch.cat([flat[r * m: r * m + c] for r, c in enumerate(counts)], dim=0)
    return out.to(torch.bool) if as_bool else out


def all_counts(n_local: int, device, group=None) -> List[int]:
    """all_gather of one int64 per rank (per-rank hit totals)."""
    world = dist.get_world_size(group)
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    bufs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(bufs, t, group=group)
    return [int(b.item()) for b in bufs]


def _align256(b: int) -> int:
    return (b + 255) // 256 * 256


def dense_layout(nray: int) -> dict:
    """Byte offsets of the dense closest-hit sections for `nray` rays (each 256-byte aligned): hit u8, front u8,
    tri i32, loc f32x3, uv f32x2; 'bytes' = total."""
    off, out = 0, {}
    for name, per_ray in (("hit", 1), ("front", 1), ("tri", 4), ("loc", 12), ("uv", 8)):
        out[name] = off
        off += _align256(per_ray * nray)
    out["bytes"] = off
    return out


def packed_layout(capacity: int, nray: int) -> dict:
    """Byte offsets of the packed-result sections for up to `capacity` hits (ray index reserved at 8 bytes) plus
    the dense hit mask of `nray` rays."""
    off, out = 0, {}
    for name, size in (("ray", 8 * capacity), ("loc", 12 * capacity), ("uv", 8 * capacity), ("tri", 4 * capacity),
                       ("front", capacity), ("hit", nray)):
        out[name] = off
        off += _align256(size)
    out["bytes"] = off
    return out


class PeerOutputs:
    """Dense closest-hit outputs for `nray` rays in SYMMETRIC memory (torch.distributed._symmetric_memory): every
    rank allocates the same buffer and learns the peer-mapped address of every other rank's copy, so a rank's
    trace kernel can store its slice of the results directly into the root's buffer over NVLink - the gather is
    fused into the kernel's own stores, there is no collective afterwards (only a barrier)."""

    def __init__(self, nray: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.nray = int(nray)
        lay = dense_layout(self.nray)
        self.off_hit, self.off_front, self.off_tri = lay["hit"], lay["front"], lay["tri"]
        self.off_loc, self.off_uv, self.bytes = lay["loc"], lay["uv"], lay["bytes"]
        self.buf = symm_mem.empty(self.bytes, dtype=torch.uint8, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]

    def addresses(self, rank: int, first_ray: int):
        """Raw addresses of ray `first_ray`'s hit / front / tri / loc / uv slots in `rank`'s buffer."""
        b = self.ptrs[rank]
        return (b + self.off_hit + first_ray, b + self.off_front + first_ray, b + self.off_tri + 4 * first_ray,
                b + self.off_loc + 12 * first_ray, b + self.off_uv + 8 * first_ray)

    def peer_slices(self, rank: int, lo: int, hi: int):
        """Tensors aliasing rays [lo, hi) of the five sections in `rank`'s buffer (peer-mapped memory)."""
        m, g = hi - lo, self.handle.get_buffer
        return (g(rank, (m,), torch.uint8, self.off_hit + lo), g(rank, (m,), torch.uint8, self.off_front + lo),
                g(rank, (m,), torch.int32, self.off_tri // 4 + lo), g(rank, (3 * m,), torch.float32, self.off_loc // 4 + 3 * lo),
                g(rank, (2 * m,), torch.float32, self.off_uv // 4 + 2 * lo))

    def local_views(self, batch):
        n, u = self.nray, self.buf
        return (u[self.off_hit:self.off_hit + n].view(torch.bool).reshape(batch),
                u[self.off_front:self.off_front + n].view(torch.bool).reshape(batch),
                u[self.off_tri:self.off_tri + 4 * n].view(torch.int32).reshape(batch),
                u[self.off_loc:self.off_loc + 12 * n].view(torch.float32).reshape(*batch, 3),
                u[self.off_uv:self.off_uv + 8 * n].view(torch.float32).reshape(*batch, 2))


class PeerPacked:
    """Packed (variable-length) results for up to `capacity` hits plus a dense hit mask for `nray` rays, in symmetric
    memory: the ranks of a sharded job scatter their hits straight into the root's copy at their global row offset
    (rt_compact_scatter_at / rt_allhits_scatter_at with peer addresses)."""

    def __init__(self, capacity: int, nray: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.capacity, self.nray = int(capacity), int(nray)
        lay = packed_layout(self.capacity, self.nray)
        self.off_ray, self.off_loc, self.off_uv, self.off_tri = lay["ray"], lay["loc"], lay["uv"], lay["tri"]
        self.off_front, self.off_hit, self.bytes = lay["front"], lay["hit"], lay["bytes"]
        self.buf = symm_mem.empty(self.bytes, dtype=torch.uint8, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]

    def addresses(self, rank: int, row: int, ray_bytes: int):
        """(front, ray_idx, tri, loc, uv) addresses of packed row `row` in `rank`'s buffer."""
        b = self.ptrs[rank]
        return (b + self.off_front + row, b + self.off_ray + ray_bytes * row, b + self.off_tri + 4 * row,
                b + self.off_loc + 12 * row, b + self.off_uv + 8 * row)

    def peer_hit_mask(self, rank: int, lo: int, hi: int) -> torch.Tensor:
        return self.handle.get_buffer(rank, (hi - lo,), torch.uint8, self.off_hit + lo)

    def peer_rows(self, rank: int, name: str, row: int, rows: int, ray_bytes: int = 4) -> torch.Tensor:
        """Tensor aliasing rows [row, row + rows) of section `name` in `rank`'s buffer (peer-mapped memory)."""
        off, dt, width = {"front": (self.off_front, torch.uint8, 1), "tri": (self.off_tri, torch.int32, 1),
                          "ray": (self.off_ray, torch.int64 if ray_bytes == 8 else torch.int32, 1),
                          "loc": (self.off_loc, torch.float32, 3), "uv": (self.off_uv, torch.float32, 2)}[name]
        item = torch.empty(0, dtype=dt).element_size()
        return self.handle.get_buffer(rank, (rows * width,), dt, off // item + row * width)

    def local_views(self, h: int, ray_bytes: int, batch):
        u = self.buf
        ray_dt = torch.int64 if ray_bytes == 8 else torch.int32
        return dict(hit=u[self.off_hit:self.off_hit + self.nray].view(torch.bool).reshape(batch),
                    front=u[self.off_front:self.off_front + h].view(torch.bool),
                    ray=u[self.off_ray:self.off_ray + ray_bytes * h].view(ray_dt),
                    tri=u[self.off_tri:self.off_tri + 4 * h].view(torch.int32),
                    loc=u[self.off_loc:self.off_loc + 12 * h].view(torch.float32).reshape(h, 3),
                    uv=u[self.off_uv:self.off_uv + 8 * h].view(torch.float32).reshape(h, 2))


class ShardedRayMeshIntersector:
    """Wraps a per-rank intersector (anything with the RayMeshIntersector query methods).

    Every method takes the FULL ray batch (replicated on all ranks, or at least the rank's own
    slice valid) and traces only this rank's slice.  With gather=True every rank returns the
    full result, identical to a single-GPU call; with gather=False it returns its slice plus the
    slice bounds.
    """

    def __init__(self, local, group=None):
        self.local = local
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    # -- construction ------------------------------------------------------------
    @classmethod
    def build(cls, vertices: torch.Tensor, faces: torch.Tensor, src: int = 0, group=None):
        """Build the BVH on `src`, broadcast the blob, attach it on every rank."""
        from triro.ray.ray_optix import OptixAccelStructureWrapper, RayMeshIntersector

        rank = dist.get_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        if rank == src:
            local = RayMeshIntersector(vertices=vertices, faces=faces)
            blob = broadcast_blob(local.as_wrapper._inner.used().contiguous(), src, group, dev)
        else:
            blob = broadcast_blob(None, src, group, dev)
            local = RayMeshIntersector.__new__(RayMeshIntersector)
            local.as_wrapper = OptixAccelStructureWrapper()
            local.as_wrapper._inner.adopt(blob)
            local.mesh_vertices = None
            local.mesh_faces = None
        # mesh AABB over all vertices is needed by contains_points: broadcast 6 floats
        aabb = torch.zeros(6, dtype=torch.float32, device=dev)
        if rank == src:
            aabb[:3] = local.mesh_aabb[0]; aabb[3:] = local.mesh_aabb[1]
        dist.broadcast(aabb, src=src, group=group)
        if rank != src:
            local.mesh_aabb = (aabb[:3].clone(), aabb[3:].clone())
            local._aabb_host = (aabb[:3].tolist(), aabb[3:].tolist())
        return cls(local, group)

    # -- helpers -----------------------------------------------------------------
    def _slice(self, origins, directions):
        n = origins.numel() // 3
        lo, hi = shard_bounds(n, self.world, self.rank)
        o, d = slice_rays(origins, directions, lo, hi)
        return n, lo, hi, o, d

    def _counts(self, n):
        return [shard_bounds(n, self.world, r)[1] - shard_bounds(n, self.world, r)[0] for r in range(self.world)]

    def _dense(self, fn: Callable, origins, directions, gather: bool):
        batch = tuple(origins.shape[:-1])
        n, lo, hi, o, d = self._slice(origins, directions)
        res = fn(o, d)
        single = not isinstance(res, tuple)
        res = (res,) if single else res
        if not gather:
            return (res[0] if single else res), (lo, hi)
        counts = self._counts(n)
        full = tuple(gather_fixed(x, counts, self.group).reshape(*batch, *x.shape[1:]) for x in res)
        return full[0] if single else full

    # -- queries -----------------------------------------------------------------
    def intersects_any(self, origins, directions, gather: bool = True):
        return self._dense(self.local.intersects_any, origins, directions, gather)

    def intersects_first(self, origins, directions, gather: bool = True):
        return self._dense(self.local.intersects_first, origins, directions, gather)

    def intersects_count(self, origins, directions, gather: bool = True):
        return self._dense(self.local.intersects_count, origins, directions, gather)

    def intersects_closest(self, origins, directions, stream_compaction: bool = False, gather: bool = True):
        if not stream_compaction:
            return self._dense(lambda o, d: self.local.intersects_closest(o, d), origins, directions, gather)
        batch = tuple(origins.shape[:-1])
        n, lo, hi, o, d = self._slice(origins, directions)
        hit, front, ray_idx, tri_idx, loc, uv = self.local.intersects_closest(o, d, stream_compaction=True)
        ray_idx = ray_idx + lo                                   # rebase to the global ray numbering
        if not gather:
            return (hit, front, ray_idx, tri_idx, loc, uv), (lo, hi)
        hit_full = gather_fixed(hit, self._counts(n), self.group).reshape(batch)
        hc = all_counts(front.shape[0], front.device, self.group)
        return (hit_full, gather_fixed(front, hc, self.group), gather_fixed(ray_idx, hc, self.group),
                gather_fixed(tri_idx, hc, self.group), gather_fixed(loc, hc, self.group), gather_fixed(uv, hc, self.group))

    TO_ROOT_MIN_CHUNK = 1 << 20      # rays per window of the trace / copy pipeline (at most 8 windows)
    _side_stream = None

    def intersects_closest_to_root(self, origins, directions, root: int = 0, outputs: "PeerOutputs | None" = None,
                                   kernel_stores: "bool | None" = None, chunks: "int | None" = None):
        """Fused trace + gather: every rank traces its ray slice and its kernel stores the results straight into
        `root`'s output tensors over NVLink (peer stores from inside k_trace; no all-gather).  Returns the dense
        5-tuple of `intersects_closest` on `root` (views of `outputs`, valid until the next call that reuses it)
        and None elsewhere.  Pass a `PeerOutputs` to reuse the symmetric allocation across calls.
        `kernel_stores=False` traces into local tensors and moves them with peer-to-peer copies instead (in `chunks`
        ray windows, a window's copies overlapping the next window's trace; default: 1 Mi-ray windows, at most 8); the default
        picks kernel stores for 2 ranks (transfer hidden under the traversal: 0.74 vs 0.95 ms per 4K frame) and copies
        beyond (many ranks' small stores contend at the root: 8 ranks 0.83 vs 0.75 ms)."""
        from triro.backend import ops as hops

        batch = tuple(origins.shape[:-1])
        n = origins.numel() // 3
        lo, hi = shard_bounds(n, self.world, self.rank)      # the slice is a ray WINDOW of the full batch: no copy
        if outputs is None or outputs.nray != n:
            outputs = PeerOutputs(n, origins.device, self.group)
        if kernel_stores is None:
            kernel_stores = self.world <= 2
        if hi > lo:
            if kernel_stores:
                hops.intersects_closest_into(self.local.as_wrapper, origins, directions, *outputs.addresses(root, lo),
                                             ray_first=lo, ray_count=hi - lo)
            else:
                # trace into local tensors and move them with bulk peer-to-peer copies; the slice is traced in `chunks`
                # ray windows and a window's five copies run on a side stream while the next window is traced (the
                # root's NVLink ingress is the bound: 8 ranks, 66 M rays 2.96 -> 2.34 ms, profiles/r2_strong_probe_n8.json)
                cur = torch.cuda.current_stream()
                if chunks is None:
                    chunks = min(8, (hi - lo) // self.TO_ROOT_MIN_CHUNK) if self.world > 2 else 1     # 8 ranks, 66 M rays: 2.96 / 2.58 / 2.40 / 2.34 / 2.69 ms with 1 / 2 / 4 / 8 / 16
                chunks = max(1, min(int(chunks), hi - lo))
                if chunks > 1 and self._side_stream is None:
                    self._side_stream = torch.cuda.Stream(device=origins.device)
                step = -(-(hi - lo) // chunks)
                for c0 in range(lo, hi, step):
                    c1 = min(c0 + step, hi)
                    res = hops.intersects_closest(self.local.as_wrapper, origins, directions, c0, c1 - c0)
                    copy_stream = cur if chunks == 1 else self._side_stream
                    if chunks > 1:
                        copy_stream.wait_stream(cur)
                    with torch.cuda.stream(copy_stream):
                        for dst, src in zip(outputs.peer_slices(root, c0, c1), res):
                            dst.copy_(src.reshape(-1).view(dst.dtype) if src.dtype == torch.bool else src.reshape(-1))
                            src.record_stream(copy_stream)
                if chunks > 1:
                    cur.wait_stream(self._side_stream)
        torch.cuda.current_stream().synchronize()      # this rank's stores have left; then everybody's have
        dist.barrier(group=self.group)
        self._peer_outputs = outputs
        return outputs.local_views(batch) if self.rank == root else None

    def _packed_for(self, packed, hits: int, nray: int, device):
        # every rank sees the same (hits, nray), so all take the same branch of this collective allocation
        if packed is None or packed.capacity < hits or packed.nray != nray:
            packed = PeerPacked(max(1024, 1 << max(hits - 1, 1).bit_length()), nray, device, self.group)
        return packed

    def intersects_closest_compact_to_root(self, origins, directions, root: int = 0, packed: "PeerPacked | None" = None,
                                           scatter_to_peer: bool = False):
        """`intersects_closest(stream_compaction=True)` of the full batch, assembled on `root` without an NCCL
        collective on the data path: every rank traces, scans and packs its slice (ray indices in the global
        numbering, int64 when the batch exceeds 2^31 rays), the per-rank hit totals are exchanged (one int64 each),
        and each rank copies its packed arrays into the root's tensors at its global row offset with plain
        peer-to-peer copies over NVLink - no padding, no concatenation pass.  `scatter_to_peer=True` instead lets the
        scatter kernel store into the root's memory directly (measured slower: small strided NVLink stores).
        Returns the 6-tuple on `root` (views of `packed`, kept in `last_packed`), None elsewhere."""
        from triro.backend import ops as hops

        batch = tuple(origins.shape[:-1])
        n = origins.numel() // 3
        lo, hi = shard_bounds(n, self.world, self.rank)
        dev = origins.device
        hit, front, tri, loc, uv = hops.intersects_closest(self.local.as_wrapper, origins, directions, lo, hi - lo)
        ws, total = hops.compact_scan(hit)
        rb = 8 if n > 2**31 - 1 else 4
        mine = None
        if not scatter_to_peer:
            mine = dict(front=torch.empty(total, dtype=torch.uint8, device=dev),
                        ray=torch.empty(total, dtype=torch.int64 if rb == 8 else torch.int32, device=dev),
                        tri=torch.empty(total, dtype=torch.int32, device=dev),
                        loc=torch.empty(3 * total, dtype=torch.float32, device=dev),
                        uv=torch.empty(2 * total, dtype=torch.float32, device=dev))
            if total > 0:
                hops.compact_scatter_at(hit, ws, front, tri, loc, uv, lo, rb, mine["front"].data_ptr(), mine["ray"].data_ptr(),
                                        mine["tri"].data_ptr(), mine["loc"].data_ptr(), mine["uv"].data_ptr())
        counts = all_counts(total, dev, self.group)
        hits, row0 = sum(counts), sum(counts[: self.rank])
        packed = self._packed_for(packed, hits, n, dev)
        if total > 0:
            if scatter_to_peer:
                hops.compact_scatter_at(hit, ws, front, tri, loc, uv, lo, rb, *packed.addresses(root, row0, rb))
            else:
                for name, t in mine.items():
                    packed.peer_rows(root, name, row0, total, rb).copy_(t)
        if hi > lo:
            packed.peer_hit_mask(root, lo, hi).copy_(hit.reshape(-1).view(torch.uint8))
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=self.group)
        self.last_packed = packed
        if self.rank != root:
            return None
        v = packed.local_views(hits, rb, batch)
        return v["hit"], v["front"], v["ray"], v["tri"], v["loc"], v["uv"]

    def intersects_location_to_root(self, origins, directions, root: int = 0, packed: "PeerPacked | None" = None):
        """`intersects_location` (all hits, <= max_hits per ray) of the full batch packed on `root` the same way:
        (loc[h,3], ray_idx[h], tri_idx[h]) on `root`, None elsewhere."""
        from triro.backend import ops as hops

        n = origins.numel() // 3
        lo, hi = shard_bounds(n, self.world, self.rank)
        dev = origins.device
        rb = 8 if n > 2**31 - 1 else 4
        # this rank's window of the batch, traced in bounded staging windows; ray indices are global already
        loc, ray, tri = hops.intersects_location(self.local.as_wrapper, origins, directions, getattr(self.local, "max_hits", 8),
                                                 0, torch.int64 if rb == 8 else torch.int32, None, lo, hi - lo)
        total = int(tri.shape[0])
        mine = dict(loc=loc.reshape(-1), ray=ray, tri=tri)
        counts = all_counts(total, dev, self.group)
        hits, row0 = sum(counts), sum(counts[: self.rank])
        packed = self._packed_for(packed, hits, n, dev)
        if total > 0:
            for name, t in mine.items():
                packed.peer_rows(root, name, row0, total, rb).copy_(t)
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=self.group)
        self.last_packed = packed
        if self.rank != root:
            return None
        v = packed.local_views(hits, rb, (n,))
        return v["loc"], v["ray"], v["tri"]

    def intersects_location(self, origins, directions, gather: bool = True):
        n, lo, hi, o, d = self._slice(origins, directions)
        loc, ray_idx, tri_idx = self.local.intersects_location(o, d)
        ray_idx = ray_idx + lo
        if not gather:
            return (loc, ray_idx, tri_idx), (lo, hi)
        hc = all_counts(loc.shape[0], loc.device, self.group)
        return gather_fixed(loc, hc, self.group), gather_fixed(ray_idx, hc, self.group), gather_fixed(tri_idx, hc, self.group)

    def intersects_id(self, origins, directions, return_locations: bool = False, multiple_hits: bool = True,
                      gather: bool = True):
        """(tri_idx, ray_idx[, loc]) like RayMeshIntersector.intersects_id; with gather=False this rank's part (ray
        indices in the global numbering) plus its slice bounds."""
        if multiple_hits:
            res = self.intersects_location(origins, directions, gather=gather)
            (loc, ray_idx, tri_idx), bounds = res if not gather else (res, None)
        else:
            res = self.intersects_closest(origins, directions, stream_compaction=True, gather=gather)
            (_, _, ray_idx, tri_idx, loc, _), bounds = res if not gather else (res, None)
        out = (tri_idx, ray_idx, loc) if return_locations else (tri_idx, ray_idx)
        return out if gather else (out, bounds)

    def contains_points(self, points, check_direction=None, gather: bool = True):
        """`contains_points` of the full point batch, sharded.  The two decisions the reference takes on the WHOLE
        batch (any point inside the AABB? any point broken?, ray_optix.py:243,269) are OR-reduced over the ranks and
        the random retry direction is drawn on rank 0 and broadcast, so N ranks return what one process returns -
        including the reference's all-False quirk branches."""
        from triro.ray.ray_optix import contains_points_flow

        n = points.numel() // 3
        lo, hi = shard_bounds(n, self.world, self.rank)
        p = points.reshape(-1, 3)[lo:hi]

        def reduce_flags(flags):
            f = flags.to(torch.int32).clone()
            dist.all_reduce(f, op=dist.ReduceOp.MAX, group=self.group)
            return tuple(int(x) for x in f.tolist())

        def draw_direction():
            d = (torch.rand(3) - 0.5).to(p.device) if self.rank == 0 else torch.empty(3, device=p.device)
            dist.broadcast(d, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            return d.tolist()

        res = contains_points_flow(self.local.contains_parity, p, check_direction, reduce_flags, draw_direction)
        if not gather:
            return res, (lo, hi)
        return gather_fixed(res, self._counts(n), self.group).reshape(points.shape[:-1])
