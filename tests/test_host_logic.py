"""Host-side logic of the Python mirror that does not need a GPU: ray descriptor marshalling
(reference fillArray, ray.cpp:151-159), input validation (ray.cpp:104-123), shard arithmetic."""
import numpy as np
import pytest
import torch

from triro import distributed as tdist
from triro.backend import ops


def test_ray_desc_right_aligns_shape_and_strides_like_fill_array():
    o = torch.zeros(5, 7, 3)
    d = torch.zeros(7, 5, 3).transpose(0, 1)
    rd, batch = ops.make_ray_desc(o, d)
    assert batch == (5, 7) and rd.nray == 35
    assert list(rd.shape) == [1, 5, 7, 3]
    assert list(rd.o_stride) == [0, 21, 3, 1]
    assert list(rd.d_stride) == [0, 3, 15, 1]
    b = torch.tensor([0.0, 0.0, 3.0]).broadcast_to(4, 6, 2, 3)
    rd, batch = ops.make_ray_desc(b, b)
    assert list(rd.shape) == [4, 6, 2, 3] and list(rd.o_stride) == [0, 0, 0, 1] and rd.nray == 48
    rd, batch = ops.make_ray_desc(torch.zeros(3), torch.zeros(3))
    assert batch == () and rd.nray == 1 and list(rd.shape) == [1, 1, 1, 3]
    rd, _ = ops.make_ray_desc(torch.zeros(0, 3), torch.zeros(0, 3))
    assert rd.nray == 0


def test_ray_desc_rejects_what_the_reference_silently_mishandles():
    # > 3 batch dims: the reference silently mis-indexes (ray.cpp:151-159); here the surplus leading dims are merged
    rd, batch = ops.make_ray_desc(torch.zeros(2, 5, 2, 2, 3), torch.zeros(2, 5, 2, 2, 3))
    assert batch == (2, 5, 2, 2) and rd.nray == 40 and list(rd.shape) == [10, 2, 2, 3] and list(rd.o_stride) == [12, 6, 3, 1]
    with pytest.raises(ValueError):
        ops.make_ray_desc(torch.zeros(4, 2), torch.zeros(4, 2))                        # last dim != 3
    with pytest.raises(ValueError):
        ops.make_ray_desc(torch.zeros(4, 3), torch.zeros(5, 3))                        # shape mismatch


def test_tensor_input_check_mirrors_reference_conditions():
    with pytest.raises(ValueError, match="cuda"):
        ops.tensor_input_check(torch.zeros(2, 3))                                      # not on the GPU
    with pytest.raises(ValueError):
        ops.tensor_input_check("not a tensor")


def test_constructor_contract():
    from triro.ray.ray_optix import RayMeshIntersector

    with pytest.raises(ValueError, match="Either 'mesh' or 'vertices' and 'faces'"):
        RayMeshIntersector()
    with pytest.raises(ValueError):
        RayMeshIntersector(vertices=torch.zeros(3, 3))


@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (10, 3), (8_294_400, 8), (1_000_000_000, 8), (7, 8)])
def test_shard_bounds_partition_the_ray_index_space(n, world):
    prev = 0
    for r in range(world):
        lo, hi = tdist.shard_bounds(n, world, r)
        assert lo == prev and lo <= hi <= n
        prev = hi
    assert prev == n
    per = -(-n // world) if world else n
    assert all(tdist.shard_bounds(n, world, r)[1] - tdist.shard_bounds(n, world, r)[0] <= per for r in range(world))


def test_slice_rays_handles_strided_and_broadcast_batches():
    d = torch.arange(2 * 3 * 4 * 3, dtype=torch.float32).reshape(2, 3, 4, 3).transpose(0, 2)
    o = torch.tensor([1.0, 2.0, 3.0]).broadcast_to(d.shape)
    so, sd = tdist.slice_rays(o, d, 5, 17)
    assert so.shape == (12, 3) and torch.equal(sd, d.reshape(-1, 3)[5:17]) and torch.all(so == torch.tensor([1.0, 2.0, 3.0]))


def test_symmetric_buffer_layouts_are_aligned_and_disjoint():
    """Section offsets of the peer-memory result buffers (triro/distributed.py): 256-byte aligned, large enough,
    non-overlapping - the kernels and peer copies of the sharded gather write through raw addresses into them."""
    from triro.distributed import dense_layout, packed_layout

    for n in (0, 1, 255, 256, 257, 90_601, 8_294_400, 3_000_000_000):
        lay = dense_layout(n)
        need = dict(hit=n, front=n, tri=4 * n, loc=12 * n, uv=8 * n)
        order = sorted(need, key=lambda k: lay[k])
        for a, b in zip(order, order[1:] + ["bytes"]):
            assert lay[a] % 256 == 0 and lay[a] + need[a] <= lay[b]
    for cap, n in ((1024, 10), (1 << 20, 90_601), (1 << 28, 1_000_000_000)):
        lay = packed_layout(cap, n)
        need = dict(ray=8 * cap, loc=12 * cap, uv=8 * cap, tri=4 * cap, front=cap, hit=n)
        order = sorted(need, key=lambda k: lay[k])
        for a, b in zip(order, order[1:] + ["bytes"]):
            assert lay[a] % 256 == 0 and lay[a] + need[a] <= lay[b]


def test_blob_header_validation_refuses_truncated_and_corrupt_blobs():
    """ops.validate_header (used by build, adopt and load): what a kernel relies on is checked on the host first."""
    import struct

    import torch

    from triro.backend import ops

    def header(n_tris=10, n_nodes=3, depth=2, nodes_off=None, parents_off=None, used=None, overflow=0, magic=ops.BLOB_MAGIC, abi=ops.ABI_VERSION):
        nodes_off = 256 + 48 * n_tris + 16 if nodes_off is None else nodes_off
        parents_off = (nodes_off + 80 * n_nodes + 255) // 256 * 256 if parents_off is None else parents_off
        used = parents_off + 4 * n_nodes if used is None else used
        raw = bytearray(256)
        struct.pack_into("<6I", raw, 0, magic, abi, n_tris, n_nodes, depth, n_nodes + 2)
        struct.pack_into("<3Q", raw, 24, 256, nodes_off, used)
        struct.pack_into("<2I", raw, 72, 0, overflow)
        struct.pack_into("<Q", raw, 80, parents_off)
        return bytes(raw), used

    raw, used = header()
    ops.validate_header(ops.parse_header(raw), used)                           # consistent: accepted
    with pytest.raises(ValueError):
        ops.validate_header(ops.parse_header(raw), used - 1)                   # tensor shorter than used_bytes
    with pytest.raises(RuntimeError):
        ops.validate_header(ops.parse_header(header(depth=ops.MAX_DEPTH + 1)[0]), 1 << 20)   # deeper than the traversal stack
    with pytest.raises(RuntimeError):
        ops.validate_header(ops.parse_header(header(overflow=1)[0]), 1 << 20)
    with pytest.raises(ValueError):
        ops.validate_header(ops.parse_header(header(nodes_off=256 + 48 * 5)[0]), 1 << 20)    # nodes overlap the triangle records
    with pytest.raises(ValueError):
        ops.validate_header(ops.parse_header(header(parents_off=256 + 48 * 10 + 16 + 80)[0]), 1 << 20)   # parents overlap the nodes
    # adopt() on a host tensor applies the same checks (a blob straight from torch.load / a gloo broadcast)
    blob = torch.zeros(used, dtype=torch.uint8)
    blob[:256] = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    assert ops.AccelStructure().adopt(blob).header["n_nodes"] == 3
    with pytest.raises(ValueError):
        ops.AccelStructure().adopt(blob[: used // 2].clone())
    bad = blob.clone(); bad[0] = 0
    with pytest.raises(ValueError):
        ops.AccelStructure().adopt(bad)


def test_trace_opts_defaults_and_knobs_are_per_call():
    from triro.backend import ops

    o = ops.trace_opts()
    assert (o.tmax, o.ray_first, o.ray_count, o.schedule) == (1.0e7, 0, 0, ops._KNOBS["schedule"])
    old = ops.set_knobs(schedule=ops.SCHED_SLOTS, tri_threshold=48)
    try:
        o = ops.trace_opts(None, 7, 9)
        assert (o.schedule, o.tri_threshold, o.ray_first, o.ray_count) == (ops.SCHED_SLOTS, 48, 7, 9)
    finally:
        ops.set_knobs(**old)
    assert ops.trace_opts().schedule == old["schedule"]
    with pytest.raises(KeyError):
        ops.set_knobs(nonsense=1)
    assert ops.allhits_window_rays(8, 1 << 20) == (1 << 20) // 128 and ops.allhits_window_rays(64, 1) == 1024


def test_contains_second_walk_shortcuts_preserve_the_reference_decisions():
    """The fused contains kernel (rt_trace_coop.cuh, retirement of MODE == kContains) replaces the second count by what
    the first count leaves open: nothing when it is 0, 'any hit?' when it is even, the full count when it is odd, and it
    may take the two directions in either order.  Exhaustive check of that algebra against the reference's formulas
    (triro/ray/ray_optix.py:262-267): contain = inside & odd(c+) & odd(c-), broken = ~(odd & odd) & (c+ == 0 | c- == 0)."""
    def reference(cp, cm, inside):
        agree = (cp & 1) and (cm & 1)
        return bool(inside and agree), bool((not agree) and (cp == 0 or cm == 0))

    def fused(first, second, inside):
        if first == 0:
            seen = 0                          # no second walk
        elif first % 2 == 0:
            seen = 1 if second > 0 else 0     # any-hit walk: stops at the first hit
        else:
            seen = second                     # full count
        agree = (first & 1) and (seen & 1)
        return bool(inside and agree), bool((not agree) and (first == 0 or seen == 0))

    for cp in range(8):
        for cm in range(8):
            for inside in (False, True):
                assert fused(cp, cm, inside) == reference(cp, cm, inside)
                assert fused(cm, cp, inside) == reference(cp, cm, inside)      # nearer side first: either order
