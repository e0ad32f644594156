"""The C-ABI library loads without a GPU, exports every symbol include/raymesh_b200.h declares,
and its size / validation entry points behave; compute entry points fail loudly (no CPU path)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "raymesh_b200.h")


@pytest.fixture(scope="module")
def lib():
    from triro.backend import build, ops

    build.build()                      # nvcc cross-compiles sm_100a without a GPU
    return ops.get_module()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src)))


def test_header_functions_are_all_exported(lib):
    names = declared_functions()
    assert len(names) >= 20 and "rt_trace_closest" in names and "rt_bvh_build" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/raymesh_b200.h but not exported"
    from triro.backend import ops

    assert set(ops.EXPORTS) == set(names), "ctypes signature table and header disagree"


def test_struct_layouts_match_the_header():
    from triro.backend import ops

    assert C.sizeof(ops.RayDesc) == 8 + 32 + 8 + 32 + 8 + 32          # rt_ray_desc
    assert C.sizeof(ops.TraceOpts) == 4 + 4 + 8 + 8 + 4 * 4 + 16     # rt_trace_opts
    assert ops.TraceOpts.ray_first.offset == 8 and ops.TraceOpts.flags.offset == 24
    o = ops.trace_opts(None, 5, 7)
    assert (o.tmax, o.ray_first, o.ray_count, o.flags & 1) == (1.0e7, 5, 7, 1) and ops.trace_opts().ray_count == 0     # reference tmax: shaders.cu:86
    hdr = bytes(range(256))
    parsed = ops.parse_header(hdr)
    assert parsed["magic"] == int.from_bytes(hdr[0:4], "little") and parsed["n_tris"] == int.from_bytes(hdr[8:12], "little")
    assert parsed["tris_offset"] == int.from_bytes(hdr[24:32], "little") and parsed["used_bytes"] == int.from_bytes(hdr[40:48], "little")
    assert parsed["bad_index_faces"] == int.from_bytes(hdr[72:76], "little")


def test_python_constants_match_the_header_defines():
    """Schedules, option flags and the ABI version the ctypes host uses are the header's #defines."""
    from triro.backend import ops

    src = open(HEADER).read()
    defs = {m.group(1): int(m.group(2).rstrip("u"), 0) for m in re.finditer(r"#define\s+(RT_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+u?|\d+u?)\b", src)}
    assert defs["RT_ABI_VERSION"] == ops.ABI_VERSION
    for name in ("AUTO", "DIRECT", "QUEUED", "COOP_COHERENT", "COOP_INCOHERENT", "SLOTS"):
        assert defs["RT_SCHED_" + name] == getattr(ops, "SCHED_" + name), name
    for name in ("SCRATCH_ZEROED", "STOP_WHEN_BROKEN", "NO_LANE_SHARING", "NO_TILE_ORDER"):
        assert defs["RT_OPT_" + name] == getattr(ops, "OPT_" + name), name
    flags = [defs[k] for k in defs if k.startswith("RT_OPT_")]
    assert len(set(flags)) == len(flags) and all(f & (f - 1) == 0 for f in flags)      # distinct single bits
    old = ops.set_knobs(no_lane_sharing=1, no_tile_order=1)
    try:
        assert ops.trace_opts().flags & (ops.OPT_NO_LANE_SHARING | ops.OPT_NO_TILE_ORDER) == ops.OPT_NO_LANE_SHARING | ops.OPT_NO_TILE_ORDER
    finally:
        ops.set_knobs(**old)
    assert ops.trace_opts().flags & (ops.OPT_NO_LANE_SHARING | ops.OPT_NO_TILE_ORDER) == 0


def test_size_queries_and_argument_validation(lib):
    ws, blob = C.c_size_t(), C.c_size_t()
    assert lib.rt_bvh_sizes(100, 200, C.byref(ws), C.byref(blob)) == 0
    assert blob.value >= 256 + 200 * 48 + (200 // 3 + 2) * 80 and ws.value > 200 * 12
    assert lib.rt_bvh_sizes(-1, 0, C.byref(ws), C.byref(blob)) == -1 and b"negative" in lib.rt_last_error()
    assert lib.rt_bvh_sizes(0, 0, C.byref(ws), C.byref(blob)) == 0 and blob.value >= 256 + 80
    st, w2 = C.c_size_t(), C.c_size_t()
    assert lib.rt_allhits_sizes(1000, 8, C.byref(st), C.byref(w2)) == 0 and st.value == 1000 * 8 * 16
    assert lib.rt_allhits_sizes(1000, 65, C.byref(st), C.byref(w2)) == -1          # RT_MAX_HITS_LIMIT = 64 (reference default 8)
    assert lib.rt_compact_sizes(5000, C.byref(w2)) == 0 and w2.value >= 256 + 2 * 8 * 3
    assert lib.rt_sort_sizes(10_000, C.byref(w2)) == 0 and w2.value >= 10_000 * 12
    assert lib.rt_abi_version() == 2
    assert not hasattr(lib, "rt_set_tmax"), "tmax is a per-call option (rt_trace_opts), not library state"
    st_ref = C.c_size_t()
    assert lib.rt_bvh_refit_sizes(200, C.byref(st_ref)) == 0 and st_ref.value >= (5 * 200 // 8) * 36


def test_compute_entry_points_fail_loudly_without_a_device(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from triro.backend import ops

    assert lib.rt_device_sm_count() == 0
    rd = ops.RayDesc()
    rd.nray = 1
    for i, s in enumerate((1, 1, 1, 3)):
        rd.shape[i] = s
    buf = (C.c_uint8 * 1024)()
    rd.origins = C.addressof(buf); rd.directions = C.addressof(buf)
    rc = lib.rt_trace_any(C.addressof(buf), C.byref(rd), None, C.addressof(buf), C.addressof(buf), None)
    assert rc == -2 and b"no CUDA device" in lib.rt_last_error()
    opts = ops.trace_opts(None, 2, 1)
    assert lib.rt_trace_any(C.addressof(buf), C.byref(rd), C.byref(opts), C.addressof(buf), C.addressof(buf), None) == -1
    assert b"ray_first" in lib.rt_last_error()                                  # window outside the 1-ray batch
    rc = lib.rt_bvh_build(C.addressof(buf), 3, C.addressof(buf), 1, None, 0, None, 0, None)
    assert rc == -1                                                            # null blob/workspace
    with pytest.raises((RuntimeError, AssertionError)):
        from triro.ray.ray_optix import RayMeshIntersector

        RayMeshIntersector(vertices=torch.zeros(3, 3), faces=torch.zeros(1, 3, dtype=torch.int32))


def test_product_never_imports_the_oracle():
    """The package must not reference oracle/ or tests/hostsim (no CPU fallback on the product path)."""
    pkg = os.path.join(ROOT, "trimesh-ray-optix_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                for needle in ("import oracle", "from oracle", "liboracle", "libhostsim", "import hostsim"):
                    assert needle not in text, f"{fn} mentions {needle}"
