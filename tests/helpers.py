"""Shared helpers of the parity tests (test infrastructure)."""
from __future__ import annotations

import numpy as np

from oracle import oracle


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype.kind == "f":
        bad = a.view(np.uint32) != b.view(np.uint32)
    else:
        bad = a != b
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} elements differ (first at {np.argwhere(bad)[0]})"


def closest_to_numpy(res):
    hit, front, tri, loc, uv = res
    return dict(hit=hit.cpu().numpy().reshape(-1).astype(np.uint8), front=front.cpu().numpy().reshape(-1).astype(np.uint8),
                tri=tri.cpu().numpy().reshape(-1), loc=loc.cpu().numpy().reshape(-1, 3), uv=uv.cpu().numpy().reshape(-1, 2))


def check_closest_vs_mirror(got: dict, mesh: oracle.OracleMesh, o, d):
    """Bit-exact comparison against the oracle's binary32 mirror (every ray, no exclusions)."""
    ref = oracle.query(mesh, o, d, oracle.MIRROR, closest_only=True, want=("hit", "front", "tri", "loc", "uv"))
    for k in ("hit", "front", "tri", "loc", "uv"):
        assert_bits_equal(got[k], ref[k], f"closest.{k} vs mirror")
    return ref


def check_closest_vs_truth(got: dict, mesh: oracle.OracleMesh, o, d, max_graze_fraction=0.02, rel_tol=1e-5):
    """Comparison against the binary64 truth: masks and indices exact outside the documented
    grazing set, locations / uv within rel_tol (relative to the scene scale, as north_star says
    1e-5 relative)."""
    ref = oracle.query(mesh, o, d, oracle.TRUTH, want=("hit", "front", "tri", "loc", "uv", "flags", "t"))
    graze = ref["flags"] != 0
    assert graze.mean() <= max_graze_fraction, f"grazing set too large: {graze.mean():.4f}"
    clean = ~graze
    for k in ("hit", "front", "tri"):
        bad = (got[k] != ref[k]) & clean
        assert not bad.any(), f"closest.{k}: {int(bad.sum())} non-grazing rays differ from the binary64 truth"
    both = clean & (ref["hit"] == 1) & (got["hit"] == 1)
    if both.any():
        scale = max(1.0, float(np.abs(mesh.vertices).max()))
        err_loc = np.abs(got["loc"][both] - ref["loc"][both]).max() / scale
        err_uv = np.abs(got["uv"][both] - ref["uv"][both]).max()
        assert err_loc <= rel_tol, f"location error {err_loc:.3e} > {rel_tol}"
        # uv are O(1) weights; tolerance scales with ray length / triangle size (binary32 conditioning)
        assert err_uv <= 2e-3, f"uv error {err_uv:.3e}"
    return ref, graze
