"""Size-independent properties of the results at (or near) the full BASELINE.json sizes, where the CPU oracle is
too slow: uv/location consistency, idempotence under re-tracing from the hit point, linearity of counts in a
duplicated mesh, any == (count > 0) == (first >= 0), compaction == boolean-mask semantics, analytic geometry."""
import numpy as np
import pytest
import torch

from triro import synth
from triro.ray.ray_optix import RayMeshIntersector

pytestmark = pytest.mark.gpu


def make(v, f):
    return RayMeshIntersector(vertices=torch.from_numpy(v), faces=torch.from_numpy(f))


def test_config2_full_size_closest_properties(cuda_device):
    """Config 2 exactly: icosphere subdiv 7 (327 680 triangles), 3840x2160 pinhole rays (8 294 400 rays)."""
    v, f = synth.icosphere(7)
    r = make(v, f)
    o, d = synth.pinhole_rays(3840, 2160, device=cuda_device)
    hit, front, tri, loc, uv = r.intersects_closest(o, d)
    assert hit.shape == (2160, 3840) and 0.33 < float(hit.float().mean()) < 0.345
    # analytic: the camera at (0,0,3) sees the unit sphere under the cone sin(theta) = 1/3; pixels well inside hit, well outside miss
    cosang = -d[..., 2]
    assert bool(hit[cosang > np.cos(np.arcsin(1 / 3) * 0.995)].all()) and not bool(hit[cosang < np.cos(np.arcsin(1 / 3) * 1.005)].any())
    assert bool(front[hit].all()) and not bool(front[~hit].any())
    assert bool((tri[~hit] == -1).all()) and float(loc[~hit].abs().sum()) == 0.0 and float(uv[~hit].abs().sum()) == 0.0
    nrm = loc[hit].norm(dim=1)
    assert float(nrm.min()) > 0.9995 and float(nrm.max()) < 1 + 1e-6                     # on the faceted unit sphere
    # location lies on the ray: cross(loc - o, d) ~ 0; and uv reconstructs it from the triangle's vertices
    rel = loc[hit] - o[hit]
    assert float(torch.linalg.cross(rel, d[hit]).norm(dim=1).max()) < 5e-6
    tv = r.mesh_vertices[r.mesh_faces[tri[hit].long()].long()]
    u = uv[hit]
    rec = u[:, :1] * tv[:, 0] + u[:, 1:] * tv[:, 1] + (1 - u[:, :1] - u[:, 1:]) * tv[:, 2]
    assert float((rec - loc[hit]).abs().max()) < 2e-6
    assert bool(((u >= -1e-6) & (u <= 1 + 1e-6)).all()) and bool((u.sum(dim=1) <= 1 + 1e-6).all())
    # consistency between the five query types
    assert torch.equal(r.intersects_first(o, d), tri) and torch.equal(r.intersects_any(o, d), hit)
    cnt = r.intersects_count(o, d)
    assert torch.equal(cnt > 0, hit) and bool((cnt[hit] == 2).float().mean() > 0.999)    # a closed sphere is crossed twice
    chit, cfront, cray, ctri, cloc, cuv = r.intersects_closest(o, d, stream_compaction=True)
    assert torch.equal(cray.long(), torch.nonzero(hit.reshape(-1)).reshape(-1)) and torch.equal(ctri, tri[hit])
    assert torch.equal(cloc, loc[hit]) and torch.equal(cuv, uv[hit]) and torch.equal(cfront, front[hit])
    # idempotence: re-tracing from just in front of the hit point along the same direction hits the same triangle
    o2 = loc[hit] - 1e-3 * d[hit]
    t2 = r.intersects_first(o2, d[hit])
    assert float((t2 == tri[hit]).float().mean()) > 0.9999


def test_count_is_linear_in_a_duplicated_mesh(cuda_device):
    """Linearity: tracing mesh U (mesh shifted far away) counts the same as mesh alone for rays that cannot reach the
    copy, and the union of two coincident copies doubles every count (1 M-triangle soup halves, 2 M rays)."""
    v, f = synth.triangle_soup(500_000, seed=21)
    r1 = make(v, f)
    o, d = synth.random_rays(2_000_000, seed=5, device=cuda_device, box=True)
    c1 = r1.intersects_count(o, d)
    v2 = np.concatenate([v, v]); f2 = np.concatenate([f, f + len(v)])
    r2 = make(v2, f2)                                                     # 1 M triangles, every triangle twice
    c2 = r2.intersects_count(o, d)
    assert torch.equal(c2, 2 * c1)
    loc, ri, ti = r2.intersects_location(o, d)
    assert loc.shape[0] == int(c2.clamp(max=8).sum()) and bool((ri[1:] >= ri[:-1]).all())
    # closest hit of the doubled mesh: same location, the smaller of the two coincident face indices wins
    h1, _, t1, l1, _ = r1.intersects_closest(o, d)
    h2, _, t2, l2, _ = r2.intersects_closest(o, d)
    assert torch.equal(h1, h2) and torch.equal(t1, t2) and torch.equal(l1, l2)


def test_config3_heightfield_any_count_properties(cuda_device):
    """Config 3 mesh (4 194 304-triangle heightfield) with 10 M of its random rays."""
    v, f = synth.heightfield(2048, 1024)
    r = make(v, f)
    o, d = synth.random_rays(10_000_000, seed=1234, device=cuda_device)
    anyh = r.intersects_any(o, d)
    cnt = r.intersects_count(o, d)
    assert torch.equal(cnt > 0, anyh)
    assert not bool(anyh[d[:, 2] > 0.0].any())                             # origins are above the terrain: upward rays miss
    # steep downward rays whose footprint stays inside the domain cross the function graph exactly once
    t_ground = (o[:, 2] + 0.1) / (-d[:, 2]).clamp(min=1e-6)
    inside = (d[:, 2] < -0.5) & ((o[:, :2] + d[:, :2] * t_ground[:, None]).abs().max(dim=1)[0] < 0.99)
    assert float((cnt[inside] == 1).float().mean()) > 0.9999 and int(inside.sum()) > 1_000_000
    hit, front, tri, loc, uv = r.intersects_closest(o[:2_000_000], d[:2_000_000])
    assert torch.equal(hit, anyh[:2_000_000])
    z_exact = torch.zeros_like(loc[hit][:, 0])
    x, y = loc[hit][:, 0].double(), loc[hit][:, 1].double()
    for k in range(4):
        fr = 2.0 ** k * np.pi
        z_exact = z_exact + (0.5 ** k) * (torch.sin(fr * x + 0.3 * k) * torch.cos(fr * y * 1.3 - 0.7 * k)).float()
    z_exact = z_exact * (0.1 / 1.875)
    assert float((loc[hit][:, 2] - z_exact).abs().max()) < 2e-4            # piecewise-linear terrain vs its closed form
