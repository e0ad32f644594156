"""triro — B200-native drop-in for lcp29/trimesh-ray-optix (same package and module names).

    from triro.ray.ray_optix import RayMeshIntersector

The reference's OptiX backend is replaced by hand-written sm_100a CUDA behind a C ABI
(include/raymesh_b200.h); see DESIGN.md.
"""
__version__ = "1.3.1+b200.1"
