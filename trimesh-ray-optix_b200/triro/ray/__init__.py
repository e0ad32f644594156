"""Import-time initialisation, as in the reference (triro/ray/__init__.py:17-22): after
`import triro.ray` the backend is ready.  Here that means the C-ABI library has been loaded
and its exports resolved; there is no OptiX context, module, pipeline or SBT to create."""
import triro.backend.ops as hops

hops.init_optix()
hops.create_optix_context()
hops.create_optix_module()
hops.create_optix_pipelines()
hops.build_sbts()

__version__ = "1.3.1+b200.1"
