"""nu_nerf_b200 -- B200-native (sm_100a) implementation of the NU-NeRF volume-rendering hot path.

Public surface (mirrors the reference's network/renderer_zerothick.py and its tracer objects):
    from nu_nerf_b200.renderer_zerothick import NeROShapeRenderer, name2renderer      # network/renderer_zerothick.py
    from nu_nerf_b200.renderer import name2renderer as name2renderer_nonzero_thickness  # network/renderer.py
    from nu_nerf_b200.tracer import optix_mesh, RayTracer, Scene
The CUDA engine is nu_nerf_b200/libnunerf_b200.so (C-ABI in include/nunerf.h); importing the engine modules
raises ImportError when it has not been built -- there is no CPU or eager fallback.
"""
__version__ = "0.1.0"
