"""Drop-in for the reference's ``neural_renderer`` package (neural_renderer/neural_renderer/__init__.py:1-13) on
librnr_b200.so: same function / class names and call signatures for the part the relighting hot path uses --
``load_obj``, ``projection``, ``vertices_to_faces``, ``vertex_attrs_to_faces``, ``lighting``, ``rasterize_rgbad`` and
``Renderer`` (camera_mode 'projection') -- plus the ``neural_renderer.cuda.*`` extension entry points.  The rasterizer is
forward-only here: no script of the reference differentiates through it ("Gradient flow").
"""
from .core import (load_obj, projection, vertices_to_faces, vertex_attrs_to_faces, lighting, look_at, look, perspective,
                   get_points_from_angles, rasterize_rgbad, rasterize, rasterize_silhouettes, rasterize_depth, Rasterize,
                   Renderer, raster_gbuffer)
from . import cuda  # noqa: F401

__version__ = '1.1.3'
name = 'neural_renderer'
