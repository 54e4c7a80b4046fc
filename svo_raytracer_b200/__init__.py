"""svo_raytracer_b200 -- B200-native (sm_100a) SVO ray-traversal path.

Drop-in for the one data-parallel hot path of dyoo47/svo-raytracer
(src/shaders/svotrace.comp behind src/engine/Renderer.java).  See DESIGN.md.

(The directory is ``svo_raytracer_b200`` -- Python package names cannot hold
the hyphen of ``svo-raytracer``.)
"""
from . import _lib  # noqa: F401
from ._lib import NO_HIT, SvoError, Frame  # noqa: F401
from .cameras import CAMERAS, camera_frame  # noqa: F401
from .renderer import (HIT_DTYPE, RAY_DTYPE, Renderer, Shader, SvoContext, build_terrain,  # noqa: F401
                       make_frame)
from .terrain import terrain_inputs  # noqa: F401
