"""Camera presets used by tests and bench (corner rays are passed directly, as
the shader receives them: uniforms 1-4 and 8, Main.java:269-273).

A is the reference's start-up pose (Camera.java:13-18 corner rays
(-+1.6, -+0.9, -1); Main.java:120 position (1.5, 1.5, 2.0)); B and C are the
two other poses SURVEY.md section 8d fixes.
"""
from __future__ import annotations

from .renderer import make_frame

_CORNERS_FWD = ((-1.6, -0.9, -1.0), (-1.6, 0.9, -1.0), (1.6, -0.9, -1.0), (1.6, 0.9, -1.0))
_CORNERS_DOWN = ((-0.8, -1.0, -0.45), (-0.8, -1.0, 0.45), (0.8, -1.0, -0.45), (0.8, -1.0, 0.45))

CAMERAS = {
    "A": ((1.5, 1.5, 2.0),) + _CORNERS_FWD,
    "B": ((1.5, 1.3, 2.0),) + _CORNERS_FWD,
    "C": ((1.5, 1.6, 1.5),) + _CORNERS_DOWN,
}


def camera_frame(name: str, **kw):
    """svo_frame for a named camera; kw as in make_frame (render_mode, frame_number, ...)."""
    pos, l1, l2, r1, r2 = CAMERAS[name]
    return make_frame(pos, l1, l2, r1, r2, **kw)
