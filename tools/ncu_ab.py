#!/usr/bin/env python
"""One launch of each given kernel variant on one bench frame, for an ncu capture:
  ncu --set full --clock-control none -k regex:k_render_tile -c 3 -o gpurun_out/ab python tools/ncu_ab.py 8192 C 0,7,8"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import svo_raytracer_b200 as svo
from svo_raytracer_b200 import _lib as L

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cam = sys.argv[2] if len(sys.argv) > 2 else "C"
variants = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "0,7,8").split(",")]
hm, mm = svo.terrain_inputs(size)
nodes = svo.build_terrain(hm, mm, size, min(size, 1024))
depth = min(13, int(np.log2(size)))
with svo.SvoContext(1920, 1080) as ctx:
    ctx.upload(nodes)
    for k in variants:
        ctx.set_option(L.OPT_KERNEL, k)
        for i, c in enumerate(cam):  # e.g. "ABC": the three frames of the bench cycle
            ctx.render(svo.camera_frame(c, frame_number=i + 1, render_mode=0, max_depth=depth))
            ctx.sync()
print("done")
