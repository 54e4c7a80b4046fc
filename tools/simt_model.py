"""SIMT divergence model of the traversal loop (no GPU needed).

Replays the bench frames warp by warp on the host build of the device source (tests/hostemu) and prints the issue
slots the traversal needs under different loop organisations, with per-path instruction counts read off the SASS of
k_render_tile<0,0,1> (cuobjdump -sass).  Used to rank kernel designs offline; the `if_if` row is the shipped loop and
can be compared with ncu's smsp__inst_executed.sum of the same frame (profiles/r01_tile_full.md).

  python tools/simt_model.py --size 2048 [--width 1920 --height 1080] [--cams A,B,C] [--costs head,push,adv,pop,tail,xh,xm,outside]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--cams", default="A,B,C")
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--casts", type=int, default=2)
    ap.add_argument("--box", type=int, default=1)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    # SASS of k_render_tile<0,0,1>: loop head 14, PUSH 40, ADVANCE 18, POP 34, BSYNC+BRA 2, exits ~8; code outside the loop per warp cast ~700
    ap.add_argument("--costs", default="14,40,18,34,2,8,10,700")
    a = ap.parse_args()
    import svo_raytracer_b200 as svo
    from hostemu import emu as E
    from oracle import oracle as O
    t0 = time.time()
    hm, mm = svo.terrain_inputs(a.size, nthreads=a.threads)
    nodes = svo.build_terrain(hm, mm, a.size, min(a.size, 1024), nthreads=a.threads)
    t1 = time.time()
    sc = E.Scene(nodes)
    t2 = time.time()
    print("world %d^3: %d bytes (build %.1f s), %d descriptors (transcode %.1f s)" % (a.size, nodes.size, t1 - t0, sc.ndesc, t2 - t1), file=sys.stderr)
    depth = min(13, int(np.log2(a.size)))
    costs = [float(v) for v in a.costs.split(",")]
    rows = []
    for i, cam in enumerate(a.cams.split(",")):
        pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
        f = O.make_frame(pos, l1, l2, r1, r2, frame_number=i + 1, render_mode=a.mode, max_depth=depth, casts=a.casts, cone_depth=11)
        r = sc.simt(f, a.width, a.height, costs, box=bool(a.box), nthreads=a.threads)
        r["camera"] = cam
        rows.append(r)
        print(json.dumps(r))
    tot = {k: sum(r[k] for r in rows) for k in rows[0] if k != "camera"}
    print("\n| organisation | issue slots (M warp-instr) | vs shipped | lane utilisation |")
    print("|---|---|---|---|")
    for k in ("if_if", "while_while", "ww_1_1", "ww_inf_1", "ww_1_inf", "ww_4_2", "longest_lane", "ideal"):
        print("| %s | %.1f | %.3f | %.1f %% |" % (k, tot[k] / 1e6, tot[k] / tot["if_if"], 100.0 * tot["ideal"] / tot[k]))
    print("\niterations/cast %.1f  (PUSH %.1f %%, ADVANCE %.1f %%, POP %.1f %%); warp iterations issuing PUSH %.1f %%, ADVANCE %.1f %%, POP %.1f %%"
          % (tot["iters"] / tot["casts"], 100.0 * tot["pushes"] / tot["iters"], 100.0 * tot["advances"] / tot["iters"],
             100.0 * tot["pops"] / tot["iters"], 100.0 * tot["warp_iters_push"] / tot["warp_iters"],
             100.0 * tot["warp_iters_adv"] / tot["warp_iters"], 100.0 * tot["warp_iters_pop"] / tot["warp_iters"]))


if __name__ == "__main__":
    main()
