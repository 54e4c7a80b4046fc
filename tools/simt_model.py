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
    # SASS of the default kernel (variant 10, 16-byte stack entries): loop head 14, PUSH 40, ADVANCE 18, POP 31 (34 with the 8-byte
    # entries of variant 0), BSYNC+BRA 2, exits ~8; code outside the loop per warp cast ~700
    ap.add_argument("--costs", default="14,40,18,31,2,8,10,700")
    ap.add_argument("--shapes", type=int, default=0, help="1: also evaluate other warp tile shapes (4x8, 16x2, 32x1)")
    ap.add_argument("--out", default="", help="write the tables as markdown to this file")
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
    md = []
    md.append("workload: %d^3 terrain, %dx%d, render mode %d, casts %d, cameras %s, content box %s; costs (head,push,adv,pop,tail,exit_hit,exit_miss,outside) = %s"
              % (a.size, a.width, a.height, a.mode, a.casts, a.cams, "on" if a.box else "off", a.costs))
    md.append("")
    md.append("| camera | model: shipped loop (M warp-instr) | casts | iterations/cast |")
    md.append("|---|---|---|---|")
    for r in rows:
        md.append("| %s | %.1f | %d | %.1f |" % (r["camera"], r["if_if"] / 1e6, r["casts"], r["iters"] / r["casts"]))
    md.append("")
    md.append("| loop organisation | issue slots (M warp-instr) | vs shipped | lane utilisation |")
    md.append("|---|---|---|---|")
    label = {"if_if": "shipped: one iteration per lane per warp iteration", "while_while": "while-while (PUSH loop, then ADVANCE/POP loop)",
             "ww_1_1": "alternate one ADVANCE/POP step and one PUSH step", "ww_inf_1": "PUSH loop + one ADVANCE/POP step",
             "ww_1_inf": "one PUSH step + ADVANCE/POP loop", "ww_4_2": "while-while bounded 4 / 2 trips",
             "longest_lane": "bound: no divergence inside an iteration (slowest lane of each warp cast)", "ideal": "bound: 32 busy lanes"}
    for k in ("if_if", "while_while", "ww_1_1", "ww_inf_1", "ww_1_inf", "ww_4_2", "longest_lane", "ideal"):
        md.append("| %s | %.1f | %.3f | %.1f %% |" % (label[k], tot[k] / 1e6, tot[k] / tot["if_if"], 100.0 * tot["ideal"] / tot[k]))
    md.append("")
    md.append("iterations/cast %.1f (PUSH %.1f %%, ADVANCE %.1f %%, ADVANCE+POP %.1f %%); warp iterations that issue PUSH %.1f %%, ADVANCE %.1f %%, POP %.1f %%"
              % (tot["iters"] / tot["casts"], 100.0 * tot["pushes"] / tot["iters"], 100.0 * tot["advances"] / tot["iters"],
                 100.0 * tot["pops"] / tot["iters"], 100.0 * tot["warp_iters_push"] / tot["warp_iters"],
                 100.0 * tot["warp_iters_adv"] / tot["warp_iters"], 100.0 * tot["warp_iters_pop"] / tot["warp_iters"]))
    md.append("")
    md.append("Casts after the first (bounce / shadow rays) of a group of G neighbouring warp tiles re-dealt to warps (M warp-instr; 'as shipped' = the same rays in their own warps):")
    md.append("")
    md.append("| G (warps) | as shipped | sorted by true length (bound) | by dir.y | by (octant, dir.y) | by octant | refill: G/2 warps, >= 8 idle | G/2, >= 16 | G/4, >= 8 | G/4, >= 16 |")
    md.append("|---|---|---|---|---|---|---|---|---|---|")
    for G in (4, 16, 64):
        md.append("| %d | " % G + " | ".join("%.1f" % (tot["regroup%d_%s" % (G, k)] / 1e6) for k in ("as_is", "by_length", "by_dy", "by_oct_dy", "by_oct"))
                  + " | " + " | ".join("%.1f" % (tot["refill%d_%s" % (G, k)] / 1e6) for k in
                                       ("half_warps_idle8", "half_warps_idle16", "quarter_warps_idle8", "quarter_warps_idle16")) + " |")
    if a.shapes:
        md.append("")
        md.append("Warp tile shape (pixels per warp), shipped loop, M warp-instr:")
        md.append("")
        md.append("| 4x8 | 8x4 (shipped) | 16x2 | 32x1 |")
        md.append("|---|---|---|---|")
        vals = []
        for tw in (4, 8, 16, 32):
            v = 0.0
            for i, cam in enumerate(a.cams.split(",")):
                pos, l1, l2, r1, r2 = svo.CAMERAS[cam]
                f = O.make_frame(pos, l1, l2, r1, r2, frame_number=i + 1, render_mode=a.mode, max_depth=depth, casts=a.casts, cone_depth=11)
                v += sc.simt(f, a.width, a.height, costs, box=bool(a.box), nthreads=a.threads, tile_w=tw)["if_if"]
            vals.append(v)
        md.append("| " + " | ".join("%.1f" % (v / 1e6) for v in vals) + " |")
    print("\n".join(md))
    if a.out:
        with open(a.out, "w") as fh:
            fh.write("\n".join(md) + "\n")


if __name__ == "__main__":
    main()
