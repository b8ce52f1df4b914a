#!/usr/bin/env python
"""Profiling driver (not a benchmark): a few C2 frames (64 x 2048 sweep, 1M-point keyframe submap, 10 GN iterations) through
mloam_frame with graph replay off, so that ncu sees every kernel of a frame as its own launch.
  MLOAM_DISABLE_GRAPHS=1 ncu --set full --clock-control none --launch-skip <frames before> --launch-count <one frame> ... python tools/profile_frame.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import synthetic as syn  # noqa: E402


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    m = bench.load_mloam()
    cfg = bench.CONFIGS["C2"]
    p = m.default_params()
    p.n_scans, p.max_outer, p.max_inner, p.map_cell, p.max_ring_points = 64, 10, 1, 0.0, 2048
    ctx = m.Context(0, p)
    wl = bench.make_workload(syn, cfg, 1, 0, 2, "keyframes")
    l0 = ctx.launch_count()
    for k in range(n_frames):
        fr = wl["frames"][k % 2]
        g = fr["groups"][0]
        pose, st = ctx.frame(g["cloud"], g["ss"], g["se"], wl["surf_map"], wl["corner_map"], fr["init"], True)
        print(f"frame {k}: launches so far {ctx.launch_count() - l0}, matches {st['n_surf']}+{st['n_corner']}")
    ctx.close()


if __name__ == "__main__":
    main()
