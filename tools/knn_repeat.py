#!/usr/bin/env python
"""Experiment: how much of a k_match_knn launch is cold-start (instruction cache / L2) rather than search work?

Runs the C2 workload's surf-feature match (k_match_knn + k_match_fit through mloam_match_from_map) N times back to
back and prints the event-timed `match` stage per launch, (a) undisturbed, (b) with an L2 flush before every launch,
(c) with the other big kernels of an LM iteration (normal equations) in between.  Not a benchmark: a diagnosis.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import synthetic as syn  # noqa: E402


def main():
    m = bench.load_mloam()
    p = m.default_params()
    p.n_scans, p.map_cell = 64, 0.26
    ctx = m.Context(0, p)
    surf_map, corner_map, frames, _ = bench.make_workload(syn, 1, 0, 1)
    f = frames[0]
    feats = ctx.extract_features(f["cloud"], f["ss"], f["se"])
    surf = ctx.voxel_downsample(feats["surf_points_less_flat"], 0.4, True)
    ctx.map_build(1, surf_map, 0.26)
    ctx.map_build(0, corner_map, 0.26)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    pose = np.asarray(f["init"], np.float64)
    valid, coeffs, _ = ctx.match_from_map(1, "s", surf, pose, want_nn=False)
    types = np.ones(valid.sum(), np.uint8)
    pts64 = surf[valid][:, :3].astype(np.float64)
    cf = coeffs[valid]
    print("features", surf.shape[0], "matched", int(valid.sum()))

    def run(label, n, pre):
        nonlocal surf
        ctx.profile(True)
        ctx.profile_reset()
        for _ in range(n):
            pre()
            ctx.match_from_map(1, "s", surf, pose, want_nn=False)
        ms, k = ctx.profile_get("match")
        fms, fk = ctx.profile_get("fit")
        ctx.profile(False)
        print(f"{label:34s} k_match_knn {1e3 * ms / max(k, 1):7.1f} us/launch   k_match_fit {1e3 * fms / max(fk, 1):6.1f} us/launch  ({k} launches)")

    def nothing():
        pass

    def do_flush():
        flush.fill_(1)
        torch.cuda.synchronize()

    def other_kernels():
        ctx.normal_equations(types, pts64, cf, 1.0, 0.1, pose)

    full = surf
    for n_sub in (32, 256, 1024, 2368, 4736, surf.shape[0]):
        surf = full[:n_sub]
        run(f"back to back, {n_sub} features", 10, nothing)
    surf = full
    for _ in range(2):
        run("back to back", 20, nothing)
        run("L2 flushed before each launch", 20, do_flush)
        run("normal equations in between", 20, other_kernels)


if __name__ == "__main__":
    main()
