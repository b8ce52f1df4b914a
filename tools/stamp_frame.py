"""Stage times of a C2 frame under GRAPH REPLAY (MLOAM_STAMP=1: one-thread %globaltimer kernels between the stages).
Usage: MLOAM_STAMP=1 python tools/stamp_frame.py [--config C2] [--frames 40]   (diagnosis; prints a table)"""
import argparse
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("MLOAM_STAMP", "1")
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import synthetic as syn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--lookahead", action="store_true", help="announce sweep k+1 with frame k (extraction on the side stream)")
    a = ap.parse_args()
    cfg = dict(bench.CONFIGS[a.config])
    m = bench.load_mloam()
    p = m.default_params()
    p.n_scans, p.max_outer, p.max_inner, p.map_cell = cfg["rings"], cfg["gn_iters"], 1, 0.0
    p.max_ring_points = cfg["horizon"]
    p.gf_method, p.gf_ratio = cfg["gf_method"], cfg["gf_ratio"]
    ctx = m.Context(0, p)
    wl = bench.make_workload(syn, cfg, 1, 0, 4, "keyframes")
    dev = torch.device("cuda", 0)
    L = cfg["lidars"]
    my = [f["groups"][0] for f in wl["frames"]]
    if L > 1:
        ctx.set_lidars(L, my[0]["ext"])
    d_surf, d_corner = torch.from_numpy(wl["surf_map"]).to(dev), torch.from_numpy(wl["corner_map"]).to(dev)
    d = [dict(cloud=torch.from_numpy(g["cloud"]).to(dev), ss=torch.from_numpy(g["ss"]).to(dev), se=torch.from_numpy(g["se"]).to(dev)) for g in my]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step(k, rebuild):
        f, g, dd = wl["frames"][k % 4], my[k % 4], d[k % 4]
        if a.lookahead:
            gn, dn = my[(k + 1) % 4], d[(k + 1) % 4]
            ctx.frame_set_next_device(dn["cloud"].data_ptr(), gn["cloud"].shape[0], dn["ss"].data_ptr(), dn["se"].data_ptr(), cfg["rings"] * L)
        return ctx.frame_device(dd["cloud"].data_ptr(), g["cloud"].shape[0], dd["ss"].data_ptr(), dd["se"].data_ptr(), cfg["rings"] * L,
                                d_surf.data_ptr(), wl["surf_map"].shape[0], d_corner.data_ptr(), wl["corner_map"].shape[0], f["init"], rebuild)

    for rb in (True, False):
        for _ in range(4):
            for k in range(4):
                step(k, rb)
    rows = {}
    order = []
    totals = []
    for k in range(a.frames):
        flush.fill_(k & 0xFF)
        torch.cuda.synchronize()
        step(k, False)
        st = ctx.debug_stamps()
        prev = 0
        seen = {}
        for label, t in st[1:]:
            seen[label] = seen.get(label, 0) + 1
            key = f"{label} #{seen[label]}"
            if key not in rows:
                rows[key] = []
                order.append(key)
            rows[key].append((t - prev) / 1e3)
            prev = t
        totals.append(st[-1][1] / 1e3)
    print(f"# {a.config}: median stage times [us] over {a.frames} regular (non-keyframe) frames, graph replay, L2 flushed before each frame")
    agg = {}
    for key in order:
        med = statistics.median(rows[key])
        print(f"{key:48s} {med:9.2f}")
        base = key.split(" #")[0]
        agg[base] = agg.get(base, 0.0) + med
    print("# sums per stage")
    for k, v in agg.items():
        print(f"{k:48s} {v:9.2f}")
    print(f"{'first stamp -> last stamp':48s} {statistics.median(totals):9.2f}")


if __name__ == "__main__":
    main()
