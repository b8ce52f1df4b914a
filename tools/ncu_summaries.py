#!/usr/bin/env python
"""Turn ncu exports (gpurun_out/) into the small text summaries committed under profiles/.

  python tools/ncu_summaries.py launches gpurun_out/launches.csv "<command>" > profiles/rNN_launches.txt
  python tools/ncu_summaries.py kernel   gpurun_out/prof.ncu-rep           > profiles/rNN_kernel.txt
  python tools/ncu_summaries.py table    gpurun_out/frame_raw.csv          > profiles/rNN_kernel_table.txt   (ncu -i x.ncu-rep --page raw --csv)
"""
import collections
import csv
import subprocess
import sys


def launches(path, cmd):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("mloam::", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {cmd}")
    print("# cold-cache, serialised per-launch device times (ncu): compare SHARES with bench.py's stage_ms_per_step, not absolutes")
    print(f"# total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:62]:62s} n={a[0]:4d} total={a[1]:9.1f}us avg={a[1] / a[0]:8.2f}us share={100 * a[1] / tot:5.1f}%")


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none --import-source on  ->  {path}")
    for r in rows[2:]:
        print("kernel:", r[idx["Kernel Name"]][:110])
        for w in WANT:
            if w in idx:
                print(f"  {w:70s} {r[idx[w]]:>16s} {units[idx[w]]}")
        st = [(h, float(r[i])) for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled") and r[i] not in ("", "n/a")]
        for h, v in sorted(st, key=lambda x: -x[1])[:6]:
            print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:6.2f} warps/issue")


# algorithmic bytes per launch of one C2 frame (SURVEY 8d x the units one launch processes)
ALG = {"k_match_knn": 1203500, "k_linearize": 425375, "k_grid_count": 10000000, "k_grid_scatter": 18000000, "k_curvature": 3276800}
PEAK_GBS = 6483.6  # MEASURED_PEAKS.json hbm_gbs


def table(path):
    rows = list(csv.reader(open(path)))
    while rows and "Kernel Name" not in rows[0]:
        rows.pop(0)
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}

    def num(r, key):
        v = r[idx[key]].replace(",", "") if key in idx else ""
        try:
            return float(v)
        except ValueError:
            return float("nan")

    units = rows[1]
    agg = collections.OrderedDict()
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("mloam::", "")
        t = num(r, "gpu__time_duration.sum")
        t = t / 1000 if units[idx["gpu__time_duration.sum"]] == "ns" else t
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = num(r, "dram__bytes_read.sum") * scale.get(units[idx["dram__bytes_read.sum"]], 1.0)
        wr = num(r, "dram__bytes_write.sum") * scale.get(units[idx["dram__bytes_write.sum"]], 1.0)
        a = agg.setdefault(name, dict(n=0, t=0.0, rd=0.0, wr=0.0, l2=0.0, l1=0.0, occ=0.0, iss=0.0, regs=0, grid="", ))
        a["n"] += 1
        a["t"] += t
        a["rd"] += rd
        a["wr"] += wr
        a["l2"] += num(r, "lts__t_sector_hit_rate.pct")
        a["l1"] += num(r, "l1tex__t_sector_hit_rate.pct")
        a["occ"] += num(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
        a["iss"] += 100.0 * num(r, "smsp__issue_active.avg.per_cycle_active")
        a["regs"] = int(num(r, "launch__registers_per_thread"))
        a["grid"] = f"{int(num(r, 'launch__grid_size')):5d} x {int(num(r, 'launch__block_size')):4d}"
    print("kernel                               n       us   dram_rd   dram_wr L2hit% L1hit%  occ% issue% regs grid x block  ALG bytes    frac")
    tot = 0.0
    for k, a in sorted(agg.items(), key=lambda x: -x[1]["t"]):
        n = a["n"]
        us = a["t"] / n
        tot += a["t"]
        alg = next((v for kk, v in ALG.items() if k.startswith(kk)), None)
        frac = (alg / (us * 1e-6) / 1e9 / PEAK_GBS) if alg else float("nan")
        print(f"{k[:34]:34s} {n:4d} {us:8.2f} {a['rd'] / n / 1e6:8.2f}M {a['wr'] / n / 1e6:8.2f}M {a['l2'] / n:6.1f} {a['l1'] / n:6.1f} {a['occ'] / n:5.1f} "
              f"{a['iss'] / n:6.1f} {a['regs']:4d} {a['grid']} {alg if alg else '-':>10} {frac:7.4f}")
    print(f"# sum over the profiled launches: {tot:.0f} us (serialised, cold)")


if __name__ == "__main__":
    if sys.argv[1] == "table":
        table(sys.argv[2])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        kernel(sys.argv[2])
