#!/usr/bin/env python
"""Turns the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/ :
   launches_<tag>.md   per-kernel totals + shares from the `--metrics gpu__time_duration.sum` launch list
   ncu_<tag>.md        key metrics of one `--set full` capture (dram bytes, duration, pipes, occupancy, stalls)
   traffic.json        dram__bytes_read+write per scan launch, consumed by bench.py's roofline.traffic
Usage: python profiles/summarize.py r1"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")


def launches(csv_path, md_path, title):
    """Per-kernel totals, grouped by (kernel, grid size): launches of one kernel on different problem sizes (bench.py's
    headline corpus and its secondary workloads) must not be averaged together (round-1 VERDICT, weak #10)."""
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    gi = h.index("Grid Size") if "Grid Size" in h else None
    agg = defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        key = (r[ki], r[gi].replace(" ", "") if gi is not None else "")
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v for _, v in agg.values())
    setup = ("fill_uniform", "fill_clustered", "build_shadow")
    search = sum(v for (k, _), (_, v) in agg.items() if not any(s in k for s in setup))
    with open(md_path, "w") as f:
        f.write(f"# {title}\n\nSource: `{os.path.basename(csv_path)}` (ncu --metrics gpu__time_duration.sum --clock-control none; "
                "cold-cache, serialised launches: compare SHARES).  One row per (kernel, grid size).\n\n"
                "| kernel | grid | launches | total us | us/launch | share of all | share of search kernels |\n|---|---|---:|---:|---:|---:|---:|\n")
        for (k, g), (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            ss = "-" if any(s in k for s in setup) else f"{100 * v / search:.1f}%"
            f.write(f"| `{k[:90]}` | {g} | {c} | {v:.1f} | {v / c:.1f} | {100 * v / tot:.1f}% | {ss} |\n")
        f.write("\n`fill_*` / `build_shadow_kernel` generate the synthetic corpus and its BF16 shadow once, outside the timed region.\n")


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__cycles_active.avg", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.per_cycle_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
]


def ncu_summary(rep, md_path, title, algorithmic_bytes=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, vals = rows[0], rows[1], rows[2]
    d = {n.split("TriageCompute.")[-1]: (vals[i], units[i]) for i, n in enumerate(h)}
    name = d.get("Kernel Name", ("?", ""))[0]
    with open(md_path, "w") as f:
        f.write(f"# {title}\n\nSource: `{os.path.basename(rep)}` (ncu --set full --clock-control none, one launch of `{name[:80]}`; "
                "numbers under the profiler are not bench values).\n\n| metric | value | unit |\n|---|---:|---|\n")
        for w in WANT:
            if w in d:
                f.write(f"| `{w}` | {d[w][0]} | {d[w][1]} |\n")
        stalls = sorted(((float(v[0] or 0), n) for n, v in d.items() if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("_per_issue_active.ratio")), reverse=True)[:6]
        if stalls:
            f.write("\nTop warp-stall reasons (warps stalled per issue-active cycle):\n\n")
            for v, n in stalls:
                f.write(f"* `{n}` = {v:.2f}\n")

    def gb(key):
        v, u = d[key]
        v = float(v.replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    traffic = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
    dur, du = d["gpu__time_duration.sum"]
    with open(md_path, "a") as f:
        f.write(f"\nDRAM traffic per launch = {traffic / 1e9:.4f} GB")
        if algorithmic_bytes:
            f.write(f" vs {algorithmic_bytes / 1e9:.4f} GB algorithmic (ratio {traffic / algorithmic_bytes:.3f})")
        f.write(f"; duration {dur} {du} under the profiler.\n")
    return traffic


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    # (csv in gpurun_out, md in profiles, title): regenerated only when the raw file is present
    LAUNCH_LISTS = [
        (f"launches_{tag}_headline.csv", f"launches_{tag}_headline.md",
         f"Launch list, round {tag[1:]}: bench.py headline (N=10M d=1024 Q=64 k=10 cosine), default path (BF16 shadow filter)"),
        (f"launches_{tag}_c2_shadow.csv", f"launches_{tag}_c2_shadow.md",
         f"Launch list, round {tag[1:]}: bench.py c2 (N=1M d=1024 Q=64 k=10 cosine), default path (BF16 shadow filter)"),
        (f"launches_{tag}_c3_shadow.csv", f"launches_{tag}_c3_shadow.md",
         f"Launch list, round {tag[1:]}: bench.py c3 (N=10M d=1024 Q=1024 k=100 inner product), default path (BF16 shadow filter)"),
        (f"launches_{tag}_c2_filter.csv", f"launches_{tag}_c2_filter.md",
         f"Launch list, round {tag[1:]}: bench.py c2 --path filter (TF32 filter over the fp32 rows)"),
        (f"launches_{tag}_simt.csv", f"launches_{tag}_c2_simt.md",
         f"Launch list, round {tag[1:]}: bench.py c2 (N=1M d=1024 Q=64 k=10 cosine), CUDA-core path (first working version)"),
    ]
    N10, D = 10_000_000, 1024
    CAPTURES = [  # (ncu-rep, md, title, algorithmic bytes per launch, traffic.json keys)
        (f"prof_{tag}_shadow_headline.ncu-rep", f"ncu_{tag}_scan_shadow_headline.md",
         "knn_scan_shadow_kernel<64> (BF16 shadow filter) — N=10M d=1024 fp32 Q=64 k=10 cosine", N10 * D * 2 + N10 * 8, ["headline:shadow"]),
        (f"prof_{tag}_shadow_q1.ncu-rep", f"ncu_{tag}_scan_shadow_q1.md",
         "knn_scan_shadow_kernel<64> (BF16 shadow filter) — N=10M d=1024 fp32 Q=1 k=10 cosine", N10 * D * 2 + N10 * 8, ["q1:shadow"]),
        (f"prof_{tag}_tc_headline.ncu-rep", f"ncu_{tag}_scan_tc_filter_headline.md",
         "knn_scan_tc_kernel<1,64> (1xTF32 filter over the fp32 rows) — N=10M d=1024 fp32 Q=64 k=10 cosine", N10 * D * 4, ["headline:filter"]),
        (f"prof_{tag}_tc3_headline.ncu-rep", f"ncu_{tag}_scan_tc_exact_headline.md",
         "knn_scan_tc_kernel<3,64> (3xTF32 exact) — N=10M d=1024 fp32 Q=64 k=10 cosine", N10 * D * 4, ["headline:tensor"]),
        (f"prof_{tag}_simt_q1.ncu-rep", f"ncu_{tag}_scan_simt_q1.md",
         "knn_scan_simt_kernel — N=10M d=1024 fp32 Q=1 k=10 cosine", N10 * D * 4, ["q1:simt", "headline:simt"]),
        (f"prof_{tag}_simt_c4.ncu-rep", f"ncu_{tag}_scan_simt_c4.md",
         "knn_scan_simt_kernel — N=10M d=768 fp16 Q=1 k=10 L2", N10 * 768 * 2, ["c4:simt"]),
    ]
    if tag == "r2":
        N2, N4 = 2_000_000, 4_000_000
        LAUNCH_LISTS.append((f"launches_{tag}_8gpu_rank0.csv", f"launches_{tag}_8gpu_rank0.md", "Launch list, round 2: rank 0 of the 8-GPU headline run"))
        CAPTURES += [
            (f"prof_{tag}_pair_c3_k10.ncu-rep", f"ncu_{tag}_scan_pair_c3_k10.md",
             "knn_scan_pair_kernel (CTA pairs, cta_group::2) — N=4M d=1024 fp32 Q=1024 k=10 dot, one launch = 2 query groups x 256", N4 * D * 2 + N4 * 8, []),
            (f"prof_{tag}_shadow_c3_k10.ncu-rep", f"ncu_{tag}_scan_shadow_c3_k10.md",
             "knn_scan_shadow_kernel<128> (single CTA, 4 query groups x 128) — N=2M d=1024 fp32 Q=1024 k=10 dot", N2 * D * 2 + N2 * 8, []),
            (f"prof_{tag}_shadow_c3_k100.ncu-rep", f"ncu_{tag}_scan_shadow_c3_k100.md",
             "knn_scan_shadow_kernel<128> (single CTA, 1 query group: the round-1 k=100 rule) — N=2M d=1024 fp32 Q=1024 k=100 dot", N2 * D * 2 + N2 * 8, []),
            (f"prof_{tag}_finish_headline.ncu-rep", f"ncu_{tag}_filter_finish_headline.md", "filter_finish_kernel — headline", None, []),
            (f"prof_{tag}_prep_headline.ncu-rep", f"ncu_{tag}_filter_prep_headline.md", "filter_prep_kernel — headline", None, []),
        ]
        N8 = 1_250_000
        LAUNCH_LISTS.append((f"launches_{tag}_shard.csv", f"launches_{tag}_shard.md",
                             "Launch list, round 2: one GPU at the 8-GPU shard shape (N=1.25M d=1024 Q=64 k=10 cosine), asynchronous API"))
        CAPTURES += [
            (f"prof_{tag}_shadow_shard.ncu-rep", f"ncu_{tag}_scan_shadow_shard.md",
             "knn_scan_shadow_kernel<64> at the 8-GPU shard shape — N=1.25M d=1024 fp32 Q=64 k=10 cosine (after the emission rework)", N8 * D * 2 + N8 * 8, []),
            (f"prof_{tag}_finish_shard.ncu-rep", f"ncu_{tag}_filter_finish_shard.md", "filter_finish_kernel — 8-GPU shard shape (a real launch, not the tail's early exit)", None, []),
            (f"prof_{tag}_prep_shard.ncu-rep", f"ncu_{tag}_filter_prep_shard.md", "filter_prep_kernel — 8-GPU shard shape", None, []),
        ]
    only = sys.argv[2] if len(sys.argv) > 2 else ""  # e.g. `summarize.py r2 shard`: only the files whose name contains it
    LAUNCH_LISTS = [x for x in LAUNCH_LISTS if only in x[0]]
    CAPTURES = [x for x in CAPTURES if only in x[0]]
    for src, dst, title in LAUNCH_LISTS:
        if os.path.exists(os.path.join(G, src)):
            launches(os.path.join(G, src), os.path.join(OUT, dst), title)
    tpath = os.path.join(OUT, "traffic.json")
    t = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for rep, md, title, algo, keys in CAPTURES:
        if os.path.exists(os.path.join(G, rep)):
            v = ncu_summary(os.path.join(G, rep), os.path.join(OUT, md), title, algo)
            for k in keys:
                t[k] = v
    json.dump(t, open(tpath, "w"), indent=1)
    print(json.dumps(t, indent=1))
