"""Turns ncu outputs brought back in gpurun_out/ into the small, committed summaries under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python scripts/ncu_summary.py full gpurun_out/prof_pair.ncu-rep profiles/r01_search_pair [bench-line.json]
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("awry::", "").replace("<unnamed>::", "anon::")
    return re.sub(r"cub::CUB_\d+_SM_\d+::", "cub::", name)[:90]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    seq = []
    for r in rows[1:]:
        k = short(r[ik])
        ns = float(r[iv].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        seq.append((k, ns, r[ig], r[ib]))
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write(f"source: `{src}`; {len(seq)} launches, {total/1e6:.2f} ms of device time\n\n")
        f.write("| kernel | launches | total ms | share | mean ms |\n|---|---:|---:|---:|---:|\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ns/1e6:.3f} | {100*ns/total:.1f}% | {ns/n/1e6:.3f} |\n")
        f.write("\n## product kernels in launch order (`anon::` = anonymous-namespace kernels of build.cu / reads.cu; cub and torch kernels and the synthetic-data generators omitted)\n\n| # | kernel | ms | grid | block |\n|---|---|---:|---|---|\n")
        for i, (k, ns, g, b) in enumerate(seq):
            if not k.startswith(("cub::", "at::", "void at")) and "gen_queries" not in k and "gen_text" not in k:
                f.write(f"| {i} | `{k}` | {ns/1e6:.3f} | {g} | {b} |\n")
    print("wrote", dst)


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "dram__sectors_read.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def full(rep, dst_prefix, bench_json=None, kernel_filter=None, traffic_json="profiles/search_kernel_traffic.json"):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": short(r[hdr.index("Kernel Name")])}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = float(r[i].replace(",", ""))
                except ValueError:
                    d[k] = r[i]
                d[k + " [unit]"] = units[i]
        out.append(d)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    starts = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
    hot = []
    # the launch to summarise: the first one matching the filter that carries DRAM counters
    sel = 0
    if kernel_filter:
        for i, d in enumerate(out):
            if kernel_filter in d["kernel"] and isinstance(d.get("dram__bytes_read.sum"), float) and d["dram__bytes_read.sum"] == d["dram__bytes_read.sum"]:
                sel = i
                break
    if starts:
        sel_s = min(sel, len(starts) - 1)
        h = srows[starts[sel_s]]
        end = starts[sel_s + 1] - 1 if len(starts) > sel_s + 1 else len(srows)
        body = [r for r in srows[starts[sel_s] + 1:end] if len(r) > h.index("Instructions Executed")]
        iS, iE, iW = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
        tot_samples = sum(int(r[iW]) for r in body if r[iW].isdigit())
        tot_inst = sum(int(r[iE]) for r in body if r[iE].isdigit())
        top = sorted(body, key=lambda r: -int(r[iW]) if r[iW].isdigit() else 0)[:12]
        hot = [{"sass": r[iS].strip(), "stall_samples": int(r[iW]), "share": int(r[iW]) / max(1, tot_samples),
                "executed": int(r[iE])} for r in top]
    first = out[sel] if out else {}
    traffic = {"kernel": first.get("kernel"),
               "dram_bytes_per_launch": first.get("dram__bytes_read.sum", 0) * (1e9 if first.get("dram__bytes_read.sum [unit]") == "Gbyte" else 1)
               + first.get("dram__bytes_write.sum", 0) * (1e6 if first.get("dram__bytes_write.sum [unit]") == "Mbyte" else 1e9 if first.get("dram__bytes_write.sum [unit]") == "Gbyte" else 1),
               "source": rep}
    summary = {"report": rep, "launches": out, "top_stall_instructions": hot,
               "warp_instructions_per_launch": tot_inst if starts else None, "traffic": traffic}
    if bench_json:
        summary["bench_line"] = json.load(open(bench_json))
    json.dump(summary, open(dst_prefix + ".json", "w"), indent=1)
    with open(dst_prefix + ".md", "w") as f:
        f.write(f"# ncu --set full summary: `{first.get('kernel')}`\n\nsource report: `{rep}` (scratch; this file is the committed summary)\n\n")
        f.write("| metric | launch 1 | unit |\n|---|---:|---|\n")
        for k in KEYS:
            if k in first:
                f.write(f"| {k} | {first[k]} | {first.get(k + ' [unit]', '')} |\n")
        f.write(f"\nDRAM traffic per launch: {traffic['dram_bytes_per_launch']/1e9:.2f} GB\n")
        f.write("\n## instructions with the most warp-stall samples\n\n| share | samples | executed | SASS |\n|---:|---:|---:|---|\n")
        for h_ in hot:
            f.write(f"| {100*h_['share']:.1f}% | {h_['stall_samples']} | {h_['executed']} | `{h_['sass']}` |\n")
    if traffic_json:
        json.dump(traffic, open(traffic_json, "w"), indent=1)
    print("wrote", dst_prefix + ".json/.md", traffic_json or "")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        # full <rep> <dst_prefix> [bench.json|-] [kernel-name filter] [traffic.json|-]
        a = sys.argv
        full(a[2], a[3], a[4] if len(a) > 4 and a[4] != "-" else None, a[5] if len(a) > 5 else None,
             (a[6] if a[6] != "-" else None) if len(a) > 6 else "profiles/search_kernel_traffic.json")
