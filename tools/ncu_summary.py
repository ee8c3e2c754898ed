"""Condense ncu exports into the small text files kept under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.csv>       # per-launch durations + per-kernel share
  python tools/ncu_summary.py kernel <raw.csv> <source.csv> <out.txt>  # key counters + hottest source lines
  python tools/ncu_summary.py traffic <out.json> <label>=<raw.csv>[@<report name>] ...   # DRAM bytes per launch, read by bench.py
(raw.csv: `ncu -i X.ncu-rep --page raw --csv`; source.csv: `--page source --print-source cuda,sass --csv`)
"""
import csv, subprocess, sys, collections

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    out = [("id", "kernel", "grid", "block", "duration_us")]
    per = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        val = float(r[-1].replace(",", "")) / (1e3 if r[-2] in ("ns", "nsecond") else 1.0)
        out.append((r[0], name, r[8], r[7], f"{val:.3f}"))
        per.setdefault(name, []).append(val)
    tot = sum(sum(v) for v in per.values())
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerows(out)
        f.write("\n# per-kernel share of the profiled launches (cold-cache, serialised: compare shares, not absolutes)\n")
        for k, v in per.items():
            f.write(f"# {k}: n={len(v)} mean={sum(v)/len(v):.2f} us share={100*sum(v)/tot:.1f}%\n")


def kernel(raw, source, dst):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        for n, r in enumerate(data):
            f.write(f"== launch {n}: {r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"   {k:70s} {r[i]} {units[i]}\n")
        f.write("\n== hottest CUDA source lines (warp-stall samples, instructions executed)\n")
        f.write(subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_lines.py"), source, "32"],
                               capture_output=True, text=True).stdout)


def traffic(dst, *pairs):
    """profiles/ncu_traffic.json: dram__bytes_read.sum / dram__bytes_write.sum per launch (mean over the captured launches)
    of each labelled kernel, with the commit the capture was taken at.  bench.py reads `roofline.traffic` from it."""
    import json, os
    def to_bytes(v, unit):
        f = float(v.replace(",", ""))
        return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    try:
        old = json.load(open(dst))
    except Exception:
        old = {"kernels": {}}
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    for pair in pairs:
        label, rest = pair.split("=", 1)
        path, _, report = rest.partition("@")
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        rd = [to_bytes(r[ir], units[ir]) for r in data]
        wr = [to_bytes(r[iw], units[iw]) for r in data]
        old["kernels"][label] = {"dram_read_bytes": sum(rd) / len(rd), "dram_write_bytes": sum(wr) / len(wr), "launches": len(data),
                                 "duration_us_under_ncu": sum(float(r[it].replace(",", "")) for r in data) / len(data) / (1e3 if units[it] in ("ns", "nsecond") else 1.0),
                                 "ncu_kernel_name": data[0][hdr.index("Kernel Name")] if "Kernel Name" in hdr else None,
                                 "report": report or os.path.basename(path), "commit": commit}
    old["commit"] = commit
    old["how"] = "ncu --set full --clock-control none, one launch at a time (cold L2, serialised); bytes are per launch"
    json.dump(old, open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
