"""Turns the raw ncu artefacts brought back in gpurun_out/ into the small, tracked summaries
under profiles/ (the .ncu-rep files themselves are too large to commit).

    python profiles/summarize.py launches gpurun_out/r01_launches.csv  profiles/r01_launches_summary.txt
    python profiles/summarize.py kernel   gpurun_out/r01_rad_prof.ncu-rep profiles/r01_rad_kernel.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        d = tot.setdefault(r[kn].split("(")[0], [0, 0.0])
        d[0] += 1
        d[1] += float(r[mv].replace(",", ""))
    s = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write(f"# per-kernel totals of {src} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)\n")
        f.write(f"{'launches':>8} {'total_us':>12} {'avg_us':>10} {'share':>7}  kernel\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[0]:8d} {v[1]/1e3:12.1f} {v[1]/1e3/v[0]:10.2f} {100*v[1]/s:6.1f}%  {k}\n")
    print(open(dst).read())


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write(f"# selected metrics of {src} (ncu --set full --clock-control none), one column per captured launch\n")
        kn = hdr.index("Kernel Name")
        f.write("kernel: " + " | ".join(r[kn] for r in data) + "\n")
        for k in KEYS + [h for h in hdr if "issue_stalled" in h and h.endswith("per_warp_active.pct")]:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k} [{units[i]}]: " + " | ".join(r[i] for r in data) + "\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
