"""tools/make_traffic.py -- profiles/traffic.json from the raw ncu pages brought back by tools/prof_round2.sh
(gpurun_out/r02_<math>_<workload>_raw.csv): per sweep kernel the DRAM bytes of one launch, the FP64 pipe's share of the
cycles and the FP64 operation slots executed per particle.  bench.py quotes these figures (with `from`) next to the
durations it measures itself; nothing here is measured under bench.py's own timing."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = {"debris_disk": 10_000_000, "circumplanetary_ring": 10_000_000, "asteroid_ensemble": 1_000_000, "planetesimal_disk": 1_000_000}


def num(s):
    return float(s.replace(",", ""))


def main(src_dir, commit):
    out = {"_what": "ncu --set full --clock-control none captures of the sweep kernels at the bench's sizes (tools/prof_round2.sh, tools/prof_final.sh); "
                    "fp64_slots_per_particle = sm__pipe_fp64_cycles_active (share of elapsed cycles) x cycles x SMs x 64 lanes / particles"}
    dst = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(dst):                 # entries whose raw pages are not in src_dir keep their recorded values
        with open(dst) as f:
            old = json.load(f)
        old.update(out)
        out = old
    cases = [(wl, n, math, f"r02_{math}_{wl}_raw.csv", f"{wl}/{math}", f"profiles/r02_kernel_{math}_{wl}.txt") for math in ("exact", "tolerance") for wl, n in N.items()]
    cases += [("planetesimal_disk", 1_000_000, math, f"r02_jacobi_fused_{math}_raw.csv", f"planetesimal_disk_JACOBI/{math}", f"profiles/r02_kernel_jacobi_fused_{math}.txt")
              for math in ("exact", "tolerance")]
    for wl, n, math, fname, key, prof in cases:
        if True:
            path = os.path.join(src_dir, fname)
            if not os.path.exists(path):
                continue
            rows = list(csv.reader(open(path)))
            hdr, units, data = rows[0], rows[1], rows[2]
            g = lambda k: data[hdr.index(k)]
            u = lambda k: units[hdr.index(k)]
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            dram = num(g("dram__bytes_read.sum"))*scale[u("dram__bytes_read.sum")] + num(g("dram__bytes_write.sum"))*scale[u("dram__bytes_write.sum")]
            cycles = num(g("sm__cycles_elapsed.avg"))
            busy_elapsed = num(g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"))/100.0
            busy_active = num(g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"))
            sms = 148
            slots = busy_elapsed*cycles*sms*64/(n + 1)
            t = num(g("gpu__time_duration.sum"))
            t_us = t if u("gpu__time_duration.sum") == "us" else (t*1e3 if u("gpu__time_duration.sum") == "ms" else t/1e3)
            out[key] = {"dram_bytes": dram, "fp64_pipe_busy_pct": busy_active, "fp64_slots_per_particle": round(slots, 1),
                        "ncu_duration_us": t_us, "kernel": g("Kernel Name"), "registers": int(num(g("launch__registers_per_thread"))),
                        "from": f"{prof}, captured at commit {commit}"}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE, text=True).stdout.strip()
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out"), sys.argv[2] if len(sys.argv) > 2 else commit)
