# GPU box: per-launch duration / registers / pipe use of a bench command's kernels -> gpurun_out/<tag>_list.csv
#   bash tools/ncu_list.sh <tag> <skip> <count> <bench args...>
tag=$1; skip=$2; count=$3; shift 3
mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/${tag}_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__grid_size,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -s $skip -c $count --csv --log-file gpurun_out/${tag}_list.csv python bench.py "$@" > gpurun_out/${tag}_ncu.log 2>&1
python - gpurun_out/${tag}_list.csv <<'PY'
import sys, csv
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; kn = h.index("Kernel Name"); mn = h.index("Metric Name"); mv = h.index("Metric Value"); idc = h.index("ID")
cur = {}
for r in rows[hi + 1:]:
    if len(r) > mv: cur.setdefault((int(r[idc]), r[kn].split("(")[0][:60]), {})[r[mn].split(".")[0].replace("__", "_")[:30]] = r[mv]
for k, v in sorted(cur.items()): print(k[0], k[1], " ".join(f"{a}={b}" for a, b in v.items()))
PY
