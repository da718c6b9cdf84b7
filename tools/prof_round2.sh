# tools/prof_round2.sh (GPU box) -- the ncu passes behind profiles/r02_*; every command ran clean without ncu first
mkdir -p gpurun_out; set -x
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err || { tail -20 gpurun_out/r02_bench_1gpu.err; exit 1; }
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_launches.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python profiles/summarize.py launches gpurun_out/r02_launches.csv gpurun_out/r02_launches_summary.txt > /dev/null
for math in exact tolerance; do
  for wl in debris_disk circumplanetary_ring asteroid_ensemble planetesimal_disk; do
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --math $math --workload $wl > gpurun_out/plain_${math}_$wl.log 2>&1 || { tail -5 gpurun_out/plain_${math}_$wl.log; exit 1; }
    ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 6 -c 1 -f -o gpurun_out/r02_${math}_$wl python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --math $math --workload $wl > gpurun_out/ncu_${math}_$wl.log 2>&1
    python profiles/summarize.py kernel gpurun_out/r02_${math}_$wl.ncu-rep gpurun_out/r02_kernel_${math}_$wl.txt > /dev/null
    ncu -i gpurun_out/r02_${math}_$wl.ncu-rep --page raw --csv > gpurun_out/r02_${math}_${wl}_raw.csv 2>/dev/null
    rm -f gpurun_out/r02_${math}_$wl.ncu-rep
  done
done
ls gpurun_out | head -80
