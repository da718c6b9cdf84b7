# tools/prof_tol.sh (GPU box) -- ncu captures of the tolerance-mode sweeps; each command ran clean without ncu first
mkdir -p gpurun_out; set -x
for wl in planetesimal_disk asteroid_ensemble circumplanetary_ring debris_disk; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --math tolerance --workload $wl > gpurun_out/tol_plain_$wl.log 2>&1 || { tail -5 gpurun_out/tol_plain_$wl.log; exit 1; }
  tail -1 gpurun_out/tol_plain_$wl.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', d['roofline']['kernel_ms'], d['roofline']['frac'])"
  ncu --set full --clock-control none --import-source on -k regex:pass_kernel_tol -s 6 -c 1 -f -o gpurun_out/r02_tol_$wl python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --math tolerance --workload $wl > gpurun_out/ncu_tol_$wl.log 2>&1
  python profiles/summarize.py kernel gpurun_out/r02_tol_$wl.ncu-rep gpurun_out/r02_kernel_tol_$wl.txt > /dev/null
  ncu -i gpurun_out/r02_tol_$wl.ncu-rep --page raw --csv > gpurun_out/r02_tol_${wl}_raw.csv 2>/dev/null
done
rm -f gpurun_out/r02_tol_asteroid_ensemble.ncu-rep gpurun_out/r02_tol_circumplanetary_ring.ncu-rep gpurun_out/r02_tol_debris_disk.ncu-rep
ls -la gpurun_out/ | tail -20
