# tools/prof_round.sh -- the ncu passes behind profiles/ (run on the GPU box: gpurun -- bash tools/prof_round.sh); each command ran clean without ncu first
mkdir -p gpurun_out; set -x
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_final4.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_final2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_final2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 6 -c 1 -f -o gpurun_out/r01_rad_final2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_rad_final2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jacobi_sweep -s 4 -c 1 -f -o gpurun_out/r01_jacobi_sweep python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --workload planetesimal_disk --coordinates JACOBI > gpurun_out/ncu_jacobi_sweep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kepler_drift -s 4 -c 1 -f -o gpurun_out/r01_kepler_drift python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_kepler.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 6 -c 1 -f -o gpurun_out/r01_trio_final2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --workload planetesimal_disk > gpurun_out/ncu_trio_final2.log 2>&1
for r in r01_rad_final2 r01_jacobi_sweep r01_kepler_drift r01_trio_final2; do python profiles/summarize.py kernel gpurun_out/$r.ncu-rep gpurun_out/$r.txt > /dev/null; done
ncu -i gpurun_out/r01_kepler_drift.ncu-rep --page details --csv > gpurun_out/r01_kepler_drift_details.csv 2>/dev/null
ncu -i gpurun_out/r01_jacobi_sweep.ncu-rep --page details --csv > gpurun_out/r01_jacobi_sweep_details.csv 2>/dev/null
# the reports themselves are ~21 MB each and gpurun brings back at most 64 MiB: keep one, summarise the rest on the box
rm -f gpurun_out/r01_jacobi_sweep.ncu-rep gpurun_out/r01_kepler_drift.ncu-rep gpurun_out/r01_trio_final2.ncu-rep
ls -la gpurun_out/
