# tools/prof_final.sh (GPU box) -- the round's final evidence: default bench line, reference arm, launch list of the same
# command under ncu, captures of the kernels that changed since tools/prof_round2.sh (the fused Jacobi pipeline in both
# arithmetic builds, the fused Wisdom-Holman sweep).  Every command ran clean without ncu first.
mkdir -p gpurun_out; set -x
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err || { tail -20 gpurun_out/r02_bench_1gpu.err; exit 1; }
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_launches.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python profiles/summarize.py launches gpurun_out/r02_launches.csv gpurun_out/r02_launches_summary.txt > /dev/null
for math in tolerance exact; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --math $math --workload planetesimal_disk --coordinates JACOBI > gpurun_out/jf_plain_$math.log 2>&1 || { tail -5 gpurun_out/jf_plain_$math.log; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:jacobi_fused -s 6 -c 1 -f -o gpurun_out/r02_jacobi_fused_$math python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --math $math --workload planetesimal_disk --coordinates JACOBI > gpurun_out/ncu_jf_$math.log 2>&1
  python profiles/summarize.py kernel gpurun_out/r02_jacobi_fused_$math.ncu-rep gpurun_out/r02_kernel_jacobi_fused_$math.txt > /dev/null
  ncu -i gpurun_out/r02_jacobi_fused_$math.ncu-rep --page raw --csv > gpurun_out/r02_jacobi_fused_${math}_raw.csv 2>/dev/null
  rm -f gpurun_out/r02_jacobi_fused_$math.ncu-rep
done
ncu --set full --clock-control none -k regex:wh_kick_drift -s 2 -c 1 -f -o gpurun_out/r02_wh python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_wh.log 2>&1
python profiles/summarize.py kernel gpurun_out/r02_wh.ncu-rep gpurun_out/r02_kernel_wh_kick_drift.txt > /dev/null
rm -f gpurun_out/r02_wh.ncu-rep
ls gpurun_out | head -40
