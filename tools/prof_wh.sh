# tools/prof_wh.sh -- launch list of the final bench + ncu capture of the fused Wisdom-Holman sweep (run on the GPU box)
mkdir -p gpurun_out; set -x
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain_final5.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_final3.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_final3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wh_kick_drift -s 4 -c 1 -f -o gpurun_out/r01_wh_kick_drift python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_wh.log 2>&1
python profiles/summarize.py kernel gpurun_out/r01_wh_kick_drift.ncu-rep gpurun_out/r01_wh_kick_drift.txt > /dev/null
rm -f gpurun_out/r01_wh_kick_drift.ncu-rep
ls -la gpurun_out/
