# tools/prof_ring.sh -- ncu capture of the J2/J4 (+) gr_potential sweep with its final CTA geometry (GPU box)
mkdir -p gpurun_out; set -x
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --workload circumplanetary_ring > gpurun_out/plain_ring.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 6 -c 1 -f -o gpurun_out/r01_ring_final3 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --workload circumplanetary_ring > gpurun_out/ncu_ring_final3.log 2>&1
python profiles/summarize.py kernel gpurun_out/r01_ring_final3.ncu-rep gpurun_out/r01_ring_final3.txt > /dev/null
rm -f gpurun_out/r01_ring_final3.ncu-rep
cat gpurun_out/r01_ring_final3.txt | head -24
