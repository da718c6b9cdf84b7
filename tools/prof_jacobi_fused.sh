# tools/prof_jacobi_fused.sh (GPU box) -- ncu captures of the fused Jacobi pipeline (jacobi_fused.cuh), both arithmetic builds;
# each command ran clean without ncu first
mkdir -p gpurun_out; set -x
for math in tolerance exact; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --math $math --workload planetesimal_disk --coordinates JACOBI > gpurun_out/jf_plain_$math.log 2>&1 || { tail -5 gpurun_out/jf_plain_$math.log; exit 1; }
  ncu --set full --clock-control none --import-source on -k regex:jacobi_fused -s 6 -c 1 -f -o gpurun_out/r02_jacobi_fused_$math python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --math $math --workload planetesimal_disk --coordinates JACOBI > gpurun_out/ncu_jf_$math.log 2>&1
  python profiles/summarize.py kernel gpurun_out/r02_jacobi_fused_$math.ncu-rep gpurun_out/r02_kernel_jacobi_fused_$math.txt > /dev/null
done
ls -la gpurun_out/ | tail
