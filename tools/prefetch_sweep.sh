# GPU box: the next-iteration operand prefetch off / L1 / L2 (tools/prefetch_variants.sh) on the four BASELINE sweeps
mkdir -p gpurun_out
fmt='import sys, json
for l in sys.stdin:
    d = json.loads(l); print(sys.argv[1], d["workload"], "exact %.4f (%.3f) tol %.4f (%.3f)" % (d["exact_ms"], d["frac_exact"], d["tolerance_ms"], d["frac_tolerance"]), "norm zero %.2g" % d["vs_exact"]["zero"]["max_normwise"])'
for rep in 1 2; do
for v in lib lib_pf0 lib_pf1; do
  REBX_B200_LIB=$PWD/reboundx_b200/$v/libreboundx_b200.so python tools/tol_sweep.py 2>/dev/null | python -c "$fmt" $v
done
done | tee gpurun_out/s2_prefetch_sweep.txt
