# GPU box: sweep the CTA-geometry variants of the tolerance kernels
mkdir -p gpurun_out
fmt='import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d["lib"], "V", d["V"], sys.argv[1], d["workload"], "exact %.4f tol %.4f frac %.3f" % (d["exact_ms"], d["tolerance_ms"], d["frac_tolerance"]), "norm zero %.2g gravity %.2g" % (d["vs_exact"]["zero"]["max_normwise"], d["vs_exact"]["gravity"]["max_normwise"]))'
python tools/tol_sweep.py > gpurun_out/tol_default.jsonl 2> gpurun_out/tol_default.err || tail -5 gpurun_out/tol_default.err
cat gpurun_out/tol_default.jsonl
for v in reboundx_b200/lib reboundx_b200/lib_tol_*; do
  for V in 1 2; do
    REBX_B200_LIB=$PWD/$v/libreboundx_b200.so REBX_B200_TOL_V=$V python tools/tol_sweep.py debris_disk circumplanetary_ring 2>/dev/null | python -c "$fmt" grid
  done
  for g in resident full; do
    REBX_B200_LIB=$PWD/$v/libreboundx_b200.so REBX_B200_TOL_GRID=$g python tools/tol_sweep.py asteroid_ensemble planetesimal_disk 2>/dev/null | python -c "$fmt" $g
  done
done | tee gpurun_out/tol_variants.txt
