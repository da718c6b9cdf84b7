# GPU box: A/B of tolerance-kernel build variants (reboundx_b200/lib_tol_*/) against the default library on the two
# 10^6-particle configurations; parity tests of the default library first.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "tolerance_mode or trio or damping or planetesimal or yarkovsky" > gpurun_out/s2_pytest_ab.log 2>&1; echo rc=$? >> gpurun_out/s2_pytest_ab.log; tail -4 gpurun_out/s2_pytest_ab.log
fmt='import sys, json
for l in sys.stdin:
    d = json.loads(l); print(sys.argv[1], d["workload"], "exact %.4f (%.3f) tol %.4f (%.3f)" % (d["exact_ms"], d["frac_exact"], d["tolerance_ms"], d["frac_tolerance"]), "normwise zero %.3g gravity %.3g" % (d["vs_exact"]["zero"]["max_normwise"], d["vs_exact"]["gravity"]["max_normwise"]))'
for rep in 1 2; do for v in reboundx_b200/lib reboundx_b200/lib_tol_*; do
  REBX_B200_LIB=$PWD/$v/libreboundx_b200.so python tools/tol_sweep.py planetesimal_disk asteroid_ensemble 2>/dev/null | python -c "$fmt" $(basename $v)
done; done | tee gpurun_out/s2_tol_ab.txt
