# tools/variant_sweep.sh -- CTA-size / occupancy variants of the sweeps (libraries built with
# make OUT=../lib_t128_K EXTRA_NVFLAGS="-DREBXB_THREADS=128 -DREBXB_HEAVY_MINBLOCKS=K"); run on the GPU box
for v in lib lib_t128_4 lib_t128_5 lib_t128_6; do
  for wl in debris_disk circumplanetary_ring asteroid_ensemble planetesimal_disk; do
    REBX_B200_LIB=$PWD/reboundx_b200/$v/libreboundx_b200.so python bench.py --workload $wl --no-cpu-baseline --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$v $wl kernel_ms %.4f frac %.3f resident_step %.4f' % (d['roofline']['kernel_ms'], d['roofline']['frac'], d['resident_step']['ms_per_step']))"
  done
done
