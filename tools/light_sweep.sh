# tools/light_sweep.sh (GPU box) -- times the variants built by tools/light_variants.sh on the headline sweep
for v in lib lib_light_f2 lib_light_f3 lib_light_h3 lib_light_f4; do
  REBX_B200_LIB=$PWD/reboundx_b200/$v/libreboundx_b200.so python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 30 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$v kernel_ms %.4f frac %.3f' % (d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
