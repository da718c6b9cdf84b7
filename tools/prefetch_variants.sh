# tools/prefetch_variants.sh -- builds the library with the next-iteration operand prefetch off / into L1 / into L2
# (REBXB_PREFETCH, pass_common.cuh) into reboundx_b200/lib_pf<0|1|2>/ for a sweep on the GPU box:
#   for v in 0 1 2; do REBX_B200_LIB=$PWD/reboundx_b200/lib_pf$v/libreboundx_b200.so python tools/tol_sweep.py; done
set -e
cd "$(dirname "$0")/../reboundx_b200/csrc"
for v in ${@:-0 1 2}; do
  out=../lib_pf$v; mkdir -p $out
  cp ../lib/api_core.o ../lib/dispatch.o ../lib/binary_io.o $out/
  touch $out/api_core.o $out/dispatch.o $out/binary_io.o
  make -s OUT=$out EXTRA_NVFLAGS="-DREBXB_PREFETCH=$v" $out/libreboundx_b200.so &
done
wait
ls -la ../lib_pf*/libreboundx_b200.so
