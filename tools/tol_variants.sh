# tools/tol_variants.sh -- builds CTA-geometry variants of the tolerance-mode sweeps (kernels_tol.cu only; the other
# objects are reused) into reboundx_b200/lib_tol_<tag>/ for a sweep on the GPU box (tools/tol_sweep.sh).
set -e
cd "$(dirname "$0")/../reboundx_b200/csrc"
build() {   # tag flags...
  out=../lib_tol_$1; shift; mkdir -p $out
  cp ../lib/kernels.o ../lib/api_core.o ../lib/dispatch.o ../lib/binary_io.o $out/
  touch $out/kernels.o $out/api_core.o $out/dispatch.o $out/binary_io.o
  make -s OUT=$out EXTRA_NVFLAGS="$*" $out/libreboundx_b200.so
}
# light x yark x trio geometry (threads:blocks)
build a -DREBXB_TOL_LIGHT_THREADS=256 -DREBXB_TOL_LIGHT_BLOCKS=2 -DREBXB_TOL_YARK_THREADS=128 -DREBXB_TOL_YARK_BLOCKS=6 -DREBXB_TOL_TRIO_THREADS=128 -DREBXB_TOL_TRIO_BLOCKS=4
build b -DREBXB_TOL_LIGHT_THREADS=128 -DREBXB_TOL_LIGHT_BLOCKS=5 -DREBXB_TOL_YARK_THREADS=128 -DREBXB_TOL_YARK_BLOCKS=5 -DREBXB_TOL_TRIO_THREADS=128 -DREBXB_TOL_TRIO_BLOCKS=5
build c -DREBXB_TOL_LIGHT_THREADS=128 -DREBXB_TOL_LIGHT_BLOCKS=6 -DREBXB_TOL_YARK_THREADS=128 -DREBXB_TOL_YARK_BLOCKS=7 -DREBXB_TOL_TRIO_THREADS=128 -DREBXB_TOL_TRIO_BLOCKS=3
build d -DREBXB_TOL_LIGHT_THREADS=128 -DREBXB_TOL_LIGHT_BLOCKS=7 -DREBXB_TOL_YARK_THREADS=256 -DREBXB_TOL_YARK_BLOCKS=4 -DREBXB_TOL_TRIO_THREADS=256 -DREBXB_TOL_TRIO_BLOCKS=3
ls ../lib_tol_*/libreboundx_b200.so
