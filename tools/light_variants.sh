# tools/light_variants.sh -- occupancy x arithmetic variants of the bit-exact light sweeps (radiation_forces, gr_potential):
# REBXB_LIGHT_MINBLOCKS = resident 256-thread CTAs per SM asked of the register allocator (4 = 64 registers, shipped),
# REBXB_FASTMATH = straight-line correctly rounded division / sqrt (exact_math.cuh) instead of the hardware sequences.
# Builds kernels.cu only; the other objects are reused.  Run tools/light_sweep.sh on the GPU box afterwards.
set -e
cd "$(dirname "$0")/../reboundx_b200/csrc"
build() {   # tag flags...
  out=../lib_light_$1; shift; mkdir -p $out
  cp ../lib/kernels_tol.o ../lib/api_core.o ../lib/dispatch.o ../lib/binary_io.o $out/
  touch $out/kernels_tol.o $out/api_core.o $out/dispatch.o $out/binary_io.o
  make -s OUT=$out EXTRA_NVFLAGS="$*" $out/libreboundx_b200.so
}
build f2 -DREBXB_FASTMATH=1 -DREBXB_LIGHT_MINBLOCKS=2 &
build f3 -DREBXB_FASTMATH=1 -DREBXB_LIGHT_MINBLOCKS=3 &
build h3 -DREBXB_FASTMATH=0 -DREBXB_LIGHT_MINBLOCKS=3 &
build f4 -DREBXB_FASTMATH=1 -DREBXB_LIGHT_MINBLOCKS=4 &
wait
ls ../lib_light_*/libreboundx_b200.so
