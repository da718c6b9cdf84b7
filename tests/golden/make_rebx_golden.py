"""Writes tests/golden/state_reference_writer.rebx with the REFERENCE's own binary writer (src/output.c compiled
unmodified into oracle/_ref/librebx_io_bridge.so) from the 12-particle state of tests/test_binary_io.py, so that
machines without the reference tree (the GPU box) still test this library's loader and writer against bytes the
reference produced.  Run in the dev container:  python tests/golden/make_rebx_golden.py"""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_binary_io as T          # noqa: E402

bridge = ctypes.CDLL(T.BRIDGE)
sim, rebx, pos = T.build_state()
out = os.path.join(HERE, "state_reference_writer.rebx")
bridge.rebx_output_binary(rebx._ptr, ctypes.c_char_p(out.encode()))
print(out, os.path.getsize(out), "bytes")
rebx.detach(); sim.free()
