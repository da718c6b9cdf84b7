"""Generates tests/golden/*.npz by running the REFERENCE's own C sources (compiled in place
into oracle/_ref/libreboundx_ref.so by oracle/Makefile) on the cases of cases.py.

Run in the dev container only (needs /root/reference to build the library):
    make -C oracle ref && python tests/golden/make_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import cases          # noqa: E402
import golden_io as gio   # noqa: E402

from oracle import oracle                  # noqa: E402
from reboundx_b200 import scenarios        # noqa: E402

if __name__ == "__main__":
    assert oracle.have_ref(), "build oracle/_ref first: make -C oracle ref"
    lib = oracle.ref_lib()
    for fn in cases.ALL:
        case = fn()
        out, msgs = scenarios.run_case(case, lib)
        gio.save(fn.__name__, case, out, sorted(set(msgs)))
        print(f"{fn.__name__:36s} N={case['state']['x'].shape[0]:5d} messages={len(set(msgs))}")
