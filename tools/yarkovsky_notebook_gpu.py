"""One-off (about 2.5 GPU-minutes): the FULL run of the reference's YarkovskyEffect.ipynb (2*10^6 Wisdom-Holman
steps, tests/yarkovsky_notebook.py) with this library's CUDA force; prints the drift next to the published one.
The test suite runs the first 10 % of it (tests/test_gpu_parity.py)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import yarkovsky_notebook as nb

out = []
for full, want in ((True, [nb.NOTEBOOK_FULL]), (False, list(nb.NOTEBOOK_SIMPLE))):
    t0 = time.time()
    da, _ = nb.drift(full, None)
    out.append({"version": "full (ye_flag 0)" if full else "simple (ye_flag +1, -1)", "drift_AU": [float(v) for v in da], "notebook_AU": want,
                "relative_difference": [float(abs(a/b - 1.0)) for a, b in zip(da, want)], "seconds": round(time.time() - t0, 1)})
    print(json.dumps(out[-1]), flush=True)
