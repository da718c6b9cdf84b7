"""Shared parity metrics for the tests.

* per-component relative error  |da_k| / |a_k|           (north_star criterion (i): <= 1e-13)
* norm-wise error               |da_k| / max_k |a_k|     (companion metric: a component that
  is 5-7 orders below |a| because the particle sits next to a coordinate axis carries the
  absolute rounding error of the large components)
"""
import numpy as np

TOL = 1e-13


def errors(got, want, skip=()):
    g = np.array(got, dtype=np.float64)
    w = np.array(want, dtype=np.float64)
    keep = np.ones(g.shape[1], dtype=bool)
    for i in skip:
        keep[i] = False
    g, w = g[:, keep], w[:, keep]
    d = np.abs(g - w)
    with np.errstate(divide="ignore", invalid="ignore"):
        comp = np.where(d == 0, 0.0, d/np.abs(w))
    norm = d/np.maximum(np.max(np.abs(w), axis=0), 1e-300)
    return {"bit_identical": bool(np.array_equal(g, w)), "max_component": float(np.max(comp)),
            "frac_component_above_tol": float(np.mean(comp > TOL)), "max_normwise": float(np.max(norm))}
