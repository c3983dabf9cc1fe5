"""Shared helpers for the oracle and GPU parity tests."""
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_g1():
    g = json.loads((GOLDEN / "g1_spline_grid_5x6.json").read_text())
    cp = np.asfortranarray(np.array(g["control_points"], dtype=np.float64).transpose(1, 2, 0))
    ev = np.asfortranarray(np.array(g["eval"], dtype=np.float64).transpose(1, 2, 0))
    return g, cp, ev


def load_g2():
    g = json.loads((GOLDEN / "g2_error_informed_indices.json").read_text())
    return g, np.array(g["refinement_indices"], dtype=np.int32)


def julia_isapprox(a, b, rtol=None):
    """Julia's `a ≈ b` on arrays: ||a-b||_2 <= sqrt(eps) * max(||a||, ||b||)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if rtol is None:
        rtol = np.sqrt(np.finfo(a.dtype if a.dtype.kind == "f" else np.float64).eps)
    na, nb = np.linalg.norm(a.ravel()), np.linalg.norm(b.ravel())
    return np.linalg.norm((a.astype(np.float64) - b.astype(np.float64)).ravel()) <= rtol * max(na, nb)


def rel_err(a, b):
    """Norm-wise relative error ||a-b|| / ||b|| in float64 (the north_star's tolerance is applied to this)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (nb if nb > 0 else 1.0))


def max_rel_err(a, b):
    """Max-norm error relative to max|b|."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    mb = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (mb if mb > 0 else 1.0))


TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}
