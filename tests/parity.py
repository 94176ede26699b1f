"""Comparison helpers shared by the parity tests (CUDA path vs CPU checker).

Tolerance (BASELINE.json north_star, FP64 build): 1e-6 degC absolute / 1e-6 relative — a value passes
when |a - b| <= ATOL + RTOL * |b|.  NA cells must match exactly (same NaN mask)."""
import numpy as np

ATOL = 1e-6
RTOL = 1e-6


def compare(got: dict, want: dict, atol: float = ATOL, rtol: float = RTOL):
    """Returns (ok, report rows).  Each row: (name, max_abs_err, max_excess_ratio, nan_mismatch)."""
    rows, ok = [], True
    for name, w in want.items():
        g = got[name]
        assert g.shape == w.shape, (name, g.shape, w.shape)
        gn, wn = np.isnan(g), np.isnan(w)
        mism = int((gn != wn).sum())
        both = ~gn & ~wn
        if both.any():
            d = np.abs(g[both] - w[both])
            lim = atol + rtol * np.abs(w[both])
            mx, ex = float(d.max()), float((d / lim).max())
        else:
            mx, ex = 0.0, 0.0
        rows.append((name, mx, ex, mism))
        if mism or ex > 1.0:
            ok = False
    return ok, rows


def fmt(rows):
    return "\n".join(f"  {n:10s} max|err|={mx:.3e}  err/tol={ex:.3e}  nan_mismatch={mm}" for n, mx, ex, mm in rows)
