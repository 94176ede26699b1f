"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's pure-R terrain preparation that feeds
the grid solver (SURVEY.md §8f NEXT-2): `.horizon` (R/internal.R:909-925), the sky-view factor
(R/internal.R:1146-1148), `.windcoef` (R/internal.R:949-968) and the 16 -> 8 direction blend of
`.windsheltera` (R/internal.R:983-989).

PARITY UNPINNED: there is no R interpreter in the build container and the reference's tests hold no
vectors for these functions, so this restatement is pinned only by reading the R source.  The R semantics
it relies on:
  * `a:b` with a non-integer `a` is the REAL sequence a, a+1, ... (length floor(b - a + 1e-10) + 1), each
    element computed as a + k in double precision;
  * a REAL subscript is truncated toward zero;
  * `.is()` of a raster is the [rows, cols] matrix (row 1 = north), NA elevations become 0, elevations
    are divided by the cell size, and the DTM is padded with 100 cells of zeros on every side.
`terra::aggregate` + `terra::resample` inside `.windsheltera` (R/internal.R:980) and `terra::terrain`
(slope / aspect) are third-party arithmetic outside /root/reference and are not restated.
"""
import numpy as np


def _shifted(dtm3, x, y, azi, step):
    """dtm3[(101 - cos(azi) step^2):(x + 100 - cos(azi) step^2), (101 + sin(azi) step^2):(y + 100 + sin(azi) step^2)]"""
    s2 = float(step * step)
    fr = 101 - np.cos(azi) * s2
    fc = 101 + np.sin(azi) * s2
    ri = np.trunc(fr + np.arange(x, dtype=np.float64)).astype(np.int64) - 1  # R is 1-based
    ci = np.trunc(fc + np.arange(y, dtype=np.float64)).astype(np.int64) - 1
    return dtm3[np.ix_(ri, ci)]


def horizon(dtm, azimuth_deg, reso):
    """.horizon (R/internal.R:909-925): tangent of the horizon angle in one direction."""
    d = np.array(dtm, dtype=np.float64, copy=True)
    d[np.isnan(d)] = 0.0
    d = d / reso
    azi = azimuth_deg * (np.pi / 180)
    x, y = d.shape
    hor = np.zeros((x, y))
    dtm3 = np.zeros((x + 200, y + 200))
    dtm3[100:x + 100, 100:y + 100] = d
    for step in range(1, 11):
        hor = np.maximum(hor, (_shifted(dtm3, x, y, azi, step) - d) / float(step * step))
    return hor


def horizon24(dtm, reso):
    """soilc$hor (R/internal.R:1142-1145): 24 directions, 15 degrees apart, layer i = (i - 1) * 15 degrees."""
    return np.stack([horizon(dtm, i * 15, reso) for i in range(24)], axis=2)


def skyview(hor):
    """soilc$svfa (R/internal.R:1146-1148)."""
    msl = np.tan(np.mean(np.arctan(hor), axis=2))
    return 0.5 * np.cos(2 * msl) + 0.5


def windcoef(dsm, direction_deg, hgt, reso):
    """.windcoef (R/internal.R:949-968): wind shelter coefficient in one direction."""
    d = np.array(dsm, dtype=np.float64, copy=True)
    d[np.isnan(d)] = 0.0
    d = d / reso
    h = hgt / reso
    azi = direction_deg * (np.pi / 180)
    x, y = d.shape
    hor = np.zeros((x, y))
    dtm3 = np.zeros((x + 200, y + 200))
    dtm3[100:x + 100, 100:y + 100] = d
    for step in range(1, 11):
        s2 = float(step * step)
        hor = np.maximum(hor, (_shifted(dtm3, x, y, azi, step) - d) / s2)
        hor = np.where(hor < (h / s2), 0.0, hor)
    return 1 - np.arctan(0.17 * 100 * hor) / 1.65


def windcoef16(dsm, hgt, reso):
    return np.stack([windcoef(dsm, i * 360 / 16, hgt, reso) for i in range(16)], axis=2)


def blend16to8(a):
    """The 16 -> 8 direction blend at the end of .windsheltera (R/internal.R:983-989)."""
    out = np.empty(a.shape[:2] + (8,))
    for i in range(1, 9):
        if i == 1:
            out[:, :, 0] = 0.5 * a[:, :, 0] + 0.25 * a[:, :, 1] + 0.25 * a[:, :, 15]
        else:
            out[:, :, i - 1] = 0.5 * a[:, :, i * 2 - 2] + 0.25 * a[:, :, i * 2 - 1] + 0.25 * a[:, :, i * 2 - 3]
    return out
