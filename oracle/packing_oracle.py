"""TEST INFRASTRUCTURE ONLY — numpy restatement of writetonc's integer packing (R/dataprep.R:1063-1260).

`atonc` (R/dataprep.R:1064-1069): `a <- aperm(a, c(2,1,3)); a <- round(a * rd, 0); as.integer(a)`, with
rd = 100 for Tz, tleaf, soilm, windspeed and 1 for relhum and the five radiation streams (:1164-1173); NA (and
NaN / Inf, for which as.integer gives NA) is written as the variable's missval -9999 by ncdf4.  R's round(x, 0)
rounds half to even, as numpy does.  PARITY UNPINNED (no R here); pinned only by reading the R source.
"""
import numpy as np

SCALE = dict(Tz=100.0, tleaf=100.0, relhum=1.0, soilm=100.0, windspeed=100.0, Rdirdown=1.0, Rdifdown=1.0, Rlwdown=1.0,
             Rswup=1.0, Rlwup=1.0)
NA = -9999


def pack(name, a):
    """[rows, cols, T] FP64 -> the integers writetonc stores, same layout (the aperm is applied by file_layout)."""
    with np.errstate(invalid="ignore"):
        s = np.asarray(a, dtype=np.float64) * SCALE[name]
        r = np.round(s)
    out = np.where(np.isfinite(s), np.clip(r, -32767, 32767), NA)
    return out.astype(np.int16)


def file_layout(a):
    """aperm(a, c(2, 1, 3)): [east, north, time] as written to the netCDF variable."""
    return np.transpose(a, (1, 0, 2))
