"""Minimal raster container and the few terra / sf operations the grid-model preparation calls.

The reference's R layer holds rasters as terra `SpatRaster`s and calls third-party routines on them
before the `.Call` boundary: `terra::terrain` (slope / aspect, R/internal.R:1124-1129),
`terra::aggregate` + `terra::resample` (wind-shelter smoothing, R/internal.R:979-981), `terra::mask`
and `sf::st_transform` of the raster centre to EPSG:4326 (R/internal.R:59-68).  terra and sf are not
part of /root/reference (DESCRIPTION lists them unpinned) and there is no R here, so these are
restatements of the *published* algorithms (Horn 1981 for slope/aspect; block mean; bilinear
interpolation between cell centres; Redfearn/OSGB inverse transverse Mercator) — PARITY UNPINNED at
this step.  The C++ boundary itself stays pinned: whatever these produce is handed identically to the
CUDA path and to the reference oracle in the tests.

Array convention: `Raster.values` has shape (nrows, ncols, nlyr); row 0 is the northern edge (terra's
cell order), so `values[:, :, k]` is what the reference's `.is(r[[k]])` returns (R/internal.R:6-15).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np


@dataclass
class Raster:
    values: np.ndarray  # (nrows, ncols, nlyr) float64
    xmin: float
    xmax: float
    ymin: float
    ymax: float
    crs: str = ""

    def __post_init__(self):
        v = np.asarray(self.values, dtype=np.float64)
        if v.ndim == 2:
            v = v[:, :, None]
        self.values = v

    # -- geometry
    @property
    def nrows(self) -> int:
        return self.values.shape[0]

    @property
    def ncols(self) -> int:
        return self.values.shape[1]

    @property
    def nlyr(self) -> int:
        return self.values.shape[2]

    @property
    def dim(self) -> Tuple[int, int, int]:
        return self.values.shape

    @property
    def res(self) -> Tuple[float, float]:
        return ((self.xmax - self.xmin) / self.ncols, (self.ymax - self.ymin) / self.nrows)

    def matrix(self, lyr: int = 0) -> np.ndarray:
        return self.values[:, :, lyr]

    def like(self, values) -> "Raster":
        return Raster(np.array(values, dtype=np.float64), self.xmin, self.xmax, self.ymin, self.ymax, self.crs)

    def crop(self, r0: int, r1: int, c0: int, c1: int) -> "Raster":
        rx, ry = self.res
        return Raster(self.values[r0:r1, c0:c1, :].copy(), self.xmin + c0 * rx, self.xmin + c1 * rx,
                      self.ymax - r1 * ry, self.ymax - r0 * ry, self.crs)

    @staticmethod
    def from_packed(obj) -> "Raster":
        """terra PackedSpatRaster (as decoded by rdata.read_rda) -> Raster."""
        d = obj.slots["definition"][0]

        def num(key):
            m = re.search(rf"{key}=([-0-9.eE+]+)", d)
            if not m:
                raise ValueError(f"PackedSpatRaster definition lacks {key}")
            return float(m.group(1))

        ncols, nrows, nlyr = int(num("ncols")), int(num("nrows")), int(num("nlyrs"))
        m = re.search(r"crs='(.*)'\)?\s*$", d, flags=re.S)
        crs = m.group(1) if m else ""
        crs = crs[:-2] if crs.endswith("')") else crs
        vals = np.asarray(obj.slots["values"], dtype=np.float64)
        # values is an [ncell, nlyr] R matrix (column-major), cells in row-major raster order
        v = vals.reshape((nrows * ncols, nlyr), order="F").reshape((nrows, ncols, nlyr))
        return Raster(v.copy(), num("xmin"), num("xmax"), num("ymin"), num("ymax"), crs)


def as_raster(x, template: Optional[Raster] = None) -> Raster:
    if isinstance(x, Raster):
        return x
    if hasattr(x, "slots") and "PackedSpatRaster" in getattr(x, "cls", []):
        return Raster.from_packed(x)
    if template is not None:
        return template.like(x)
    raise TypeError("expected a Raster, a PackedSpatRaster or an array with a template")


# ---------------------------------------------------------------------------------------------
# terra::mask, terra::terrain, terra::aggregate, terra::resample (published algorithms)
# ---------------------------------------------------------------------------------------------
def mask(r: Raster, msk: Raster) -> Raster:
    """Cells that are NA in (the first layer of) `msk` become NA in every layer of `r`."""
    v = r.values.copy()
    v[np.isnan(msk.values[:, :, 0]), :] = np.nan
    return r.like(v)


def terrain(dtm: Raster, v: str = "slope", unit: str = "degrees") -> Raster:
    """Horn (1981) 8-neighbour slope / aspect on a projected raster; edge cells and cells with an NA
    neighbour are NA (terra::terrain(neighbors = 8)).  Aspect is the downslope direction, clockwise
    from north; a flat cell has aspect 90 degrees (atan2(0, 0) = 0 in terra's formula)."""
    z = dtm.values[:, :, 0]
    dx, dy = dtm.res
    out = np.full(z.shape, np.nan)
    if z.shape[0] >= 3 and z.shape[1] >= 3:
        a, b, c = z[:-2, :-2], z[:-2, 1:-1], z[:-2, 2:]
        d, f = z[1:-1, :-2], z[1:-1, 2:]
        g, h, i = z[2:, :-2], z[2:, 1:-1], z[2:, 2:]
        e = z[1:-1, 1:-1]
        dzdx = ((c + 2.0 * f + i) - (a + 2.0 * d + g)) / (8.0 * dx)   # towards the east
        dzdy = ((a + 2.0 * b + c) - (g + 2.0 * h + i)) / (8.0 * dy)   # towards the north
        if v == "slope":
            res = np.arctan(np.sqrt(dzdx * dzdx + dzdy * dzdy))
        elif v == "aspect":
            # downslope vector = (-dzdx, -dzdy); bearing clockwise from north = atan2(east, north)
            res = np.mod(0.5 * np.pi - np.arctan2(-dzdy, -dzdx), 2.0 * np.pi)
            flat = (dzdx == 0.0) & (dzdy == 0.0)
            res = np.where(flat, 0.5 * np.pi, res)
        else:
            raise ValueError("terrain: v must be 'slope' or 'aspect'")
        res = np.where(np.isnan(e), np.nan, res)
        out[1:-1, 1:-1] = res
    if unit == "degrees":
        out = out * (180.0 / np.pi)
    return dtm.like(out)


def aggregate_mean(r: Raster, fact: int, na_rm: bool = False) -> Raster:
    """Block mean over fact x fact cells (terra::aggregate(fun = "mean")); ragged edge blocks are
    averaged over the cells they have and the extent grows to whole blocks."""
    fact = int(fact)
    nr, nc, nl = r.dim
    onr, onc = -(-nr // fact), -(-nc // fact)
    out = np.empty((onr, onc, nl))
    for i in range(onr):
        for j in range(onc):
            blk = r.values[i * fact:(i + 1) * fact, j * fact:(j + 1) * fact, :].reshape(-1, nl)
            out[i, j, :] = np.nanmean(blk, axis=0) if na_rm else blk.mean(axis=0)
    rx, ry = r.res
    return Raster(out, r.xmin, r.xmin + onc * fact * rx, r.ymax - onr * fact * ry, r.ymax, r.crs)


def resample_bilinear(src: Raster, dst: Raster) -> Raster:
    """Bilinear interpolation of `src` at the cell centres of `dst` (terra::resample(method =
    "bilinear")).  Outside the hull of the source cell centres the nearest centre line is used (values
    are held constant over the outer half cell)."""
    sx, sy = src.res
    dxr, dyr = dst.res
    xc = dst.xmin + (np.arange(dst.ncols) + 0.5) * dxr
    yc = dst.ymax - (np.arange(dst.nrows) + 0.5) * dyr
    fx = (xc - src.xmin) / sx - 0.5   # fractional source column of each destination column
    fy = (src.ymax - yc) / sy - 0.5
    fx = np.clip(fx, 0.0, src.ncols - 1.0)
    fy = np.clip(fy, 0.0, src.nrows - 1.0)
    x0 = np.minimum(np.floor(fx).astype(int), max(src.ncols - 2, 0))
    y0 = np.minimum(np.floor(fy).astype(int), max(src.nrows - 2, 0))
    x1 = np.minimum(x0 + 1, src.ncols - 1)
    y1 = np.minimum(y0 + 1, src.nrows - 1)
    wx = (fx - x0)[None, :, None]
    wy = (fy - y0)[:, None, None]
    v = src.values
    top = v[y0][:, x0, :] * (1.0 - wx) + v[y0][:, x1, :] * wx
    bot = v[y1][:, x0, :] * (1.0 - wx) + v[y1][:, x1, :] * wx
    return Raster(top * (1.0 - wy) + bot * wy, dst.xmin, dst.xmax, dst.ymin, dst.ymax, dst.crs)


# ---------------------------------------------------------------------------------------------
# sf::st_transform(raster centre -> EPSG:4326)  (R/internal.R:59-68)
# ---------------------------------------------------------------------------------------------
def _wkt_param(crs: str, name: str) -> Optional[float]:
    m = re.search(rf'PARAMETER\["{re.escape(name)}",\s*([-0-9.eE+]+)', crs)
    return float(m.group(1)) if m else None


def _inverse_tmerc(E, N, a, invf, lat0, lon0, k0, FE, FN):
    """Inverse transverse Mercator (Ordnance Survey 'A guide to coordinate systems in Great Britain',
    annexe C): easting / northing -> geodetic latitude / longitude on the projection's own ellipsoid."""
    E, N = np.asarray(E, dtype=np.float64), np.asarray(N, dtype=np.float64)
    f = 1.0 / invf
    b = a * (1.0 - f)
    e2 = (a * a - b * b) / (a * a)
    n = (a - b) / (a + b)
    phi0, lam0 = math.radians(lat0), math.radians(lon0)

    def M(phi):
        dp, sp = phi - phi0, phi + phi0
        return b * k0 * ((1 + n + 1.25 * n * n + 1.25 * n ** 3) * dp
                         - (3 * n + 3 * n * n + 2.625 * n ** 3) * np.sin(dp) * np.cos(sp)
                         + (1.875 * n * n + 1.875 * n ** 3) * np.sin(2 * dp) * np.cos(2 * sp)
                         - (35.0 / 24.0) * n ** 3 * np.sin(3 * dp) * np.cos(3 * sp))

    phi = (N - FN) / (a * k0) + phi0
    for _ in range(100):
        m = M(phi)
        if np.all(np.abs(N - FN - m) < 1e-6):
            break
        phi = (N - FN - m) / (a * k0) + phi
    s2 = np.sin(phi) ** 2
    nu = a * k0 / np.sqrt(1 - e2 * s2)
    rho = a * k0 * (1 - e2) * (1 - e2 * s2) ** -1.5
    eta2 = nu / rho - 1.0
    t = np.tan(phi)
    sec = 1.0 / np.cos(phi)
    VII = t / (2 * rho * nu)
    VIII = t / (24 * rho * nu ** 3) * (5 + 3 * t * t + eta2 - 9 * t * t * eta2)
    IX = t / (720 * rho * nu ** 5) * (61 + 90 * t * t + 45 * t ** 4)
    X = sec / nu
    XI = sec / (6 * nu ** 3) * (nu / rho + 2 * t * t)
    XII = sec / (120 * nu ** 5) * (5 + 28 * t * t + 24 * t ** 4)
    XIIA = sec / (5040 * nu ** 7) * (61 + 662 * t * t + 1320 * t ** 4 + 720 * t ** 6)
    dE = E - FE
    lat = phi - VII * dE ** 2 + VIII * dE ** 4 - IX * dE ** 6
    lon = lam0 + X * dE - XI * dE ** 3 + XII * dE ** 5 - XIIA * dE ** 7
    return np.degrees(lat), np.degrees(lon)


def latlong_from_xy(crs: str, x: float, y: float) -> Tuple[float, float]:
    """(lat, long) in degrees of a point given in the raster's CRS.  Geographic CRSs pass through;
    transverse Mercator CRSs (e.g. the bundled rasters' British National Grid definition) are
    inverted on their own ellipsoid.  No datum shift is applied (the bundled CRS has an unknown datum,
    for which PROJ applies none either)."""
    if not crs or re.search(r"^\s*(GEOGCRS|GEOGCS)\[", crs) or "+proj=longlat" in crs:
        return y, x
    if "Transverse Mercator" in crs or "Transverse_Mercator" in crs or "+proj=tmerc" in crs:
        m = re.search(r'ELLIPSOID\["[^"]*",\s*([-0-9.eE+]+),\s*([-0-9.eE+]+)', crs)
        if not m:
            m = re.search(r'SPHEROID\["[^"]*",\s*([-0-9.eE+]+),\s*([-0-9.eE+]+)', crs)
        if not m:
            raise ValueError("cannot find the ellipsoid in the CRS definition")
        a, invf = float(m.group(1)), float(m.group(2))
        get = lambda *names: next((v for v in (_wkt_param(crs, n) for n in names) if v is not None), None)  # noqa: E731
        lat0 = get("Latitude of natural origin", "latitude_of_origin")
        lon0 = get("Longitude of natural origin", "central_meridian")
        k0 = get("Scale factor at natural origin", "scale_factor")
        FE = get("False easting", "false_easting")
        FN = get("False northing", "false_northing")
        if None in (lat0, lon0, k0, FE, FN):
            raise ValueError("incomplete transverse Mercator parameters in the CRS definition")
        return _inverse_tmerc(x, y, a, invf, lat0, lon0, k0, FE, FN)
    raise ValueError("unsupported CRS: supply lat / long explicitly")


def latlong_from_raster(r: Raster) -> Tuple[float, float]:
    """ref .latlongfromraster (R/internal.R:59-68): centre of the extent, transformed to EPSG:4326."""
    lat, lon = latlong_from_xy(r.crs, 0.5 * (r.xmin + r.xmax), 0.5 * (r.ymin + r.ymax))
    return float(lat), float(lon)


def latslons_from_raster(r: Raster) -> Tuple[np.ndarray, np.ndarray]:
    """ref .latslonsfromr (R/internal.R:86-99): latitude / longitude of every cell centre, [nrows, ncols]."""
    rx, ry = r.res
    xc = r.xmin + (np.arange(r.ncols) + 0.5) * rx
    yc = r.ymax - (np.arange(r.nrows) + 0.5) * ry
    X, Y = np.meshgrid(xc, yc)
    lat, lon = latlong_from_xy(r.crs, X, Y)
    return np.asarray(lat, dtype=np.float64), np.asarray(lon, dtype=np.float64)
