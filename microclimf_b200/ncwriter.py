"""`writetonc` (R/dataprep.R:1063-1260): hourly model outputs as a netCDF file of integers.

The reference writes NetCDF-4 through `ncdf4` (variables `prec = "integer"`, `missval = -9999`, deflate level 9),
dimensions east / north / time, values `as.integer(round(x * rd))` with rd = 100 for temperatures, soil moisture and
wind speed and 1 for humidity and radiation, array order `aperm(a, c(2, 1, 3))`, plus a `crs` variable carrying the WKT.
Here the integers come straight from the kernels' packed sink (`mcf_runmicro_packed`) and the container is NetCDF-3
classic through `scipy.io.netcdf_file` (neither netCDF4 nor HDF5 bindings exist in this environment, so there is no
deflate): same dimensions, variable names, units, long names, missing value, time encoding and CRS attributes.

Two of the reference's own quirks are NOT reproduced because they are plainly unintended: its `ncvar_put` calls for the
radiation streams test for names that are never in `vars` (`"raddir" %in% vars`, R/dataprep.R:1168-1172) and the
soil-moisture put refers to an undefined handle (`nccew`, :1166), so the reference leaves those variables empty or
errors; this writer fills every variable it defines.
"""
from __future__ import annotations

import numpy as np

from .spatial import Raster

_LONG = {
    "Tz": ("Air temperature at height {h} m", "deg C x 100"), "tleaf": ("Leaf temperature at height {h} m", "deg C x 100"),
    "relhum": ("Relative humidity at height {h} m", "Percentage"),
    "soilm": ("Soil surface moisture", "Volume percentage soil moisture in top 10 cm of soil"),
    "windspeed": ("Wind speed at height {h} m", "m/s x 100"),
    "Rdirdown": ("Downward direct shortwave radiation", "W/m^2"), "Rdifdown": ("Downward diffuse shortwave radiation", "W/m^2"),
    "Rlwdown": ("Downward longwave radiation", "W/m^2"), "Rswup": ("Upward shortwave radiation", "W/m^2"),
    "Rlwup": ("Upward longwave radiation", "W/m^2"),
}
DEFAULT_VARS = {  # R/dataprep.R:1112, 1180, 1236
    "above": ("Tz", "tleaf", "relhum", "windspeed", "Rdirdown", "Rdifdown", "Rlwdown", "Rswup", "Rlwup"),
    "surface": ("Tz", "soilm", "Rdirdown", "Rdifdown", "Rlwdown", "Rswup", "Rlwup"),
    "below": ("Tz", "soilm"),
}


def writetonc(mout_packed, fileout: str, dtm: Raster, reqhgt: float, vars=None) -> None:
    """`mout_packed`: runmicro(..., packed=True) — int16 arrays [rows, cols, hours] (-9999 = NA) and `tme`."""
    from scipy.io import netcdf_file

    regime = "above" if reqhgt > 0 else ("surface" if reqhgt == 0 else "below")
    vars = tuple(vars) if vars is not None else DEFAULT_VARS[regime]
    rx, ry = dtm.res
    est = dtm.xmin + rx / 2 + rx * np.arange(dtm.ncols)
    nth = dtm.ymin + ry / 2 + ry * np.arange(dtm.nrows)
    tme = np.asarray(mout_packed["tme"]).astype("datetime64[s]").astype(np.int64) / 3600.0
    with netcdf_file(fileout, "w", version=2) as nc:
        nc.createDimension("east", est.size)
        nc.createDimension("north", nth.size)
        nc.createDimension("time", tme.size)
        for name, vals, units, longname in (("east", est, "metres", "Eastings"), ("north", nth, "metres", "Northings"),
                                            ("time", tme, "hours since 1970-01-01 00:00", "time")):
            v = nc.createVariable(name, "d", (name,))
            v[:] = vals
            v.units = units
            v.long_name = longname
        nc.variables["time"].standard_name = "time"
        nc.variables["time"].calendar = "gregorian"
        crs = nc.createVariable("crs", "i", ())
        crs.data[...] = 1
        crs.crs_wkt = dtm.crs
        crs.grid_mapping_name = "longitude_latitude"
        for name in vars:
            if name not in mout_packed:
                continue
            longname, units = _LONG[name]
            if regime == "surface" and name == "Tz":
                longname = "Soil surface temperature"
            if regime == "below" and name == "Tz":
                longname = f"Soil temperature at depth {abs(reqhgt)} m"
            a = np.asarray(mout_packed[name])
            # ncdf4 stores the first dimension (east) fastest: the on-disk order is [time, north, east]
            v = nc.createVariable(name, "i", ("time", "north", "east"))
            v[:] = np.ascontiguousarray(np.transpose(a, (2, 0, 1))).astype(np.int32)
            v.long_name = longname.format(h=reqhgt)
            v.units = units
            v.missing_value = np.int32(-9999)
            v._FillValue = np.int32(-9999)
            v.grid_mapping = "crs"
