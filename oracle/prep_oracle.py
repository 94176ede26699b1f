"""TEST INFRASTRUCTURE ONLY — numpy restatement of what `.runmodel2Cpp` / `.runmodel4Cpp` do to the coarse-grid
climate before the `.Call` (R/internal.R:1219-1277): expand every coarse series to the fine raster (`.cca`,
R/internal.R:523-542: terra::resample, bilinear), derive es / ea / tdew from the resampled temperature and relative
humidity (`.satvap` :501, `.dewpoint` :509), apply the altitude correction to pressure and temperature (:1226-1245,
`.lapserate` :546) and rebuild the wind speed from its resampled components (:1250-1259).

The result is the fine-array problem the reference's runmicro2Cpp / runmicro4Cpp take, so the kernels' fused
interpolation (mcf_problem.clim_rows > 0) can be checked against the compiled reference.

PARITY UNPINNED for the resample step itself: terra is third-party and absent; bilinear interpolation between
cell centres with values held constant over the outer half cell is its published behaviour.
"""
import numpy as np

from microclimf_b200.problem import GridProblem


def _weights(n_fine, n_coarse, f0, df):
    f = np.clip(f0 + df * np.arange(n_fine), 0.0, n_coarse - 1.0)
    i0 = np.minimum(np.floor(f).astype(int), max(n_coarse - 2, 0))
    i1 = np.minimum(i0 + 1, n_coarse - 1)
    return i0, i1, f - i0


def resample(p: GridProblem, coarse_flat):
    """[crows, ccols, T] (R order, flat) -> [T, ncells] (ncells in R order: i + rows * j)."""
    T = p.tsteps
    a = np.asarray(coarse_flat).reshape(T, p.clim_cols, p.clim_rows)  # [k, cj, ci]
    y0, y1, wy = _weights(p.rows, p.clim_rows, p.clim_row0, p.clim_drow)
    x0, x1, wx = _weights(p.cols, p.clim_cols, p.clim_col0, p.clim_dcol)
    wxb, wyb = wx[None, :, None], wy[None, None, :]
    top = a[:, x0][:, :, y0] * (1.0 - wxb) + a[:, x1][:, :, y0] * wxb   # [k, j, i]
    bot = a[:, x0][:, :, y1] * (1.0 - wxb) + a[:, x1][:, :, y1] * wxb
    return (top * (1.0 - wyb) + bot * wyb).reshape(T, p.cols * p.rows)


def resample_cells(p: GridProblem, coarse_flat, pick):
    """The same interpolation for the cells `pick` only (flat R-order indices): [T, len(pick)].  Used to check rasters too
    large to expand whole (BASELINE configs[4]: 4096 x 4096 cells)."""
    T = p.tsteps
    a = np.asarray(coarse_flat).reshape(T, p.clim_cols, p.clim_rows)  # [k, cj, ci]
    y0, y1, wy = _weights(p.rows, p.clim_rows, p.clim_row0, p.clim_drow)
    x0, x1, wx = _weights(p.cols, p.clim_cols, p.clim_col0, p.clim_dcol)
    i, j = np.asarray(pick) % p.rows, np.asarray(pick) // p.rows
    wxc, wyc = wx[j][None, :], wy[i][None, :]
    top = a[:, x0[j], y0[i]] * (1.0 - wxc) + a[:, x1[j], y0[i]] * wxc
    bot = a[:, x0[j], y1[i]] * (1.0 - wxc) + a[:, x1[j], y1[i]] * wxc
    return top * (1.0 - wyc) + bot * wyc


def _satvap(tc):
    es = 0.61078 * np.exp(17.27 * tc / (tc + 237.3))
    ei = 0.61078 * np.exp(21.875 * tc / (tc + 265.5))
    return np.where(tc < 0, ei, es)


def _dewpoint(ea, tc):
    e0 = 611.2 / 1000
    L = (2.501 * 10 ** 6) - (2340 * tc)
    it = 1 / 273.15 - (461.5 / L) * np.log(ea / e0)
    Tdew = 1 / it - 273.15
    e0 = 610.78 / 1000
    L = 2.834 * 10 ** 6
    it = 1 / 273.15 - (461.5 / L) * np.log(ea / e0)
    Tfrost = 1 / it - 273.15
    return np.where(Tdew < 0, Tfrost, Tdew)


def _lapserate(tc, ea, pk):
    rv = 0.622 * ea / (pk - ea)
    return 9.8076 * (1 + (2501000 * rv) / (287 * (tc + 273.15))) / (1003.5 + (0.622 * 2501000 ** 2 * rv) / (287 * (tc + 273.15) ** 2))


def materialise_coarse(p: GridProblem, pick=None) -> GridProblem:
    """The fine-array (reference layout) problem equivalent to a coarse-grid problem.  With `pick` (flat R-order cell
    indices): the same for those cells only, as a len(pick) x 1 raster — cells are independent, so a sample of a raster
    too large to expand whole is the same problem for the sampled cells."""
    assert p.coarse
    q = p.replace(clim_rows=0, clim_cols=0, altcorrect=0)
    q.arrays = {n: a for n, a in p.arrays.items() if n not in ("relhum", "wu", "wv", "elevd", "pfac")}
    rs = (lambda a: resample(p, a)) if pick is None else (lambda a: resample_cells(p, a, pick))
    sel = (lambda a: np.asarray(a)) if pick is None else (lambda a: np.asarray(a)[pick])
    tc = rs(p.arrays["temp"])
    rh = rs(p.arrays["relhum"])
    es = _satvap(tc)
    ea = es * rh / 100
    tdew = _dewpoint(ea, tc)
    pk = rs(p.arrays["pres"])
    if p.altcorrect:
        pk = pk * sel(p.arrays["pfac"])[None, :]
        elevd = sel(p.arrays["elevd"])[None, :]
        tcdif = elevd * (5 / 1000) if p.altcorrect == 1 else _lapserate(tc, ea, pk) * elevd
        tc = tcdif + tc
    wu, wv = rs(p.arrays["wu"]), rs(p.arrays["wv"])
    fine = dict(temp=tc, es=es, ea=ea, tdew=tdew, pres=pk, windspeed=np.sqrt(wu ** 2 + wv ** 2))
    for n in ("swdown", "difrad", "lwdown", "p_soilm", "p_G", "p_umu", "p_kp", "p_muGp", "p_dtrp", "p_Tg", "p_Tbp"):
        if n in p.arrays:
            fine[n] = rs(p.arrays[n])
    if pick is not None:  # static layers of the sampled cells
        nc = p.ncells
        q.rows, q.cols, q.twi_mean = len(pick), 1, None
        for n in list(q.arrays):
            ln = p.expected_len(n)
            if ln % nc == 0 and ln >= nc and n not in fine and ln != p.tsteps:
                q.arrays[n] = np.ascontiguousarray(np.asarray(q.arrays[n]).reshape(ln // nc, nc)[:, pick].ravel())
    for n, a in fine.items():
        q.arrays[n] = np.ascontiguousarray(a).ravel()
    q.validate()
    return q
