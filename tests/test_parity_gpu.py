"""Parity of the CUDA path (through the C ABI) against the CPU checker on seeded synthetic inputs.

Checker: oracle/_ref (the unmodified reference C++, src/microclimfCpp.cpp, compiled against the Rcpp
stand-in) when present, else the C restatement.  Tolerance: 1e-6 abs + 1e-6 rel (tests/parity.py)."""
import numpy as np
import pytest

import parity
from microclimf_b200 import _abi, _lib, api, synth
from oracle import pyoracle

pytestmark = pytest.mark.gpu
KIND = "ref" if pyoracle.have_ref() else "oracle"


def _check(p, out_mask=None):
    want = pyoracle.runmicro(p, out_mask=out_mask, kind=KIND)
    got = api.run_problem(p, out=out_mask)
    assert set(got) == set(want)
    ok, rows = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rows)
    return got


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("reqhgt", [0.05, 0.0, -0.1, 1.0, 5.0, 40.0])
def test_runmicro_modes_heights(mode, reqhgt):
    """Every driver (ref runmicro1..4Cpp :2052/:2340/:2624/:2926) x below-canopy, surface, below-ground,
    mid-canopy and above-canopy heights; grid with NA cells, bare cells, x == 1, clump == 0, flat cells."""
    p = synth.make_problem(37, 29, 24 * 5, reqhgt=reqhgt, mode=mode, nlyr=3, zref=45.0 if reqhgt > 30 else 30.0)
    _check(p)


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("reqhgt", [-0.02, -0.5, -3.0])
@pytest.mark.parametrize("complete", [True, False])
def test_below_ground_branches(mode, reqhgt, complete):
    """Tbelowgroundv (ref :1474-1539): n <= 48 direct window, n > 48 daily path, n >= T series mean;
    incomplete-series blends (nb <= 1, <= 24, < hiy, >= hiy)."""
    p = synth.make_problem(19, 23, 24 * 12, reqhgt=reqhgt, mode=mode, nlyr=2, complete=complete)
    _check(p)


def test_deep_soil_full_mean():
    """n >= tsteps: every hour is the series mean (ref :1488-1493)."""
    p = synth.make_problem(8, 8, 48, reqhgt=-2.0, mode=1)
    _check(p)


@pytest.mark.parametrize("tsteps", [24, 30, 47, 49, 71])
def test_ragged_hours(tsteps):
    """ndays = tsteps / 24 (ref :2116): trailing hours are never computed and stay NA."""
    p = synth.make_problem(9, 7, tsteps, reqhgt=0.05, mode=1)
    got = _check(p)
    nd = tsteps // 24
    assert np.isnan(got["Tz"][:, :, nd * 24:]).all()
    p = synth.make_problem(9, 7, tsteps, reqhgt=-0.1, mode=2)
    _check(p)


def test_output_mask_and_na_bits():
    """out[] gating (ref :2131-2151) and R NA_real_ payload in skipped cells."""
    p = synth.make_problem(16, 16, 48, reqhgt=0.05, mode=1)
    mask = [True, False, False, True, False, True, False, True, False, True]
    got = _check(p, mask)
    assert set(got) == {n for n, m in zip(_abi.OUT_NAMES, mask) if m}
    hgt = p.arrays["hgt"].reshape(p.rows, p.cols, order="F")
    na = np.isnan(hgt)
    assert na.any()
    bits = got["Tz"].view(np.uint64)[na]
    assert (bits == _abi.NA_REAL_BITS).all()
    # masks R applies for reqhgt == 0 and < 0 (R/internal.R:1159-1166)
    _check(synth.make_problem(12, 12, 48, reqhgt=0.0, mode=1), [1, 0, 0, 1, 0, 1, 1, 1, 1, 1])
    _check(synth.make_problem(12, 12, 48, reqhgt=-0.1, mode=1), [1, 0, 0, 1, 0, 0, 0, 0, 0, 0])


def test_single_cell_and_single_column():
    _check(synth.make_problem(1, 1, 48, reqhgt=0.05, mode=1, zref=30.0))
    _check(synth.make_problem(130, 1, 24, reqhgt=0.05, mode=3, nlyr=1))
    _check(synth.make_problem(1, 131, 24, reqhgt=1.0, mode=2))


def test_all_na_grid():
    p = synth.make_problem(6, 5, 24, reqhgt=0.05, mode=1)
    p.arrays["hgt"] = np.full(p.ncells, np.nan)
    got = _check(p)
    assert all(np.isnan(v).all() for v in got.values())


@pytest.mark.parametrize("mode", [1, 3])
@pytest.mark.parametrize("reqhgt", [0.05, 0.0, 1.0])
def test_real_raster_corner_inputs(mode, reqhgt):
    """Input patterns of the reference's bundled rasters that seeded uniform draws never produce: bare cells whose
    leaf reflectance / transmittance stay NA (canopycondCpp then returns Gs = 9999.99, ref :463-464), clumping
    factors down to 1e-62, centimetre-high vegetation below reqhgt, leaf reflectance 1e-4, x up to 3."""
    p = synth.make_problem(31, 23, 24 * 3, reqhgt=reqhgt, mode=mode, nlyr=2, seed=77)
    rng = np.random.default_rng(5)
    nl = p.nlyr if p.layered else 1
    hgt = p.arrays["hgt"].reshape(nl, -1)
    pai = p.arrays["pai"].reshape(nl, -1)
    bare = (hgt[0] == 0)
    assert bare.sum() > 10
    nan_lr = bare & (rng.random(bare.size) < 0.6)
    veg_nan = (~bare) & ~np.isnan(hgt[0]) & (rng.random(bare.size) < 0.03)  # NA reflectance under a real canopy too
    for n in ("leafr", "leaft"):
        a = p.arrays[n].reshape(nl, -1)
        a[:, nan_lr | veg_nan] = np.nan
    tiny = (~bare) & (rng.random(bare.size) < 0.15)
    p.arrays["clump"].reshape(nl, -1)[:, tiny] = 10.0 ** rng.uniform(-62, -3, tiny.sum())
    short = (~bare) & ~np.isnan(hgt[0]) & (rng.random(bare.size) < 0.15)
    hgt[:, short] = rng.uniform(0.008, 0.04, short.sum())
    pai[:, short] = rng.uniform(0.008, 0.05, short.sum())
    for l in range(nl):
        pa, ld = synth.foliage_density(max(reqhgt, 0.0), hgt[l], pai[l])
        above = ~(max(reqhgt, 0.0) < hgt[l])
        p.arrays["paia"].reshape(nl, -1)[l] = np.where(above | bare, 0.0, pa)
        p.arrays["leafden"].reshape(nl, -1)[l] = np.where(above | bare, 0.0, ld)
    lowr = (~bare) & (rng.random(bare.size) < 0.1)
    p.arrays["leafr"].reshape(nl, -1)[:, lowr & ~veg_nan] = 1e-4
    p.arrays["leaft"].reshape(nl, -1)[:, lowr & ~veg_nan] = 5e-5
    p.arrays["x"].reshape(nl, -1)[:, rng.random(bare.size) < 0.1] = 3.0
    _check(p)


def test_latitude_classes_and_seasons():
    """Stomatal classes by |lat| (ref stomparamsCpp :391-440) and polar day / night solar geometry."""
    for lat, doy in ((10.0, 80), (-35.0, 355), (65.0, 172), (78.0, 355)):
        p = synth.make_problem(20, 20, 72, reqhgt=0.5, mode=1, lat=lat, lon=20.0, start_doy=doy)
        _check(p)
    p = synth.make_problem(20, 20, 72, reqhgt=0.5, mode=2, lat=-12.0, lon=140.0, start_doy=10)
    _check(p)


def test_layer_spans_and_errors():
    """dfsel spans (ref :2629-2639): uneven spans, a gap before the first layer, a trailing partial day;
    a span shorter than a day raises like the reference's Rcpp::stop."""
    p = synth.make_problem(10, 11, 24 * 9, reqhgt=0.05, mode=3, nlyr=3)
    p.lyr_st = np.array([24, 72, 150], dtype=np.int32)
    p.lyr_ed = np.array([71, 149, 215], dtype=np.int32)
    _check(p)
    p.lyr_ed = np.array([40, 149, 215], dtype=np.int32)
    with pytest.raises(_lib.McfError) as ei:
        api.run_problem(p)
    assert ei.value.code == _abi.MCF_ERR_ARG and "Too many layers" in ei.value.msg
    with pytest.raises(RuntimeError):
        pyoracle.runmicro(p, kind=KIND)


def test_year_long_series_band_equivalence():
    """A full year for a small band, solved whole and as two column bands with the twi mean supplied
    (the multi-GPU sharding contract): identical results."""
    from microclimf_b200 import bands

    p = synth.make_problem(8, 10, 8760, reqhgt=0.05, mode=1)
    whole = _check(p)
    s, n = bands.twi_partial_host(p.arrays["twi"], p.tfact)
    parts = []
    for c0, c1 in bands.band_ranges(p.cols, 2):
        b = p.band(c0, c1)
        b.twi_mean = s / n
        parts.append(api.run_problem(b))
    for nm in whole:
        glued = np.concatenate([pt[nm] for pt in parts], axis=1)
        np.testing.assert_allclose(glued, whole[nm], rtol=1e-12, atol=1e-12, equal_nan=True)


def test_twi_partial_matches_numpy():
    import ctypes as C

    from microclimf_b200 import bands

    rng = np.random.default_rng(3)
    twi = np.exp(rng.uniform(0.5, 3.0, 100_003))
    twi[::97] = np.nan
    L = _lib.lib()
    s, n = C.c_double(), C.c_int64()
    err = C.create_string_buffer(256)
    rc = L.mcf_twi_partial(twi.ctypes.data_as(C.POINTER(C.c_double)), twi.size, 1.5, C.byref(s), C.byref(n), err, 256)
    assert rc == 0, err.value
    hs, hn = bands.twi_partial_host(twi, 1.5)
    assert n.value == hn
    assert abs(s.value - hs) <= 1e-9 * abs(hs)


def test_device_window_ring_matches_whole():
    """mcf_runmicro_dev with a partial window into a 24-hour ring == the same hours of the whole run."""
    import torch

    p = synth.make_problem(33, 17, 24 * 6, reqhgt=0.05, mode=1)
    whole = api.run_problem(p)
    dp = p.to_device()
    nc = p.ncells
    outs = [torch.full((24 * nc,), -1.0, dtype=torch.float64, device="cuda") for _ in range(10)]
    api.run_problem_dev(dp, outs, window=(2, 3, 48, 24))  # days 2..4; the ring ends up holding day 4
    torch.cuda.synchronize()
    for nm, t in zip(_abi.OUT_NAMES, outs):
        ring = t.cpu().numpy().reshape(p.rows, p.cols, 24, order="F")
        np.testing.assert_array_equal(ring, whole[nm][:, :, 96:120])


@pytest.mark.parametrize("window", [(0, 5, 24, 48), (2, 2, 49, 24), (1, 1, -1, 24), (0, 1, 0, 23), (5, 2, 120, 24)])
def test_device_window_rejected_before_any_write(window):
    """A window origin beyond the window's first hour would make the kernels' (k - hour0) % ring negative and index in
    front of the output buffers: every sink rejects it (MCF_ERR_ARG) and leaves the buffers untouched."""
    import torch

    p = synth.make_problem(9, 7, 24 * 6, reqhgt=0.05, mode=1)
    dp = p.to_device()
    ring = max(window[3], 24)
    outs = [torch.full((ring * p.ncells,), -1.0, dtype=torch.float64, device="cuda") for _ in range(10)]
    outs16 = [torch.full((ring * p.ncells,), 7, dtype=torch.int16, device="cuda") for _ in range(10)]
    outsf = [torch.full((ring * p.ncells,), -1.0, dtype=torch.float32, device="cuda") for _ in range(10)]
    for fn, bufs in ((api.run_problem_dev, outs), (api.run_problem_packed_dev, outs16), (api.run_problem_f32_dev, outsf)):
        with pytest.raises(_lib.McfError) as ei:
            fn(dp, bufs, window=window)
        assert ei.value.code == _abi.MCF_ERR_ARG
    torch.cuda.synchronize()
    assert all(bool((t == -1.0).all()) for t in outs + outsf) and all(bool((t == 7).all()) for t in outs16)


@pytest.mark.parametrize("ring_hours,hour0", [(31, 48), (24, 37), (50, 0)])
def test_device_window_ring_seam_inside_a_day(ring_hours, hour0):
    """A ring whose seam falls inside a day block (ring_hours not a multiple of 24, or a window origin that is not a
    day boundary): pass 1 walks the ring forwards and pass 2 backwards across the seam; every slot must hold the hour
    (slot + hour0) mod ring of the window's last ring_hours hours — FP64, packed and FP32 sinks."""
    import torch

    from oracle import packing_oracle

    p = synth.make_problem(29, 13, 24 * 6, reqhgt=0.05, mode=1)
    whole = api.run_problem(p)
    dp = p.to_device()
    nc = p.ncells
    block0, nblocks = 2, 3
    k_last = (block0 + nblocks) * 24 - 1
    hours = np.arange(max(block0 * 24, k_last - ring_hours + 1), k_last + 1)
    slots = (hours - hour0) % ring_hours
    outs = [torch.full((ring_hours * nc,), -1.0, dtype=torch.float64, device="cuda") for _ in range(10)]
    api.run_problem_dev(dp, outs, window=(block0, nblocks, hour0, ring_hours))
    outs16 = [torch.full((ring_hours * nc,), 7, dtype=torch.int16, device="cuda") for _ in range(10)]
    api.run_problem_packed_dev(dp, outs16, window=(block0, nblocks, hour0, ring_hours))
    outsf = [torch.full((ring_hours * nc,), -1.0, dtype=torch.float32, device="cuda") for _ in range(10)]
    api.run_problem_f32_dev(dp, outsf, window=(block0, nblocks, hour0, ring_hours))
    torch.cuda.synchronize()
    for nm, t, t16, tf in zip(_abi.OUT_NAMES, outs, outs16, outsf):
        ring = t.cpu().numpy().reshape(p.rows, p.cols, ring_hours, order="F")
        np.testing.assert_array_equal(ring[:, :, slots], whole[nm][:, :, hours], err_msg=nm)
        ring16 = t16.cpu().numpy().reshape(p.rows, p.cols, ring_hours, order="F")
        assert np.array_equal(ring16[:, :, slots], packing_oracle.pack(nm, whole[nm][:, :, hours])), nm
        ringf = tf.cpu().numpy().reshape(p.rows, p.cols, ring_hours, order="F")[:, :, slots].astype(np.float64)
        want = whole[nm][:, :, hours]
        m = np.isfinite(want)
        assert np.array_equal(np.isfinite(ringf), m), nm
        # FP32 build: coarse agreement is enough to tell a misplaced hour (the FP32 tests hold its tolerances)
        assert np.all(np.abs(ringf[m] - want[m]) <= 0.5 + 0.02 * np.abs(want[m])), nm


def test_physical_ranges_wrapper_scenario():
    """Known-range bounds in the spirit of the reference's own test (tests/testthat/
    test-microclimatemodel_wrapper.R:40-48 parameters, 82-90 bounds), re-expressed on the grid kernels:
    a uniform short canopy (h = 0.5, pai = 2, x = 1, clump = 0.1) at reqhgt 0.05."""
    p = synth.make_problem(8, 8, 48, reqhgt=0.05, mode=1, zref=2.0, start_doy=79)
    nc = p.ncells
    hgt = np.full(nc, 0.5)
    pai = np.full(nc, 2.0)
    paia, leafden = synth.foliage_density(0.05, hgt, pai)
    for k, v in dict(hgt=hgt, pai=pai, x=np.ones(nc), clump=np.full(nc, 0.1), leafr=np.full(nc, 0.4),
                     leaft=np.full(nc, 0.2), leafd=np.full(nc, 0.05), gsmax=np.full(nc, 0.13), gref=np.full(nc, 0.15),
                     paia=paia, leafden=leafden, slope=np.zeros(nc), aspect=np.zeros(nc), svfa=np.ones(nc),
                     hor=np.zeros(nc * 24), wsa=np.ones(nc * 8)).items():
        p.arrays[k] = np.ascontiguousarray(v, dtype=np.float64)
    got = _check(p)
    tair = p.arrays["temp"][None, None, :]
    assert np.abs(got["Tz"] - tair).max() <= 8.0
    assert got["relhum"].max() <= 100.0 and got["relhum"].min() > 10.0
    ratio = got["windspeed"] / p.arrays["windspeed"][None, None, :]
    assert ratio.min() > 0.02 and ratio.max() < 0.5
    sw = p.arrays["swdown"][None, None, :]
    day = sw[0, 0] > 50
    assert (got["Rswup"][:, :, day] / sw[:, :, day]).max() < 0.3
    lw = p.arrays["lwdown"][None, None, :]
    r = got["Rlwdown"] / lw
    assert r.min() > 0.85 and r.max() < 1.4


def test_kernel_math_accuracy():
    """The kernels' own branch-free exp / log / pow / division / sqrt (csrc/mcf_math.cuh) against numpy
    (glibc, < 1 ulp) over the argument ranges the physics uses.  Budgets (csrc/mcf_math.cuh): exp 1e-13
    (32-entry table + degree-5 polynomial), 1/x and x/y 2e-12 (20-bit MUFU seed + one Newton step), log 2e-13, sqrt 1e-15, pow 5e-11."""
    rng = np.random.default_rng(11)
    n = 200_000
    x = np.concatenate([rng.uniform(-700, 700, n), rng.uniform(-2, 2, n), [-708.0, 709.0, 0.0, -1e-300]])
    np.testing.assert_allclose(api.math_eval(3, x), np.exp(x), rtol=1e-13, atol=0)
    assert api.math_eval(3, np.array([-1e4, -np.inf]))[0] < 1e-300
    assert np.isnan(api.math_eval(3, np.array([np.nan]))[0])
    x = np.concatenate([rng.uniform(-1000, 1000, n), [-1021.0, 1023.0, 0.0, 0.5]])
    np.testing.assert_allclose(api.math_eval(4, x), np.exp2(x), rtol=1e-13, atol=0)
    x = np.concatenate([np.exp(rng.uniform(-300, 300, n)), rng.uniform(0.5, 2.0, n), [1.0, 2.0, 0.5, 1e-300]])
    # table-driven log: absolute error <= 3e-16 where |log x| < 1 (every call site feeds an exponential or a sum)
    np.testing.assert_allclose(api.math_eval(5, x), np.log(x), rtol=2e-13, atol=3e-16)
    x = np.exp(rng.uniform(-200, 200, n)) * rng.choice([-1.0, 1.0], n)
    np.testing.assert_allclose(api.math_eval(0, x), 1.0 / x, rtol=2e-12, atol=0)
    y = np.exp(rng.uniform(-100, 100, n))
    np.testing.assert_allclose(api.math_eval(1, x, y), x / y, rtol=2e-12, atol=0)
    x = np.concatenate([np.exp(rng.uniform(-300, 300, n)), [0.0, 1.0, 4.0]])
    np.testing.assert_allclose(api.math_eval(2, x), np.sqrt(x), rtol=1e-15, atol=0)
    x = np.concatenate([rng.uniform(-50, 50, n), rng.uniform(-1, 1, n) * 1e-3, [0.0, np.pi / 4, np.pi / 2, -np.pi, 7.0]])
    np.testing.assert_allclose(api.math_eval(7, x), np.sin(x), rtol=0, atol=4e-15)  # two-term reduction: |k| * 6e-17
    np.testing.assert_allclose(api.math_eval(8, x), np.cos(x), rtol=0, atol=4e-15)
    x = np.exp(rng.uniform(-30, 10, n))
    for e in (0.2, 0.2672778, -4.5, -0.733):
        np.testing.assert_allclose(api.math_eval(6, x, e), np.power(x, e), rtol=5e-11, atol=0)


def test_full_size_band_sampled_against_oracle():
    """A bench-sized band slice (2048 x 1024 cells, device-resident, 24-hour ring) checked through
    size-independent properties: (1) determinism: two runs are bit-identical; (2) cell independence:
    300 randomly sampled cells, re-solved on the CPU as a 300 x 1 raster with the band's twi mean, match
    the big run; (3) NA cells stay NA, every other value is finite."""
    import torch

    from microclimf_b200 import bands

    rows, cols, T = 2048, 1024, 48
    p = synth.make_problem(rows, cols, T, reqhgt=0.05, mode=1, seed=99)
    s, n = bands.twi_partial_host(p.arrays["twi"], p.tfact)
    p.twi_mean = s / n
    dp = p.to_device()
    nc = p.ncells
    outs = [torch.empty(T * nc, dtype=torch.float64, device="cuda") for _ in range(10)]
    api.run_problem_dev(dp, outs)
    torch.cuda.synchronize()
    first = [o.clone() for o in outs]
    api.run_problem_dev(dp, outs)
    torch.cuda.synchronize()
    for a, b in zip(first, outs):
        assert torch.equal(a.view(torch.int64), b.view(torch.int64))
    rng = np.random.default_rng(4)
    pick = np.sort(rng.choice(nc, 300, replace=False))
    sub = synth.make_problem(300, 1, T, reqhgt=0.05, mode=1, seed=99)
    for name, arr in p.arrays.items():
        ln = p.expected_len(name)
        if ln % nc == 0 and name not in ("year", "month", "day", "hour", "winddir") and ln >= nc and \
                (name in _abi.VEG_FIELDS or name in _abi.SOIL_FIELDS):
            sub.arrays[name] = np.ascontiguousarray(arr.reshape(ln // nc, nc)[:, pick].ravel())
        else:
            sub.arrays[name] = arr
    sub.twi_mean = p.twi_mean
    sub.validate()
    kind = "oracle" if pyoracle.have_oracle() else None
    if kind is None:
        pytest.skip("the C restatement (honours has_twi_mean) is not built")
    want = pyoracle.runmicro(sub, kind=kind)
    got = {}
    na = np.isnan(p.arrays["hgt"])
    for nm, o in zip(_abi.OUT_NAMES, outs):
        full = o.view(T, nc)
        got[nm] = full[:, torch.from_numpy(pick).cuda()].cpu().numpy().T.reshape(300, 1, T)
        v = full.cpu().numpy()
        assert np.isnan(v[:, na]).all() and np.isfinite(v[:, ~na]).all(), nm
    ok, rws = parity.compare(got, want)
    assert ok, "\n" + parity.fmt(rws)


def test_host_path_variants_agree(monkeypatch):
    """mcf_runmicro's three copy-back routes give identical results: pageable destination (staged through
    pinned slots by copy threads), pinned destination (direct), and the time-streaming path used when the
    outputs exceed device memory (forced here with the MCF_FORCE_STREAM_BLOCKS test hook), including a
    layered problem whose day-blocks leave gaps."""
    import torch

    for mode, kw in ((1, {}), (3, dict(nlyr=3))):
        p = synth.make_problem(61, 47, 24 * 11, reqhgt=0.05, mode=mode, **kw)
        if mode == 3:
            p.lyr_st = np.array([24, 96, 192], dtype=np.int32)
            p.lyr_ed = np.array([71, 167, 263], dtype=np.int32)
        base = api.run_problem(p)  # pageable numpy outputs
        n = p.ncells * p.tsteps
        pinned_t = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(10)]
        pinned = api.run_problem(p, out_buffers=[t.numpy() for t in pinned_t])
        for nm in base:
            np.testing.assert_array_equal(base[nm].view(np.uint64), pinned[nm].view(np.uint64))
        for nb in ("1", "3"):
            monkeypatch.setenv("MCF_FORCE_STREAM_BLOCKS", nb)
            streamed = api.run_problem(p)
            monkeypatch.delenv("MCF_FORCE_STREAM_BLOCKS")
            for nm in base:
                np.testing.assert_array_equal(base[nm].view(np.uint64), streamed[nm].view(np.uint64))
