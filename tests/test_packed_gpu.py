"""Packed integer sink (SURVEY.md NEXT-4): mcf_runmicro_packed stores, from inside the kernels, the integers the
reference's writetonc writes (R/dataprep.R:1064-1069, 1164-1173).  Integer work => the bar is bit-exact against
the packing of the same FP64 values; against the packing of the REFERENCE's FP64 values a value within the FP64
tolerance of a rounding boundary (x.5) may land one unit away, and only there."""
import numpy as np
import pytest

from microclimf_b200 import _abi, api, synth
from oracle import packing_oracle, pyoracle

pytestmark = pytest.mark.gpu
KIND = "ref" if pyoracle.have_ref() else "oracle"


def _check_packed(p, out_mask=None):
    fp = api.run_problem(p, out=out_mask)
    pk = api.run_problem_packed(p, out=out_mask)
    want = pyoracle.runmicro(p, out_mask=out_mask, kind=KIND)
    assert set(pk) == set(fp)
    for name, a in pk.items():
        assert a.dtype == np.int16 and a.shape == fp[name].shape
        # bit-exact against the packing of the FP64 path's own values
        assert np.array_equal(a, packing_oracle.pack(name, fp[name])), name
        # against the reference: NA mask identical, differences of at most one unit and only at x.5 boundaries
        w = packing_oracle.pack(name, want[name])
        assert np.array_equal(a == packing_oracle.NA, w == packing_oracle.NA), name
        d = np.abs(a.astype(np.int32) - w.astype(np.int32))
        assert d.max() <= 1, (name, int(d.max()))
        if d.any():
            s = want[name][d > 0] * packing_oracle.SCALE[name]
            assert np.all(np.abs(np.abs(s - np.floor(s)) - 0.5) < 1e-4), name
            assert d.mean() < 1e-4
    return pk


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("reqhgt", [0.05, 0.0, -0.1, 5.0])
def test_packed_modes_heights(mode, reqhgt):
    p = synth.make_problem(37, 29, 24 * 4, reqhgt=reqhgt, mode=mode, nlyr=2)
    _check_packed(p)


def test_packed_output_mask_and_na():
    p = synth.make_problem(16, 12, 48, reqhgt=0.05, mode=1)
    mask = [True, False, True, False, True, False, False, True, False, True]
    pk = _check_packed(p, out_mask=mask)
    assert set(pk) == {n for n, m in zip(_abi.OUT_NAMES, mask) if m}
    na_cells = np.isnan(p.arrays["hgt"].reshape(p.cols, p.rows).T)
    assert na_cells.any() and np.all(pk["Tz"][na_cells] == _abi.PACKED_NA)
    # ragged tail: hours beyond the last whole day stay NA
    p = synth.make_problem(9, 7, 24 + 11, reqhgt=0.05, mode=1)
    pk = _check_packed(p)
    assert np.all(pk["Tz"][:, :, 24:] == _abi.PACKED_NA)


def test_packed_streaming_path_and_pageable(monkeypatch):
    """The host path's three routes (resident, streamed in day chunks, pageable destinations) give the same bytes."""
    p = synth.make_problem(40, 33, 24 * 6, reqhgt=0.05, mode=1)
    base = api.run_problem_packed(p)
    monkeypatch.setenv("MCF_FORCE_STREAM_BLOCKS", "2")
    streamed = api.run_problem_packed(p)
    monkeypatch.delenv("MCF_FORCE_STREAM_BLOCKS")
    for k in base:
        assert np.array_equal(base[k], streamed[k]), k


def test_packed_device_ring():
    import torch

    p = synth.make_problem(32, 24, 24 * 4, reqhgt=0.05, mode=1)
    whole = api.run_problem_packed(p)
    d = p.to_device()
    ring = [torch.empty(p.ncells * 24, dtype=torch.int16, device="cuda") for _ in range(10)]
    for day in range(4):
        api.run_problem_packed_dev(d, ring, window=(day, 1, day * 24, 24))
        torch.cuda.synchronize()
        for nm, t in zip(_abi.OUT_NAMES, ring):
            got = t.cpu().numpy().reshape((p.rows, p.cols, 24), order="F")
            assert np.array_equal(got, whole[nm][:, :, day * 24:(day + 1) * 24]), (nm, day)
