"""World-size-2 gloo test of the multi-GPU sharding logic on CPU: each rank takes its
column band, the (sum, count) of log(twi)/tfact is all-reduced, the band is solved (here by the CPU
checker, the test's stand-in for the per-rank GPU solve) and the gathered bands must equal the
whole-raster solve.  This is the exact host logic bench.py and a runmicro_big driver run per GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from microclimf_b200 import bands, synth
    from oracle import pyoracle as po

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = synth.make_problem(9, 12, 48, reqhgt=0.05, mode=2)
        kind = "oracle" if po.have_oracle() else "ref"
        if kind == "ref":
            # The reference has no twi-mean argument (it always subtracts the mean of the raster it is given,
            # src/microclimfCpp.cpp:993-1004), so for it to stand in for the per-rank solve the two bands are
            # given the same twi pattern: band mean == whole-raster mean.  The C restatement honours
            # has_twi_mean like the product and needs no such trick.
            twi = p.arrays["twi"].reshape(p.rows, p.cols, order="F").copy()
            twi[:, 6:] = twi[:, :6]
            p.arrays["twi"] = np.ascontiguousarray(twi.ravel(order="F"))
        b = bands.shard(p, rank, world)
        whole_s, whole_n = bands.twi_partial_host(p.arrays["twi"], p.tfact)
        assert abs(b.twi_mean - whole_s / whole_n) < 1e-14
        c0, c1 = bands.band_ranges(p.cols, world)[rank]
        assert b.cols == c1 - c0
        out = po.runmicro(b, kind=kind)
        parts = bands.gather_bands({k: v for k, v in out.items()}, p.rows, p.cols, world)
        if rank == 0:
            glued = {k: np.concatenate([pt[k] for pt in parts], axis=1) for k in parts[0]}
            whole = po.runmicro(p, kind=kind)
            err = max(float(np.nanmax(np.abs(glued[k] - whole[k]))) if np.isfinite(whole[k]).any() else 0.0
                      for k in whole)
            nanok = all(np.array_equal(np.isnan(glued[k]), np.isnan(whole[k])) for k in whole)
            q.put((err, nanok))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not (pyoracle.have_ref() or pyoracle.have_oracle()), reason="no CPU checker built")
def test_two_rank_band_sharding_matches_whole_raster():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    err, nanok = q.get(timeout=10)
    assert nanok
    assert err < 1e-9


def _worker_plane(rank, world, port, q):
    """Data plane of a multi-GPU runmicro_big on gloo: rank 0 holds the whole problem, bands.scatter_problem hands every
    rank its band (replicated series broadcast, per-cell arrays point-to-point), the band is solved by the CPU checker
    (stand-in for the per-rank GPU solve) and bands.gather_rasters reassembles the whole raster on rank 0."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from microclimf_b200 import bands, bigrun, synth
    from oracle import pyoracle as po

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = []
        for mode in (3, 2):
            root = synth.make_problem(7, 11, 72, reqhgt=0.05, mode=mode, nlyr=3) if rank == 0 else None
            band, (c0, c1), (R, C) = bands.scatter_problem(root, 0)
            assert (R, C) == (7, 11) and band.cols == c1 - c0 and band.rows == 7
            assert band.twi_mean is None
            band.twi_mean = bigrun._twi_mean(band)
            hb = band._clone_meta()
            hb.arrays = {n: (a.numpy() if hasattr(a, "numpy") else a) for n, a in band.arrays.items()}
            hb.validate()
            out = po.runmicro(hb, kind="oracle")
            import torch

            T = hb.tsteps
            flat = torch.from_numpy(np.ascontiguousarray(out["Tz"].transpose(2, 1, 0)).reshape(-1))  # [T][band cols][rows]
            full = bands.gather_rasters(flat, T, R, C, 0)
            if rank == 0:
                whole = root.replace()
                s, n = bands.twi_partial_host(root.arrays["twi"], root.tfact)
                assert abs(band.twi_mean - s / n) < 1e-14
                whole.twi_mean = band.twi_mean  # (the all-reduced sum may differ from numpy's in the last bit)
                want = po.runmicro(whole, kind="oracle")["Tz"]
                got = full.reshape(T, C, R).transpose(2, 1, 0)
                res.append(bool(np.array_equal(got, want, equal_nan=True)))
            else:
                assert full is None
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not pyoracle.have_oracle(), reason="the C restatement (honours has_twi_mean) is not built")
def test_two_rank_scatter_solve_gather_is_bit_identical_to_whole_raster():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_plane, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=10) == [True, True]


def test_window_planner_and_day_blocks():
    from microclimf_b200 import bigrun, synth

    p = synth.make_problem(3, 3, 24 * 10 + 5, mode=1)
    blocks = bigrun.day_blocks(p)
    assert blocks == [(24 * d, 0) for d in range(10)]
    assert bigrun.plan_windows(blocks, 4) == [(0, 4, 0), (4, 4, 96), (8, 2, 192)]
    q = synth.make_problem(3, 3, 240, mode=3, nlyr=3)
    q.lyr_st, q.lyr_ed = [0, 96, 192], [47, 167, 239]  # a gap after day 2: windows never span it
    assert bigrun.plan_windows(bigrun.day_blocks(q), 4) == [(0, 2, 0), (2, 3, 96), (5, 2, 192)]
