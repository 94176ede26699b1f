"""Randomised corner-input parity sweep (dev aid; GPU box): problems whose per-cell parameters and forcing are drawn
from the full ranges `checkinputs` admits (R/dataprep.R:206-397), with special values mixed in, run through the CUDA
path and the CPU checker.  Prints every output whose error exceeds the 1e-6 tolerance, with the offending cell."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import parity
from microclimf_b200 import api, synth
from oracle import pyoracle

KIND = "ref" if (pyoracle.have_ref() and os.environ.get("FUZZ_REF")) else "oracle"


def pick(rng, n, lo, hi, specials=(), p_special=0.15, log=False):
    v = np.exp(rng.uniform(np.log(lo), np.log(hi), n)) if log else rng.uniform(lo, hi, n)
    if specials:
        m = rng.random(n) < p_special
        v[m] = rng.choice(np.asarray(specials, dtype=float), m.sum())
    return v


def mutate(p, rng):
    nl = p.nlyr if p.layered else 1
    nc = p.ncells
    A = p.arrays
    zref = p.zref
    hgt = pick(rng, nc * nl, 0.005, zref * 0.98, specials=(0.0, 0.01, 0.05, p.reqhgt if p.reqhgt > 0 else 0.3, zref * 0.999), log=True)
    pai = pick(rng, nc * nl, 0.002, 15.0, specials=(0.0, 0.001, 1.0), log=True)
    bare = (hgt == 0) | (pai == 0)
    hgt[bare] = 0.0
    pai[bare] = 0.0
    na = rng.random(nc) < 0.03
    hgt.reshape(nl, nc)[:, na] = np.nan
    A["hgt"], A["pai"] = hgt, pai
    A["x"] = np.tile(pick(rng, nc, 0.01, 10.0, specials=(1.0, 0.0, 0.5), log=True), nl)
    A["gsmax"] = np.tile(pick(rng, nc, 0.001, 2.0, specials=(0.0, 999.99, 1000.0), p_special=0.05), nl)
    lr = pick(rng, nc, 1e-4, 0.95, specials=(0.0, 0.5))
    lt = lr * rng.uniform(0, 1, nc) * np.minimum(1.0, (1 - lr) / np.maximum(lr, 1e-9))
    nanr = rng.random(nc) < 0.04  # NA reflectance (bare cells of real rasters; also under canopy here)
    lr[nanr] = np.nan
    A["leafr"], A["leaft"] = np.tile(lr, nl), np.tile(np.minimum(lt, 1 - lr), nl)
    A["clump"] = pick(rng, nc * nl, 1e-60, 0.99, specials=(0.0, 0.95, 1e-300, 1.0), log=True)
    A["leafd"] = np.tile(pick(rng, nc, 1e-3, 5.0, log=True), nl)
    pa, ld = [], []
    for l in range(nl):
        h_, p_ = hgt.reshape(nl, nc)[l], pai.reshape(nl, nc)[l]
        a_, d_ = synth.foliage_density(max(p.reqhgt, 0.0), h_, p_)
        above = ~(max(p.reqhgt, 0.0) < h_)
        pa.append(np.where(above | (h_ == 0), 0.0, a_))
        ld.append(np.where(above | (h_ == 0), 0.0, d_))
    A["paia"], A["leafden"] = np.concatenate(pa), np.concatenate(ld)
    A["gref"] = pick(rng, nc, 0.01, 0.95, specials=(1.0, 0.999, 0.0), p_special=0.06)
    A["slope"] = pick(rng, nc, 0.0, 89.0, specials=(0.0, 90.0))
    A["aspect"] = pick(rng, nc, 0.0, 360.0, specials=(0.0, 180.0, 360.0))
    A["twi"] = pick(rng, nc, 0.05, 1e4, specials=(1.0,), log=True)
    A["svfa"] = pick(rng, nc, 0.0, 1.0, specials=(0.0, 1.0))
    A["wsa"] = pick(rng, nc * 8, 0.0, 1.0, specials=(0.0, np.nan, 1.0), p_special=0.1)
    A["hor"] = pick(rng, nc * 24, 0.0, 5.0, specials=(0.0,), p_special=0.3)
    # forcing extremes (modes 1/3: per hour; modes 2/4: per cell-hour)
    n = A["temp"].size
    T = p.tsteps
    scale = lambda a, lo, hi: lo + (hi - lo) * rng.random(a.size)  # noqa: E731
    if rng.random() < 0.5:
        tc = scale(A["temp"], -50, 65)
        es = np.where(tc > 0, 0.61078 * np.exp(17.27 * tc / (tc + 237.3)), 0.61078 * np.exp(21.875 * tc / (tc + 265.5)))
        rh = pick(rng, n, 1.0, 100.0, specials=(100.0, 0.5))
        A["temp"], A["es"], A["ea"] = tc, es, es * rh / 100
        A["tdew"] = synth._dewpoint_r(A["ea"], tc)
        A["pres"] = scale(A["pres"], 87, 108)
        sw = np.where(A["swdown"] > 0, pick(rng, n, 0.0, 1350.0, specials=(1e-3, 1350.0)), 0.0)
        A["swdown"] = sw
        A["difrad"] = sw * pick(rng, n, 0.0, 1.0, specials=(0.0, 1.0))
        A["lwdown"] = scale(A["lwdown"], 50, 600)
        A["windspeed"] = pick(rng, n, 0.01, 100.0, specials=(0.0,), p_special=0.02, log=True)
        A["p_umu"] = pick(rng, n, 0.05, 3.0)
        A["p_G"] = scale(A["p_G"], -300, 300)
        A["p_soilm"] = pick(rng, n, 0.01, 0.6, specials=(0.091, 0.419, 0.0))
        A["p_dtrp"] = pick(rng, n, 0.01, 60.0, specials=(0.0,), p_special=0.02)
        A["p_kp"] = pick(rng, n, 0.1, 3.0)
        A["p_muGp"] = pick(rng, n, 0.01, 0.3)
    p.validate()


def main():
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    nprob = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    bad_total = 0
    for it in range(nprob):
        rng = np.random.default_rng(seed0 + it)
        mode = int(rng.choice([1, 2, 3, 4]))
        reqhgt = float(rng.choice([0.05, 0.0, -0.1, 1.0, 5.0, 0.3, 20.0]))
        zref = float(rng.choice([2.0, 10.0, 30.0, 60.0]))
        if reqhgt >= zref:
            reqhgt = 0.05
        ndays = int(rng.integers(1, 4))
        p = synth.make_problem(int(rng.integers(5, 40)), int(rng.integers(5, 40)), 24 * ndays, reqhgt=reqhgt, mode=mode,
                               seed=seed0 + it, nlyr=int(rng.integers(1, ndays + 1)), zref=zref, lat=float(rng.uniform(-80, 80)),
                               lon=float(rng.uniform(-180, 180)), complete=bool(rng.random() < 0.5), start_doy=int(rng.integers(0, 360)))
        mutate(p, rng)
        want = pyoracle.runmicro(p, kind=KIND)
        got = api.run_problem(p)
        ok, rows = parity.compare(got, want)
        if not ok:
            bad_total += 1
            print(f"--- seed {seed0 + it} mode {mode} reqhgt {reqhgt} zref {zref} {p.rows}x{p.cols}x{p.tsteps}")
            print(parity.fmt(rows))
            for name, w in want.items():
                g = got[name]
                with np.errstate(invalid="ignore"):
                    d = np.abs(g - w)
                d[np.isnan(d)] = 0
                mism = np.isnan(g) != np.isnan(w)
                badm = (d > 1e-6 + 1e-6 * np.abs(np.nan_to_num(w))) | mism
                if badm.any():
                    i, j, k = np.argwhere(badm)[0]
                    c = i + p.rows * j
                    nl = p.nlyr if p.layered else 1
                    cell = {n: float(p.arrays[n].reshape(nl, -1)[0, c]) for n in ("hgt", "pai", "x", "gsmax", "leafr", "leaft", "clump", "leafd", "paia")}
                    cell.update({n: float(p.arrays[n][c]) for n in ("gref", "slope", "aspect", "twi", "svfa")})
                    print(f"  {name}: {int(badm.sum())} bad; first cell ({i},{j}) hour {k}: got {g[i, j, k]} want {w[i, j, k]}  {cell}")
                    break
    print("problems with mismatches:", bad_total, "of", nprob, "checker:", KIND)


if __name__ == "__main__":
    main()
