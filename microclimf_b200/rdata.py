"""Reader for R's serialised workspace files (`.rda` / `.RData`, format RDX2/RDX3, XDR encoding).

The reference ships its example inputs (`data/*.rda`: `climdata`, `dtmcaerth`, `vegp`, `soilc`,
`habitats`, `soilparameters`, `soilparamsp`; documented in /root/reference/R/data.R:1-145) as bz2
compressed RDX workspaces whose rasters are terra `PackedSpatRaster` S4 objects (a `definition` string
and a `values` matrix).  There is no R in this build's environment, so the host side reads them
directly: this module decodes the serialisation stream described in "R Internals", section 1.8
(`serialize.c`: a pairlist of named values, each item a 32-bit flag word followed by its payload, all
big-endian) into plain Python objects:

    numeric / integer / logical vectors -> numpy arrays (NA_integer_ / NA logical -> masked by caller)
    character vectors                   -> list[str | None]
    lists (VECSXP)                      -> list, or dict-like `RList` when it has names
    data.frame                          -> RList with `.attrs['class'] == ['data.frame']`
    S4 objects                          -> RS4 (`.cls`, `.slots`)
    factors / POSIXct                   -> array with `.attrs`

Only what the bundled files use is implemented; anything else raises `ValueError` naming the SEXP
type, rather than guessing.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct
from typing import Any, Dict, List, Optional

import numpy as np

# SEXP type codes (Rinternals.h) and serialisation pseudo-types (serialize.c)
_NILSXP, _SYMSXP, _LISTSXP, _CLOSXP, _ENVSXP, _LANGSXP = 0, 1, 2, 3, 4, 6
_CHARSXP, _LGLSXP, _INTSXP, _REALSXP, _CPLXSXP, _STRSXP, _VECSXP, _EXPRSXP = 9, 10, 13, 14, 15, 16, 19, 20
_RAWSXP, _S4SXP = 24, 25
_REFSXP, _NILVALUE, _GLOBALENV, _EMPTYENV, _BASEENV, _ALTREP, _ATTRLIST, _ATTRLANG = 255, 254, 253, 242, 241, 238, 239, 240
_NAMESPACESXP, _PACKAGESXP, _MISSINGARG, _UNBOUND = 249, 250, 251, 252

NA_INTEGER = -2147483648


class RArray(np.ndarray):
    """numpy array carrying R attributes (dim, names, class, levels, tzone ...) in `.attrs`."""

    def __new__(cls, a, attrs=None):
        obj = np.asarray(a).view(cls)
        obj.attrs = dict(attrs or {})
        return obj

    def __array_finalize__(self, obj):
        self.attrs = dict(getattr(obj, "attrs", {}) or {})


class RList(list):
    """An R list; named lists can be indexed by name."""

    def __init__(self, items=(), attrs=None):
        super().__init__(items)
        self.attrs: Dict[str, Any] = dict(attrs or {})

    @property
    def names(self) -> Optional[List[str]]:
        return self.attrs.get("names")

    def __getitem__(self, k):
        if isinstance(k, str):
            n = self.names or []
            return super().__getitem__(n.index(k))
        return super().__getitem__(k)

    def get(self, k, default=None):
        n = self.names or []
        return super().__getitem__(n.index(k)) if k in n else default

    def keys(self):
        return list(self.names or [])

    def __contains__(self, k):
        if isinstance(k, str):
            return k in (self.names or [])
        return super().__contains__(k)


class RStrings(list):
    def __init__(self, items=(), attrs=None):
        super().__init__(items)
        self.attrs: Dict[str, Any] = dict(attrs or {})


class RS4:
    def __init__(self, attrs):
        cls = attrs.pop("class", None)
        self.cls = list(cls) if cls is not None else []
        self.slots = attrs

    def __repr__(self):
        return f"RS4({self.cls}, slots={list(self.slots)})"


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.p = 0
        self.refs: List[Any] = []

    def i32(self) -> int:
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def length(self) -> int:
        n = self.i32()
        if n == -1:  # long vector: two 32-bit halves
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def raw(self, n: int) -> bytes:
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def attributes(self) -> Dict[str, Any]:
        attrs: Dict[str, Any] = {}
        pl = self.item()
        for tag, val in pl or []:
            attrs[tag] = val
        for k in ("names", "class", "levels", "row.names", "tzone"):
            if k in attrs and isinstance(attrs[k], RStrings):
                attrs[k] = list(attrs[k])
        return attrs

    def item(self) -> Any:
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == _NILVALUE or t == _NILSXP:
            return None
        if t in (_GLOBALENV, _EMPTYENV, _BASEENV, _MISSINGARG, _UNBOUND):
            return None
        if t == _REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == _SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if t in (_NAMESPACESXP, _PACKAGESXP):
            self.i32()
            n = self.i32()
            v = [self.item() for _ in range(n)]
            self.refs.append(v)
            return v
        if t in (_LISTSXP, _LANGSXP, _ATTRLIST, _ATTRLANG):
            # a pairlist: returned as [(tag, value), ...]
            out = []
            while True:
                attrs = self.attributes() if has_attr else None  # noqa: F841 (pairlist attributes are unused here)
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                t = flags & 0xFF
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
                if t in (_NILVALUE, _NILSXP):
                    return out
                if t not in (_LISTSXP, _LANGSXP, _ATTRLIST, _ATTRLANG):
                    raise ValueError(f"rdata: unexpected pairlist tail type {t}")
        if t == _CHARSXP:
            n = self.i32()
            if n == -1:
                return None  # NA_character_
            raw = self.raw(n)
            enc = "latin-1" if flags & (1 << 14) else "utf-8"
            return raw.decode(enc, errors="replace")
        if t in (_LGLSXP, _INTSXP):
            n = self.length()
            a = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int32)
            self.p += 4 * n
            attrs = self.attributes() if has_attr else {}
            if t == _LGLSXP:
                attrs.setdefault("_logical", True)
            return RArray(a, attrs)
        if t == _REALSXP:
            n = self.length()
            a = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
            self.p += 8 * n
            attrs = self.attributes() if has_attr else {}
            return RArray(a, attrs)
        if t == _CPLXSXP:
            n = self.length()
            a = np.frombuffer(self.b, dtype=">c16", count=n, offset=self.p).astype(np.complex128)
            self.p += 16 * n
            attrs = self.attributes() if has_attr else {}
            return RArray(a, attrs)
        if t == _RAWSXP:
            n = self.length()
            a = np.frombuffer(self.raw(n), dtype=np.uint8)
            attrs = self.attributes() if has_attr else {}
            return RArray(a, attrs)
        if t == _STRSXP:
            n = self.length()
            v = [self.item() for _ in range(n)]
            attrs = self.attributes() if has_attr else {}
            return RStrings(v, attrs)
        if t in (_VECSXP, _EXPRSXP):
            n = self.length()
            v = [self.item() for _ in range(n)]
            attrs = self.attributes() if has_attr else {}
            return RList(v, attrs)
        if t == _S4SXP:
            attrs = self.attributes() if has_attr else {}
            return RS4(attrs)
        if t == _ALTREP:
            info = self.item()
            state = self.item()
            attr = self.item()
            return self._altrep(info, state, attr)
        raise ValueError(f"rdata: unsupported SEXP type {t} at byte {self.p}")

    @staticmethod
    def _altrep(info, state, attr):
        cls = info[0][1] if info else None
        attrs = {tag: val for tag, val in (attr or [])}
        if cls == "compact_intseq":
            n, start, step = (int(state[0]), int(state[1]), int(state[2]))
            return RArray(np.arange(start, start + n * step, step, dtype=np.int32), attrs)
        if cls == "compact_realseq":
            n, start, step = (int(state[0]), float(state[1]), float(state[2]))
            return RArray(start + step * np.arange(n, dtype=np.float64), attrs)
        if cls in ("wrap_real", "wrap_integer", "wrap_logical", "wrap_string"):
            return state[0][1] if isinstance(state, list) else state
        if cls == "deferred_string":
            src = state[0][1] if isinstance(state, list) else state
            return RStrings([None if (isinstance(x, float) and np.isnan(x)) else repr(x) for x in np.asarray(src)], attrs)
        raise ValueError(f"rdata: unsupported ALTREP class {cls!r}")


def _decompress(raw: bytes) -> bytes:
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


def read_rda(path: str) -> Dict[str, Any]:
    """Load an `.rda` workspace: {object name: value}."""
    with open(path, "rb") as f:
        buf = _decompress(f.read())
    if buf[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError(f"{path}: not an RDX2/RDX3 workspace (magic {buf[:5]!r})")
    r = _Reader(buf)
    r.p = 5
    fmt = r.raw(2)
    if fmt != b"X\n":
        raise ValueError(f"{path}: only XDR serialisation is supported (format {fmt!r})")
    version = r.i32()
    r.i32()  # writer R version
    r.i32()  # minimal reader R version
    if version == 3:
        n = r.i32()
        r.raw(n)  # native encoding name
    top = r.item()
    return {tag: val for tag, val in (top or [])}


# ---------------------------------------------------------------------------------------------
# helpers to turn decoded objects into what the host layer works with
# ---------------------------------------------------------------------------------------------
def as_matrix(a) -> np.ndarray:
    """R matrix / array (column-major with a `dim` attribute) -> numpy array of that shape."""
    dim = getattr(a, "attrs", {}).get("dim")
    v = np.asarray(a)
    if dim is None:
        return v.copy()
    return v.reshape(tuple(int(d) for d in np.asarray(dim)), order="F").copy()


def dataframe_columns(df: RList) -> Dict[str, Any]:
    out = {}
    for name, col in zip(df.names or [], df):
        if isinstance(col, RArray):
            arr = np.asarray(col)
            if col.attrs.get("levels") is not None:  # factor -> strings
                lev = col.attrs["levels"]
                out[name] = [None if k == NA_INTEGER else lev[k - 1] for k in arr]
                continue
            if arr.dtype == np.int32:
                arr = np.where(arr == NA_INTEGER, np.nan, arr.astype(np.float64)) if (arr == NA_INTEGER).any() else arr
            out[name] = arr.copy()
        elif isinstance(col, RList) and "POSIXlt" in (col.attrs.get("class") or []):
            out[name] = posixlt_fields(col)
        else:
            out[name] = col
    return out


def posixlt_fields(x: RList) -> Dict[str, np.ndarray]:
    """POSIXlt list -> dict of its integer fields (sec, min, hour, mday, mon, year ...)."""
    return {n: np.asarray(v, dtype=np.float64) for n, v in zip(x.names or [], x) if isinstance(v, np.ndarray)}
