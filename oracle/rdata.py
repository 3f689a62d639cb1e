"""TEST INFRASTRUCTURE ONLY -- minimal reader for R's XDR serialization (RDX2/RDX3).

Just enough of the format to pull the S4 `dgCMatrix` out of the reference's fixture
`data/movielens100k.RData` (R/data.R:1-20; 943 x 1682, nnz = 100000, values 1..5), which is
what the reference's tests use (tests/testthat.R:9, tests/testthat/test-wrmf.R:6-7).
Used once, in this container, by tests/golden/make_golden.py to write the npz fixture that
travels with the repo; nothing in the product imports it.
"""
import bz2
import gzip
import struct

import numpy as np


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.o = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def raw(self, n):
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def length(self):
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + lo
        return n

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == 254:  # NILVALUE
            return None
        if t == 253 or t == 242:  # global env / empty env
            return "<env>"
        if t == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 9:  # CHARSXP
            n = self.i32()
            return None if n == -1 else self.raw(n).decode("utf-8", "replace")
        if t in (2, 6):  # pairlist / LANGSXP
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                nflags = struct.unpack_from(">i", self.b, self.o)[0]
                nt = nflags & 0xFF
                if nt == 254:
                    self.i32()
                    break
                if nt not in (2, 6):
                    out.append(("<cdr>", self.item()))
                    break
                flags = self.i32()
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return out
        if t in (13, 10):  # INTSXP / LGLSXP
            n = self.length()
            v = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
        elif t == 14:  # REALSXP
            n = self.length()
            v = np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64)
        elif t == 16:  # STRSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t == 19:  # VECSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t == 25:  # S4SXP
            v = {}
        else:
            raise NotImplementedError("SEXP type %d at offset %d" % (t, self.o))
        if has_attr:
            attrs = self.item()
            d = {k: val for k, val in attrs}
            if t == 25:
                v = d
            else:
                v = (v, d) if d else v
        return v


def load_rdata(path):
    raw = open(path, "rb").read()
    if raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    assert raw[:5] in (b"RDX2\n", b"RDX3\n"), raw[:8]
    r = _Reader(raw)
    r.o = 5
    assert r.raw(2) == b"X\n"
    version = r.i32()
    r.i32()
    r.i32()
    if version == 3:
        n = r.i32()
        r.raw(n)
    top = r.item()  # pairlist of (name, value)
    return {k: v for k, v in top}


def _strip(v):
    return v[0] if isinstance(v, tuple) else v


def load_movielens100k(path="/root/reference/data/movielens100k.RData"):
    """Returns (i, p, x, dim) of the dgCMatrix (0-based `i`, column pointers `p`)."""
    obj = load_rdata(path)["movielens100k"]
    i = _strip(obj["i"]).astype(np.int32)
    p = _strip(obj["p"]).astype(np.int32)
    x = _strip(obj["x"]).astype(np.float64)
    dim = _strip(obj["Dim"]).astype(np.int32)
    return i, p, x, dim


if __name__ == "__main__":
    i, p, x, dim = load_movielens100k()
    print(dim, len(i), p[:4], p[-1], x[:8], x.min(), x.max())
