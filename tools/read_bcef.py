"""Read the reference package's BCEF data set (data/BCEF.rda: an xz-compressed RDX3 / XDR workspace holding one
data.frame, man/BCEF.rda.Rd:16-20) without R, and write it as data/bcef.npz (git-ignored; travels to the GPU box with
the snapshot).  Only the handful of SEXP types a data.frame needs are decoded.

    python tools/read_bcef.py [/root/reference/data/BCEF.rda] [data/bcef.npz]

Columns: x, y (km, Albers), FCH (forest canopy height), PTC (percent tree cover), holdout (0/1): 188,717 rows.
"""
import lzma
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class XdrReader:
    """R's XDR serialization (src/main/serialize.c of R): big-endian, one flags word per item."""

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
        if n == -1:  # long vector: two words
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t == 254:  # NILVALUE_SXP
            return None
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
        if t in (2, 6):  # LISTSXP / LANGSXP: attr, tag, car, then the cdr as the next item
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                out.append((tag, self.item()))
                flags = self.i32()
                t2 = flags & 0xFF
                if t2 == 254:
                    return out
                if t2 not in (2, 6):
                    raise ValueError(f"pairlist continues with type {t2}")
                has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t == 238:  # ALTREP_SXP: info pairlist, state, attributes -- compact sequences are all a data.frame uses
            info = self.item()
            state = self.item()
            attr = self.item()
            cls = info[0][1] if info else ""
            if cls in ("compact_intseq", "compact_realseq"):
                n, start, step = (float(v) for v in state)
                return {"data": start + step * np.arange(int(n)), "attr": attr}
            return {"data": state, "attr": attr}
        if t == 14:  # REALSXP
            n = self.length()
            v = np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64)
        elif t in (13, 10):  # INTSXP / LGLSXP
            n = self.length()
            v = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
        elif t == 16:  # STRSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t == 19:  # VECSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        else:
            raise ValueError(f"SEXP type {t} not handled at offset {self.o}")
        attr = self.item() if has_attr else None
        return {"data": v, "attr": attr} if attr is not None else v


def read_rda(path):
    buf = lzma.open(path).read()
    if buf[:5] != b"RDX3\n" or buf[5:7] != b"X\n":
        raise ValueError("not an RDX3 XDR workspace")
    r = XdrReader(buf)
    r.o = 7
    version, _writer, _min_reader = r.i32(), r.i32(), r.i32()
    if version == 3:
        r.raw(r.i32())  # native encoding name
    return r.item()  # pairlist of (name, value)


def bcef_columns(path):
    top = read_rda(path)
    name, df = top[0]
    cols = df["data"]
    names = dict(df["attr"])["names"]
    names = names["data"] if isinstance(names, dict) else names
    out = {}
    for nm, col in zip(names, cols):
        out[nm] = np.asarray(col["data"] if isinstance(col, dict) else col, dtype=np.float64)
    return name, out


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/BCEF.rda"
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "data", "bcef.npz")
    name, cols = bcef_columns(src)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    np.savez_compressed(dst, **cols)
    n = len(next(iter(cols.values())))
    print(f"{name}: {n} rows, columns {list(cols)} -> {dst}")
    for k, v in cols.items():
        print(f"  {k:8s} min {v.min():.4f} max {v.max():.4f} mean {v.mean():.4f}")
    if "holdout" in cols:
        print(f"  holdout = 0: {(cols['holdout'] == 0).sum()}, holdout = 1: {(cols['holdout'] == 1).sum()}")


if __name__ == "__main__":
    main()
