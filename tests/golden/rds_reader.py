"""Minimal reader for R's XDR serialisation (version 2/3) as written by saveRDS(compress="xz").

Test infrastructure only: used by make_golden.py to turn the reference's golden ``.rds``
fixtures (inst/unitTests/*.rds) into plain ``.npz`` files that can travel to the GPU box.
Supports exactly the node types those fixtures contain: lists, pairlists (attributes),
real / integer / logical / string vectors, symbols and back-references.
"""
import gzip
import lzma
import bz2
import struct

import numpy as np

NILVALUE_SXP, REFSXP = 254, 255
NA_INTEGER = -2147483648


class RObj:
    """A decoded R value plus its attributes (names, dim, class, ...)."""

    def __init__(self, value, attrs=None):
        self.value = value
        self.attrs = attrs or {}

    def names(self):
        nm = self.attrs.get("names")
        return list(nm.value) if nm is not None else None

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.value[self.names().index(key)]
        return self.value[key]

    def keys(self):
        return self.names()

    def as_array(self):
        a = np.asarray(self.value)
        dim = self.attrs.get("dim")
        if dim is not None:
            a = a.reshape(tuple(int(d) for d in dim.value), order="F")
        return a


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
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if t == NILVALUE_SXP:
            return None
        if t == REFSXP:
            return self.refs[(flags >> 8) - 1]
        if t == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 9:  # CHARSXP
            n = self.i32()
            return None if n == -1 else self.raw(n).decode("utf-8", "replace")
        if t in (2, 6):  # LISTSXP / LANGSXP -> dict of tag -> value (attributes)
            out = {}
            while True:
                if has_attr:
                    self.item()
                tag = self.item() if has_tag else None
                car = self.item()
                out[tag if tag is not None else len(out)] = car
                save = self.o
                flags = self.i32()
                t2 = flags & 0xFF
                if t2 == NILVALUE_SXP:
                    break
                if t2 not in (2, 6):  # dotted tail
                    self.o = save
                    out[len(out)] = self.item()
                    break
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
            return out
        if t == 238:  # ALTREP_SXP: (class info, state, attributes)
            info = self.item()
            state = self.item()
            attrs = self.item()
            cls = list(info.values())[0] if isinstance(info, dict) else str(info)
            if cls in ("compact_intseq", "compact_realseq"):
                n, start, step = (float(x) for x in state.value)
                v = start + step * np.arange(int(n))
                v = v.astype(np.int32 if cls == "compact_intseq" else np.float64)
                return RObj(v, attrs)
            if cls.startswith("wrap_"):
                inner = state[0] if isinstance(state, dict) else state.value[0]
                return RObj(inner.value, attrs or inner.attrs)
            if cls == "deferred_string":  # as.character(<numeric>) kept lazy
                arg = state[0] if isinstance(state, dict) else state
                return RObj(["%g" % x for x in arg.value], attrs)
            raise ValueError("unsupported ALTREP class %r" % cls)
        if t == 10 or t == 13:  # LGLSXP / INTSXP
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
        elif t == 24:  # RAWSXP
            n = self.length()
            v = np.frombuffer(self.raw(n), dtype=np.uint8).copy()
        else:
            raise ValueError("unsupported SEXP type %d at offset %d" % (t, self.o))
        attrs = self.item() if has_attr else None
        return RObj(v, attrs)


def read_rds(path):
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    elif raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    elif raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    if raw[:2] != b"X\n":
        raise ValueError("not an XDR serialisation")
    r = _Reader(raw)
    r.o = 2
    version = r.i32()
    r.i32()  # writer version
    r.i32()  # min reader version
    if version == 3:
        n = r.i32()
        r.raw(n)  # native encoding
    return r.item()
