"""Minimal reader of R's serialisation format (saveRDS, XDR, version 2/3) -- enough for the data frames
and lists of /root/reference/tests/testthat/fixtures/simu_data_reference.rds.  stdlib only."""
import gzip
import struct


class _R:
    def __init__(self, buf):
        self.b, self.p, self.refs = buf, 0, []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def f64(self, n):
        v = struct.unpack_from(">%dd" % n, self.b, self.p)
        self.p += 8 * n
        return list(v)

    def ints(self, n):
        v = struct.unpack_from(">%di" % n, self.b, self.p)
        self.p += 4 * n
        return list(v)

    def length(self):
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + lo
        return n

    def item(self):
        flags = self.i32()
        t, has_attr, has_tag = flags & 0xFF, bool(flags & (1 << 9)), bool(flags & (1 << 10))
        if t == 254:                       # NILVALUE
            return None
        if t == 255:                       # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 1:                         # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 2:                         # LISTSXP (pairlist): returns dict tag -> value
            out = {}
            while True:
                attr = self.item() if has_attr else None  # noqa: F841
                tag = self.item() if has_tag else None
                out[tag if tag is not None else len(out)] = self.item()
                flags = self.i32()
                t, has_attr, has_tag = flags & 0xFF, bool(flags & (1 << 9)), bool(flags & (1 << 10))
                if t == 254:
                    return out
                if t != 2:
                    raise ValueError("unexpected pairlist tail type %d" % t)
        if t == 9:                         # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.p:self.p + n].decode("utf-8", "replace")
            self.p += n
            return s
        if t in (10, 13):                  # LGLSXP, INTSXP
            v = self.ints(self.length())
            v = [None if x == -2147483648 else x for x in v]
        elif t == 14:                      # REALSXP
            v = self.f64(self.length())
        elif t == 16:                      # STRSXP
            v = [self.item() for _ in range(self.length())]
        elif t == 19:                      # VECSXP
            v = [self.item() for _ in range(self.length())]
        else:
            raise ValueError("unsupported SEXP type %d at byte %d" % (t, self.p))
        attrs = self.item() if has_attr else None
        if attrs:
            if t == 19 and "names" in attrs:
                d = {"__names__": attrs["names"], "__class__": attrs.get("class")}
                for k, x in zip(attrs["names"], v):
                    d[k] = x
                return d
            if t == 13 and "levels" in attrs:      # factor
                return [attrs["levels"][x - 1] if x is not None else None for x in v]
        return v


def read_rds(path):
    raw = gzip.open(path, "rb").read()
    if raw[:2] != b"X\n":
        raise ValueError("not an XDR RDS file")
    r = _R(raw)
    r.p = 2
    version = r.i32()
    r.i32(); r.i32()
    if version == 3:
        n = r.i32()
        r.p += n
    return r.item()
