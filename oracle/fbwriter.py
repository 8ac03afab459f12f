"""ORACLE — TEST INFRASTRUCTURE ONLY.

Minimal FlatBuffers writer (no `flatbuffers` module in this image), just enough to emit a
TFLite v3 (`TFL3`) model.  Objects are written back-to-front so every uoffset points forward.
Offsets inside the builder are "bytes from the end of the buffer".
"""
import struct


class Builder:
    def __init__(self):
        self.chunks = []  # reversed order (first pushed = last in file)
        self.size = 0
        self.minalign = 8

    def _push(self, b):
        if b:
            self.chunks.append(bytes(b))
            self.size += len(b)

    def _align(self, a, extra):
        """pad so that after `extra` more bytes the write head is `a`-aligned"""
        self.minalign = max(self.minalign, a)
        pad = (-(self.size + extra)) % a
        if pad:
            self._push(b"\0" * pad)

    def string(self, s):
        data = s.encode() + b"\0"
        self._align(4, len(data) + 4)
        self._push(data)
        self._push(struct.pack("<I", len(data) - 1))
        return self.size

    def vector(self, data, count, align=4):
        """data: raw little-endian element bytes"""
        a = max(4, align)
        self._align(a, len(data))
        self._push(data)
        self._push(struct.pack("<I", count))
        return self.size

    def vector_of(self, fmt, values):
        values = list(values)
        data = struct.pack("<%d%s" % (len(values), fmt), *values) if values else b""
        return self.vector(data, len(values), struct.calcsize(fmt))

    def vector_of_offsets(self, offs):
        offs = list(offs)
        n = len(offs)
        self._align(4, 4 * n)
        v_off = self.size + 4 * n + 4  # from-end offset of the length field once written
        elems = []
        for i, t in enumerate(offs):
            e_off = v_off - 4 - 4 * i
            elems.append(struct.pack("<I", e_off - t))
        self._push(b"".join(elems))
        self._push(struct.pack("<I", n))
        assert self.size == v_off
        return self.size

    def table(self, fields):
        """fields: list of (id, kind, value); kind is a struct format char for scalars or 'o' for an
        offset (value = from-end offset of the target, 0/None = absent)."""
        fields = [(i, k, v) for (i, k, v) in fields if not (k == "o" and not v)]
        sized = []
        for i, k, v in fields:
            sz = 4 if k == "o" else struct.calcsize(k)
            sized.append((sz, i, k, v))
        sized.sort(key=lambda t: (-t[0], t[1]))
        pos = 4
        layout = {}
        maxal = 4
        for sz, i, k, v in sized:
            pos = (pos + sz - 1) // sz * sz
            layout[i] = (pos, k, v, sz)
            pos += sz
            maxal = max(maxal, sz)
        tsize = (pos + 3) // 4 * 4
        nslots = (max(layout) + 1) if layout else 0
        # vtable first (higher address than the table => negative soffset)
        vt = [4 + 2 * nslots, tsize] + [layout[i][0] if i in layout else 0 for i in range(nslots)]
        vbytes = struct.pack("<%dH" % len(vt), *vt)
        self._align(2, len(vbytes))
        self._push(vbytes)
        vt_off = self.size
        self._align(maxal, tsize)
        t_off = self.size + tsize
        body = bytearray(tsize)
        struct.pack_into("<i", body, 0, vt_off - t_off)
        for i, (p, k, v, sz) in layout.items():
            if k == "o":
                struct.pack_into("<I", body, p, (t_off - p) - v)
            else:
                struct.pack_into("<" + k, body, p, v)
        self._push(bytes(body))
        assert self.size == t_off
        return t_off

    def finish(self, root, ident=b"TFL3"):
        self._align(self.minalign, 8)
        self._push(ident)
        total = self.size + 4
        self._push(struct.pack("<I", total - root))
        return b"".join(reversed(self.chunks))
