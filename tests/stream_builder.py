"""TEST INFRASTRUCTURE: builds VALID DINT codeword streams for arbitrary integer lists.

A deliberately simple greedy parser against a reference dictionary file (longest
dictionary entry first, run codewords for zeros, u16/u32 exceptions otherwise). It is not
the reference's optimal encoder (include/dint/dint_codecs.hpp:145-255) and is never used
to make golden data — any stream it emits is checked through the oracle (and the real
reference when available) before a GPU result is compared with it. Its purpose is to
reach corners the test collection does not: 32-bit exceptions, payload words that equal
0/1, exceptions straddling 32-symbol tiles, back-to-back runs, 8-bit blocks.
"""
import numpy as np

RUN_CODES = {256: 2, 128: 3, 64: 4, 32: 5, 16: 6}  # single_dictionary.hpp:46-55
RESERVED = 7


def vbyte(v):
    """TightVariableByte::encode_single (block_codecs.hpp:26-86): terminator has bit 7 set"""
    out = bytearray()
    while v >= 128:
        out.append(v & 127)
        v >>= 7
    out.append(v | 128)
    return bytes(out)


class DictImage:
    """entries of a dictionary file as python tuples, per class"""

    def __init__(self, kind, data):
        w = np.frombuffer(data, dtype="<u4")
        self.kind = kind
        if kind == 0:  # rectangular: m_size, rows of 17
            m = int(w[0])
            rows = w[1:1 + m * 17].reshape(m, 17)
            self.classes = [[tuple(int(x) for x in r[:min(int(r[16]), 16)]) if i >= RESERVED else None
                             for i, r in enumerate(rows)]]
        else:
            if kind == 1:
                no, nt = int(w[1]), int(w[2])
                starts = [0]
                offs = w[3:3 + no]
                table = w[3 + no:3 + no + nt]
            else:
                ns, no, nt = int(w[1]), int(w[2]), int(w[3])
                starts = [int(x) for x in w[4:4 + ns]]
                offs = w[4 + ns:4 + ns + no]
                table = w[4 + ns + no:4 + ns + no + nt]
            table = np.concatenate([table, np.zeros(16, dtype="<u4")])
            self.classes = []
            for c, b in enumerate(starts):
                e = starts[c + 1] if c + 1 < len(starts) else len(offs)
                ents = []
                for i in range(b, e):
                    if i - b < RESERVED:
                        ents.append(None)
                        continue
                    so = int(offs[i])
                    off, ln = so & 0xFFFFFF, (so >> 24) + 1
                    ents.append(tuple(int(x) for x in table[off:off + ln]))
                self.classes.append(ents)
        self.lookup = []
        for ents in self.classes:
            m = {}
            for i, t in enumerate(ents):
                if t is not None and i < 65536 and t not in m:
                    m[t] = i
            self.lookup.append(m)
        self.lookup8 = []
        for ents in self.classes:
            m = {}
            for i, t in enumerate(ents[:256]):
                if t is not None and t not in m:
                    m[t] = i
            self.lookup8.append(m)


def encode_stream(di, values, cls=0, sym=2, use_runs=True, use_dict=True):
    """greedy parse of `values` into a u16 (sym=2) or u8 (sym=1) codeword stream of class `cls`"""
    lut = di.lookup[cls] if sym == 2 else di.lookup8[cls]
    out = bytearray()
    i, n = 0, len(values)
    vals = [int(v) for v in values]

    def put(idx):
        out.extend(int(idx).to_bytes(sym, "little"))

    while i < n:
        done = False
        if use_runs and vals[i] == 0:
            z = 0
            while i + z < n and vals[i + z] == 0 and z < 256:
                z += 1
            for ln in (256, 128, 64, 32, 16):
                if z >= ln:
                    put(RUN_CODES[ln])
                    i += ln
                    done = True
                    break
        if done:
            continue
        if use_dict:
            for ln in (16, 8, 4, 2, 1):
                if i + ln <= n:
                    idx = lut.get(tuple(vals[i:i + ln]))
                    if idx is not None:
                        put(idx)
                        i += ln
                        done = True
                        break
        if done:
            continue
        v = vals[i]
        if v < 65536:
            put(0)
            out.extend(v.to_bytes(2, "little"))
        else:
            put(1)
            out.extend(v.to_bytes(4, "little"))
        i += 1
    return bytes(out)


def encode_vroom_list(di, values, block_size=256, selectors=None, **kw):
    """header + payload of one vroom list (vroom_env/jobs.hpp:89-95). multi: `selectors`
    gives the selector byte of each block (default: class 0, 16-bit)."""
    values = [int(v) for v in values]
    hdr = vbyte(len(values)) + vbyte(sum(values) & 0xFFFFFFFF)
    if di.kind != 2:
        return hdr + encode_stream(di, values, **kw)
    out = bytearray(hdr)
    nb = (len(values) + block_size - 1) // block_size
    for b in range(nb):
        sel = selectors[b % len(selectors)] if selectors else 0
        blk = values[b * block_size:(b + 1) * block_size]
        out.append(sel)
        out.extend(encode_stream(di, blk, cls=sel % 6, sym=2 if sel < 6 else 1, **kw))
    return bytes(out)


# ---- binary interpolative code (include/ds2i/interpolative_coding.hpp:9-76 bit_writer) ----
class _BitWriter:
    def __init__(self):
        self.acc = 0
        self.nbits = 0

    def write(self, bits, ln):
        if ln:
            self.acc |= (bits & ((1 << ln) - 1)) << self.nbits
            self.nbits += ln

    def write_int(self, val, u):
        b = u.bit_length() - 1
        m = (1 << (b + 1)) - u
        if val < m:
            self.write(val, b)
        else:
            val += m
            self.write(val >> 1, b)
            self.write(val & 1, 1)

    def write_interpolative(self, a, lo_i, n, low, high):
        if n == 0:
            return
        h = n // 2
        val = a[lo_i + h]
        self.write_int(val - low, high - low + 1)
        self.write_interpolative(a, lo_i, h, low, val)
        self.write_interpolative(a, lo_i + h + 1, n - h - 1, val, high)

    def bytes(self):
        return self.acc.to_bytes((self.nbits + 7) // 8, "little")


def encode_interpolative_block(values, sum_known):
    """interpolative_block::encode (block_codecs.hpp:108-128); returns (bytes, sum_of_values)"""
    pre = np.cumsum(np.asarray(values, dtype=np.uint64)).tolist()
    s = int(pre[-1])
    out = bytearray()
    if not sum_known:
        out.extend(vbyte(s))
    bw = _BitWriter()
    bw.write_interpolative(pre, 0, len(values) - 1, 0, s)
    out.extend(bw.bytes())
    return bytes(out), s
