"""TEST INFRASTRUCTURE ONLY -- pure-Python restatement of the nn sub-stream codec (csrc/lbdrn_fpz.cpp), itself a restatement of
the published fpzip algorithm (Lindstrom & Isenburg 2006; LLNL fpzip 1.x: pcmap.h, pcencoder.inl, rcqsmodel.cpp,
rcencoder.inl), for the reference's call sites encode.py:129 / decode.py:113.

PARITY UNPINNED against the real library (PyPI fpzip 1.2.4 is absent and cannot be fetched): this file pins the C++ codec
against an independent, slow, integer-by-integer statement of the same algorithm (identical bytes), and pins the VALUE MAP --
the only part of fpzip the decoder's arithmetic depends on -- against the published PCmap<float, bits> transform.
Python integers are masked to 32 bits where the C code relies on unsigned wrap-around."""
import struct

M32 = 0xFFFFFFFF


def map_forward(x, bits):
    """PCmap<float, bits>::forward (pcmap.h): complement, keep the top `bits` bits, fold the sign: order preserving."""
    r = struct.unpack("<I", struct.pack("<f", x))[0]
    shift = 32 - bits
    r = (~r) & M32
    r >>= shift
    if shift + 1 < 32:
        r ^= ((-(r >> (bits - 1))) & M32) >> (shift + 1)
    return r


def map_inverse(r, bits):
    shift = 32 - bits
    if shift + 1 < 32:
        r ^= ((-(r >> (bits - 1))) & M32) >> (shift + 1)
    r = (~r) & M32
    r = (r << shift) & M32
    return struct.unpack("<f", struct.pack("<I", r))[0]


class QsModel:
    def __init__(self, n, bits=16, period=0x400):
        self.n, self.bits, self.target = n, bits, period
        self.cumf = [0] * (n + 1)
        self.cumf[n] = 1 << bits
        self.rescale, self.nextleft, self.incr, self.left = (n >> 4) | 2, 0, 0, 0
        f, m = divmod(self.cumf[n], n)
        self.symf = [f + 1] * m + [f] * (n - m) + [0]
        self.update()

    def update(self):
        if self.nextleft:
            self.incr += 1
            self.left, self.nextleft = self.nextleft, 0
            return
        if self.rescale != self.target:
            self.rescale = min(self.rescale * 2, self.target)
        cf = missing = self.cumf[self.n]
        for i in range(self.n - 1, -1, -1):
            t = self.symf[i]
            cf -= t
            self.cumf[i] = cf
            t = (t >> 1) | 1
            missing -= t
            self.symf[i] = t
        self.incr, self.nextleft = divmod(missing, self.rescale)
        self.left = self.rescale - self.nextleft

    def bump(self, s):
        if not self.left:
            self.update()
        self.left -= 1
        self.symf[s] += self.incr


class Encoder:
    def __init__(self):
        self.out, self.low, self.range = bytearray(), 0, M32

    def put(self, k):
        for _ in range(k):
            self.out.append(self.low >> 24)
            self.low = (self.low << 8) & M32

    def normalize(self):
        while not ((self.low ^ ((self.low + self.range) & M32)) >> 24):
            self.put(1)
            self.range = (self.range << 8) & M32
        if not (self.range >> 16):
            self.put(2)
            self.range = (-self.low) & M32

    def shift(self, s, nb):
        self.range >>= nb
        self.low = (self.low + self.range * s) & M32
        self.normalize()

    def bits(self, s, nb):
        if nb > 16:
            self.shift(s & 0xFFFF, 16)
            s >>= 16
            nb -= 16
        self.shift(s, nb)

    def symbol(self, s, m):
        l, r = m.cumf[s], m.cumf[s + 1] - m.cumf[s]
        m.bump(s)
        self.range >>= m.bits
        self.low = (self.low + self.range * l) & M32
        self.range = (self.range * r) & M32
        self.normalize()


def compress(values, precision):
    bits = 32 if precision == 0 else precision
    e = Encoder()
    for ch in (ord("f"), ord("p"), ord("z"), 0):
        e.bits(ch, 8)
    e.bits(0x0110, 16)
    e.bits(1, 8)
    e.bits(0, 1)
    e.bits(0 if bits == 32 else bits, 7)
    for v in (len(values), 1, 1, 1):
        e.bits(v, 32)
    wide = bits > 8
    symbols = 2 * bits + 1 if wide else 2 * (1 << bits) - 1
    bias = bits if wide else (1 << bits) - 1
    m = QsModel(symbols)
    pred = 0.0
    for x in values:
        a, p = map_forward(float(x), bits), map_forward(pred, bits)
        if not wide:
            e.symbol(bias + a - p, m)
        elif p < a:
            d = a - p
            k = d.bit_length() - 1
            e.symbol(bias + 1 + k, m)
            e.bits(d - (1 << k), k)
        elif p > a:
            d = p - a
            k = d.bit_length() - 1
            e.symbol(bias - 1 - k, m)
            e.bits(d - (1 << k), k)
        else:
            e.symbol(bias, m)
        pred = map_inverse(a, bits)
    e.put(4)
    return bytes(e.out)
