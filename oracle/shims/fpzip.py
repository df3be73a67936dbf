"""TEST INFRASTRUCTURE ONLY -- stand-in for PyPI `fpzip==1.2.4` (reference requirements.txt:2), absent here.

Call sites it serves: reference encode.py:129 `fpzip.compress(params, precision=prec, order='C')` and
decode.py:113 `fpzip.decompress(bytes, order='C')[0][0][0]`.

PARITY UNPINNED for this dependency: the real library is a Lorenzo-predictor + range coder whose payload
bytes cannot be reproduced without it.  What matters for the decoder's arithmetic is the VALUE map of its
lossy mode, restated here from the published algorithm (LLNL fpzip 1.x `pcmap.h`, PCmap<float,bits>):
a float32 is bit-complemented, shifted right by 32-prec, sign-folded, coded, and the inverse undoes these
steps exactly -- i.e. the decoded value is the input with its low (32-prec) bits cleared (truncation
toward zero; prec=16 gives bf16-exact values).  The container below stores those truncated words with a
generic entropy coder (zlib), so stream *sizes* are only indicative.
"""
import struct
import zlib

import numpy as np

_MAGIC = b"fpz\x05"


def truncate(params, precision):
    """Value map of fpzip's lossy float32 mode: keep the top `precision` bits of each word."""
    a = np.ascontiguousarray(params, dtype=np.float32)
    if precision in (0, 32):
        return a.copy()
    mask = np.uint32((0xFFFFFFFF << (32 - precision)) & 0xFFFFFFFF)
    return (a.view(np.uint32) & mask).view(np.float32)


def compress(data, precision=0, order="C"):
    a = np.asarray(data)
    if a.dtype != np.float32:
        raise TypeError("shim handles float32 only")
    prec = 32 if precision == 0 else int(precision)
    words = truncate(a.reshape(-1), prec).view(np.uint32) >> np.uint32(32 - prec)
    # split into byte planes so the generic coder sees the correlated high bytes together
    planes = b"".join(((words >> np.uint32(8 * k)) & np.uint32(0xFF)).astype(np.uint8).tobytes()
                      for k in range((prec + 7) // 8))
    return _MAGIC + struct.pack("<BI", prec, a.size) + zlib.compress(planes, 9)


def decompress(blob, order="C"):
    if blob[:4] != _MAGIC:
        raise ValueError("not a shim fpzip stream")
    prec, n = struct.unpack("<BI", blob[4:9])
    planes = np.frombuffer(zlib.decompress(blob[9:]), dtype=np.uint8).reshape(-1, n)
    words = np.zeros(n, dtype=np.uint32)
    for k in range(planes.shape[0]):
        words |= planes[k].astype(np.uint32) << np.uint32(8 * k)
    vals = (words << np.uint32(32 - prec)).view(np.float32)
    return vals.reshape(1, 1, 1, n)
