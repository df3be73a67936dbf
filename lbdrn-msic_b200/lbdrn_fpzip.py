"""Stand-in for the `fpzip` Python package at the reference's two call sites -- encode.py:129
`fpzip.compress(params, precision=prec, order='C')` and decode.py:113 `fpzip.decompress(bytes, order='C')[0][0][0]` -- backed by
the C++ codec of liblbdrn_b200 (csrc/lbdrn_fpz.cpp: PCmap value map, 1-D Lorenzo predictor, adaptive range coder; N2 of
SURVEY.md 8f).  encode.py / decode.py use it only when `import fpzip` fails, so an installation that has the real library keeps
producing and reading the reference's streams.

Same call shapes as fpzip 1.2.x for the flat parameter vector: `compress` takes a float32 array (any shape; coded as one flat
C-order vector, which is what the reference passes) and returns bytes; `decompress` returns a 4-D array (1, 1, 1, n) so that
`[0][0][0]` yields the vector.  Byte compatibility with the real library's payload is unverified (see the codec's header)."""
import ctypes as C

import numpy as np

import lbdrn_cabi as cabi


def _lib():
    return cabi.load()


def compress(data, precision=0, order="C"):
    a = np.asarray(data)
    if a.dtype != np.float32:
        raise TypeError("lbdrn_fpzip codes float32 arrays (the nn sub-stream); got %s" % a.dtype)
    a = np.ascontiguousarray(a.reshape(-1, order=order))
    lib = _lib()
    cap = int(lib.lbdrn_fpz_bound(a.size))
    out = np.empty(cap, dtype=np.uint8)
    nbytes = C.c_int64(0)
    cabi.check(lib.lbdrn_fpz_compress(a.ctypes.data_as(C.c_void_p), a.size, int(precision), out.ctypes.data_as(C.c_void_p), cap,
                                      C.byref(nbytes)))
    return out[:nbytes.value].tobytes()


def decompress(blob, order="C"):
    buf = np.frombuffer(blob, dtype=np.uint8)
    lib = _lib()
    n, prec = C.c_int64(0), C.c_int32(0)
    cabi.check(lib.lbdrn_fpz_header(buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(n), C.byref(prec)))
    out = np.empty(n.value, dtype=np.float32)
    cabi.check(lib.lbdrn_fpz_decompress(buf.ctypes.data_as(C.c_void_p), buf.size, out.ctypes.data_as(C.c_void_p), n.value))
    return out.reshape((1, 1, 1, n.value))
