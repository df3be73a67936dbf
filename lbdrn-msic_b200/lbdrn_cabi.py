"""ctypes binding of liblbdrn_b200.so (C ABI declared in include/lbdrn.h).

This is the only place host code touches the native library.  There is no CPU fallback: if the library cannot be
loaded, or a compute entry point is called without a CUDA device, an exception is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LBDRN_LIB", os.path.join(HERE, "liblbdrn_b200.so"))

OK, E_INVALID, E_UNSUPPORTED, E_CUDA, E_NOMEM = 0, -1, -2, -3, -4
USE_COORDINATES, EMBEDDING, USE_COLORS, RELATIVE, ACT_RELU = 1, 2, 4, 8, 16
U8, U16 = 0, 1
PATH_AUTO, PATH_PRECISE, PATH_TENSOR, PATH_TENSOR_FASTSIN, PATH_TENSOR_FASTSIN2 = 0, 1, 2, 3, 4

# every symbol include/lbdrn.h declares (tests check the library exports exactly these)
SYMBOLS = ["lbdrn_version", "lbdrn_last_error", "lbdrn_dim_in", "lbdrn_param_count", "lbdrn_has_tensor_path",
           "lbdrn_launch_count", "lbdrn_selftest_tc_gemm", "lbdrn_selftest_tc_gemm2", "lbdrn_selftest_tc_gemm3", "lbdrn_split", "lbdrn_sse_u16", "lbdrn_max_shifted", "lbdrn_randperm", "lbdrn_host_randperm", "lbdrn_host_randperm32", "lbdrn_host_randperm32_progress", "lbdrn_decode", "lbdrn_predict",
           "lbdrn_eval_sse", "lbdrn_train_create", "lbdrn_train_destroy", "lbdrn_train_set_params",
           "lbdrn_train_get_params", "lbdrn_train_steps", "lbdrn_train_grad", "lbdrn_train_apply",
           "lbdrn_fpz_bound", "lbdrn_fpz_compress", "lbdrn_fpz_header", "lbdrn_fpz_decompress"]


class LbdrnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liblbdrn_b200 error {code}: {msg}")
        self.code = code


class LbdrnDesc(C.Structure):
    _fields_ = [("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("K", C.c_int32), ("D", C.c_int32),
                ("bc", C.c_int32), ("nl", C.c_int32), ("flags", C.c_uint32), ("w0", C.c_float),
                ("n_freq", C.c_int32), ("msb_max", C.c_uint32), ("msb_dtype", C.c_int32),
                ("row0", C.c_int32), ("row1", C.c_int32), ("buf_row0", C.c_int32), ("buf_rows", C.c_int32),
                ("path", C.c_int32), ("reserved", C.c_int32 * 3), ("msb_max_dev", C.c_void_p)]


class LbdrnTrainCfg(C.Structure):
    _fields_ = [("batch_size", C.c_int32), ("world_size", C.c_int32), ("rank", C.c_int32),
                ("reserved0", C.c_int32), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("reserved", C.c_int32 * 4)]


_lib = None


def load(build_if_missing=True):
    """Load (building in-tree first if the .so is absent and nvcc is available) and prototype the library."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and build_if_missing:
        import build_ext
        build_ext.build()
    if not os.path.exists(LIB_PATH):
        raise LbdrnError(E_CUDA, f"{LIB_PATH} not found: build it with `python lbdrn-msic_b200/build_ext.py` "
                                 "(no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    P, vp, i32, i64, f64 = C.POINTER, C.c_void_p, C.c_int32, C.c_int64, C.c_double
    D = P(LbdrnDesc)
    proto = {
        "lbdrn_version": (i32, []),
        "lbdrn_last_error": (C.c_char_p, []),
        "lbdrn_dim_in": (i32, [D]),
        "lbdrn_param_count": (i64, [D]),
        "lbdrn_has_tensor_path": (i32, [D]),
        "lbdrn_launch_count": (i64, []),
        "lbdrn_selftest_tc_gemm": (i32, [vp, vp, vp, i32, vp]),
        "lbdrn_selftest_tc_gemm2": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
        "lbdrn_selftest_tc_gemm3": (i32, [vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
        "lbdrn_sse_u16": (i32, [vp, vp, i64, vp, vp]),
        "lbdrn_split": (i32, [vp, i64, i32, i32, vp, vp, vp]),
        "lbdrn_max_shifted": (i32, [vp, i64, i32, vp, vp]),
        "lbdrn_randperm": (i32, [i64, C.c_uint64, vp, vp]),
        "lbdrn_host_randperm": (i32, [i64, C.c_uint64, vp]),
        "lbdrn_host_randperm32": (i32, [i64, C.c_uint64, vp]),
        "lbdrn_host_randperm32_progress": (i32, [i64, C.c_uint64, vp, vp]),
        "lbdrn_decode": (i32, [D, vp, vp, vp, vp, vp]),
        "lbdrn_predict": (i32, [D, vp, vp, vp, vp, vp]),
        "lbdrn_eval_sse": (i32, [D, vp, vp, vp, vp, vp, vp]),
        "lbdrn_train_create": (i32, [D, P(LbdrnTrainCfg), P(vp)]),
        "lbdrn_train_destroy": (i32, [vp]),
        "lbdrn_train_set_params": (i32, [vp, vp, vp]),
        "lbdrn_train_get_params": (i32, [vp, vp, vp]),
        "lbdrn_train_steps": (i32, [vp, vp, vp, vp, vp, i64, i32, i64, f64, vp, vp]),
        "lbdrn_train_grad": (i32, [vp, vp, vp, vp, vp, i32, i32, vp, vp]),
        "lbdrn_train_apply": (i32, [vp, vp, i64, f64, vp]),
        "lbdrn_fpz_bound": (i64, [i64]),
        "lbdrn_fpz_compress": (i32, [vp, i64, i32, vp, i64, P(i64)]),
        "lbdrn_fpz_header": (i32, [vp, i64, P(i64), P(i32)]),
        "lbdrn_fpz_decompress": (i32, [vp, i64, vp, i64]),
    }
    for name, (res, args) in proto.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise LbdrnError(rc, load().lbdrn_last_error().decode("utf-8", "replace"))
    return rc


def flag_bits(use_coordinates, embedding, use_colors, relative, relu=False):
    return ((USE_COORDINATES if use_coordinates else 0) | (EMBEDDING if embedding else 0) |
            (USE_COLORS if use_colors else 0) | (RELATIVE if relative else 0) | (ACT_RELU if relu else 0))


def make_desc(C_, H, W, K, D, bc, nl, flags, msb_max, msb_u16, row0=0, row1=None, buf_row0=0, buf_rows=None,
              w0=30.0, n_freq=12, path=PATH_AUTO, msb_max_dev=None):
    """msb_max_dev: optional int32/uint32 CUDA tensor holding the exact MSB.max(); `msb_max` is then an upper bound."""
    d = LbdrnDesc()
    d.C, d.H, d.W, d.K, d.D, d.bc, d.nl = C_, H, W, K, D, bc, nl
    d.flags, d.w0, d.n_freq = flags, w0, n_freq
    d.msb_max, d.msb_dtype = int(msb_max), (U16 if msb_u16 else U8)
    d.row0, d.row1 = row0, (H if row1 is None else row1)
    d.buf_row0, d.buf_rows = buf_row0, (H if buf_rows is None else buf_rows)
    d.path = path
    d.msb_max_dev = None if msb_max_dev is None else msb_max_dev.data_ptr()
    return d


def stream_ptr():
    """Current torch CUDA stream as a void* for the `stream` argument."""
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device (or NULL) pointer of a torch tensor."""
    return C.c_void_p(0 if t is None else t.data_ptr())
