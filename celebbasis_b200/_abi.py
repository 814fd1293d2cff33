"""ctypes signatures of every non-GEMM entry point of include/celebbasis_b200.h."""
import ctypes as C

_p, _i, _l, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float

SIGS = {}


def declare(lib):
    for name, argtypes in SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
