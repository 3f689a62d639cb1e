"""ctypes binding of libb200als.so (include/b200als.h).  The library is the product; this module
only loads it and declares the prototypes.  There is no fallback: if the shared object is missing
or a compute call is made without a CUDA device, an exception is raised."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200als.so")

OK, EINVAL, ECUDA, ENCCL, ENOTSPD, EUNSUPPORTED = 0, 1, 2, 3, 4, 5
CHOLESKY, CONJUGATE_GRADIENT, NNLS = 0, 1, 2
IMPLICIT, EXPLICIT = 0, 1
ITEMS, USERS = 0, 1


class B200AlsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200als error %d: %s" % (code, msg))
        self.code = code


class Csc(C.Structure):
    """struct b200als_csc (mirrors MappedCSC<double>, inst/include/mapped_csc.hpp:8-29)."""
    _fields_ = [("n_rows", C.c_int32), ("n_cols", C.c_int32), ("nnz", C.c_int64), ("ptr", C.c_void_p),
                ("idx", C.c_void_p), ("val_f64", C.c_void_p), ("val_f32", C.c_void_p)]


class Options(C.Structure):
    _fields_ = [("feedback", C.c_int), ("solver", C.c_int), ("cg_steps", C.c_int), ("dynamic_lambda", C.c_int),
                ("lambda_", C.c_double), ("kernel", C.c_int), ("reserved", C.c_int * 7)]


_lib = None

_PROTOS = {
    "b200als_last_error": (C.c_char_p, []),
    "b200als_version": (C.c_int, []),
    "b200als_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "b200als_set_device": (C.c_int, [C.c_int]),
    "b200als_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "b200als_host_free": (C.c_int, [C.c_void_p]),
    "b200als_timer_start": (C.c_int, []),
    "b200als_timer_stop": (C.c_int, [C.POINTER(C.c_float)]),
    "b200als_launch_count": (C.c_ulonglong, []),
    "b200als_initialize_biases_float": (C.c_int, [C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 8 +
                                        [C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "b200als_initialize_biases_double": (C.c_int, [C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 8 +
                                         [C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "b200als_als_implicit_float": (C.c_int, [C.POINTER(Csc), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                             C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                             C.c_int, C.POINTER(C.c_double)]),
    "b200als_als_implicit_double": (C.c_int, [C.POINTER(Csc), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                              C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                              C.c_int, C.POINTER(C.c_double)]),
    "b200als_als_explicit_float": (C.c_int, [C.POINTER(Csc), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                             C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int,
                                             C.POINTER(C.c_double)]),
    "b200als_als_explicit_double": (C.c_int, [C.POINTER(Csc), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                              C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int,
                                              C.POINTER(C.c_double)]),
    "b200als_gram_float": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_void_p]),
    "b200als_top_product": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "b200als_default_options": (None, [C.POINTER(Options)]),
    "b200als_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Csc), C.POINTER(Csc), C.c_int32, C.c_int32, C.c_int,
                                 C.POINTER(Options)]),
    "b200als_destroy": (C.c_int, [C.c_void_p]),
    "b200als_build_missing_orientation": (C.c_int, [C.c_void_p]),
    "b200als_get_orientation": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "b200als_set_factors": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "b200als_get_factors": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "b200als_init_factors": (C.c_int, [C.c_void_p, C.c_uint64]),
    "b200als_randomize_factors": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_float, C.c_float]),
    "b200als_half_iteration": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "b200als_fit": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.POINTER(C.c_int)]),
    "b200als_transform": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]),
    "b200als_exchange_mode": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "b200als_set_bias": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "b200als_row_plan": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "b200als_last_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                      C.POINTER(C.c_float)]),
    "b200als_comm_unique_id": (C.c_int, [C.c_void_p]),
    "b200als_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "b200als_comm_destroy": (C.c_int, []),
    "b200als_comm_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200als_set_shard": (C.c_int, [C.c_void_p, C.c_int, C.c_int32, C.c_int32]),
    "b200als_synth_csr_host": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_int, C.c_int64, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200als_create_synthetic": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_uint64, C.c_int, C.POINTER(Options)]),
    "b200als_create_synthetic_ex": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_uint64, C.c_int, C.POINTER(Options), C.c_int, C.c_int]),
}


def lib():
    """Load libb200als.so or fail loudly (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). rsparse_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_PROTOS)


def check(rc):
    if rc != OK:
        raise B200AlsError(rc, lib().b200als_last_error().decode("utf-8", "replace"))


def device_count():
    n = C.c_int(0)
    lib().b200als_device_count(C.byref(n))
    return n.value


def vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_csc(n_rows, ptr, idx, val):
    """Build a struct b200als_csc over numpy arrays (kept alive by the returned tuple)."""
    ptr = np.ascontiguousarray(ptr, dtype=np.int32)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    if val.dtype == np.float32:
        val = np.ascontiguousarray(val)
        s = Csc(int(n_rows), len(ptr) - 1, len(idx), vp(ptr), vp(idx), None, vp(val))
    else:
        val = np.ascontiguousarray(val, dtype=np.float64)
        s = Csc(int(n_rows), len(ptr) - 1, len(idx), vp(ptr), vp(idx), vp(val), None)
    return s, (ptr, idx, val)


def pinned_empty(shape, dtype):
    """numpy array over page-locked host memory from b200als_host_alloc (freed when garbage collected)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    p = C.c_void_p(None)
    check(lib().b200als_host_alloc(n * dtype.itemsize, C.byref(p)))
    buf = (C.c_char * (n * dtype.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    _pinned_keep[arr.ctypes.data] = p

    return arr


_pinned_keep = {}


def pinned_free(arr):
    p = _pinned_keep.pop(arr.ctypes.data, None)
    if p is not None:
        lib().b200als_host_free(p)
