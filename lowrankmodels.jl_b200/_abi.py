"""ctypes binding of the C ABI (include/glrm_b200.h) — the Python twin of the Julia `ccall` shim
(julia/LowRankModelsB200.jl).  Loading fails loudly when the CUDA library has not been built: the
product has no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GLRMB200_LIB", os.path.join(_HERE, "csrc", "libglrm_b200.so"))  # env: tuning builds

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class Problem(C.Structure):
    _fields_ = [
        ("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int64), ("d", C.c_int64),
        ("loss_code", c_int32_p), ("loss_param", c_double_p),
        ("rx_count", C.c_int64), ("rx_code", c_int32_p), ("rx_param", c_double_p),
        ("ry_count", C.c_int64), ("ry_code", c_int32_p), ("ry_param", c_double_p),
        ("obs_full", C.c_int32), ("dense_A", c_double_p),
        ("row_ptr", c_int64_p), ("row_idx", c_int32_p), ("row_val", c_double_p),
        ("col_ptr", c_int64_p), ("col_idx", c_int32_p), ("col_val", c_double_p),
        ("rx_payload_ptr", c_int64_p), ("rx_payload", c_double_p),
        ("ry_payload_ptr", c_int64_p), ("ry_payload", c_double_p),
    ]


class Params(C.Structure):
    _fields_ = [
        ("stepsize", C.c_double), ("max_iter", C.c_int32), ("inner_iter_X", C.c_int32),
        ("inner_iter_Y", C.c_int32), ("abs_tol", C.c_double), ("rel_tol", C.c_double),
        ("min_stepsize", C.c_double),
    ]


class SparseParams(C.Structure):
    _fields_ = [("stepsize", C.c_double), ("max_iter", C.c_int32), ("inner_iter", C.c_int32),
                ("abs_tol", C.c_double), ("min_stepsize", C.c_double)]


class Profile(C.Structure):
    _fields_ = [
        ("setup_ms", C.c_double), ("update_x_ms", C.c_double), ("update_y_ms", C.c_double),
        ("reduce_ms", C.c_double), ("comm_ms", C.c_double), ("loop_ms", C.c_double),
        ("x_launches", C.c_int64), ("y_launches", C.c_int64), ("other_launches", C.c_int64),
        ("x_trials", C.c_int64), ("y_trials", C.c_int64),
        ("iterations", C.c_int32), ("reserved", C.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


Handle = C.c_void_p

# every symbol include/glrm_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("glrmb200_version", C.c_int, []),
    ("glrmb200_last_error", C.c_char_p, []),
    ("glrmb200_device_count", C.c_int, [c_int32_p]),
    ("glrmb200_create", C.c_int, [C.POINTER(Handle), C.POINTER(Problem), C.c_int32, C.c_int32, C.c_int32]),
    ("glrmb200_create_ex", C.c_int, [C.POINTER(Handle), C.POINTER(Problem), C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    ("glrmb200_comm_unique_id", C.c_int, [C.POINTER(C.c_uint8)]),
    ("glrmb200_comm_init", C.c_int, [Handle, C.POINTER(C.c_uint8)]),
    ("glrmb200_ipc_export", C.c_int, [Handle, C.POINTER(C.c_uint8)]),
    ("glrmb200_ipc_open", C.c_int, [Handle, C.POINTER(C.c_uint8)]),
    ("glrmb200_comm_barrier", C.c_int, [Handle]),
    ("glrmb200_shard", C.c_int, [Handle, c_int64_p, c_int64_p, c_int64_p, c_int64_p]),
    ("glrmb200_fit", C.c_int, [Handle, C.POINTER(Params), c_double_p, c_double_p, c_double_p, c_double_p,
                                C.c_int32, c_int32_p, C.POINTER(Profile)]),
    ("glrmb200_fit_sparse", C.c_int, [Handle, C.POINTER(SparseParams), c_double_p, c_double_p, c_double_p, c_double_p,
                                       C.c_int32, c_int32_p, C.POINTER(Profile)]),
    ("glrmb200_objective", C.c_int, [Handle, c_double_p, c_double_p, C.c_int32, c_double_p]),
    ("glrmb200_set_reg_scale", C.c_int, [Handle, C.c_double]),
    ("glrmb200_set_obs", C.c_int, [Handle, c_int64_p, c_int32_p, c_double_p, c_int64_p, c_int32_p, c_double_p]),
    ("glrmb200_upload_factors", C.c_int, [Handle, c_double_p, c_double_p]),
    ("glrmb200_fit_resident", C.c_int, [Handle, C.POINTER(Params), c_double_p, c_double_p, C.c_int32,
                                         c_int32_p, C.POINTER(Profile)]),
    ("glrmb200_download_factors", C.c_int, [Handle, c_double_p, c_double_p]),
    ("glrmb200_impute", C.c_int, [Handle, c_double_p, c_double_p, c_int32_p, c_double_p, c_double_p]),
    ("glrmb200_error_metric", C.c_int, [Handle, c_double_p, c_double_p, c_int32_p, c_double_p, C.c_int32, c_double_p]),
    ("glrmb200_get_stepsizes", C.c_int, [Handle, c_double_p, c_double_p]),
    ("glrmb200_destroy", C.c_int, [Handle]),
    ("glrmb200_plan_shards", C.c_int, [c_int64_p, C.c_int64, C.c_int32, c_int64_p]),
    ("glrmb200_plan_dense_rows", C.c_int, [C.c_int64, C.c_int32, c_int64_p]),
]

_lib = None


class GLRMB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"glrmb200 error {code}: {msg}")
        self.code = code


def load(path):
    """dlopen a build of the engine and type every symbol of include/glrm_b200.h."""
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build the CUDA engine first "
            "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")
    L = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(L, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return L


def lib():
    """Load csrc/libglrm_b200.so (built by __graft_entry__.build() / csrc/build.sh)."""
    global _lib
    if _lib is None:
        _lib = load(LIB_PATH)
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().glrmb200_last_error()
        raise GLRMB200Error(rc, msg.decode() if msg else "")


def dptr(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def i32ptr(a):
    return a.ctypes.data_as(c_int32_p) if a is not None else None


def i64ptr(a):
    return a.ctypes.data_as(c_int64_p) if a is not None else None
