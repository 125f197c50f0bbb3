"""ctypes binding of libnvd.so — prototypes mirror include/nvd.h one to one."""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.dirname(_HERE)
ROOT_DIR = os.path.dirname(PKG_DIR)
LIB_PATH = os.environ.get("NVD_LIB", os.path.join(PKG_DIR, "libnvd.so"))
HEADER_PATH = os.path.join(ROOT_DIR, "include", "nvd.h")

# enums of include/nvd.h
OK = 0
CODE_57_REF, CODE_133171, CODE_57_CAUSAL = 0, 1, 2
F32, BF16, I32_TRUNC, BF16_TRUNC = 0, 1, 2, 3
ACT_IDENTITY, ACT_TANH, ACT_RELU = 0, 1, 2
PREC_FP32, PREC_BF16_TC = 0, 1
MC_STAGED, MC_FUSED = 0, 1
COUNT_PRED_IS_MSB_CLASS = 1
COMM_ID_BYTES = 128
COMM_IPC_BYTES = 64

CODES = {"57": CODE_57_REF, "57_ref": CODE_57_REF, "133171": CODE_133171, "57_causal": CODE_57_CAUSAL}
DTYPES = {"f32": F32, "bf16": BF16, "i32_trunc": I32_TRUNC, "bf16_trunc": BF16_TRUNC}
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16_TC}


class NvdError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libnvd error {status}: {message}")
        self.status = status


class McConfig(C.Structure):
    _fields_ = [("struct_bytes", C.c_int32), ("model", C.c_int32), ("code", C.c_int32), ("L", C.c_int32),
                ("precision", C.c_int32), ("mode", C.c_int32), ("y_dtype", C.c_int32), ("reserved", C.c_int32),
                ("frame0", C.c_uint64), ("frames", C.c_uint64)]


class TrainConfig(C.Structure):
    _fields_ = [("struct_bytes", C.c_int32), ("loss", C.c_int32), ("learning_rate", C.c_float), ("beta", C.c_float),
                ("l2_layer_mask", C.c_uint32), ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float)]


LOSSES = {"softmax_ce": 0, "sigmoid_ce": 1}


class McBitwiseConfig(C.Structure):
    _fields_ = [("struct_bytes", C.c_int32), ("model", C.c_int32), ("code", C.c_int32), ("L", C.c_int32),
                ("n_state_bits", C.c_int32), ("state", C.c_int32), ("bit_position", C.c_int32), ("precision", C.c_int32),
                ("mode", C.c_int32), ("trunc", C.c_int32), ("frame0", C.c_uint64), ("frames", C.c_uint64)]


class McLstmConfig(C.Structure):
    _fields_ = [("struct_bytes", C.c_int32), ("model", C.c_int32), ("code", C.c_int32), ("L", C.c_int32),
                ("fast_math", C.c_int32), ("teacher_forced", C.c_int32), ("trunc", C.c_int32), ("reserved", C.c_int32),
                ("frame0", C.c_uint64), ("frames", C.c_uint64)]


_vp, _i, _i64, _u64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float
_PROTOS = {
    "nvd_version": ([], C.c_int),
    "nvd_last_error": ([], C.c_char_p),
    "nvd_device_count": ([C.POINTER(_i)], _i),
    "nvd_ctx_create": ([_i, _u64, C.POINTER(_vp)], _i),
    "nvd_ctx_destroy": ([_vp], _i),
    "nvd_set_seed": ([_vp, _u64], _i),
    "nvd_sync": ([_vp], _i),
    "nvd_device_info": ([_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(C.c_size_t)], _i),
    "nvd_alloc": ([_vp, C.c_size_t, C.POINTER(_vp)], _i),
    "nvd_free": ([_vp, _vp], _i),
    "nvd_host_alloc": ([_vp, C.c_size_t, C.POINTER(_vp)], _i),
    "nvd_host_free": ([_vp, _vp], _i),
    "nvd_memcpy_h2d": ([_vp, _vp, _vp, C.c_size_t], _i),
    "nvd_memcpy_d2h": ([_vp, _vp, _vp, C.c_size_t], _i),
    "nvd_memset": ([_vp, _vp, _i, C.c_size_t], _i),
    "nvd_event_create": ([_vp, C.POINTER(_vp)], _i),
    "nvd_event_record": ([_vp, _vp], _i),
    "nvd_event_elapsed_ms": ([_vp, _vp, _vp, C.POINTER(_f)], _i),
    "nvd_event_destroy": ([_vp, _vp], _i),
    "nvd_source": ([_vp, _i, _u64, _i64, _vp], _i),
    "nvd_encode": ([_vp, _i, _vp, _i, _i64, _vp], _i),
    "nvd_channel": ([_vp, _vp, _i, _i64, _f, _i, _u64, _i, _vp], _i),
    "nvd_mlp_load": ([_vp, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i)], _i),
    "nvd_mlp_free": ([_vp, _i], _i),
    "nvd_mlp_forward": ([_vp, _i, _vp, _i, _i64, _i, _vp, _vp, _i], _i),
    "nvd_mlp_tc_supported": ([_vp, _i, C.POINTER(_i)], _i),
    "nvd_mlp_param": ([_vp, _i, _i, _i, C.POINTER(_vp)], _i),
    "nvd_window_decode": ([_vp, _i, _vp, _i, _i64, _i64, _i64, _i, _vp, _vp, _i], _i),
    "nvd_tc_tanh_model": ([C.POINTER(_i), C.POINTER(_f), C.POINTER(_f)], _i),
    "nvd_mlp_decode_host": ([_vp, _i, _vp, _i64, _i, _vp, _vp], _i),
    "nvd_lstm_load": ([_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i)], _i),
    "nvd_lstm_free": ([_vp, _i], _i),
    "nvd_lstm_forward": ([_vp, _i, _vp, _i, _vp, _i, _i, _i64, _i, _vp, _vp], _i),
    "nvd_count_errors": ([_vp, _vp, _vp, _i, _i, _i64, _i, _vp], _i),
    "nvd_mc_run": ([_vp, C.POINTER(McConfig), C.POINTER(_f), C.POINTER(C.c_int32), _i, _vp], _i),
    "nvd_mc_run_host": ([_vp, C.POINTER(McConfig), _f, _i, _vp, _vp, _vp], _i),
    "nvd_mc_stage_ms": ([_vp, C.POINTER(_f)], _i),
    "nvd_mc_set_stage_timing": ([_vp, _i], _i),
    "nvd_mc_run_bitwise": ([_vp, C.POINTER(McBitwiseConfig), C.POINTER(_f), C.POINTER(C.c_int32), _i, _vp], _i),
    "nvd_mc_run_lstm": ([_vp, C.POINTER(McLstmConfig), C.POINTER(_f), C.POINTER(C.c_int32), _i, _vp], _i),
    "nvd_count_bit_errors": ([_vp, _vp, _vp, _i, _i, _i64, _vp], _i),
    "nvd_count_row_errors": ([_vp, _vp, _i, _vp, _i, _i, _i64, _vp], _i),
    "nvd_prepend_state": ([_vp, _vp, _i, _i, _i, _i64, _vp], _i),
    "nvd_train_begin": ([_vp, _i, C.POINTER(TrainConfig), _i64, C.POINTER(_i)], _i),
    "nvd_train_step": ([_vp, _i, _vp, _vp, _i64, C.POINTER(_f)], _i),
    "nvd_train_end": ([_vp, _i, C.POINTER(_vp), C.POINTER(_vp)], _i),
    "nvd_comm_unique_id": ([_vp], _i),
    "nvd_comm_init": ([_vp, _i, _i, _vp], _i),
    "nvd_comm_destroy": ([_vp], _i),
    "nvd_comm_info": ([_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)], _i),
    "nvd_comm_allreduce_u64": ([_vp, _vp, C.c_size_t], _i),
    "nvd_comm_p2p_handle": ([_vp, _vp], _i),
    "nvd_comm_p2p_open": ([_vp, _i, _i, _vp], _i),
    "nvd_comm_p2p_close": ([_vp], _i),
    "nvd_mc_sweep": ([_vp, C.POINTER(McConfig), C.POINTER(_f), C.POINTER(C.c_int32), _i, _u64, _u64, _u64, _vp], _i),
    "nvd_debug_tc_timing": ([_vp, _i, _u64, _vp], _i),
    "nvd_debug_umma_probe_ts": ([_vp, _i, _i, _i, _vp, _vp, _vp], _i),
    "nvd_debug_umma_probe": ([_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp], _i),
}

_lib = None


def header_symbols() -> list[str]:
    """Every function declared in include/nvd.h."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nvd_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    """Load libnvd.so (built by __graft_entry__.build()); fail loudly — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NvdError(-2, f"{LIB_PATH} not built: run `python __graft_entry__.py` (nvcc, sm_100a). "
                           "There is no CPU fallback for the decode path.")
    lib = C.CDLL(LIB_PATH)
    for name in header_symbols():
        if not hasattr(lib, name):
            raise NvdError(-2, f"libnvd.so does not export {name} declared in include/nvd.h")
        if name not in _PROTOS:
            raise NvdError(-2, f"no ctypes prototype for {name}")
    for name, (argtypes, restype) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != OK:
        raise NvdError(status, load().nvd_last_error().decode("utf-8", "replace"))


def tc_tanh_model():
    """(poly_pairs_of_8, coef[6], inv_range) of the tensor-core path's two-pipe tanh (include/nvd.h)."""
    n, r = C.c_int(0), C.c_float(0)
    coef = (C.c_float * 6)()
    check(load().nvd_tc_tanh_model(C.byref(n), coef, C.byref(r)))
    return n.value, [float(c) for c in coef], float(r.value)


def device_count() -> int:
    n = C.c_int(0)
    try:
        check(load().nvd_device_count(C.byref(n)))
    except NvdError:
        return 0
    return n.value
