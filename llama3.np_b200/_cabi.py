"""ctypes binding of libllama3_b200.so (include/llama3_b200.h).  No fallback: if the library
is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# L3_LIB_VARIANT=name selects libllama3_b200_name.so, a build of the same sources with extra -D switches
# (`build.py --variant name -DFOO`): A/B measurement of compile-time kernel variants, never a fallback.
_VARIANT = os.environ.get("L3_LIB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, f"libllama3_b200_{_VARIANT}.so" if _VARIANT else "libllama3_b200.so")

L3_OK, L3_EINVAL, L3_ECUDA, L3_ESTATE, L3_ENOMEM, L3_ENCCL = 0, -1, -2, -3, -4, -5
DTYPE_F32, DTYPE_BF16 = 0, 1
FLAG_NO_GRAPH, FLAG_NO_TENSORCORE, FLAG_NO_PDL, FLAG_NO_MEGA = 1, 2, 4, 8


class L3Config(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_layers", C.c_int32), ("n_heads", C.c_int32),
                ("n_kv_heads", C.c_int32), ("vocab_size", C.c_int32), ("max_seq_len", C.c_int32),
                ("max_batch_size", C.c_int32), ("hidden_dim", C.c_int32), ("norm_eps", C.c_float),
                ("dtype", C.c_int32), ("device", C.c_int32), ("tp_rank", C.c_int32),
                ("tp_world", C.c_int32), ("flags", C.c_int32)]


_P = C.c_void_p
_I = C.c_int
_F32P = C.POINTER(C.c_float)
_F64P = C.POINTER(C.c_double)
_I32P = C.POINTER(C.c_int32)
_I64P = C.POINTER(C.c_int64)

# name -> (restype, argtypes); must list every symbol the header declares (tests check this)
SIGNATURES = {
    "l3_version": (C.c_char_p, []),
    "l3_last_error": (C.c_char_p, [_P]),
    "l3_device_count": (_I, [C.POINTER(_I)]),
    "l3_create": (_I, [C.POINTER(L3Config), C.POINTER(_P)]),
    "l3_load_weight": (_I, [_P, C.c_char_p, _F32P, _I64P, _I]),
    "l3_fill_random": (_I, [_P, C.c_uint64]),
    "l3_save_packed": (_I, [_P, C.c_char_p, C.c_char_p]),
    "l3_load_packed": (_I, [_P, C.c_char_p, C.c_char_p]),
    "l3_packed_info": (_I, [C.c_char_p, C.POINTER(L3Config), C.c_char_p, _I]),
    "l3_set_rope_tables": (_I, [_P, _F64P, _F64P]),
    "l3_finalize": (_I, [_P]),
    "l3_destroy": (_I, [_P]),
    "l3_reset_cache": (_I, [_P]),
    "l3_tp_init": (_I, [_P, C.c_char_p]),
    "l3_nccl_unique_id": (_I, [C.c_char_p]),
    "l3_forward": (_I, [_P, _I32P, _I, _I, _I, _F32P, _I64P]),
    "l3_forward_dev": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "l3_generate_greedy": (_I, [_P, _I32P, _I, _I, _I, _I64P]),
    "l3_generate_greedy_dev": (_I, [_P, _P, _I, _I, _I, _P]),
    "l3_generate_begin": (_I, [_P, _I32P, _I, _I]),
    "l3_generate_begin_ex": (_I, [_P, _I32P, _I, _I, _I]),
    "l3_generate_next": (_I, [_P, _I64P]),
    "l3_generate_ragged": (_I, [_P, _I32P, _I32P, _I, _I, _I, _I, _I, _I64P]),
    "l3_read_cache": (_I, [_P, _I, _F32P, _F32P]),
    "l3_op_rmsnorm": (_I, [_I, _F32P, _F32P, C.c_float, _I, _I, _F32P]),
    "l3_op_linear": (_I, [_I, _F32P, _F32P, _I, _I, _I, _I, _I, _F32P]),
    "l3_op_rope": (_I, [_I, _F32P, _F64P, _F64P, _I, _I, _I, _I, _I, _F32P]),
    "l3_op_swiglu": (_I, [_I, _F32P, _F32P, C.c_int64, _F32P]),
    "l3_op_attention": (_I, [_I, _F32P, _F32P, _F32P, _I, _I, _I, _I, _I, _I, _I, _I, _F32P]),
    "l3_op_argmax": (_I, [_I, _F32P, _I, _I, _I64P]),
    "l3_debug_tc_timeline": (_I, [_I, _I, C.POINTER(C.c_uint64)]),
    "l3_sync": (_I, [_P]),
    "l3_timer_start": (_I, [_P]),
    "l3_timer_stop": (_I, [_P, _F32P]),
    "l3_dev_alloc": (_I, [_P, C.c_int64, C.POINTER(_P)]),
    "l3_dev_free": (_I, [_P, _P]),
    "l3_memcpy_h2d": (_I, [_P, _P, _P, C.c_int64]),
    "l3_memcpy_d2h": (_I, [_P, _P, _P, C.c_int64]),
    "l3_flush_l2": (_I, [_P]),
    "l3_launch_count": (_I, [_P, _I64P, _I]),
    "l3_bench_kernel": (_I, [_P, _I, _I, _I, _I, _F32P]),
    "l3_debug_mega_timeline": (_I, [_P, C.POINTER(C.c_uint64), C.c_int64]),
    "l3_debug_stack": (_I, [_P, _I, _P, C.c_int64]),
    "l3_bench_gemv": (_I, [_I, _I, _I, _I, _I, _I, _F32P]),
    "l3_probe_mma": (_I, [_I, _I, _I, _I, _I, _I, _F64P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load (once) the in-tree library.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python llama3.np_b200/build.py` "
                "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, handle=None) -> None:
    if rc == L3_OK:
        return
    msg = lib().l3_last_error(handle)
    msg = msg.decode() if msg else ""
    if rc == L3_EINVAL:
        raise ValueError(msg or "invalid argument")
    names = {L3_ECUDA: "CUDA", L3_ESTATE: "state", L3_ENOMEM: "out of memory", L3_ENCCL: "NCCL"}
    raise RuntimeError(f"llama3_b200 {names.get(rc, rc)} error: {msg}")


def f32p(a):
    return a.ctypes.data_as(_F32P)


def f64p(a):
    return a.ctypes.data_as(_F64P)


def i32p(a):
    return a.ctypes.data_as(_I32P)


def i64p(a):
    return a.ctypes.data_as(_I64P)
