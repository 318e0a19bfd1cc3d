"""ctypes binding of libcoordx_b200.so (include/coordx_b200.h).

The library is the product; there is NO fallback: if it cannot be loaded, every
hot-path call raises.  Loading the library needs no GPU (symbols only), so the
CPU test-suite can check the ABI surface.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('COORDX_B200_LIB') or os.path.join(HERE, 'libcoordx_b200.so')   # override: tuning builds only

MAX_BRANCHES = 4
MAX_LAYERS = 12
ACT_SINE, ACT_RELU = 0, 1
ACTS = {'sine': 0, 'relu': 1, 'sigmoid': 2, 'tanh': 3, 'selu': 4, 'softplus': 5, 'elu': 6}     # COORDX_ACT_*
FUSE_SUM, FUSE_PROD = 0, 1
LOSS_MSE, LOSS_BCE = 0, 1
PREC_FP32, PREC_BF16, PREC_FP16 = 0, 1, 2
PRECISIONS = {'fp32': PREC_FP32, 'bf16': PREC_BF16, 'fp16': PREC_FP16}
ERR_UNSUPPORTED = -2

FloatP = C.c_void_p     # device pointers travel as integers (tensor.data_ptr())


class Net(C.Structure):
    """struct coordx_net"""
    _fields_ = [
        ('n_branches', C.c_int32),
        ('branch_in', C.c_int32 * MAX_BRANCHES),
        ('pre_width', C.c_int32),
        ('fuse_width', C.c_int32),
        ('hidden', C.c_int32),
        ('n_pre', C.c_int32),
        ('n_layers', C.c_int32),
        ('layer_in', C.c_int32 * MAX_LAYERS),
        ('layer_out', C.c_int32 * MAX_LAYERS),
        ('act', C.c_int32),
        ('fusion_op', C.c_int32),
        ('fusion_before_act', C.c_int32),
        ('has_code', C.c_int32),
        ('act_scale', C.c_float),
        ('first_w', FloatP * MAX_BRANCHES),
        ('first_b', FloatP * MAX_BRANCHES),
        ('layer_w', FloatP * MAX_LAYERS),
        ('layer_b', FloatP * MAX_LAYERS),
        ('code', FloatP),
    ]


class Grads(C.Structure):
    """struct coordx_grads"""
    _fields_ = [
        ('first_w', FloatP * MAX_BRANCHES),
        ('first_b', FloatP * MAX_BRANCHES),
        ('layer_w', FloatP * MAX_LAYERS),
        ('layer_b', FloatP * MAX_LAYERS),
        ('code', FloatP),
    ]


# name -> (restype, argtypes); mirrors include/coordx_b200.h one to one
_I64P = C.POINTER(C.c_int64)
_PP = C.POINTER(C.c_void_p)
SIGNATURES = {
    'coordx_abi_version': (C.c_int, []),
    'coordx_last_error': (C.c_char_p, []),
    'coordx_launch_count': (C.c_ulonglong, []),
    'coordx_profile_enable': (C.c_int, [C.c_int]),
    'coordx_profile_last_ms': (C.c_int, [C.POINTER(C.c_float)]),
    'coordx_profile_set_trace': (C.c_int, [C.c_void_p]),
    'coordx_profile_set_bwd_trace': (C.c_int, [C.c_void_p, C.c_int]),
    'coordx_trunk_tc_supported': (C.c_int, [C.POINTER(Net), C.c_int]),
    'coordx_branch_workspace_bytes': (C.c_size_t, [C.POINTER(Net), C.c_int, C.c_int64, C.c_int]),
    'coordx_branch_forward': (C.c_int, [C.POINTER(Net), C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_size_t, C.c_void_p]),
    'coordx_branch_backward': (C.c_int, [C.POINTER(Net), C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(Grads),
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'coordx_branch_forward_save': (C.c_int, [C.POINTER(Net), C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                             C.c_size_t, C.c_void_p]),
    'coordx_branch_backward_saved': (C.c_int, [C.POINTER(Net), C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                               C.POINTER(Grads), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'coordx_trunk_workspace_bytes': (C.c_size_t, [C.POINTER(Net), _I64P, C.c_int64, C.c_int64, C.c_int, C.c_int]),
    'coordx_trunk_forward': (C.c_int, [C.POINTER(Net), _PP, _I64P, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_size_t, C.c_void_p]),
    'coordx_gather_rows': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'coordx_trunk_forward_loss': (C.c_int, [C.POINTER(Net), _PP, _I64P, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    'coordx_trunk_backward': (C.c_int, [C.POINTER(Net), _PP, _I64P, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(Grads),
                                        _PP, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    'coordx_loss_forward_backward': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p,
                                               C.c_void_p, C.c_void_p]),
    'coordx_adam_step': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float,
                                   C.c_float, C.c_float, C.c_int64, C.c_void_p, C.c_void_p]),
    'coordx_optim_step': (C.c_int, [C.c_int, _PP, _PP, _PP, _PP, _I64P, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                    C.c_float, C.c_float, C.c_void_p]),
    'coordx_interp_fuse': (C.c_int, [C.c_int, _PP, _I64P, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'coordx_infer_grid_host_workspace_bytes': (C.c_size_t, [C.POINTER(Net), _I64P, C.c_int]),
    'coordx_infer_grid_host': (C.c_int, [C.POINTER(Net), _PP, _I64P, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                         C.c_void_p]),
}

_lock = threading.Lock()
_lib = None


class CoordxError(RuntimeError):
    pass


def load():
    """dlopen the in-tree library (building it is `python -m coordx_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise CoordxError(
                f"{LIB_PATH} is missing: build it with `python -m coordx_b200.build` "
                "(nvcc, sm_100a).  coordx_b200 has no CPU / eager fallback for the split-MLP hot path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        if lib.coordx_abi_version() != 3:
            raise CoordxError('libcoordx_b200.so ABI version mismatch')
        _lib = lib
    return _lib


def check(status, what):
    if status != 0:
        msg = load().coordx_last_error().decode(errors='replace')
        raise CoordxError(f'{what} failed ({status}): {msg}')


def i64_array(vals):
    return (C.c_int64 * len(vals))(*[int(v) for v in vals])


def ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])
