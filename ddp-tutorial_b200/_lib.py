"""ctypes binding of libb2ddp.so (include/b2ddp.h).  Fails loudly when the library is
missing or a call errors -- there is no CPU or PyTorch fallback for the hot path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb2ddp.so")

MODEL_NET, MODEL_RESNET18, MODEL_RESNET50 = 0, 1, 2
PREC_FP32, PREC_BF16 = 0, 1
IPC_HANDLE_BYTES = 64


class B2Error(RuntimeError):
    pass


_lib = None

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_SIGS = {
    "b2_version": (_i, []),
    "b2_last_error": (C.c_char_p, []),
    "b2_launch_count": (_i64, []),
    "b2_trace_start": (_i, [_i]),
    "b2_trace_stop": (_i, [_vp, _i, C.POINTER(_i)]),
    "b2_model_create": (_i, [_i, _i, _i, _i, _i, C.POINTER(_vp)]),
    "b2_model_destroy": (_i, [_vp]),
    "b2_model_param_count": (_i64, [_vp]),
    "b2_model_num_tensors": (_i, [_vp]),
    "b2_model_tensor_info": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(_i64), C.POINTER(_i),
                                  C.POINTER(_i64), C.POINTER(_i64)]),
    "b2_model_buffer_count": (_i64, [_vp]),
    "b2_model_num_buffers": (_i, [_vp]),
    "b2_model_buffer_info": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(_i64), C.POINTER(_i64)]),
    "b2_model_steps_tracked": (_i64, [_vp]),
    "b2_model_bind": (_i, [_vp, _vp, _vp, _vp]),
    "b2_model_sync_weights": (_i, [_vp, _vp]),
    "b2_model_mark_weights_dirty": (_i, [_vp]),
    "b2_forward": (_i, [_vp, _vp, _vp, _vp]),
    "b2_backward": (_i, [_vp, _vp, _vp]),
    "b2_model_set_graphs": (_i, [_vp, _i]),
    "b2_model_profile": (_i, [_vp, _i]),
    "b2_model_profile_collect": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i)]),
    "b2_model_time_conv_kernels": (_i, [_vp, _i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i), _vp]),
    "b2_model_check": (_i, [_vp]),
    "b2_model_num_ops": (_i, [_vp]),
    "b2_model_read_activation": (_i, [_vp, _i, _i, _vp, _i64, C.POINTER(_i64), C.POINTER(_i), _vp]),
    "b2_preprocess_u8": (_i, [_vp, _vp, _i, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _vp]),
    "b2_argmax_count": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "b2_conv_test": (_i, [_i] * 11 + [_vp, _vp, _vp, _vp, _vp]),
    "b2_conv_bn_test": (_i, [_i] * 9 + [_vp] * 12),
    "b2_ce_loss_grad": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "b2_ce_loss_grad_running": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "b2_sgd": (_i, [_vp, _vp, _vp, _i64, _f, _f, _f, _i, _vp, _vp]),
    "b2_model_bind_bf16": (_i, [_vp, _vp]),
    "b2_model_mark_weights_clean": (_i, [_vp]),
    "b2_comm_init": (_i, [_i, _i, _i]),
    "b2_comm_finalize": (_i, []),
    "b2_comm_alloc": (_i, [_i64, C.POINTER(_vp)]),
    "b2_comm_free": (_i, [_vp]),
    "b2_comm_export": (_i, [_vp, _vp]),
    "b2_comm_import": (_i, [_vp, C.POINTER(_vp)]),
    "b2_comm_unimport": (_i, [_vp]),
    "b2_comm_register": (_i, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _i64]),
    "b2_comm_barrier": (_i, [_vp]),
    "b2_comm_flag_bytes": (_i, []),
    "b2_vmm_query": (_i, [_i, C.POINTER(_i), C.POINTER(_i64)]),
    "b2_vmm_alloc": (_i, [_i, _i64, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i)]),
    "b2_vmm_import": (_i, [_i, _i, _i64, C.POINTER(_vp)]),
    "b2_mc_create": (_i, [_i, _i64, C.POINTER(_i64), C.POINTER(_i)]),
    "b2_mc_import": (_i, [_i, C.POINTER(_i64)]),
    "b2_mc_add_device": (_i, [_i64, _i]),
    "b2_mc_bind_and_map": (_i, [_i64, _i64, _i64, _i, C.POINTER(_vp)]),
    "b2_vmm_close_fd": (_i, [_i]),
    "b2_vmm_release_all": (_i, []),
    "b2_comm_register_multicast": (_i, [_vp, _vp]),
    "b2_comm_set_nvls": (_i, [_i]),
    "b2_comm_nvls_active": (_i, []),
    "b2_comm_check": (_i, []),
    "b2_allreduce": (_i, [_i64, _f, _i, _vp]),
    "b2_allreduce_range": (_i, [_i64, _i64, _f, _i, _vp]),
    "b2_allreduce_sgd": (_i, [_vp, _i64, _f, _f, _i, _i, _vp, _vp]),
    "b2_allreduce_sgd_range": (_i, [_vp, _i64, _i64, _f, _f, _i, _i, _vp, _vp]),
    "b2_model_set_ddp": (_i, [_vp, _i, _vp, _f, _f, _i, _i64]),
    "b2_model_set_fused_sgd": (_i, [_vp, _i, _vp, _f, _f, _i64]),
    "b2_model_ddp_first_step": (_i, [_vp, _i]),
    "b2_model_num_buckets": (_i, [_vp]),
    "b2_model_bucket_plan": (_i, [_vp, _i64, _i, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i)]),
    "b2_model_workspace_bytes": (_i64, [_vp]),
    "b2_model_op_info": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "b2_graph_begin": (_i, [_vp]),
    "b2_graph_end": (_i, [_vp, C.POINTER(_vp)]),
    "b2_graph_launch": (_i, [_vp, _vp]),
    "b2_graph_destroy": (_i, [_vp]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


def lib() -> C.CDLL:
    """The loaded library.  Raises B2Error (never falls back) if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2Error(
                f"{LIB_PATH} is missing: the CUDA extension is not built.  Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc); "
                "there is no CPU fallback for this path.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().b2_last_error()
        raise B2Error(f"{what or 'b2 call'} failed (rc={rc}): {msg.decode() if msg else '?'}")


def call(name: str, *args):
    """Call an int-returning entry point and raise on a non-zero return code."""
    check(getattr(lib(), name)(*args), name)
