"""ctypes binding of libneuromah_b200.so (C ABI: include/neuromah_b200.h).

The product path has NO fallback: if the shared library is missing or a call
fails, an exception is raised.  Error mapping follows SURVEY.md 8(b): bad
arguments -> ValueError, CUDA failures -> RuntimeError (the reference's C++
raised std::runtime_error -> RuntimeError), out of memory -> MemoryError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NM_LIB_PATH") or os.path.join(_HERE, "libneuromah_b200.so")

NM_OK, NM_ERR_BAD_ARG, NM_ERR_CUDA, NM_ERR_NCCL, NM_ERR_OOM, NM_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_pvp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every int-returning function is status-checked.
_SIGS = {
    "nm_last_error": (C.c_char_p, []),
    "nm_version": (_i, []),
    "nm_device_count": (_i, [C.POINTER(_i)]),
    "nm_init": (_i, [_i]),
    "nm_device_props": (_i, [_i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz)]),
    "nm_malloc": (_i, [_pvp, _sz]),
    "nm_free": (_i, [_vp]),
    "nm_host_alloc": (_i, [_pvp, _sz]),
    "nm_host_free": (_i, [_vp]),
    "nm_memset": (_i, [_vp, _i, _sz, _vp]),
    "nm_h2d": (_i, [_vp, _vp, _sz, _vp]),
    "nm_d2h": (_i, [_vp, _vp, _sz, _vp]),
    "nm_d2d": (_i, [_vp, _vp, _sz, _vp]),
    "nm_stream_create": (_i, [_pvp]),
    "nm_stream_destroy": (_i, [_vp]),
    "nm_stream_sync": (_i, [_vp]),
    "nm_device_sync": (_i, []),
    "nm_event_create": (_i, [_pvp]),
    "nm_event_destroy": (_i, [_vp]),
    "nm_event_record": (_i, [_vp, _vp]),
    "nm_event_sync": (_i, [_vp]),
    "nm_event_elapsed_ms": (_i, [_vp, _vp, C.POINTER(_f)]),
    "nm_stream_wait_event": (_i, [_vp, _vp]),
    "nm_graph_begin": (_i, [_vp]),
    "nm_graph_end": (_i, [_vp, _pvp]),
    "nm_graph_abort": (_i, [_vp]),
    "nm_graph_launch": (_i, [_vp, _vp]),
    "nm_graph_destroy": (_i, [_vp]),
    "nm_profiler_start": (_i, []),
    "nm_profiler_stop": (_i, []),
    "nm_launch_count": (_i, [C.POINTER(C.c_ulonglong)]),
    "nm_conv2d_fprop_f32": (_i, [_vp, _vp, _vp, _vp] + [_i] * 12 + [_vp]),
    "nm_conv2d_dgrad_f32": (_i, [_vp, _vp, _vp] + [_i] * 11 + [_vp]),
    "nm_conv2d_wgrad_f32_workspace": (_sz, [_i] * 11),
    "nm_conv2d_wgrad_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _sz] + [_i] * 11 + [_vp]),
    "nm_relu_fwd_f32": (_i, [_vp, _vp, _sz, _vp]),
    "nm_relu_bwd_f32": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "nm_maxpool2d_fwd_f32": (_i, [_vp, _vp, _vp] + [_i] * 12 + [_vp]),
    "nm_maxpool2d_bwd_f32": (_i, [_vp, _vp, _vp] + [_i] * 12 + [_vp]),
    "nm_dense_fprop_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "nm_dense_dgrad_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "nm_dense_wgrad_f32_workspace": (_sz, [_i, _i, _i]),
    "nm_dense_wgrad_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _f, _f, _f, _vp]),
    "nm_softmax_fwd_f32": (_i, [_vp, _vp, _i, _i, _vp]),
    "nm_softmax_bwd_f32": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "nm_softmax_ce_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "nm_cce_bwd_f32": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp]),
    "nm_sigmoid_fwd_f32": (_i, [_vp, _vp, _sz, _vp]),
    "nm_sigmoid_bwd_f32": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "nm_loss_fwd_f32": (_i, [_i, _vp, _vp, _vp, _i, _i, _vp]),
    "nm_loss_bwd_f32": (_i, [_i, _vp, _vp, _vp, _i, _i, _vp]),
    "nm_adam_step_f32": (_i, [_vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _f, _f, _i, _vp]),
    "nm_adam_state_init": (_i, [_vp, C.c_double, C.c_double, C.c_double, C.c_double, _f, C.c_longlong, _vp]),
    "nm_adam_step_dev_f32": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _i, _vp]),
    "nm_adam_advance_f32": (_i, [_vp, _vp]),
    "nm_adam_step_advance_dev_f32": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _i, _vp]),
    "nm_sgd_step_f32": (_i, [_vp, _vp, _vp, _sz, _f, _f, _i, _vp, _vp]),
    "nm_rmsprop_step_f32": (_i, [_vp, _vp, _vp, _sz, _f, _f, _f, _i, _vp, _vp]),
    "nm_adagrad_step_f32": (_i, [_vp, _vp, _vp, _sz, _f, _f, _i, _vp, _vp]),
    "nm_nadam_step_f32": (_i, [_vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _f, _f, _i, _vp]),
    "nm_dropout_fwd_f32": (_i, [_vp, _vp, _vp, _sz, _f, C.c_ulonglong, C.c_ulonglong, _vp]),
    "nm_mul_f32": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "nm_scale_f32": (_i, [_vp, _vp, _sz, _f, _i, _vp]),
    "nm_u8_to_f32_div": (_i, [_vp, _vp, _sz, _f, _vp]),
    "nm_minmax_f32": (_i, [_vp, _sz, _vp, _vp]),
    "nm_histogram_f32": (_i, [_vp, _sz, C.c_double, C.c_double, _i, _vp, _vp]),
    "nm_tc_pack_conv_weights_bf16": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "nm_tc_conv3x3_fprop_pool_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_conv3x3_dgrad_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_conv3x3_wgrad_workspace": (_sz, [_i, _i]),
    "nm_tc_conv3x3_wgrad_bf16": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_conv3x3_bwd_workspace": (_sz, [_i, _i]),
    "nm_tc_conv3x3_bwd_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_conv3x3_bwd_reduce_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_conv3x3_bwd_pool_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_pack_conv_weights_gen_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "nm_tc_pack_input_nhwc_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_conv3x3_gen_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_conv3x3_wgrad_gen_workspace": (_sz, [_i, _i, _i, _i, _i, _i]),
    "nm_tc_conv3x3_wgrad_gen_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_maxpool2x2_fwd_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_maxpool2x2_bwd_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_dense_dgrad_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "nm_tc_dense_dgrad_mask_bf16": (_i, [_vp, _vp, _vp, _f, _vp, _i, _i, _i, _vp]),
    "nm_tc_to_bf16": (_i, [_vp, _vp, _sz, _vp]),
    "nm_tc_u8_to_bf16": (_i, [_vp, _f, _vp, _sz, _vp]),
    "nm_tc_pack_dense_gen_bf16": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "nm_tc_dense_fprop_gen_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, C.c_ulonglong, C.c_ulonglong, _vp, _vp]),
    "nm_tc_dense_dgrad_gen_bf16": (_i, [_vp, _vp, _vp, _f, _vp, _i, _i, _i, _vp]),
    "nm_tc_dense_wgrad_gen_workspace": (_sz, [_i, _i, _i]),
    "nm_tc_pack_conv_weights_gen_x3": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "nm_tc_pack_input_nhwc_x3": (_i, [_vp, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_pack_input_nhwc_u8_x3": (_i, [_vp, _f, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_conv3x3_gen_x3": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_conv3x3_wgrad_gen_x3": (_i, [_vp, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_maxpool2x2_fwd_x3": (_i, [_vp, _sz, _vp, _sz, _vp, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_merge3_nhwc_to_nchw_f32": (_i, [_vp, _sz, _vp, _i, _i, _i, _i, _vp]),
    "nm_tc_split3_nchw_to_nhwc": (_i, [_vp, _vp, _sz, _i, _i, _i, _i, _vp]),
    "nm_tc_split3_f32": (_i, [_vp, _vp, _f, _vp, _sz, _sz, _vp]),
    "nm_tc_split3_u8": (_i, [_vp, _f, _vp, _sz, _sz, _vp]),
    "nm_tc_pack_dense_gen_x3": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "nm_tc_dense_fprop_gen_x3": (_i, [_vp, _sz, _vp, _vp, _vp, _sz, _vp, _i, _i, _i, _i, _f, C.c_ulonglong, C.c_ulonglong, _vp, _vp]),
    "nm_tc_dense_dgrad_gen_x3": (_i, [_vp, _sz, _vp, _vp, _f, _vp, _sz, _i, _i, _i, _vp]),
    "nm_tc_dense_wgrad_gen_x3": (_i, [_vp, _sz, _vp, _sz, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _f, _f, _f, _vp]),
    "nm_tc_dense_wgrad_gen_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _f, _f, _f, _vp]),
    "nm_tc_dense_bwd_pool_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_pack_conv1_weights_bf16": (_i, [_vp, _vp, _i, _vp]),
    "nm_tc_conv1_fwd_pool_bf16": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "nm_tc_conv1_fwd_pool_u8_bf16": (_i, [_vp, _f, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "nm_tc_conv1_fwd_pool_pack_bf16": (_i, [_vp, _i, _f, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "nm_tc_conv1_bwd_u8_bf16": (_i, [_vp, _f, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _vp]),
    "nm_tc_pack_input_nhwc_u8_bf16": (_i, [_vp, _f, _vp, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_conv1_bwd_workspace": (_sz, [_i]),
    "nm_tc_conv1_bwd_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _vp]),
    "nm_tc_pack_dense_weights": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "nm_tc_pack_cnn_weights_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "nm_tc_pack_dlogits_bf16": (_i, [_vp, _vp, _i, _i, _vp]),
    "nm_tc_dense_softmax_ce_workspace": (_sz, [_i, _i]),
    "nm_tc_dense_softmax_ce_bf16": (_i, [_vp] * 11 + [_sz, _i, _i, _i, _f, _vp]),
    "nm_tc_dense_bwd_unpool_workspace": (_sz, [_i]),
    "nm_tc_dense_bwd_unpool_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "nm_tc_dense_wgrad_workspace": (_sz, [_i, _i]),
    "nm_tc_dense_wgrad_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _f, _f, _f, _vp]),
    "nm_comm_load": (_i, [C.c_char_p]),
    "nm_comm_unique_id": (_i, [_vp]),
    "nm_comm_init": (_i, [_pvp, _vp, _i, _i]),
    "nm_comm_destroy": (_i, [_vp]),
    "nm_comm_allreduce_sum_f32": (_i, [_vp, _vp, _sz, _vp]),
    "nm_comm_broadcast_f32": (_i, [_vp, _vp, _sz, _i, _vp]),
    "nm_dp_create": (_i, [_pvp, _i, _i, _sz]),
    "nm_dp_export": (_i, [_vp, _vp]),
    "nm_dp_connect": (_i, [_vp, _vp]),
    "nm_dp_set_timeout_ms": (_i, [_vp, C.c_uint]),
    "nm_dp_allreduce_sum_f32": (_i, [_vp, _vp, _sz, _vp]),
    "nm_dp_allreduce_adam_dev_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _i, _i, _vp]),
    "nm_dp_chunk_floats": (_i, []),
    "nm_dp_push_f32": (_i, [_vp, _vp, _sz, _i, _i, _vp]),
    "nm_dp_collect_adam_dev_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _i, _i, _i, _vp]),
    "nm_dp_status": (_i, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_uint)]),
    "nm_dp_destroy": (_i, [_vp]),
}

_UNCHECKED = {"nm_last_error", "nm_version", "nm_dp_chunk_floats", "nm_conv2d_wgrad_f32_workspace", "nm_dense_wgrad_f32_workspace",
              "nm_tc_conv3x3_wgrad_workspace", "nm_tc_conv1_bwd_workspace", "nm_tc_dense_wgrad_workspace",
              "nm_tc_dense_softmax_ce_workspace", "nm_tc_dense_bwd_unpool_workspace", "nm_tc_conv3x3_bwd_workspace",
              "nm_tc_conv3x3_wgrad_gen_workspace", "nm_tc_dense_wgrad_gen_workspace"}


class NeuromahB200Error(RuntimeError):
    pass


def _raise(name: str, rc: int, dll):
    msg = dll.nm_last_error().decode("utf-8", "replace")
    text = f"{name} failed ({rc}): {msg}"
    if rc == NM_ERR_BAD_ARG:
        raise ValueError(text)
    if rc == NM_ERR_OOM:
        raise MemoryError(text)
    raise NeuromahB200Error(text)


class _Lib:
    """Lazily loaded library; attribute access returns status-checked callables."""

    def __init__(self):
        self._dll = None

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(neuromah_b200 has no CPU fallback)")
        dll = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGS.items():
            fn = getattr(dll, name)          # AttributeError if the .so is stale
            fn.restype, fn.argtypes = res, args
        self._dll = dll
        return dll

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        dll = self.load()
        fn = getattr(dll, name)
        if name in _UNCHECKED:
            return fn

        def checked(*a, _fn=fn, _name=name):
            rc = _fn(*a)
            if rc != 0:
                _raise(_name, rc, dll)
            return rc
        checked.__name__ = name
        setattr(self, name, checked)
        return checked


lib = _Lib()


def declared_symbols(header_path: str | None = None):
    """Names of every function declared in include/neuromah_b200.h (for the ABI test)."""
    import re
    if header_path is None:
        header_path = os.path.join(os.path.dirname(_HERE), "include", "neuromah_b200.h")
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nm_[a-z0-9_]+)\s*\(", text)))
