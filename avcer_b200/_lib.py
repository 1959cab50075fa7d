"""ctypes binding of libavcer_b200.so (the C ABI declared in include/avcer_b200.h).

There is deliberately no CPU fallback: importing this module without the built library, or
calling an op without a B200 visible, raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int16, c_int32, c_int64, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# AVCER_LIB: alternative build of the same library (A/B measurements of kernel changes); default is the in-tree build
LIB_PATH = os.environ.get("AVCER_LIB") or os.path.join(_HERE, "libavcer_b200.so")
# the same sources compiled with -DAVCER_HALF: IEEE half as the 16-bit storage type (precision "fp16")
LIB_PATH_FP16 = os.path.join(_HERE, "libavcer_b200_fp16.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2


class AvcerError(RuntimeError):
    pass


class ContractDesc(ctypes.Structure):
    """Mirror of avcer_contract_desc (include/avcer_b200.h)."""

    _fields_ = [
        ("a", c_void_p),
        ("a_dim", c_int64 * 5),
        ("a_stride", c_int64 * 5),
        ("wt", c_void_p),
        ("bias", c_void_p),
        ("residual", c_void_p),
        ("out", c_void_p),
        ("out_stride", c_int64 * 3),
        ("res_stride", c_int64 * 3),
        ("W", c_int32),
        ("H", c_int32),
        ("NB", c_int32),
        ("cin", c_int32),
        ("cout", c_int32),
        ("taps_w", c_int32),
        ("taps_h", c_int32),
        ("off_w", c_int32),
        ("off_h", c_int32),
        ("tap_h_in_dim4", c_int32),
        ("group_cin_shift", c_int32),
        ("a_strip", c_int32),
        ("wt_packed", c_void_p),
        ("act", c_int32),
        ("res_after_act", c_int32),
        ("dtype", c_int32),
        ("out_f32", c_int32),
        ("a_step", c_int32),
        ("reverse_tiles", c_int32),
    ]


# name -> (restype, argtypes); every symbol include/avcer_b200.h declares.
_SIGNATURES = {
    "avcer_last_error": (c_char_p, []),
    "avcer_version": (c_int, []),
    "avcer_storage_type": (c_char_p, []),
    "avcer_device_check": (c_int, []),
    "avcer_num_sms": (c_int, []),
    "avcer_set_sm_limit": (c_int, [c_int]),
    "avcer_debug_set_trace": (c_int, [c_void_p]),
    "avcer_preprocess_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "avcer_preprocess_maps": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "avcer_contract": (c_int, [POINTER(ContractDesc), c_void_p]),
    "avcer_last_contract_kernel": (c_int, []),
    "avcer_fuse_compound": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_double), POINTER(c_double), c_int, c_int, c_void_p, c_int64, c_void_p]),
    "avcer_fuse_compound_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, POINTER(c_double), POINTER(c_double), c_int, c_int, c_void_p, c_int64, c_void_p]),
    "avcer_compound_scores": (c_int, [c_void_p, c_int64, c_int, c_int, POINTER(c_int32), POINTER(c_double), c_int, c_int, c_void_p, c_void_p]),
    "avcer_weight_search_confusion": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "avcer_fused_argmax": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "avcer_softmax7": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "avcer_softmax7_f64": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "avcer_window_to_frame_mean": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "avcer_window_to_frame_mean_f64": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "avcer_gather_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "avcer_gather_rows_f64": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "avcer_stem_pool": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "avcer_stem_pool_ld": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p]),
    "avcer_stem_pool_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p]),
    "avcer_maxpool3x3s2": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "avcer_avgpool": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "avcer_small_linear": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "avcer_lstm_cell": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "avcer_gru_cell": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "avcer_split_bf16x3": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_int64, c_void_p]),
    "avcer_pcm16_resample": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "avcer_audio_normalize_windows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "avcer_w2v_conv0_tc": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int64, c_void_p]),
    "avcer_w2v_conv0_ln_gelu": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "avcer_layernorm": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_float, c_int, c_void_p, c_int64, c_int, c_void_p]),
    "avcer_subsample_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_int, c_void_p]),
    "avcer_add_rows": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    "avcer_attention": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int, c_void_p]),
    "avcer_maxpool1d5_relu": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "avcer_avgpool1d_relu": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "avcer_jpeg_decode": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "avcer_read_files": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int]),
    "avcer_det_stem": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "avcer_det_prepare": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "avcer_maxpool3x3s2p1": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "avcer_upsample_add": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "avcer_det_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "avcer_cast": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p]),
}

_libs = {}


def load(kind: str = "bf16") -> ctypes.CDLL:
    """Load a build of the shared library and bind every exported symbol (raises if it is missing).
    kind "bf16": libavcer_b200.so (bfloat16 storage; also every fp32 / fp64 / integer kernel); kind "fp16":
    libavcer_b200_fp16.so (the same kernels with IEEE half storage)."""
    lib = _libs.get(kind)
    if lib is not None:
        return lib
    path = LIB_PATH if kind == "bf16" else LIB_PATH_FP16
    if not os.path.exists(path):
        raise AvcerError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C avcer_b200/csrc`). avcer_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.avcer_storage_type().decode() != kind:
        raise AvcerError(f"{path} stores {lib.avcer_storage_type().decode()}, expected {kind}")
    _libs[kind] = lib
    return lib


def loaded():
    return list(_libs.values())


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc: int) -> None:
    if rc != 0:
        raise AvcerError(load().avcer_last_error().decode("utf-8", "replace"))


def require_device() -> None:
    check(load().avcer_device_check())
