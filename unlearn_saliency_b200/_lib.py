"""ctypes binding of libsalun.so (the C ABI declared in include/salun.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  Nothing here (or anywhere in this package) imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsalun.so")
# Two builds of the same sources and the same C ABI (csrc/Makefile):
#   "bf16"  : activations / tensor-core operands stored as bf16 (the fast path)
#   "split" : every activation element is a (hi, lo) bf16 pair (16 significand bits) and every product runs as
#             hi*hi + hi*lo + lo*hi + lo*lo on the tensor cores: fp32-class results for the precision-critical
#             passes (saliency-mask generation), DESIGN.md section 4
LIB_PATHS = {"bf16": LIB_PATH, "split": os.path.join(_HERE, "csrc", "libsalun_split.so")}

_lib = None
_libs: dict = {}


class salun_resnet_cfg(C.Structure):
    _fields_ = [
        ("depth", C.c_int), ("num_classes", C.c_int), ("image_size", C.c_int), ("max_batch", C.c_int),
        ("mean", C.c_float * 3), ("std", C.c_float * 3), ("bn_eps", C.c_float), ("bn_momentum", C.c_float),
        ("imagenet_stem", C.c_int),
    ]


class salun_topk_info(C.Structure):
    _fields_ = [
        ("thr_key", C.c_uint32),
        ("thr_value", C.c_float),
        ("n_greater", C.c_int64),
        ("n_equal", C.c_int64),
    ]


_P = C.c_void_p
_I64 = C.c_int64
_F = C.c_float

# name -> argtypes (restype is always int unless listed in _RESTYPES)
_SIGNATURES = {
    "salun_version": [],
    "salun_launch_count": [],
    "salun_profile_begin": [],
    "salun_debug_role_timing": [_P],
    "salun_profile_end": [_P, _P, _P],
    "salun_last_error": [],
    "salun_ctx_create": [C.c_int, C.POINTER(_P)],
    "salun_ctx_destroy": [_P],
    "salun_saliency_accumulate": [_P, C.POINTER(_P), C.POINTER(_I64), C.c_int, _P, _P, _P],
    "salun_saliency_accumulate_flat": [_P, _P, _P, _I64, _P, _P],
    "salun_abs_inplace": [_P, _P, _I64, _P],
    "salun_topk_mask": [_P, _P, _I64, _I64, _P, _P, C.POINTER(salun_topk_info), _P],
    "salun_topk_mask_multi": [_P, _P, _I64, C.POINTER(C.c_int64), C.c_int, C.POINTER(_P), C.POINTER(_P),
                              C.POINTER(salun_topk_info), _P],
    "salun_pack_mask": [_P, _P, _I64, _P, _P],
    "salun_unpack_mask": [_P, _P, _I64, _P, _P],
    "salun_apply_mask": [_P, _P, _P, _I64, _P],
    "salun_masked_sgd_step": [_P, _P, _P, _P, _P, _I64, _F, _F, _F, _P],
    "salun_dp_shard": [_I64, C.c_int, C.c_int, C.POINTER(_I64), C.POINTER(_I64)],
    "salun_dp_masked_sgd_step": [_P, C.POINTER(_P), C.POINTER(_P), _P, _P, _I64, C.c_int, C.c_int, _F, _F, _F, _P],
    "salun_dp_grad_reduce_sumsq": [_P, C.POINTER(_P), _P, _P, _I64, C.c_int, C.c_int, _P],
    "salun_dp_masked_adam_step": [_P, C.POINTER(_P), C.POINTER(_P), _P, _P, _P, _P, _I64, C.c_int, C.c_int, _F, _F, _F, _F,
                                  _F, _I64, _F, _P, _P],
    "salun_grad_sumsq": [_P, _P, _I64, _P, _P],
    "salun_l1_penalty_grad": [_P, _P, _P, _I64, _F, _P, _P],
    "salun_clip_coef": [_P, _P, _F, _P, _P],
    "salun_augment_batch": [_P, _P, _I64, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P],
    "salun_eval_logits": [_P, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P],
    # op-level entry points (salun_ops.cu)
    "salun_act_bytes": [],
    "salun_wop_k": [],
    "salun_op_f32_to_act": [_P, _P, _I64, _P, _I64, _I64, _I64, _P],
    "salun_op_act_to_f32": [_P, _P, _I64, _P, _I64, _I64, _I64, _P],
    "salun_op_nchw_to_padded": [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P],
    "salun_op_padded_to_nchw": [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P],
    "salun_op_rows_to_nchw": [_P, _P, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P],
    "salun_op_prep_weight": [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P],
    "salun_op_conv": [_P, _P, C.c_int, _P, _P, _P, C.c_int, _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                      C.c_int, _P],
    "salun_op_conv_s2": [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P],
    "salun_op_groupnorm": [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _F, C.c_int, _P],
    "salun_op_upsample2": [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P],
    "salun_op_concat": [_P, _P, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P],
    "salun_op_linear_f32": [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P],
    "salun_sd_timestep_embedding": [_P, _P, _P, C.c_int, C.c_int, _F, _P],
    "salun_sd_layernorm": [_P, _P, _P, _P, _P, _I64, C.c_int, _F, _P],
    "salun_sd_geglu": [_P, _P, _P, _I64, C.c_int, _P],
    "salun_sd_attention_ws_bytes": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int],
    "salun_op_groupnorm_ws_floats": [C.c_int],
    "salun_op_set_scratch": [_P, _P, _I64],
    "salun_sd_attention": [_P, _P, _I64, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P],
    "salun_sd_attention_ld": [_P, _P, _I64, _P, C.c_int, _P, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_int, _P],
    "salun_ddim_step": [_P, _P, _P, _P, _P, _P, _P, _F, _F, C.c_int, C.c_int, _P, _P, _P],
    "salun_masked_adam_step": [_P, _P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _I64, _P, _P],
    # tcgen05 GEMM / convolution entry points (salun_gemm.cu)
    "salun_gemm_bf16_tn": [_P, _P, _P, _P, _P, _I64, _I64, _I64, _P],
    "salun_gemm2_bf16_tn": [_P, _P, _P, _P, _P, _I64, _I64, _I64, _P],
    "salun_conv_fwd_bf16": [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P],
    "salun_conv_rw_fwd_bf16": [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P],
    "salun_conv_wgrad_bf16": [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P],
    # ResNet engine (salun_resnet.cu)
    "salun_resnet_param_count": [_P],
    "salun_resnet_bn_channels": [_P],
    "salun_resnet_create": [_P, _P, _P, _P, _P, _P, C.POINTER(_P)],
    "salun_resnet_destroy": [_P],
    "salun_resnet_forward_backward": [_P, _P, _P, C.c_int, C.c_int, _F, _P, _P, _P],
    "salun_resnet_forward": [_P, _P, C.c_int, _P, _P],
    "salun_resnet_syncbn_doubles": [_P],
    "salun_resnet_enable_syncbn": [_P, C.POINTER(_P), C.POINTER(_P), C.c_int, C.c_int],
}
_RESTYPES = {"salun_last_error": C.c_char_p, "salun_launch_count": C.c_longlong, "salun_resnet_param_count": C.c_int64,
             "salun_resnet_bn_channels": C.c_int64, "salun_resnet_syncbn_doubles": C.c_int64,
             "salun_sd_attention_ws_bytes": C.c_int64, "salun_op_groupnorm_ws_floats": C.c_int64}


def exported_symbols():
    """Every symbol include/salun.h declares (used by the CPU-side ABI test)."""
    return sorted(set(_SIGNATURES) | set(_EXTRA_SIGNATURES))


_EXTRA_SIGNATURES: dict = {}


def register_signatures(sigs: dict, restypes: dict | None = None):
    """Other modules of the package (gemm / resnet engine) register their entry points here."""
    _EXTRA_SIGNATURES.update(sigs)
    if restypes:
        _RESTYPES.update(restypes)
    for l in _libs.values():
        _bind(l, sigs)


def _bind(lib, sigs):
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it: fail loudly
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)


def available_precisions():
    """precision modes whose library has been built"""
    return [k for k, p in LIB_PATHS.items() if os.path.exists(p)]


def lib(precision: str = "bf16"):
    """The library of one precision mode.  The two builds export the same symbols: they are loaded RTLD_LOCAL (and
    linked -Bsymbolic), so each keeps its own kernels; context handles (plain device workspaces) are interchangeable."""
    global _lib
    l = _libs.get(precision)
    if l is None:
        if precision not in LIB_PATHS:
            raise ValueError(f"unknown precision {precision!r} (known: {sorted(LIB_PATHS)})")
        path = LIB_PATHS[precision]
        if not os.path.exists(path):
            raise RuntimeError(
                f"{os.path.basename(path)} not found at {path}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C unlearn_saliency_b200/csrc`). There is no CPU / PyTorch fallback."
            )
        l = C.CDLL(path, mode=C.RTLD_LOCAL)
        _bind(l, _SIGNATURES)
        _bind(l, _EXTRA_SIGNATURES)
        _libs[precision] = l
        if precision == "bf16":
            _lib = l
    return l


def check(rc: int, what: str = "", lib_=None):
    if rc != 0:
        msgs = []
        for l in ([lib_] if lib_ is not None else list(_libs.values())):
            m = l.salun_last_error()
            if m:
                msgs.append(m.decode())
        raise RuntimeError(f"libsalun {what} failed (status {rc}): {' | '.join(msgs) if msgs else '?'}")
