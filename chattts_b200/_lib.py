"""ctypes binding of ``libchattts_b200.so`` (the C ABI in include/chattts_b200.h).

There is deliberately no fallback: if the library is missing or no CUDA device is present
the product path raises (north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib = None


class GptConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "hidden_size", "intermediate_size", "num_layers", "num_heads", "num_kv_heads", "head_dim", "num_vq",
        "num_audio_tokens", "num_text_tokens", "max_positions")] + [
        ("rms_eps", C.c_float), ("max_batch", C.c_int32), ("max_context", C.c_int32)]


class GptLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "layer0", "layer_stride", "wqkv", "wo", "wgate_up", "wdown", "ln1", "ln2", "final_norm", "head_code",
        "head_text", "emb_code", "emb_text", "rope_cos", "rope_sin", "total")]


class SamplerConfig(C.Structure):
    _fields_ = [
        ("temperature", C.c_float * 8), ("top_p", C.c_float), ("top_k", C.c_int32),
        ("min_tokens_to_keep", C.c_int32), ("penalty_on", C.c_int32), ("penalty_lut", C.c_float * 32),
        ("past_window", C.c_int32), ("penalty_max_ids", C.c_int32), ("greedy", C.c_int32),
        ("eos_token", C.c_int32), ("min_new_token", C.c_int32), ("top_p_removed_max", C.c_float),
        ("has_removed_max", C.c_int32), ("philox_seed", C.c_uint64)]


class GptStatus(C.Structure):
    _fields_ = [("steps_done", C.c_int32), ("all_finished", C.c_int32),
                ("any_finished_first_step", C.c_int32), ("reserved", C.c_int32)]


class ConvStackConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "idim", "odim", "hidden", "n_layer", "bn_dim", "kernel", "dilation", "out_dim", "vq_dim", "vq_groups",
        "vq_residual", "vq_levels")]


class VocosConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("input_channels", "dim", "intermediate_dim", "num_layers", "n_fft",
                                         "hop_length")]


#: every symbol include/chattts_b200.h declares (checked by tests/test_abi.py)
EXPORTS = (
    "ctb_abi_version", "ctb_last_error", "ctb_launch_count", "ctb_gpt_layout_query", "ctb_gpt_create",
    "ctb_gpt_destroy", "ctb_gpt_begin", "ctb_gpt_decode", "ctb_gpt_status_query", "ctb_gpt_profile_kernel", "ctb_gpt_debug_trace", "ctb_gpt_embed_prompt", "ctb_sample",
    "ctb_dvae_blob_floats", "ctb_vocos_blob_floats", "ctb_decoder_create", "ctb_decoder_destroy",
    "ctb_dvae_decode", "ctb_vocos_decode",
    "ctb_dvae_encoder_blob_floats", "ctb_dvae_encoder_create", "ctb_dvae_encoder_destroy", "ctb_dvae_encode",
)


class CtbError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building in-tree first if needed) and type the C ABI."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if build_if_missing:
            # build() is a no-op when the digest of csrc/ + include/ matches the stamp beside the library, so an
            # edited kernel is never silently ignored; a box without nvcc keeps using the library that travelled
            try:
                _build.build()
            except Exception:
                if not os.path.exists(path):
                    raise
        if not os.path.exists(path):
            raise CtbError(f"{path} missing: run `python -m chattts_b200.build`")
        lib = C.CDLL(path)
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        lib.ctb_abi_version.restype = C.c_int
        lib.ctb_last_error.restype = C.c_char_p
        lib.ctb_launch_count.restype = C.c_uint64
        lib.ctb_gpt_layout_query.argtypes = [C.POINTER(GptConfig), C.POINTER(GptLayout)]
        lib.ctb_gpt_create.argtypes = [C.POINTER(GptConfig), vp, C.POINTER(vp)]
        lib.ctb_gpt_destroy.argtypes = [vp]
        lib.ctb_gpt_begin.argtypes = [vp, i32, i32, vp, vp, C.POINTER(SamplerConfig), vp, i32, i32, vp, vp, vp]
        lib.ctb_gpt_decode.argtypes = [vp, i32, vp]
        lib.ctb_gpt_status_query.argtypes = [vp, C.POINTER(GptStatus), vp, vp, vp]
        lib.ctb_gpt_profile_kernel.argtypes = [vp, i32, vp]
        lib.ctb_gpt_debug_trace.argtypes = [vp, vp, i32]
        lib.ctb_gpt_embed_prompt.argtypes = [vp, vp, vp, i32, i32, vp, vp]
        lib.ctb_sample.argtypes = [vp, i32, i32, i32, C.POINTER(SamplerConfig), vp, vp, i32, i32, i32, vp, vp]
        lib.ctb_dvae_blob_floats.argtypes = [C.POINTER(ConvStackConfig)]
        lib.ctb_dvae_blob_floats.restype = i64
        lib.ctb_vocos_blob_floats.argtypes = [C.POINTER(VocosConfig)]
        lib.ctb_vocos_blob_floats.restype = i64
        lib.ctb_decoder_create.argtypes = [C.POINTER(ConvStackConfig), vp, C.POINTER(VocosConfig), vp, i32, i32,
                                           C.POINTER(vp)]
        lib.ctb_decoder_destroy.argtypes = [vp]
        lib.ctb_dvae_decode.argtypes = [vp, vp, i32, i32, i32, vp, vp]
        lib.ctb_vocos_decode.argtypes = [vp, vp, i32, i32, vp, vp]
        lib.ctb_dvae_encoder_blob_floats.argtypes = [C.POINTER(ConvStackConfig)]
        lib.ctb_dvae_encoder_blob_floats.restype = i64
        lib.ctb_dvae_encoder_create.argtypes = [C.POINTER(ConvStackConfig), vp, i64, C.POINTER(vp)]
        lib.ctb_dvae_encoder_destroy.argtypes = [vp]
        lib.ctb_dvae_encode.argtypes = [vp, vp, i64, vp, i32, C.POINTER(i32), vp, vp, vp]
        if lib.ctb_abi_version() != 3:
            raise CtbError("ABI version mismatch")
        _lib = lib
        return lib


def check(rc: int) -> None:
    if rc != 0:
        raise CtbError(f"chattts_b200 error {rc}: {load().ctb_last_error().decode()}")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise CtbError("chattts_b200 needs a CUDA device (sm_100a); there is no CPU path")
