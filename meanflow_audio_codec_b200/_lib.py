"""ctypes binding of ``libmfac.so`` (the C ABI declared in ``include/mfac.h``).

The product path has no fallback: if the shared library is missing or a CUDA device
is absent when a compute entry point is called, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmfac.so"


class MfacError(RuntimeError):
    pass


class MlpDims(C.Structure):
    _fields_ = [("D", C.c_int32), ("L", C.c_int32), ("C", C.c_int32), ("nb", C.c_int32)]


class ImfConfig(C.Structure):
    _fields_ = [
        ("noise_min", C.c_float), ("noise_max", C.c_float),
        ("time_mean", C.c_float), ("time_std", C.c_float),
        ("data_proportion", C.c_float), ("loss_c", C.c_float),
        ("use_weighted_loss", C.c_int32),
        ("seed", C.c_uint64), ("step", C.c_uint64), ("row_offset", C.c_uint64),
        ("step_dev", C.c_void_p),
        ("method", C.c_int32), ("gamma", C.c_float), ("uniform_time", C.c_int32),
        ("rows_r_equals_t", C.c_int64),
    ]


class AdamWConfig(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("lr", "b1", "b2", "eps", "weight_decay")]


GRAD_READY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64)


class ImfAux(C.Structure):
    _fields_ = ([(n, C.c_void_p) for n in ("v", "u", "dudt", "per_example", "e", "t", "r")] +
                [("grad_ready", GRAD_READY_FN), ("grad_ready_user", C.c_void_p)])


class Dense(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p)]


class MixerDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("D", "C", "nb", "tokens", "channels", "token_mix", "channel_mix", "latent_flat")]


class MixerBlockW(C.Structure):
    _fields_ = [(n, Dense) for n in ("input_proj", "adaln1", "tok1", "tok2", "adaln2", "ch1", "ch2", "output_proj")]


class MixerWeights(C.Structure):
    _fields_ = [("blocks", C.POINTER(MixerBlockW)), ("latent_proj", Dense)]


class ConvDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("D", "C", "nb", "S", "channels", "bottleneck", "latent_flat")]


class ConvBlockW(C.Structure):
    _fields_ = ([(n, Dense) for n in ("input_proj1", "input_proj2", "conditioning", "output_proj1", "output_proj2")] +
                [(n, C.c_void_p) for n in ("conv3_w", "conv3_b", "pw1_w", "pw1_b", "grn_gamma", "grn_beta", "pw2_w", "pw2_b",
                                           "layer_scale")])


class ConvWeights(C.Structure):
    _fields_ = [("blocks", C.POINTER(ConvBlockW)), ("latent_proj", Dense)]


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_F = C.c_float

# name -> (restype, argtypes); mirrors include/mfac.h one to one
PROTOTYPES = {
    "mfac_version": (C.c_int, []),
    "mfac_status_string": (C.c_char_p, [C.c_int]),
    "mfac_mdct_num_frames": (_I64, [_I64, _I32, _I32]),
    "mfac_imdct_length": (_I64, [_I64, _I32, _I32]),
    "mfac_mdct_f32": (C.c_int, [_P, _P, _I64, _I64, _I32, _I32, _P]),
    "mfac_imdct_f32": (C.c_int, [_P, _P, _I64, _I64, _I32, _I32, _P]),
    "mfac_mdct_strided_f32": (C.c_int, [_P, _I64, _I64, _P, _I64, _I64, _I64, _I64, _I32, _I32, _P]),
    "mfac_imdct_strided_f32": (C.c_int, [_P, _I64, _I64, _P, _I64, _I64, _I64, _I64, _I32, _I32, _P]),
    "mfac_mlp_param_count": (_I64, [C.POINTER(MlpDims)]),
    "mfac_mlp_param_offset": (C.c_int, [C.POINTER(MlpDims), _I32, _I32, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "mfac_mlp_shadow_bytes": (C.c_size_t, [C.POINTER(MlpDims)]),
    "mfac_mlp_cast_params": (C.c_int, [C.POINTER(MlpDims), _P, _P, _P]),
    "mfac_workspace_bytes": (C.c_size_t, [_I32, C.POINTER(MlpDims), _I64]),
    "mfac_mlp_encode": (C.c_int, [C.POINTER(MlpDims), _P, _P, _P, _P, _I64, _P, C.c_size_t, _P]),
    "mfac_mlp_forward": (C.c_int, [C.POINTER(MlpDims), _P, _P, _P, _P, _P, _P, _I64, _P, C.c_size_t, _P]),
    "mfac_imf_loss_grad": (C.c_int, [C.POINTER(MlpDims), C.POINTER(ImfConfig), _P, _P, _P, _P, _P, _P, _P, _P,
                                     C.POINTER(ImfAux), _I64, _P, C.c_size_t, _P]),
    "mfac_imf_train_step": (C.c_int, [C.POINTER(MlpDims), C.POINTER(ImfConfig), C.POINTER(AdamWConfig), _P, _P, _P, _P, _I64, _P, _P,
                                      _P, _P, _P, _P, _P, _P, C.POINTER(ImfAux), _I64, _I32, _P, C.c_size_t, _P]),
    "mfac_imf_loss_grad_audio": (C.c_int, [C.POINTER(MlpDims), C.POINTER(ImfConfig), _P, _P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P,
                                           C.POINTER(ImfAux), _I64, _P, C.c_size_t, _P]),
    "mfac_imf_train_step_audio": (C.c_int, [C.POINTER(MlpDims), C.POINTER(ImfConfig), C.POINTER(AdamWConfig), _P, _P, _P, _P, _I64, _P,
                                            _P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P, C.POINTER(ImfAux), _I64, _I32, _P,
                                            C.c_size_t, _P]),
    "mfac_adamw_step": (C.c_int, [C.POINTER(MlpDims), _P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _F, _P]),
    "mfac_adamw_step_dev": (C.c_int, [C.POINTER(MlpDims), _P, _P, _P, _P, _P, _P, _P, _F, _F, _F, _F, _F, _F, _P]),
    "mfac_sample": (C.c_int, [C.POINTER(MlpDims), _P, _P, _P, _P, _I32, _I32, _F, C.c_uint64, _P, _I64, _P,
                              C.c_size_t, _P]),
    "mfac_mixer_workspace_bytes": (C.c_size_t, [C.POINTER(MixerDims), _I64]),
    "mfac_mixer_forward": (C.c_int, [C.POINTER(MixerDims), C.POINTER(MixerWeights), _P, _P, _P, _P, _I64, _P, C.c_size_t, _P]),
    "mfac_conv_workspace_bytes": (C.c_size_t, [C.POINTER(ConvDims), _I64]),
    "mfac_conv_forward": (C.c_int, [C.POINTER(ConvDims), C.POINTER(ConvWeights), _P, _P, _P, _P, _I64, _P, C.c_size_t, _P]),
    "mfac_comm_unique_id": (C.c_int, [_P]),
    "mfac_comm_init": (C.c_int, [_P, _I32, _I32]),
    "mfac_comm_allreduce_sum_f32": (C.c_int, [_P, _I64, _P]),
    "mfac_comm_destroy": (C.c_int, []),
    "mfac_set_concurrency_max_rows": (C.c_int, [_I32]),
    "mfac_uses_concurrent_schedule": (C.c_int, [C.POINTER(MlpDims), _I64]),
    "mfac_debug_gemm_bf16": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _I32, _I32, _I32, _P]),
    "mfac_debug_set_simt_gemm": (C.c_int, [_I32]),
    "mfac_debug_set_pair_gemm": (C.c_int, [_I32]),
    "mfac_debug_set_stream_k": (C.c_int, [_I32]),
    "mfac_debug_phase_marks": (C.c_int, [_I32]),
    "mfac_debug_phase_collect": (C.c_int, [C.POINTER(_I32), C.POINTER(C.c_float), _I32]),
    "mfac_debug_counters": (C.c_int, [C.POINTER(_I64)]),
    "mfac_profile_enable": (C.c_int, [_I32]),
    "mfac_profile_collect": (C.c_int, [C.POINTER(_I64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

WS_FORWARD, WS_LOSS_GRAD, WS_SAMPLE = 0, 1, 2
LOSS_IMPROVED_MEAN_FLOW, LOSS_MEAN_FLOW, LOSS_FLOW_MATCHING = 0, 1, 2
SAMPLE_HEUN, SAMPLE_MF = 0, 1

_lib = None


def lib() -> C.CDLL:
    """Load ``libmfac.so`` (once).  Raises MfacError when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise MfacError(
                f"{LIB_PATH} not found: build it with `python -m meanflow_audio_codec_b200.build` "
                "(there is no CPU or PyTorch fallback)")
        l = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().mfac_status_string(int(status)).decode()
        raise MfacError(f"libmfac {what} failed: {msg} (status {status})")


layout_epoch = 0   # bumped whenever a knob that changes workspace layouts is turned (cached workspaces are keyed on it)


def set_concurrency_max_rows(rows: int) -> None:
    """Batches of up to ``rows`` rows run ``mfac_imf_loss_grad``'s concurrent schedule (include/mfac.h); 0 switches it off."""
    global layout_epoch
    check(lib().mfac_set_concurrency_max_rows(int(rows)), "set_concurrency_max_rows")
    layout_epoch += 1


def launches() -> int:
    n = _I64(0)
    lib().mfac_debug_counters(C.byref(n))
    return int(n.value)


PROF_FAMILIES = ("gemm_tcgen05", "mdct512", "imdct512", "adamw")


def profile_enable(on: bool) -> None:
    check(lib().mfac_profile_enable(1 if on else 0), "profile_enable")


def profile_collect() -> dict:
    """{family: {"launches", "ms", "work"}}; synchronises the device."""
    n = len(PROF_FAMILIES)
    launches, ms, work = (_I64 * n)(), (C.c_double * n)(), (C.c_double * n)()
    check(lib().mfac_profile_collect(launches, ms, work), "profile_collect")
    return {f: {"launches": int(launches[i]), "ms": float(ms[i]), "work": float(work[i])} for i, f in enumerate(PROF_FAMILIES)}


def require_cuda(t, name: str = "input"):
    import torch

    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor on a CUDA device, got {type(t)}")
    if not t.is_cuda:
        raise MfacError(f"{name} must live on a CUDA device: libmfac has no CPU path")
    return t


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
