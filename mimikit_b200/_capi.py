"""ctypes binding of the C ABI declared in include/mmk_b200.h.

There is no CPU fallback: if the CUDA library is missing this module raises on first use, and every compute
entry point fails when no CUDA device is present.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int64, c_size_t, c_ubyte, c_ulonglong,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libmmk_b200.so")
_lib = None


class MmkError(RuntimeError):
    pass


class DeviceInfo(Structure):
    _fields_ = [("sm_count", c_int), ("cc_major", c_int), ("cc_minor", c_int), ("max_smem_optin", c_int),
                ("l2_bytes", c_int)]


class LaunchInfo(Structure):
    _fields_ = [("cluster_size", c_int), ("n_stages", c_int), ("group_size", c_int), ("threads", c_int),
                ("smem_bytes", c_int), ("sm_used", c_int)]


_fpp = POINTER(POINTER(c_float))


class WaveNetDesc(Structure):
    _fields_ = [("n_layers", c_int), ("dilated_dim", c_int), ("skips_dim", c_int), ("head_hidden", c_int),
                ("q_levels", c_int), ("min_temperature", c_float), ("dilations", POINTER(c_int)),
                ("embedding", POINTER(c_float)),
                ("conv_dil_w", _fpp), ("conv_dil_b", _fpp), ("conv_skip_w", _fpp), ("conv_skip_b", _fpp),
                ("conv_res_w", _fpp), ("conv_res_b", _fpp),
                ("head_w1", POINTER(c_float)), ("head_b1", POINTER(c_float)),
                ("head_w2", POINTER(c_float)), ("head_b2", POINTER(c_float))]


class WaveNetDescEx(Structure):
    _fields_ = [("base", WaveNetDesc), ("kernel_sizes", POINTER(c_int)), ("layerwise_inputs", c_int),
                ("head_hidden_layers", c_int), ("head_wh", POINTER(c_float)), ("head_bh", POINTER(c_float)),
                ("aff_res_w", _fpp), ("aff_res_b", _fpp), ("act_f", c_int), ("act_g", c_int)]


class SampleRNNDesc(Structure):
    _fields_ = [("n_tiers", c_int), ("frame_sizes", POINTER(c_int)), ("hidden_dim", c_int), ("head_hidden", c_int),
                ("q_levels", c_int), ("min_temperature", c_float),
                ("in_w", _fpp), ("in_b", _fpp), ("w_ih", _fpp), ("w_hh", _fpp), ("b_ih", _fpp), ("b_hh", _fpp),
                ("up_w", _fpp), ("up_b", _fpp),
                ("conv_w", POINTER(c_float)), ("conv_b", POINTER(c_float)),
                ("head_w1", POINTER(c_float)), ("head_b1", POINTER(c_float)),
                ("head_w2", POINTER(c_float)), ("head_b2", POINTER(c_float))]


class SampleRNNDescEx(Structure):
    _fields_ = [("base", SampleRNNDesc), ("rnn_type", c_int), ("n_rnn", c_int),
                ("w_ih", _fpp), ("w_hh", _fpp), ("b_ih", _fpp), ("b_hh", _fpp),
                ("head_hidden_layers", c_int), ("head_wh", POINTER(c_float)), ("head_bh", POINTER(c_float)),
                ("need_set_hidden", c_int), ("compute_mode", c_int)]


# name -> (restype, argtypes); every symbol include/mmk_b200.h declares
PROTOTYPES = {
    "mmk_abi_version": (c_int, []),
    "mmk_last_error": (c_char_p, []),
    "mmk_get_device_info": (c_int, [POINTER(DeviceInfo)]),
    "mmk_mulaw_compress": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_float, c_void_p]),
    "mmk_mulaw_compress_u8": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_float, c_void_p]),
    "mmk_mulaw_prepare": (c_int, [c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "mmk_remove_dc_scratch_bytes": (c_size_t, [c_int64, c_int64]),
    "mmk_remove_dc": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "mmk_remove_dc_recomputed_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "mmk_normalize_inf": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "mmk_normalize_mulaw_compress": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_float,
                                             c_void_p]),
    "mmk_mulaw_expand": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_float, c_void_p]),
    "mmk_stft_mag_mel": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                 c_int, c_void_p, c_void_p]),
    "mmk_stft_n_frames": (c_int, [c_int64, c_int, c_int, c_int, c_int, POINTER(c_int64), POINTER(c_int64)]),
    "mmk_mel_filterbank": (c_int, [c_int, c_int, c_float, c_float, c_int, c_void_p]),
    "mmk_mel_apply": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "mmk_wavenet_create": (c_int, [POINTER(WaveNetDesc), c_int, POINTER(c_void_p)]),
    "mmk_wavenet_create_ex": (c_int, [POINTER(WaveNetDesc), c_int, c_int, POINTER(c_void_p)]),
    "mmk_wavenet_create_cfg": (c_int, [POINTER(WaveNetDescEx), c_int, c_int, POINTER(c_void_p)]),
    "mmk_tc_gemm_check": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "mmk_wavenet_destroy": (c_int, [c_void_p]),
    "mmk_wavenet_rf": (c_int, [c_void_p]),
    "mmk_wavenet_sync_check": (c_int, [c_void_p, c_void_p]),
    "mmk_wavenet_launch_info": (c_int, [c_void_p, POINTER(LaunchInfo)]),
    "mmk_wavenet_run": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p, c_int,
                                c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mmk_wavenet_generate": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_int, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "mmk_samplernn_create": (c_int, [POINTER(SampleRNNDesc), c_int, POINTER(c_void_p)]),
    "mmk_samplernn_create_ex": (c_int, [POINTER(SampleRNNDescEx), c_int, POINTER(c_void_p)]),
    "mmk_samplernn_set_hidden": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "mmk_samplernn_destroy": (c_int, [c_void_p]),
    "mmk_samplernn_launch_info": (c_int, [c_void_p, POINTER(LaunchInfo)]),
    "mmk_samplernn_sync_check": (c_int, [c_void_p, c_void_p]),
    "mmk_samplernn_run": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                  c_int64, c_int, c_int, c_void_p, c_int, c_void_p, c_int64, c_int64, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "mmk_samplernn_generate": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_int,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
}


def lib():
    """Loads libmmk_b200.so (built by mimikit_b200.build / __graft_entry__.build()); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MmkError(f"CUDA library not built: {LIB_PATH} is missing. Run `python -m mimikit_b200.build` "
                           f"(needs nvcc). There is no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError here = header and library disagree
            fn.restype, fn.argtypes = res, args
        if handle.mmk_abi_version() != 1:
            raise MmkError("libmmk_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().mmk_last_error()
        raise MmkError(msg.decode() if msg else f"mmk error {rc}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise MmkError("mimikit_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def fptr(t):
    """float* to a contiguous fp32 CPU tensor / numpy array (kept alive by the caller)."""
    return ctypes.cast(c_void_p(t.data_ptr() if hasattr(t, "data_ptr") else t.ctypes.data), POINTER(c_float))


def device_info():
    require_cuda()
    info = DeviceInfo()
    check(lib().mmk_get_device_info(ctypes.byref(info)))
    return info
