"""ctypes binding of libttt_b200.so (C-ABI in include/ttt_b200.h).  No fallback: if the library is missing or a call
fails this raises -- the product path never routes through torch eager or the oracle."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TTT_B200_LIB") or os.path.join(_HERE, "lib", "libttt_b200.so")
_lib = None

_vp, _fp, _i = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int

_SIGS = {
    "ttt_b200_version": ([], ctypes.c_int),
    "ttt_b200_last_error": ([], ctypes.c_char_p),
    "ttt_b200_mlp_forward": ([_vp] * 4 + [_fp] * 2 + [_fp] * 4 + [_fp] * 4 + [_fp] * 4 + [_vp] + [_i] * 4 + [_vp], ctypes.c_int),
    "ttt_b200_mlp_backward_workspace_bytes": ([_i, _i, _i], ctypes.c_size_t),
    "ttt_b200_mlp_backward": ([_vp] * 4 + [_fp] * 2 + [_fp] * 4 + [_vp] + [_fp] * 6 + [_vp] * 4 + [_vp, ctypes.c_size_t] + [_i] * 4 + [_vp], ctypes.c_int),
    "ttt_b200_mlp_backward_seeded": ([_vp] * 4 + [_fp] * 2 + [_fp] * 4 + [_vp] + [_fp] * 4 + [_fp] * 6 + [_vp] * 4 + [_vp, ctypes.c_size_t]
                                     + [_i] * 4 + [_vp], ctypes.c_int),
    "ttt_b200_linear_forward": ([_vp] * 4 + [_fp] * 2 + [_fp] * 2 + [_fp] * 4 + [_vp] + [_i] * 4 + [_vp], ctypes.c_int),
    "ttt_b200_linear_backward_workspace_bytes": ([_i] * 4, ctypes.c_size_t),
    "ttt_b200_linear_backward": ([_vp] * 4 + [_fp] * 4 + [_vp] + [_fp] * 5 + [_vp] * 4 + [ctypes.c_size_t] + [_i] * 4 + [_vp],
                                 ctypes.c_int),
    "ttt_b200_attention_forward": ([_vp] * 4 + [_i] * 3 + [ctypes.c_float, _vp], ctypes.c_int),
    "ttt_b200_attention_forward_lse": ([_vp] * 4 + [_fp] + [_i] * 3 + [ctypes.c_float, _vp], ctypes.c_int),
    "ttt_b200_attention_backward": ([_vp] * 5 + [_fp] * 2 + [_vp] * 3 + [_i] * 3 + [ctypes.c_float, _vp], ctypes.c_int),
    "ttt_b200_process_input": ([_vp] * 3 + [_fp] * 5 + [_vp] + [_vp] * 4 + [_i] * 5 + [ctypes.c_float, _vp], ctypes.c_int),
    "ttt_b200_process_input_backward": ([_vp] * 3 + [_fp] * 4 + [_vp] + [_vp] * 3 + [_fp] + [_vp] * 3 + [_fp] * 3 + [_i] * 5
                                        + [ctypes.c_float, _vp], ctypes.c_int),
    "ttt_b200_output_norm": ([_vp, _fp, _fp, _vp, _vp] + [_i] * 3 + [ctypes.c_float, _vp], ctypes.c_int),
    "ttt_b200_output_norm_backward": ([_vp, _fp, _vp, _vp, _vp, _fp, _fp] + [_i] * 3 + [ctypes.c_float, _vp], ctypes.c_int),
    "ttt_b200_gate_forward": ([_vp, _vp, _fp, _fp, _vp, _vp] + [_i] * 6 + [_vp], ctypes.c_int),
    "ttt_b200_gate_backward": ([_vp, _vp, _vp, _fp, _fp, _vp, _vp, _fp, _fp] + [_i] * 6 + [_vp], ctypes.c_int),
    "ttt_b200_qk_norm_rope": ([_vp, _vp, _fp, _fp, _fp, _fp, _vp, _vp] + [_i] * 4 + [ctypes.c_float, _vp], ctypes.c_int),
    "ttt_b200_qk_norm_rope_backward": ([_vp, _vp, _fp, _fp, _fp, _vp, _vp, _vp, _vp, _fp, _fp] + [_i] * 4 + [ctypes.c_float, _vp],
                                       ctypes.c_int),
    "ttt_b200_ln_affine": ([_vp, _fp, _fp, _vp] + [_i] * 4 + [ctypes.c_float, _vp], ctypes.c_int),
    "ttt_b200_ln_affine_backward": ([_vp, _fp, _vp, _vp, _fp, _fp] + [_i] * 4 + [ctypes.c_float, _vp], ctypes.c_int),
    "ttt_b200_gate_add": ([_vp, _vp, _fp, _vp] + [_i] * 4 + [_vp], ctypes.c_int),
    "ttt_b200_gate_add_backward": ([_vp, _vp, _fp, _vp, _fp] + [_i] * 4 + [_vp], ctypes.c_int),
}

# development probes (include/ttt_b200_debug.h) live in their own library; the production .so exports none of them
DEBUG_LIB_PATH = os.environ.get("TTT_B200_SELFTEST_LIB") or os.path.join(_HERE, "lib", "libttt_b200_selftest.so")
_DEBUG_SIGS = {
    "ttt_b200_debug_last_error": ([], ctypes.c_char_p),
    "ttt_b200_debug_umma": ([_i, _vp, _vp, _fp, _i, _i, _vp], ctypes.c_int),
    "ttt_b200_debug_dsmem": ([_i, _i, _i, _fp, _vp], ctypes.c_int),
    "ttt_b200_debug_spin": ([_i, _i, ctypes.c_longlong, _i, _i, _fp, ctypes.c_longlong, _vp], ctypes.c_int),
}
_debug_lib = None


class TTTB200Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TTTB200Error(
                f"{LIB_PATH} not found: build it with `python ttt_video_dit_b200/build.py` "
                "(there is no CPU / eager fallback for this path)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (args, res) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError if the ABI symbol is missing
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def debug_lib():
    """libttt_b200_selftest.so (UMMA descriptor self-test, interference spin kernel)."""
    global _debug_lib
    if _debug_lib is None:
        if not os.path.exists(DEBUG_LIB_PATH):
            raise TTTB200Error(f"{DEBUG_LIB_PATH} not found: build it with `python ttt_video_dit_b200/build.py`")
        L = ctypes.CDLL(DEBUG_LIB_PATH)
        for name, (args, res) in _DEBUG_SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _debug_lib = L
    return _debug_lib


def set_timing_buffer(buf):
    """Phase-timing builds only (TTT_B200_LIB=.../libttt_b200_dbg.so): returns 1 if the loaded library records timings."""
    fn = getattr(lib(), "ttt_b200_debug_set_timing_buffer", None)
    if fn is None:
        return 0
    fn.argtypes, fn.restype = [_vp], ctypes.c_int
    return fn(ptr(buf))


def exported_symbols():
    return sorted(_SIGS)


def debug_exported_symbols():
    return sorted(_DEBUG_SIGS)


def check(code, what):
    if code != 0:
        msg = lib().ttt_b200_last_error()
        raise TTTB200Error(f"{what} failed with code {code}: {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream(t=None):
    """cudaStream_t of torch's current stream on the device that owns tensor ``t`` (default: the current device).  The
    stream must belong to the tensor's device: torch's "current stream" is per device."""
    import torch
    dev = t.device if t is not None else None
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
