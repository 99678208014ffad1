"""ctypes binding of libsyntalker_b200.so (include/syntalker_b200.h).

The library is the product: if it is missing this module raises at import of any compute entry point --
there is no CPU or torch fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsyntalker_b200.so")

ST_VARIANT = {"beatx": 0, "beatx_motionclip": 1, "h3d": 2}
ST_ENGINE_SIMT, ST_ENGINE_TC = 0, 1
ST_MODE_DDPM, ST_MODE_DDIM = 0, 1
ST_CFG_NONE, ST_CFG_TEXT, ST_CFG_TWO, ST_CFG_BODYPART, ST_CFG_BODYPART1 = 0, 1, 2, 3, 4
ST_FLAG_UNCOND, ST_FLAG_UNCOND_AUDIO = 1, 2
ST_COEF_STRIDE = 5

# every symbol include/syntalker_b200.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "st_last_error", "st_abi_version", "st_launch_count", "st_set_engine", "st_get_engine", "st_model_set_engine", "st_vq_set_engine", "st_set_graphs", "st_set_pdl", "st_debug_timeline", "st_debug_timeline_select", "st_debug_timeline_select2", "st_debug_trace", "st_debug_probe", "st_debug_cond_taps",
    "st_model_create", "st_model_destroy", "st_vq_create", "st_vq_destroy", "st_vq_out_dim",
    "st_schedule_create", "st_schedule_destroy", "st_cond_encode", "st_denoise", "st_sample", "st_sample_chunk", "st_sample_begin", "st_sample_run", "st_sample_end",
    "st_rvq_decode", "st_pose_assemble_330", "st_pose_assemble_623", "st_sample_to_tokens", "st_pose_330_to_aa165", "st_moments_accumulate", "st_l1div_accumulate",
    "st_rvq_encode", "st_generate_330", "st_generate_330_host", "st_generate_330_host_begin", "st_generate_330_host_wait", "st_generate_long_330", "st_selftest_gemm", "st_selftest_conv", "st_bench_gemm", "st_profile_begin", "st_profile_end",
]


class StTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


class StCond(C.Structure):
    _fields_ = [("audio", C.c_void_p), ("word", C.c_void_p), ("seed", C.c_void_p), ("style", C.c_void_p * 3)]


class StGuidance(C.Structure):
    _fields_ = [("mode", C.c_int), ("flags", C.c_int), ("scale", C.c_void_p), ("scale2", C.c_void_p),
                ("audio_scale", C.c_float), ("prompt_scale", C.c_float)]


class StHostInputs(C.Structure):
    _fields_ = [("audio", C.c_void_p), ("word", C.c_void_p), ("seed", C.c_void_p), ("style", C.c_void_p * 3),
                ("x_init", C.c_void_p), ("noise_tape", C.c_void_p), ("jaw_aa", C.c_void_p),
                ("mean", C.c_void_p), ("std", C.c_void_p), ("trans_mean", C.c_void_p), ("trans_std", C.c_void_p)]


class StError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(or syntalker_b200/csrc/build.sh). There is no fallback path.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    L.st_last_error.restype = C.c_char_p
    L.st_abi_version.restype = i32
    L.st_launch_count.restype = i64
    L.st_set_engine.argtypes = [i32]
    L.st_get_engine.restype = i32
    L.st_model_set_engine.argtypes = [vp, i32]
    L.st_vq_set_engine.argtypes = [vp, i32]
    L.st_set_graphs.argtypes = [i32]
    L.st_set_pdl.argtypes = [i32]
    L.st_debug_timeline.argtypes = [vp]
    L.st_debug_timeline_select.argtypes = [i32, i32]
    L.st_debug_timeline_select2.argtypes = [i32, i32]
    L.st_debug_trace.argtypes = [vp]
    L.st_debug_cond_taps.argtypes = [vp, vp, vp, i32, vp]
    L.st_model_create.argtypes = [C.POINTER(StTensor), i32, i32, C.POINTER(vp)]
    L.st_model_destroy.argtypes = [vp]
    L.st_model_destroy.restype = None
    L.st_vq_create.argtypes = [C.POINTER(StTensor), i32, i32, C.POINTER(vp)]
    L.st_vq_destroy.argtypes = [vp]
    L.st_vq_destroy.restype = None
    L.st_vq_out_dim.argtypes = [vp]
    L.st_schedule_create.argtypes = [i32, i32, vp, vp, C.POINTER(vp)]
    L.st_schedule_destroy.argtypes = [vp]
    L.st_schedule_destroy.restype = None
    L.st_cond_encode.argtypes = [vp, C.POINTER(StCond), i32, vp]
    L.st_denoise.argtypes = [vp, vp, vp, C.POINTER(StGuidance), vp, i32, vp]
    L.st_sample.argtypes = [vp, vp, C.POINTER(StGuidance), vp, vp, i32, vp, vp]
    L.st_sample_chunk.argtypes = [vp]
    L.st_sample_begin.argtypes = [vp, vp, C.POINTER(StGuidance), vp, i32, vp]
    L.st_sample_run.argtypes = [vp, i32, vp, vp]
    L.st_sample_end.argtypes = [vp, vp, vp]
    L.st_rvq_decode.argtypes = [vp, vp, i64, f32, i32, i32, vp, vp, vp, vp]
    L.st_rvq_encode.argtypes = [vp, vp, i32, i32, vp, vp]
    L.st_generate_long_330.argtypes = [vp, vp, C.POINTER(StGuidance), vp, vp, vp, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp, i32, i32, f32, vp, vp, vp, vp]
    L.st_pose_assemble_330.argtypes = [vp] * 8 + [i32, i32, vp, vp, vp]
    L.st_pose_assemble_623.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    L.st_sample_to_tokens.argtypes = [vp, i32, i32, f32, vp, vp]
    L.st_pose_330_to_aa165.argtypes = [vp, i64, vp, vp]
    L.st_moments_accumulate.argtypes = [vp, i64, i32, vp, vp]
    L.st_l1div_accumulate.argtypes = [vp, i32, i32, vp, vp]
    L.st_generate_330.argtypes = [vp, vp, C.POINTER(StGuidance), vp, vp, vp, C.POINTER(StCond), vp, vp, vp, vp, i32, f32, vp, vp, vp, vp]
    L.st_generate_330_host.argtypes = [vp, vp, C.POINTER(StGuidance), vp, vp, vp, C.POINTER(StHostInputs), i32, f32,
                                       vp, vp, vp, vp]
    L.st_generate_330_host_begin.argtypes = [vp, vp, C.POINTER(StGuidance), vp, vp, vp, C.POINTER(StHostInputs), i32, f32,
                                             vp, vp, vp, i32, vp]
    L.st_generate_330_host_wait.argtypes = [vp, i32]
    L.st_profile_end.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64)]
    L.st_bench_gemm.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, C.POINTER(C.c_double)]
    L.st_selftest_gemm.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp]
    L.st_selftest_conv.argtypes = [i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    if L.st_abi_version() != 1:
        raise StError("libsyntalker_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise StError(f"syntalker_b200 error {rc}: {lib().st_last_error().decode()}")


def tensor_array(named):
    """dict name -> contiguous fp32 CPU torch tensor  ==>  (StTensor array, keepalive list)."""
    import torch
    keep, arr = [], (StTensor * len(named))()
    for i, (k, v) in enumerate(named.items()):
        v = v.detach().to(torch.float32).contiguous().cpu()
        keep.append(v)
        arr[i] = StTensor(k.encode(), v.data_ptr(), v.numel())
    return arr, keep


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def set_engine(name: str):
    check(lib().st_set_engine({"simt": ST_ENGINE_SIMT, "tc": ST_ENGINE_TC}[name]))


def get_engine() -> str:
    return "tc" if lib().st_get_engine() == ST_ENGINE_TC else "simt"


def launch_count() -> int:
    return int(lib().st_launch_count())


def profile_begin():
    check(lib().st_profile_begin())


def profile_end():
    ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
    check(lib().st_profile_end(C.byref(ms), C.byref(fl), C.byref(n)))
    return ms.value, fl.value, n.value
