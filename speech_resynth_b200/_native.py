"""ctypes binding of libsrb.so (the C ABI declared in include/srb.h).

There is no fallback: if the shared library is missing or was not built for the GPU in use, importing the
compute path raises.  torch is used only to own device memory and to name the current stream.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_void_p
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsrb.so")
if os.environ.get("SRB_DEBUG_LIB"):   # tools/trace_kernels.py: instrumented build of the same sources
    LIB_PATH = os.environ["SRB_DEBUG_LIB"]
# the same sources built with -DSRB_SPLIT: tight-precision mode (bf16 operands split into hi + lo, see include/srb.h)
TIGHT_LIB_PATH = os.path.join(_HERE, "libsrb_tight.so")

_P = c_void_p
_I = c_int32
_L = c_int64
_F = c_float

# name -> argument ctypes (everything returns int except srb_last_error)
_PROTOTYPES = {
    "srb_version": [],
    "srb_device_arch": [],
    "srb_embed_gather": [_P, _P, _P, _L, _I, _I, _P],
    "srb_unit_lengths": [_P, _P, _I, _I, _P],
    "srb_unit_extents": [_P, _P, _P, _I, _I, _P],
    "srb_stage_inputs": [_P, _P, _P, _P, _P, _P, _L, _P, _L, _I, _I, _I, _F, _I, _P],
    "srb_time_cond_table": [_P, _I, _P, _P, _P, _P, _I, _P, _P, _P],
    "srb_rotary_table": [_P, _I, _P, _P, _P],
    "srb_prior_prepare": [_P, _P, _L, _F, _I, _P],
    "srb_log_mel": [_P, _L, _I, _I, _P, _P, _P, _P, _P, _I, _P],
    "srb_duration_predict": [_P, _P, _F, _P, _P, _I, _I, _I, _P],
    "srb_length_regulate": [_P, _P, _P, _I, _I, _I, _I, _P],
    "srb_cfm_embed": [_P, _P, _P, _P, _I, _I, _P],
    "srb_cfm_posconv_norm": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "srb_cfm_qkv_rope": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "srb_cfm_qk_rope": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "srb_cfm_v_transposed": [_P, _P, _P, _L, _P],
    "srb_cfm_qk_rope_vt": [_P, _P, _P, _P, _P, _P, _L, _P, _P, _I, _I, _P],
    "srb_cfm_attention_tc": [_P, _I, _P, _L, _P, _P, _P, _I, _I, _P],
    "srb_cfm_attn_out_norm": [_P, _P, _P, _P, _P, _P, _I, _I, _P],
    "srb_cfm_ffn_glu": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "srb_cfm_ffn_out_norm": [_P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _P],
    "srb_cfm_pred_euler": [_P, _P, _F, _P, _P, _P, _P, _I, _F, _F, _F, _P, _I, _I, _P],
    "srb_hifigan_conv": [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P],
    "srb_hifigan_pair_fused": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "srb_hifigan_conv_res_act": [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _I, _I, _I, _I, _F, _F, _P],
    "srb_hifigan_upsample": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "srb_hifigan_mrf_fused": [_P, _P, _P, _P, _I, _I, _I, _F, _F, _P],
    "srb_hifigan_post": [_P, _P, _F, _P, _I, _I, _P, _P],
    "srb_cfm_attention_simt": [_P, _P, _P, _I, _I, _P],
    "srb_hifigan_mean3": [_P, _P, _P, _P, _L, _I, _F, _F, _P],
    "srb_split_factor": [],
    "srb_kmeans_assign": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P, _I, _P],
    "srb_split_bf16": [_P, _P, _L, _I, _P, _P],
    "srb_kmeans_scores_argmax": [_P, _P, _P, _P, _L, _I, _I, _P],
    "srb_kmeans_decode": [_P, _P, _L, _I, _P, _I, _P],
    "srb_unique_consecutive": [_P, _P, _P, _P, _P, _I, _I, _P],
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES) + ("srb_last_error", "srb_hifigan_mrf_phases")


class NativeLibraryError(RuntimeError):
    pass


_libs: dict = {}
launch_count = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)
# when set to a list, every call appends (name, shape_tag, start_event, end_event): per-op device timing for
# bench.py's roofline line and tools/profile_ops.py (never enabled on the product path)
profile_log: Optional[list] = None


def load(tight: bool = False) -> ctypes.CDLL:
    """Load libsrb.so (or, tight=True, libsrb_tight.so) and declare prototypes.  Raises NativeLibraryError when absent."""
    lib = _libs.get(tight)
    if lib is not None:
        return lib
    path = TIGHT_LIB_PATH if tight else LIB_PATH
    if not os.path.exists(path):
        raise NativeLibraryError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(speech_resynth_b200 has no CPU or PyTorch fallback)"
        )
    lib = ctypes.CDLL(path)
    for name, argtypes in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int32
    lib.srb_last_error.argtypes = []
    lib.srb_last_error.restype = c_char_p
    lib.srb_hifigan_mrf_phases.argtypes = [c_int32]
    lib.srb_hifigan_mrf_phases.restype = c_int32
    want = 3 if tight else 1
    if lib.srb_split_factor() != want:
        raise NativeLibraryError(f"{path} reports split factor {lib.srb_split_factor()}, expected {want}")
    _libs[tight] = lib
    return lib


def last_error(tight: bool = False) -> str:
    return load(tight).srb_last_error().decode("utf-8", "replace")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "native ops take contiguous CUDA tensors"
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args, flops: float = 0.0, nbytes: float = 0.0, tight: bool = False) -> None:
    """Invoke an entry point on the current torch stream; non-zero status -> RuntimeError with the C-side text.
    `flops` / `nbytes` are the op's algorithmic work (used only by the profiling log); tight=True calls the
    tight-precision build of the same entry point."""
    global launch_count
    lib = load(tight)
    if profile_log is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    rc = getattr(lib, name)(*args, stream_ptr())
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {last_error(tight)}")
    if profile_log is not None:
        ev1.record()
        profile_log.append((name, tuple(a for a in args if isinstance(a, int) and 0 <= a < (1 << 24)), ev0, ev1, flops, nbytes))
    launch_count += 1


def require_blackwell() -> None:
    lib = load()
    if not torch.cuda.is_available():
        raise NativeLibraryError("speech_resynth_b200 needs a CUDA device (sm_100a); none is visible")
    arch = lib.srb_device_arch()
    if arch != 100:
        raise NativeLibraryError(f"libsrb.so is built for sm_100a only; current device reports sm_{arch}")
