"""ctypes binding of the C-ABI library ``libsom_b200.so`` (include/som_b200.h).

The library is built in-tree by :func:`build` (``nvcc -gencode arch=compute_100a,code=sm_100a``);
there is no JIT cache and no fallback: if the shared object is missing or a call fails the error is
raised to the caller.  PyTorch only supplies device pointers and the current stream.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
# (SOM_B200_LIB: an alternative build of the same sources, for A/B measurements of kernel variants)
LIB_PATH = os.environ.get("SOM_B200_LIB") or os.path.join(_HERE, "libsom_b200.so")
_SRC = [os.path.join(_HERE, "csrc", "som_b200.cu")]
_DEPS = _SRC + [os.path.join(_HERE, "csrc", "som_gemm.cuh"), os.path.join(_ROOT, "include", "som_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]

# name -> (restype, argtypes); mirrors include/som_b200.h one to one.
_P = c_void_p
SIGNATURES = {
    "som_b200_abi_version": (c_int, []),
    "som_last_error": (c_char_p, []),
    "som_launch_count": (c_int64, []),
    "som_launch_count_reset": (None, []),
    "som_set_tuning": (None, [c_int, c_int]),
    "som_set_cta_group": (None, [c_int]),
    "som_set_debug": (None, [c_int]),
    "som_set_debug_times": (None, [_P]),
    "som_prep_rows": (c_int, [_P, c_int64, c_int64, c_int64, c_int, _P, _P, c_int64, _P, _P]),
    "som_bmu_init": (c_int, [_P, c_int64, _P]),
    "som_fwd_distances": (c_int, [_P, _P, c_int64, _P, _P, _P, c_int64, _P, c_int64, c_int64, c_int64, c_int,
                                  c_int64, _P, c_int64, _P, _P, c_int64, _P]),
    "som_gemm_workspace_floats": (c_int64, []),
    "som_set_streamk": (None, [c_int]),
    "som_set_pdl": (None, [c_int]),
    "som_bmu_decode": (c_int, [_P, c_int64, c_int64, _P, _P, _P]),
    "som_neighbourhood": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, _P, c_int64, _P]),
    "som_loss_scratch_floats": (c_int64, [c_int64, c_int64]),
    "som_weighted_loss": (c_int, [_P, c_int64, _P, _P, c_int64, c_int64, c_int64, _P, c_float, _P, _P, _P]),
    "som_weighted_loss_grad": (c_int, [_P, _P, c_int64, c_int64, c_int64, _P, _P, c_float, _P, c_int64, _P]),
    "som_bwd_coeffs": (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int64, c_int, _P, _P, _P, _P, c_int64,
                               _P, _P, _P, _P, _P]),
    "som_bwd_dx": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, c_int64, _P, _P, c_int64, c_int64, c_int64,
                           _P, c_int64, _P, c_int64, _P]),
    "som_bwd_dw": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, c_int64, _P, _P, c_int64, c_int64, c_int64,
                           _P, c_int64, _P, c_int64, _P]),
    "som_forward": (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_int64,
                            _P, _P, _P, _P, _P, _P, c_int64, _P, c_int64, _P, _P, c_int64, _P, c_int64, _P]),
    "som_loss_fused_scratch_floats": (c_int64, [c_int64, c_int64]),
    "som_loss_fused_parts": (c_int, [c_int64, c_int64, _P, _P]),
    "som_loss_fused": (c_int, [_P, c_int64, _P, _P, c_int, c_int, c_int64, c_int64, c_int64, _P, c_float, c_int,
                               _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "som_bmu_decode_scaled": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, c_int64, c_int, _P]),
    "som_debug_gemm_f16": (c_int, [_P, _P, c_int64, c_int, _P, _P, c_int64, c_int, c_int64, c_int64, c_int64,
                                   c_int, c_int, c_int, _P, c_int64, _P, c_int64, _P]),
    "som_backward_dw": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, c_int64, _P, c_int64, _P, _P,
                                c_int64, c_int64, c_int64, c_int, _P, c_int64, c_int, c_int, _P, c_int64, _P]),
    "som_backward_dx": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, c_int64, _P, c_int64, _P, _P,
                                c_int64, c_int64, c_int64, c_int, _P, c_int64, c_int, c_int, _P, c_int64, _P]),
    "som_backward_fused": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, c_int64, _P, c_int64, _P, c_int64,
                                   _P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int64, c_int64, c_int,
                                   _P, c_int64, c_int, _P, c_int64, c_int, c_int, _P, _P, _P, c_int64, _P]),
    "som_stream_wait_value": (c_int, [_P, c_uint32, _P]),
    "som_stream_write_value": (c_int, [_P, c_uint32, _P]),
    "som_adamw_step": (c_int, [_P, c_int64, _P, c_int64, _P, _P, c_int64, c_int64, c_int64, _P, c_double, c_double,
                               c_double, c_double, c_int, _P, _P, c_int64, _P, _P]),
    "som_allreduce_mean_nvls": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, _P]),
    "som_allreduce_nvls": (c_int, [_P, _P, c_int64, c_int, c_int, c_float, c_int, _P]),
    "som_nvls_flag_words": (c_int64, [c_int]),
    "som_debug_schedule": (c_int, [c_int64, c_int64, c_int64, c_int64, c_int, c_int, _P]),
    "som_debug_gemm": (c_int, [_P, _P, c_int64, c_int, _P, _P, c_int64, c_int, c_int64, c_int64, c_int64,
                               c_int, c_int, c_int, _P, c_int64, _P, c_int64, _P]),
}


class SomError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def _needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in _DEPS if os.path.exists(p))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into ``vit_som_b200/libsom_b200.so`` (in-tree)."""
    if not force and not _needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise SomError("nvcc not found: cannot build libsom_b200.so (no fallback path exists)")
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    cmd = [nvcc, *NVCC_FLAGS, "-o", tmp, *_SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise SomError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def lib() -> ctypes.CDLL:
    """Load (once) and return the C-ABI library with typed signatures."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise SomError(
                    f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(the SOM hot path has no CPU or eager fallback)")
            handle = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().som_last_error()
        raise SomError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    """Raw handle of the current CUDA stream of ``device`` (default: the current device)."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream
