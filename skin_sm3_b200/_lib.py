"""ctypes binding of libsm3_b200.so (the C ABI declared in include/sm3_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a RuntimeError
is raised.  Build it with ``python -m skin_sm3_b200.build`` (or ``make -C skin_sm3_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SM3_LIB_PATH") or os.path.join(_HERE, "lib", "libsm3_b200.so")
CSRC = os.path.join(_HERE, "csrc")

F32, F16, BF16 = 0, 1, 2
ALGO_AUTO, ALGO_SIMT, ALGO_TC = 0, 1, 2
_DTYPES = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}

_lib = None
_lock = threading.Lock()

_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/sm3_b200.h (tests check this)
SIGNATURES = {
    "sm3_version": (_i, []),
    "sm3_last_error": (C.c_char_p, []),
    "sm3_device_supported": (_i, []),
    "sm3_l2norm_fwd": (_i, [_vp, _i64, _vp, _i64, _i, _i, _vp, _i, _vp, _f, _vp]),
    "sm3_l2norm_bwd": (_i, [_vp, _i, _f, _vp, _i, _vp, _f, _vp, _i64, _vp, _i64, _i, _i, _vp]),
    "sm3_infonce_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "sm3_infonce_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "sm3_infonce_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "sm3_infonce_bwd_packed": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "sm3_peer_scatter_rows": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_vp), _i, _vp]),
    "sm3_peer_scatter_stats": (_i, [_vp, _vp, _vp, _i, _i, _i, C.POINTER(_vp), _i, _vp]),
    "sm3_peer_multicast_rows": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "sm3_peer_multicast_stats": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "sm3_peer_signal": (_i, [C.POINTER(_vp), _i, _i, _i, C.c_uint, _vp]),
    "sm3_peer_wait": (_i, [_vp, _i, _i, C.c_uint, _vp]),
    "sm3_infonce_remote_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "sm3_infonce_fwd_remote": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sm3_infonce_bwd_remote_packed": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sm3_infonce_loss": (_i, [_vp, _vp, _i64, _f, _vp, _i, _vp, _vp, _vp]),
    "sm3_multihead_ce_workspace_bytes": (_sz, [_i64, _i]),
    "sm3_multihead_ce": (_i, [_vp, _i, _vp, _i64, _i, C.POINTER(_i), C.POINTER(_f), _f, _i, _i64, _vp, _vp, _f,
                              _vp, _sz, _vp]),
    "sm3_bce_workspace_bytes": (_sz, [_i64, _i]),
    "sm3_bce_logits": (_i, [_vp, _i, _vp, _i, _vp, _i64, _i, _vp, _vp, _f, _vp, _sz, _vp]),
    "sm3_infonce_fwd_symmetric": (_i, [_i, _i, _vp]),
    "sm3_debug_sym_enumerate": (C.c_longlong, [_i, _i, _i, _i, _vp, C.c_longlong]),
    "sm3_debug_mr_workspace": (_sz, [_i, _i, _i]),
    "sm3_debug_mr_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, C.c_uint, _vp, _vp, _sz, _vp]),
    "sm3_debug_mr_fold": (_i, [_vp, _i, _i, _i, _f, _vp, _vp, _vp, C.c_uint, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sm3_sim_topk": (_i, [_vp, _vp, _i64, _i64, _i, _i, _i, _i64, _vp, _vp, _vp]),
    "sm3_sim_topk_workspace_bytes": (_sz, [_i64, _i64, _i]),
    "sm3_sim_topk_ws": (_i, [_vp, _vp, _i64, _i64, _i, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "sm3_sim_topk_mat_workspace_bytes": (_sz, [_i64, _i64, _i]),
    "sm3_sim_topk_mat": (_i, [_vp, _vp, _i64, _i64, _i, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "sm3_proto_heads_supported": (_i, [_i, _i, _i]),
    "sm3_proto_heads_fwd": (_i, [_vp, _i, _i, _i64, _i, _vp, _i, _vp, _i, _f, _vp, _vp, _vp, _vp]),
    "sm3_proto_heads_bwd": (_i, [_vp, _i, _i, _i64, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "sm3_kmeans_supported": (_i, [_i, _i, _i]),
    "sm3_kmeans_workspace_bytes": (_sz, [_i64, _i, _i]),
    "sm3_kmeans_assign": (_i, [_vp, _i64, _i, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sm3_kmeans_update": (_i, [_vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "sm3_proj_tail_supported": (_i, [_i, _i, _i]),
    "sm3_proj_tail_workspace_bytes": (_sz, [_i64, _i]),
    "sm3_proj_tail_gemm": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "sm3_proj_tail_bn_l2": (_i, [_vp, _i64, _i, _vp, _f, _f, _f, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sm3_proj_tail_bwd1": (_i, [_vp, _i, _i64, _vp, _vp, _f, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _sz, _vp]),
    "sm3_proj_tail_bwd2": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _i, _i64, _i, _vp, _i, _vp]),
    "sm3_infonce_host_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "sm3_infonce_host": (_i, [_vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "sm3_host_pipe_scratch_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sm3_host_pipe_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _vp, _sz]),
    "sm3_host_pipe_submit": (_i64, [_vp, _vp, _vp, _f, _vp, _vp, _vp]),
    "sm3_host_pipe_wait": (_i, [_vp, _i64]),
    "sm3_host_pipe_destroy": (_i, [_vp]),
    "sm3_host_pipe_peer_scratch_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sm3_host_pipe_create_peer": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _vp, _sz]),
    "sm3_host_pipe_submit_peer": (_i64, [_vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _vp, C.POINTER(_vp), _vp, C.POINTER(_vp),
                                         _vp, C.POINTER(_vp), C.c_uint, _i]),
    "sm3_infonce_step_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "sm3_infonce_step": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "sm3_infonce_step_multi_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "sm3_infonce_step_multi": (_i, [_i, C.POINTER(_vp), C.POINTER(_vp), _i, _i, _i, _f, C.POINTER(_f), _vp,
                                    C.POINTER(_vp), C.POINTER(_vp), _vp, _sz, _i, _vp]),
    "sm3_infonce_step_peer_scratch_bytes": (_sz, [_i, _i, _i]),
    "sm3_infonce_step_peer": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, C.POINTER(_vp), _vp,
                                   C.POINTER(_vp), _vp, C.POINTER(_vp), C.c_uint, _i, _vp, _sz, _vp, _vp]),
    "sm3_scale_grads": (_i, [C.POINTER(_vp), C.POINTER(_vp), _i, _i64, _i, _i, _vp, _vp]),
    "sm3_stage_timing": (_i, [_i]),
    "sm3_stage_timing_read": (_i, [_vp, _i]),
    "sm3_stage_timing_names": (C.c_char_p, []),
    "sm3_debug_reload_env": (None, []),
    "sm3_debug_infonce_fwd_ordered_workspace": (_sz, [_i, _i, _i]),
    "sm3_debug_infonce_fwd_ordered": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, C.c_uint, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sm3_debug_umma_rate": (_i, [_i, _i, _i, _i, _vp]),
    "sm3_debug_umma_probe": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
}


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into skin_sm3_b200/lib/libsm3_b200.so (nvcc cross-compiles
    without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0 or not os.path.exists(LIB_PATH):
        raise RuntimeError("building libsm3_b200.so failed (see output above)")
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -m skin_sm3_b200.build`. "
                    "skin_sm3_b200 has no CPU or PyTorch fallback.")
            l = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(l, name)
                fn.restype = res
                fn.argtypes = args
            _lib = l
    return _lib


def last_error() -> str:
    return lib().sm3_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> int:
    if rc < 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")
    return rc


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype} (float32 / float16 / bfloat16 only)") from None


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("skin_sm3_b200 operates on CUDA tensors only (no CPU path); got a CPU tensor")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("all tensors must live on the same CUDA device")
    return dev


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()
