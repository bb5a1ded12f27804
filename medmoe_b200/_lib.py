"""ctypes binding of libmedmoe_b200.so — the C-ABI declared in include/medmoe_b200.h.

There is no fallback: if the library is missing it is built in-tree with nvcc; if a call
fails the C-ABI's error text is raised as RuntimeError (the reference's convention is
Python exceptions, e.g. src/data/unimed_datamodule.py:75-78).
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

_LIB = None
_LOCK = threading.Lock()

c_int, c_ll, c_f, c_vp = C.c_int, C.c_longlong, C.c_float, C.c_void_p

# name -> (restype, argtypes).  Kept in the order of include/medmoe_b200.h.
SIGNATURES = {
    "mm_last_error": (C.c_char_p, []),
    "mm_abi_version": (c_int, []),
    "mm_device_sm_count": (c_int, []),
    "mm_launch_count": (c_ll, []),
    "mm_trace_enable": (None, [c_int]),
    "mm_trace_collect": (c_int, [C.c_char_p, c_int]),
    "mm_router_topk": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                               c_f, c_vp]),
    "mm_router_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                              c_vp, c_vp, c_vp]),
    "mm_dispatch_build": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                  c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mm_dispatch_rows": (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                 c_vp, c_vp]),
    "mm_undispatch_rows": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mm_cast_f32_bf16": (c_int, [c_vp, c_vp, c_ll, c_vp]),
    "mm_transpose_cast_f32_bf16": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp]),
    "mm_grouped_gemm_rows": (c_int, [c_vp, c_ll, c_int, c_ll, c_vp, c_int, c_int, c_ll, c_vp, c_int, c_int, c_int,
                                     c_vp, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_int, c_vp, c_f, c_int, c_vp]),
    "mm_grouped_gemm_rows_rank1": (c_int, [c_vp, c_ll, c_int, c_ll, c_vp, c_int, c_int, c_ll, c_vp, c_int, c_int,
                                           c_vp, c_vp, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_vp, c_int, c_vp]),
    "mm_grouped_gemm_wgrad": (c_int, [c_vp, c_ll, c_int, c_ll, c_vp, c_ll, c_int, c_ll, c_vp, c_int, c_int, c_int,
                                      c_vp, c_vp]),
    "mm_grouped_gemm_wgrad_colsum": (c_int, [c_vp, c_ll, c_int, c_ll, c_vp, c_ll, c_int, c_ll, c_vp, c_int, c_int, c_int,
                                             c_vp, c_vp, c_vp]),
    "mm_expert_b2b_fwd_supported": (c_int, [c_int, c_int, c_int]),
    "mm_expert_b2b_fwd": (c_int, [c_vp, c_ll, c_int, c_ll, c_vp, c_int, c_int, c_ll, c_vp, c_vp, c_int, c_ll, c_vp, c_vp, c_int,
                                  c_int, c_vp, c_ll, c_vp, c_ll, c_int, c_vp]),
    "mm_combine_num_token_blocks": (c_int, [c_int]),
    "mm_combine_num_row_blocks": (c_int, [c_vp]),
    "mm_combine_num_runs": (c_int, [c_int]),
    "mm_combine_num_part_blocks": (c_int, [c_int, c_vp]),
    "mm_combine_bwd_z_scratch_floats": (c_ll, [c_int, c_vp, c_int]),
    "mm_interp_softmax_combine_fwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp,
                                              c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int,
                                              c_int, c_int, c_ll, c_int, c_vp]),
    "mm_interp_softmax_combine_bwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp,
                                              c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp,
                                              c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "mm_combine_bwd_global_supported": (c_int, [c_int, c_vp, c_int]),
    "mm_combine_bwd_tc_supported": (c_int, [c_int, c_vp, c_int]),
    "mm_debug_force_cuda_core_dut": (None, [c_int]),
    "mm_debug_gemm_pair": (None, [c_int]),
    "mm_local_scores_softmax_exp": (c_int, [c_vp, c_ll, c_int, c_ll, c_vp, c_int, c_ll, c_vp, c_f, c_vp, c_ll, c_vp]),
    "mm_local_softmax_exp_fwd": (c_int, [c_vp, c_ll, c_vp, c_ll, c_ll, c_int, c_int, c_vp, c_f, c_vp]),
    "mm_local_softmax_exp_bwd": (c_int, [c_vp, c_ll, c_vp, c_ll, c_ll, c_int, c_int, c_vp, c_f, c_vp]),
    "mm_local_cos_lse_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_f, c_int, c_vp, c_vp, c_ll, c_vp]),
    "mm_local_cos_lse_bwd": (c_int, [c_vp, c_ll, c_vp, c_ll, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_f, c_int,
                                     c_vp, c_vp, c_ll, c_vp, c_vp]),
    "mm_interp_softmax_combine_bwd_tc": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp, c_vp,
                                                 c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_ll, c_vp, c_vp, c_vp, c_vp,
                                                 c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mm_interp_softmax_combine_bwd_global": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp,
                                                     c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                                     c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mm_gloria_workspace_floats": (c_ll, [c_int]),
    "mm_gloria_global_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_f, c_f, c_vp, c_vp, c_vp]),
    "mm_gloria_global_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_f, c_f, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mm_infonce_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mm_infonce_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_f, c_vp, c_vp,
                               c_vp, c_vp, c_vp, c_int, c_vp]),
    "mm_pack_expert_params": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "mm_infonce_fused_supported": (c_int, [c_int, c_int, c_int]),
    "mm_infonce_fused_workspace_bytes": (c_ll, [c_int, c_int, c_int]),
    "mm_infonce_fused_fwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_f, c_vp, c_vp, c_vp, c_vp,
                                     c_vp, c_vp]),
    "mm_infonce_fused_bwd": (c_int, [c_int, c_int, c_int, c_vp, c_int, c_vp, c_f, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                     c_vp, c_int, c_vp]),
    "mm_l2_normalize_fwd": (c_int, [c_vp, c_int, c_int, c_f, c_vp, c_vp, c_vp]),
    "mm_l2_normalize_bwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_f, c_vp, c_vp]),
    "mm_zeroshot_argmax": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_f, c_vp, c_vp, c_vp]),
    "mm_dispatch_group_map": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "mm_expert_b2b_fwd_gather": (c_int, [c_vp, c_ll, c_int, c_ll, c_vp, c_int, c_int, c_ll, c_vp, c_vp, c_int, c_ll, c_vp, c_vp, c_int,
                                         c_int, c_vp, c_ll, c_vp, c_ll, c_int, c_vp, c_vp]),
    "mm_grouped_gemm_wgrad_colsum_gather": (c_int, [c_vp, c_ll, c_int, c_ll, c_vp, c_ll, c_int, c_ll, c_vp, c_int, c_int, c_int,
                                                    c_vp, c_vp, c_vp, c_vp]),
    "mm_grouped_gemm_rows_scatter": (c_int, [c_vp, c_ll, c_int, c_ll, c_vp, c_int, c_int, c_ll, c_vp, c_int, c_int, c_vp, c_vp,
                                             c_ll, c_ll, c_vp, c_int, c_vp]),
    "mm_p2p_workspace_bytes": (c_ll, [c_int, c_ll]),
    "mm_p2p_alloc": (c_int, [c_ll, c_vp, c_vp]),
    "mm_p2p_open": (c_int, [c_vp, c_vp]),
    "mm_p2p_close": (c_int, [c_vp]),
    "mm_p2p_free": (c_int, [c_vp]),
    "mm_p2p_all_gather": (c_int, [c_vp, c_vp, c_ll, c_vp, c_int, c_int, c_vp]),
    "mm_p2p_reduce_scatter_f32": (c_int, [c_vp, c_vp, c_ll, c_vp, c_int, c_int, c_vp]),
}

# entry points that return a plain value, not an mm_status
_VALUE_FUNCS = {"mm_trace_enable", "mm_trace_collect", "mm_last_error", "mm_abi_version", "mm_device_sm_count", "mm_launch_count", "mm_combine_num_token_blocks",
                "mm_combine_num_row_blocks", "mm_combine_num_runs", "mm_combine_num_part_blocks", "mm_combine_bwd_z_scratch_floats", "mm_gloria_workspace_floats",
                "mm_combine_bwd_global_supported", "mm_combine_bwd_tc_supported",
                "mm_infonce_fused_supported", "mm_infonce_fused_workspace_bytes", "mm_expert_b2b_fwd_supported",
                "mm_debug_force_cuda_core_dut", "mm_debug_gemm_pair", "mm_p2p_workspace_bytes"}


def library_path() -> Path:
    import os
    if os.environ.get("MEDMOE_LIB"):      # tuning experiments only: a variant built with MEDMOE_LIB_OUT (medmoe_b200/build.py)
        return Path(os.environ["MEDMOE_LIB"]).resolve()
    return Path(__file__).resolve().parent / "lib" / "libmedmoe_b200.so"


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building in-tree if needed) the C-ABI library and attach signatures."""
    global _LIB
    if _LIB is not None:
        return _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = library_path()
        if not path.exists():
            if not build_if_missing:
                raise RuntimeError(f"{path} is missing: run `python -m medmoe_b200.build` (no fallback path exists)")
            from . import build as _build
            _build.build()
        lib = C.CDLL(str(path))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError here == header/library drift: fail loudly
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def last_error() -> str:
    return load().mm_last_error().decode("utf-8", "replace")


class EventProfiler:
    """Optional per-call CUDA-event timing on the launching stream (bench.py's roofline evidence).
    Enabled by assigning an instance to `_lib.PROFILER`; costs two event records per C call."""

    def __init__(self):
        self.records = []   # (label, start_event, end_event)

    def summary(self):
        """label -> (calls, total_ms). Call after a device synchronize."""
        out = {}
        for label, a, b in self.records:
            n, t = out.get(label, (0, 0.0))
            out[label] = (n + 1, t + a.elapsed_time(b))
        return out


PROFILER = None


def trace_collect():
    """name -> (calls, total_ms) of the kernels traced inside composite C entry points since mm_trace_enable(1)."""
    buf = C.create_string_buffer(1 << 16)
    n = load().mm_trace_collect(buf, len(buf))
    out = {}
    for line in buf.raw[:n].decode().splitlines():
        name, calls, ms = line.split()
        out[name] = (int(calls), float(ms))
    return out


def call(name: str, *args, label: str = None):
    """Call an mm_status entry point; raise RuntimeError(mm_last_error()) on failure."""
    lib = load()
    fn = getattr(lib, name)
    if name in _VALUE_FUNCS:
        return fn(*args)
    prof = PROFILER
    if prof is not None:
        import torch
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        rc = fn(*args)
        b.record()
        prof.records.append((label or name, a, b))
    else:
        rc = fn(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {last_error()}")
    return rc


def ptr(t) -> int:
    """Device (or host) address of a torch tensor, None -> NULL."""
    return 0 if t is None else t.data_ptr()


def host_i32(values):
    """HOST int32 array argument (ctypes owns the memory for the duration of the call)."""
    return (C.c_int32 * len(values))(*[int(v) for v in values])


def host_ptrs(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
