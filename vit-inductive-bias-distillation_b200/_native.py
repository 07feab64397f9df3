"""ctypes binding of libbasd_b200.so (the C ABI declared in include/basd_b200.h).

There is no CPU or PyTorch fallback: if the library is missing the first kernel call
raises, loudly, with the build command.  Tensors stay torch-owned; only data pointers,
sizes and the current CUDA stream cross the boundary.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbasd_b200.so")

F32, BF16 = 0, 1
_DT = {torch.float32: F32, torch.bfloat16: BF16}

_p, _i, _l, _f = C.c_void_p, C.c_int, C.c_long, C.c_float

# name -> argument ctypes (return type is int unless listed in _RET)
_SIG = {
    "basd_sgemm_batched": [_i, _i, _i, _i, _i, _p, _i, _i, _l, _p, _p, _i, _l, _p, _i, _l, _i, _f, _p, _f, _p],
    "basd_rough_means": [_p, _i, _i, _l, _i, _l, _p, _p],
    "basd_merge_shifted_stats": [_p, _p, _p, _p, _p, _i, _i, _i, _l, _p],
    "basd_token_gram_simt_workspace_floats": [_l, _i],
    "basd_token_gram_simt": [_p, _i, _l, _i, _p, _p, _p, _p, _p],
    "basd_pivoted_cholesky": [_p, _i, _i, _l, _p, _i, _l, _i, _f, _p, _p, _p],
    "basd_jacobi_rows": [_p, _i, _i, _i, _l, _i, _p, _f, _i, _p, _p],
    "basd_jacobi_rows_counted": [_p, _i, _i, _i, _l, _i, _p, _f, _i, _p, _p, _p],
    "basd_jacobi_rows_ex": [_p, _i, _i, _i, _l, _i, _p, _p, _f, _f, _i, _p, _p, _p],
    "basd_jacobi_rows_ranked": [_p, _i, _i, _i, _l, _i, _p, _f, _i, _p, _p, _p],
    "basd_rows_normalize": [_p, _i, _i, _i, _l, _p, _i, _l, _p, _i, _i, _i, _f, _p, _p],
    "basd_rowdot": [_p, _i, _l, _p, _i, _l, _i, _i, _i, _p, _p],
    "basd_center_gram": [_p, _p, _i, _f, _p, _i, _p],
    "basd_rotate_stats_f64_workspace_bytes": [_i, _i, _i],
    "basd_rotate_stats_f64": [_p, _i, _i, _p, _p, _i, C.c_double, _p, _p, _p, _p],
    "basd_mp_rank": [_p, _i, _l, _i, _p, _p, _i, _p],
    "basd_mp_rank_secular": [_p, _p, _i, _l, _i, _p, _p, _i, _p],
    "basd_expand_ranks": [_p, _i, _i, _p, _p],
    "basd_mask_block": [_p, _p, _i, _p, _i, _p],
    "basd_angle_distance": [_p, _p, _p, _i, _i, _i, _p, _p],
    "basd_mix_weights": [_p, _p, _i, _i, _p, _p, _p],
    "basd_mix_weights_bwd": [_p, _p, _p, _p, _i, _i, _f, _p, _p, _p],
    "basd_scale_rows_dsigma": [_p, _p, _p, _p, _p, _i, _i, _i, _p],
    "basd_omega_accumulate": [_p, _p, _p, _i, _i, _i, _p, _p],
    "basd_symmetrize_add": [_p, _i, _p, _i, _p],
    "basd_shift_diag": [_p, _i, _f, _i, _p],
    "basd_projector_complement": [_p, _i, _p, _p, _i, _p],
    "basd_place_complement": [_p, _p, _p, _i, _i, _p],
    "basd_attn_rows": [_p, _i, _i, _i, _i, _i, _i, _p, _p],
    "basd_mix_interp": [_p, _i, _i, _p, _i, _i, _i, _i, _i, _p, _i, _p],
    "basd_mix_flat": [_p, _i, _i, _p, _i, _l, _p, _p],
    "basd_mix_rows": [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p],
    "basd_weight_grad_slices": [],
    "basd_weight_grad": [_p, _i, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p, _p],
    "basd_weighted_center": [_p, _i, _l, _p, _l, _i, _i, _p, _l, _i, _p],
    "basd_extract_diag": [_p, _i, _i, _l, _i, _p, _p],
    "basd_procrustes_rows_finish": [_p, _i, _i, _l, _p, _i, _i, _l, _i, _f, _f, _i, _i, _p, _p, _p, _p],
    "basd_procrustes_grad_prep": [_p, _p, _p, _p, _i, _i, _i, _l, _i, _l, _i, _p, _p, _p, _p, _p, _p, _p,
                                  _p, _i, _p],
    "basd_procrustes_direct_grad": [_p, _p, _p, _i, _i, _i, _p],
    "basd_scale_out": [_p, _p, _i, _l, _f, _p, _p],
    "basd_geo_reduce": [_p, _i, _i, _p, _i, _p, _p, _p],
    "basd_gemm_tc3_supported": [_i, _i, _i, _i, _i, _i, _l, _l, _l],
    "basd_gemm_tc3_batched": [_i, _i, _i, _i, _i, _p, _i, _l, _p, _i, _l, _p, _i, _l, _i, _f, _p, _p],
    "basd_gemm_tc3_batched_ex": [_i, _i, _i, _i, _i, _p, _i, _i, _l, _p, _i, _l, _p, _i, _i, _l, _i, _f, _p,
                                 _p, _p],
    "basd_cast_out": [_p, _p, _i, _l, _p],
}
_RET = {"basd_token_gram_simt_workspace_floats": _l, "basd_rotate_stats_f64_workspace_bytes": _l,
        }
# optional symbols (present once the tcgen05 Gram is built)
_OPTIONAL = {
    "basd_token_gram_tc_workspace_bytes": ([_l, _i], _l),
    "basd_token_gram_tc": ([_p, _l, _i, _p, _p, _p, _p, _p], _i),
}

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    """Loads (once) and returns the ctypes handle; raises if the .so was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: the BASD CUDA kernels are not built and there is no "
            "fallback path. Build with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C vit-inductive-bias-distillation_b200/csrc`.")
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIG.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _RET.get(name, _i)
    for name, (args, ret) in _OPTIONAL.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = ret
    _lib = lib
    return lib


def exported_symbols() -> list[str]:
    return list(_SIG)


def has(name: str) -> bool:
    return hasattr(load(), name)


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"BASD kernels take float32 or bfloat16 tensors, got {t.dtype}") from None


def ptr(t) -> int | None:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("BASD kernels need CUDA tensors (there is no CPU path)")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# Optional per-call CUDA-event timing (bench.py's roofline leg): when `_timeline` is a list,
# every C-ABI call is bracketed by events recorded on the launching stream.
_timeline = None
launch_count = 0
gemm_flops = {}          # entry point -> fp32-equivalent flops of its launches while the timeline is on


def start_timeline():
    global _timeline
    _timeline = []


def stop_timeline():
    """Returns [(entry point, milliseconds)] for the calls since start_timeline()."""
    global _timeline
    tl, _timeline = _timeline, None
    torch.cuda.synchronize()
    return [(n, a.elapsed_time(b)) for n, a, b in tl]


def call(name: str, *args):
    global launch_count
    launch_count += 1
    if _timeline is not None:
        if name.startswith("basd_gemm_tc3_batched"):      # args: ta, tb, m, n, k, ... batch
            m, n, k = args[2], args[3], args[4]
            batch = args[14] if name == "basd_gemm_tc3_batched" else args[16]
            gemm_flops[name] = gemm_flops.get(name, 0.0) + 2.0 * m * n * k * batch
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        rc = getattr(load(), name)(*args)
        b.record()
        _timeline.append((name, a, b))
    else:
        rc = getattr(load(), name)(*args)
    if rc != 0:
        msg = f"{name} failed with code {rc}"
        if rc > 0:
            msg += " (cudaError_t)"
        raise RuntimeError(msg)
