"""ctypes binding of libpasnl_b200.so (the C ABI declared in include/pasnl.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is missing or a symbol
cannot be resolved, importing an op raises immediately (``PasnlLibraryError``) -- it never degrades silently.
Build it in-tree with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C pointasnl_b200/csrc``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpasnl_b200.so")

c_int, c_long, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_long, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
P = c_void_p  # every device pointer travels as void*

# name -> (restype, argtypes); mirrors include/pasnl.h one to one (tests/test_abi.py checks the two against each other)
SIGNATURES = {
    "pasnl_version": (ctypes.c_char_p, []),
    "pasnl_error_string": (ctypes.c_char_p, [c_int]),
    "pasnl_fps_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pasnl_fps": (c_int, [c_int, c_int, c_int, P, P, P, P]),
    "pasnl_gather_point": (c_int, [c_int, c_int, c_int, P, P, P, P]),
    "pasnl_gather_point_grad": (c_int, [c_int, c_int, c_int, P, P, P, P]),
    "pasnl_gather_rows": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P]),
    "pasnl_gather_rows_grad": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P]),
    "pasnl_ball_query": (c_int, [c_int, c_int, c_int, c_float, c_int, P, P, P, P, P]),
    "pasnl_group_point": (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, P]),
    "pasnl_group_point_grad": (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, P]),
    "pasnl_knn": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "pasnl_group_concat": (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, P, c_int, P, P, P]),
    "pasnl_group_concat_grad": (c_int, [c_int, c_int, c_int, c_int, c_int, P, P, c_int, c_int, P, P, P, P]),
    "pasnl_pointconv_fwd": (c_int, [c_int, c_int, c_int, P, P, P, P]),
    "pasnl_pointconv_bwd": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P]),
    "pasnl_pointconv_fwd_bn": (c_int, [c_int, c_int, c_int, P, P, P, P, P, c_int, P]),
    "pasnl_pointconv_bwd_bn": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, P, c_int, P]),
    "pasnl_maxpool_k": (c_int, [c_long, c_int, c_int, P, P, P]),
    "pasnl_maxpool_k_grad": (c_int, [c_long, c_int, c_int, P, P, P, P, P]),
    "pasnl_three_nn": (c_int, [c_int, c_int, c_int, P, P, P, P, P]),
    "pasnl_three_interpolate": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "pasnl_three_interpolate_grad": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "pasnl_as_core_fwd": (c_int, [c_int, c_int, c_int, P, P, c_float, P, P, P]),
    "pasnl_as_core_bwd": (c_int, [c_int, c_int, c_int, P, P, P, c_float, P, P, P, P]),
    "pasnl_as_resample_fwd": (c_int, [c_int, c_int, c_int, P, P, c_long, c_long, P, c_long, c_long, P, P, P, P]),
    "pasnl_as_resample_bwd": (c_int, [c_int, c_int, c_int, P, P, c_long, c_long, P, c_long, c_long, P, P, P, P, P, P]),
    "pasnl_nl_core_fwd": (c_int, [c_int, c_int, c_int, c_int, P, P, c_float, P, P, P]),
    "pasnl_nl_core_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "pasnl_nl_core_bwd": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P, c_float, P, P, P, P, P]),
    "pasnl_gemm_rows_tf32x3": (c_int, [c_long, c_int, c_int, P, c_long, P, c_long, P, c_long, P, P]),
    "pasnl_gemm_rows_ws_eligible": (c_int, [c_long, c_int, c_int, P, c_long]),
    "pasnl_gemm_rows_ws_wide_eligible": (c_int, [c_long, c_int, c_int, P, c_long]),
    "pasnl_gemm_rows_ws": (c_int, [c_long, c_int, c_int, P, c_long, P, c_long, c_int, P, c_long, P, P, P, c_int, P]),
    "pasnl_gemm_rows_tma_eligible": (c_int, [c_long, c_int, c_int, P, c_long, P, c_long]),
    "pasnl_gemm_rows_tma": (c_int, [c_long, c_int, c_int, P, c_long, P, c_long, P, c_long, P, P]),
    "pasnl_gemm_wgrad_tf32x3": (c_int, [c_long, c_int, c_int, P, c_long, P, c_long, P, c_long, c_int, P]),
    "pasnl_gemm_skinny": (c_int, [c_int, c_int, c_int, P, c_long, P, c_long, c_int, P, c_long, P]),
    "pasnl_wgrad_skinny": (c_int, [c_int, c_int, c_int, P, c_long, P, c_long, P, c_long, P]),
    "pasnl_add_n": (c_int, [c_long, P, P, P, P, P]),
    "pasnl_concat_cols": (c_int, [c_long, c_int, c_int, P, c_long, P, c_long, P, P]),
    "pasnl_split_cols": (c_int, [c_long, c_int, c_int, P, P, P, P]),
    "pasnl_dropout": (c_int, [c_long, P, c_float, ctypes.c_ulonglong, P, P, P]),
    "pasnl_softmax_ce_workspace_bytes": (c_size_t, []),
    "pasnl_softmax_ce_fwd": (c_int, [c_long, c_int, P, P, P, c_float, P, P, P, P]),
    "pasnl_softmax_ce_bwd": (c_int, [c_long, c_int, P, P, P, c_float, P, P, P, P, P]),
    "pasnl_adam_l2": (c_int, [c_long, P, P, P, P, P, c_float, c_float, c_float, c_float, c_float, c_float, P, P, P, P, P]),
    "pasnl_maxpool_k_mask": (c_int, [c_long, c_int, c_int, P, P, P, P]),
    "pasnl_maxpool_k_scatter": (c_int, [c_long, c_int, c_int, P, P, P, P]),
    "pasnl_add_cols": (c_int, [c_long, c_int, P, P, c_long, P]),
    "pasnl_take_cols": (c_int, [c_long, c_int, P, c_long, P, P]),
    "pasnl_transpose_batched": (c_int, [c_int, P, c_int, P]),
    "pasnl_bias_add": (c_int, [c_long, c_int, P, P, P, P]),
    "pasnl_col_sum": (c_int, [c_long, c_int, P, P, P]),
    "pasnl_grid_subsample_workspace_bytes": (c_size_t, [c_int]),
    "pasnl_grid_subsample": (c_int, [c_int, P, c_int, P, c_int, P, c_float, P, P, P, P, P, P]),
    "pasnl_prob_sample": (c_int, [c_int, c_int, c_int, P, P, P, P, P]),
    "pasnl_selection_sort": (c_int, [c_int, c_int, c_int, c_int, P, P, P, P]),
    "pasnl_gemm_wgrad_tma_eligible": (c_int, [c_long, c_int, c_int, P, c_long, P, c_long]),
    "pasnl_gemm_wgrad_tma": (c_int, [c_long, c_int, c_int, P, c_long, P, c_long, P, c_long, c_int, P, P, c_int, P]),
    "pasnl_bn_finalize": (c_int, [c_long, c_int, P, P, P, c_float, P, P, P, c_float, c_int, P, P, P, P, P]),
    "pasnl_bn_workspace_bytes": (c_size_t, [c_int]),
    "pasnl_bn_act_fwd": (c_int, [c_long, c_int, P, P, P, c_float, c_int, P, P, P, P, P, P, c_float, c_int, c_int, P, P]),
    "pasnl_bn_act_bwd": (c_int, [c_long, c_int, P, P, P, P, P, P, c_float, c_int, P, P, P, c_int, P, P]),
}


class PasnlLibraryError(RuntimeError):
    pass


class PasnlError(RuntimeError):
    """A pasnl_* entry point returned non-zero."""


_lib = None
launch_count = 0   # CUDA kernels launched by pasnl_* entry points so far (bench.py reports the delta as gpu_launches)
profiler = None    # optional object with before(name, args) / after(name, args), set by bench.py's per-kernel pass

# kernels (incl. memset nodes) each entry point enqueues; everything not listed launches exactly one kernel
KERNELS_PER_CALL = {
    "pasnl_gather_point_grad": 2, "pasnl_gather_rows_grad": 2, "pasnl_group_point_grad": 2,
    "pasnl_three_interpolate_grad": 2, "pasnl_group_concat_grad": 4, "pasnl_nl_core_bwd": 3,
    "pasnl_bn_act_fwd": 3, "pasnl_bn_act_bwd": 3, "pasnl_gemm_wgrad_tf32x3": 2, "pasnl_softmax_ce_fwd": 2,
    "pasnl_adam_l2": 2, "pasnl_col_sum": 2, "pasnl_grid_subsample": 15, "pasnl_prob_sample": 2,
}


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PasnlLibraryError(
            f"{LIB_PATH} not found: the sm_100a CUDA library is not built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no fallback path.")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise PasnlLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise PasnlLibraryError(f"{LIB_PATH} does not export {name}; rebuild the library") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def version():
    return load().pasnl_version().decode()


def check(code, what):
    if code != 0:
        msg = load().pasnl_error_string(int(code)).decode()
        raise PasnlError(f"{what} failed with code {code}: {msg}")


def call(name, *args):
    """Invoke a compute entry point and raise on a non-zero return code."""
    global launch_count
    launch_count += KERNELS_PER_CALL.get(name, 1)
    fn = getattr(load(), name)
    if profiler is None:
        check(fn(*args), name)
    else:
        profiler.before(name, args)
        check(fn(*args), name)
        profiler.after(name, args)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
