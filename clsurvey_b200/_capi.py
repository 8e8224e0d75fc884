"""ctypes binding of include/clb.h (the C-ABI boundary).  No torch types cross this boundary: tensors are passed as
raw device pointers (`tensor.data_ptr()`), the CUDA stream as an integer handle.

There is NO fallback: if libclb.so is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLB_LIB_PATH") or os.path.join(_HERE, "_lib", "libclb.so")   # CLB_LIB_PATH: A/B builds

_lib = None

c_f, c_i, c_i64, c_d, c_p, c_sz = (ctypes.c_float, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p,
                                   ctypes.c_size_t)

# name -> argtypes (restype int unless listed in _RESTYPE)
SIGNATURES = {
    "clb_last_error": [],
    "clb_version": [],
    "clb_sm_count": [c_p],
    "clb_set_matmul_mode": [c_i],
    "clb_get_matmul_mode": [],
    "clb_launch_count": [],
    "clb_memset_zero": [c_p, c_sz, c_p],
    "clb_stream_create": [c_p],
    "clb_conv2d_fwd": [c_p, c_p, c_p, c_p, c_p] + [c_i] * 10 + [c_p],
    "clb_conv2d_dgrad": [c_p, c_p, c_p, c_p] + [c_i] * 9 + [c_p],
    "clb_conv2d_wgrad_ws": [c_i] * 9,
    "clb_conv2d_wgrad": [c_p, c_p, c_p, c_p, c_p, c_sz] + [c_i] * 9 + [c_p],
    "clb_linear_ws": [c_i, c_i, c_i],
    "clb_linear_fwd": [c_p, c_p, c_p, c_p, c_p, c_sz, c_i, c_i, c_i, c_i, c_p],
    "clb_linear_dgrad": [c_p, c_p, c_p, c_p, c_sz, c_i, c_i, c_i, c_p],
    "clb_linear_wgrad": [c_p, c_p, c_p, c_p, c_p, c_sz, c_i, c_i, c_i, c_p],
    "clb_relu_bwd": [c_p, c_p, c_p, c_i64, c_p],
    "clb_maxpool_fwd": [c_p, c_p, c_p] + [c_i] * 6 + [c_p],
    "clb_maxpool_bwd": [c_p, c_p, c_p, c_p] + [c_i] * 6 + [c_p],
    "clb_adaptive_avgpool_fwd": [c_p, c_p] + [c_i] * 6 + [c_p],
    "clb_adaptive_avgpool_bwd": [c_p, c_p] + [c_i] * 6 + [c_p],
    "clb_mask_mul": [c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "clb_softmax_loss": [c_p, c_i, c_i, c_i, c_p, c_i, c_i, c_f, c_p, c_p, c_p, c_p],
    "clb_planes_conv_supported": [c_i] * 8,
    "clb_planes_weights": [c_p] * 5 + [c_i, c_i, c_p],
    "clb_planes_weights_batch": [c_i] + [c_p] * 9,
    "clb_planes_defer_join": [c_i],
    "clb_planes_join": [c_p],
    "clb_planes_linear_supported": [c_i, c_i],
    "clb_planes_linear_fwd": [c_p] * 7 + [c_i] * 4 + [c_p],
    "clb_planes_linear_dgrad": [c_p] * 7 + [c_i] * 3 + [c_p],
    "clb_planes_linear_wgrad_ws": [c_i] * 3,
    "clb_planes_linear_wgrad": [c_p] * 7 + [c_sz] + [c_i] * 4 + [c_p, c_f, c_f, c_p],
    "clb_planes_pool_fwd_flat": [c_p] * 5 + [c_i] * 4 + [c_p],
    "clb_planes_pool_bwd_flat": [c_p] * 6 + [c_i] * 4 + [c_p],
    "clb_planes_to_f32": [c_p, c_p, c_p, c_i64, c_p],
    "clb_planes_from_f32": [c_p, c_p, c_p, c_p, c_i64, c_p],
    "clb_planes_conv_fwd": [c_p] * 7 + [c_i] * 6 + [c_p],
    "clb_planes_conv_dgrad": [c_p] * 7 + [c_i] * 5 + [c_p],
    "clb_planes_conv_wgrad_ws": [c_i] * 5,
    "clb_planes_conv_wgrad": [c_p] * 7 + [c_sz] + [c_i] * 6 + [c_p, c_f, c_f, c_p],
    "clb_planes_conv1_supported": [c_i] * 8,
    "clb_planes_conv1_pool_fwd": [c_p] * 6 + [c_i] * 5 + [c_p],
    "clb_planes_conv1_ws": [],
    "clb_planes_conv1_pool_bwd": [c_p] * 8 + [c_sz] + [c_i] * 5 + [c_p],
    "clb_planes_pool_fwd": [c_p] * 6 + [c_i] * 4 + [c_p],
    "clb_planes_pool_fwd_nchw": [c_p] * 4 + [c_i] * 4 + [c_p],
    "clb_planes_pool_bwd": [c_p] * 8 + [c_i] * 4 + [c_p],
    "clb_planes_pool_bwd_nchw": [c_p] * 5 + [c_i] * 4 + [c_p],
    "clb_sgd_penalty_step": [c_p, c_p, c_p, c_p, c_p, c_i64, c_i64, c_f, c_f, c_f, c_f, c_f, c_i, c_p],
    "clb_si_step": [c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_f, c_f, c_f, c_f, c_f, c_i, c_p],
    "clb_fisher_accum": [c_p, c_p, c_f, c_i64, c_p],
    "clb_mas_accum": [c_p, c_p, c_f, c_f, c_i64, c_p],
    "clb_si_consolidate": [c_p, c_p, c_p, c_p, c_f, c_i64, c_p],
    "clb_imm_merge_accum": [c_p, c_p, c_p, c_p, c_i64, c_i, c_p],
    "clb_axpby": [c_p, c_p, c_p, c_f, c_i64, c_p],
    "clb_gem_dots_gram": [c_p, c_p, c_i64, c_i64, c_p, c_i, c_p, c_p, c_p],
    "clb_gem_solve_qp": [c_p, c_p, c_i, c_d, c_d, c_p, c_p, c_p],
    "clb_gem_solve_qp_host": [c_p, c_p, c_i, c_d, c_d, c_p, c_p],
    "clb_gem_project": [c_p, c_p, c_i64, c_i64, c_p, c_i, c_p, c_p, c_p],
    "clb_nccl_unique_id": [c_p],
    "clb_nccl_init": [c_p, c_i, c_i, c_p],
    "clb_nccl_allreduce_f32": [c_p, c_p, c_i64, c_p],
    "clb_nccl_destroy": [c_p],
}
_RESTYPE = {"clb_last_error": ctypes.c_char_p, "clb_conv2d_wgrad_ws": c_sz, "clb_planes_conv_wgrad_ws": c_sz, "clb_planes_linear_wgrad_ws": c_sz, "clb_planes_conv1_ws": c_sz, "clb_launch_count": ctypes.c_ulonglong, "clb_linear_ws": c_sz}


class ClbError(RuntimeError):
    pass


def lib():
    """Load libclb.so (once). Raises if it has not been built -- the product has no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ClbError("libclb.so not found at %s -- run `python -m clsurvey_b200.build` "
                           "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if the symbol is missing
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, c_i)
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise ClbError("%s failed (rc=%d): %s" % (what, rc, lib().clb_last_error().decode()))


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise ClbError("%s failed (rc=%d): %s" % (name, rc, lib().clb_last_error().decode()))


_PRIVATE_STREAMS = {}


def private_stream(name):
    """A process-wide torch stream (one per name and device) backed by clb_stream_create: never shared with torch's pooled
    streams, hence never the stream a CUDA graph is being captured on."""
    import ctypes
    import torch
    key = (name, torch.cuda.current_device())
    st = _PRIVATE_STREAMS.get(key)
    if st is None:
        h = ctypes.c_void_p()
        call("clb_stream_create", ctypes.byref(h))
        st = torch.cuda.ExternalStream(h.value)
        _PRIVATE_STREAMS[key] = st
    return st
