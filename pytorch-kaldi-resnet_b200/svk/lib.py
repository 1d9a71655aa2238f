"""ctypes binding of libsvk.so (include/svk.h).  The product path fails loudly when the library is missing: there is
no torch / CPU fallback anywhere in this package."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVK_LIB_PATH") or os.path.join(_HERE, "libsvk.so")     # override: A/B runs of kernel variants

F32, BF16 = 0, 1
IMPL_SIMT, IMPL_TCGEN05 = 0, 1


class BnBwdFuse(Structure):
    """Mirror of svk_bn_bwd_fuse."""
    _fields_ = [("mask", c_void_p), ("c", c_void_p), ("mean", c_void_p), ("rstd", c_void_p), ("sums", c_void_p)]


class ConvDesc(Structure):
    """Mirror of svk_conv_desc."""
    _fields_ = [("N", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Ho", c_int), ("Wo", c_int),
                ("Cout", c_int), ("R", c_int), ("stride", c_int), ("dtype", c_int), ("impl", c_int)]


def make_conv_desc(N, H, W, Cin, Cout, R, stride, dtype, impl):
    return ConvDesc(N, H, W, Cin, (H - 1) // stride + 1, (W - 1) // stride + 1, Cout, R, stride, dtype, impl)


_P = c_void_p
_I = c_int
_L = c_longlong
_F = c_float
_D = POINTER(ConvDesc)

# name -> argtypes, exactly as declared in include/svk.h (tests check every symbol is exported)
SIGNATURES = {
    "svk_pack_conv_weight": [_P, _P, _P, _I, _I, _I, _I, _P],
    "svk_pack_conv_weights_batched": [_P, _P, _P, _P, _I, _L, _I, _P],
    "svk_bn_train_act_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _F, _F, _I, _P, _L, _I, _I, _P],
    "svk_conv2d_fwd": [_D, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P],
    "svk_conv2d_dgrad": [_D, _P, _P, _P, _P, _P, _P, _P],
    "svk_conv2d_dgrad_bn": [_D, _P, _P, _P, _P, _P, _P, ctypes.POINTER(BnBwdFuse), _P],
    "svk_downsample_dgrad_bn": [_D, _P, _P, _D, _P, _P, _P, ctypes.POINTER(BnBwdFuse), _P],
    "svk_relu_mask_inplace": [_P, _P, _L, _I, _P],
    "svk_conv2d_wgrad": [_D, _P, _P, _P, _P, ctypes.c_size_t, _P],
    "svk_stem_conv_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P],
    "svk_stem_conv_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "svk_channel_stats": [_P, _L, _I, _I, _P, _P],
    "svk_bn_finalize": [_P, _L, _I, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P],
    "svk_bn_eval_coeffs": [_P, _P, _P, _P, _F, _I, _P, _P, _P],
    "svk_bn_act_fwd": [_P, _P, _P, _P, _P, _P, _I, _P, _L, _I, _I, _P],
    "svk_bn_bwd_reduce": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "svk_bn_bwd_apply": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P],
    "svk_add_masked": [_P, _P, _P, _P, _L, _I, _P],
    "svk_add_strided2": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "svk_statspool_fwd": [_P, _P, _I, _I, _I, _I, _I, _P, _I, _P],
    "svk_statspool_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "svk_sgemm": [_P, _L, _L, _P, _L, _L, _P, _L, _I, _I, _I, _F, _F, _P, _P],
    "svk_colsum": [_P, _P, _I, _I, _L, _P],
    "svk_l2norm_rows_fwd": [_P, _P, _P, _I, _I, _F, _P],
    "svk_l2norm_rows_bwd": [_P, _P, _P, _P, _I, _I, _P],
    "svk_aam_margin_fwd": [_P, _P, _P, _I, _I, _F, _F, _F, _F, _F, _P],
    "svk_aam_margin_bwd": [_P, _P, _P, _I, _I, _I, _F, _F, _F, _F, _P],
    "svk_aam_ce_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _F, _F, _F, _I, _P, ctypes.c_size_t, _P],
    "svk_aam_ce_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _F, _F, _I, _P, ctypes.c_size_t, _P],
    "svk_gemm_tf32": [_P, _L, _I, _P, _L, _I, _P, _L, _I, _I, _I, _P, _P, ctypes.c_size_t, _P],
    "svk_ce_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _P],
    "svk_ce_bwd": [_P, _P, _P, _P, _F, _P, _I, _I, _P],
    "svk_sgd_step": [_P, _P, _P, _L, _F, _F, _F, _F, _P],
    "svk_cast": [_P, _P, _L, _I, _I, _P],
    "svk_cosine_score_pairs": [_P, _P, _P, _P, _P, _P, _L, _I, _P],
    "svk_topk_meanstd": [_P, _I, _I, _I, _P, _P, _P],
    "svk_snorm_apply": [_P, _P, _P, _P, _P, _P, _P, _P, _L, _P],
    "svk_sort_pairs_f64": [_P, _P, _P, _P, _L, _P, ctypes.c_size_t, _P],
    "svk_det_metrics": [_P, _L, _L, c_double, c_double, c_double, _P, _P, ctypes.c_size_t, _P],
    "svk_segment_mean": [_P, _P, _P, _I, _I, _P, _P],
    "svk_col_mean": [_P, _L, _I, _P, _P, ctypes.c_size_t, _P],
}

_lib = None


class SvkError(RuntimeError):
    pass


def load():
    """Load libsvk.so once.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SvkError("libsvk.so not found at %s — build it with `python pytorch-kaldi-resnet_b200/csrc/build.py` "
                       "(or __graft_entry__.build()); there is no fallback path" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.svk_version.restype = c_int
    lib.svk_last_error_string.restype = c_char_p
    lib.svk_launch_count.restype = c_longlong
    lib.svk_conv2d_wgrad_workspace_bytes.restype = ctypes.c_size_t
    lib.svk_conv2d_wgrad_workspace_bytes.argtypes = [_D]
    lib.svk_gemm_tf32_workspace_bytes.restype = ctypes.c_size_t
    lib.svk_gemm_tf32_workspace_bytes.argtypes = [_I, _I, _I]
    lib.svk_aam_ce_workspace_bytes.restype = ctypes.c_size_t
    lib.svk_aam_ce_workspace_bytes.argtypes = [_I, _I, _I]
    for name, argtypes in (("svk_sort_pairs_f64_workspace_bytes", [_L]), ("svk_det_metrics_workspace_bytes", [_L]),
                           ("svk_col_mean_workspace_bytes", [_L, _I])):
        getattr(lib, name).restype = ctypes.c_size_t
        getattr(lib, name).argtypes = argtypes
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    _lib = lib
    return lib


def launch_count():
    return int(load().svk_launch_count())


def check(rc, name):
    if rc != 0:
        msg = load().svk_last_error_string().decode("utf-8", "replace")
        raise SvkError("%s failed (code %d): %s" % (name, rc, msg))


class _Caller(object):
    """`call.svk_xxx(args...)` -> runs the C function and raises SvkError on a non-zero return code.
    With PROFILE set to a list (bench.py), every call is bracketed by CUDA events on the current stream and
    (name, tag, start_event, end_event) is appended — per-kernel device times measured live, no profiler attached."""

    def __getattr__(self, name):
        fn = getattr(load(), name)

        def wrapped(*args):
            if PROFILE is None:
                rc = fn(*args)
                if rc != 0:
                    check(rc, name)
                if TRACE is not None:
                    TRACE(name)
                return
            import torch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            tag = ""
            if args and isinstance(args[0], ConvDesc):
                d = args[0]
                tag = "%dx%d %d->%d k%d s%d N%d" % (d.H, d.W, d.Cin, d.Cout, d.R, d.stride, d.N)
                for d in args[1:]:          # a second conv in the same call (svk_downsample_dgrad_bn)
                    if isinstance(d, ConvDesc):
                        tag += " + %dx%d %d->%d k%d s%d N%d" % (d.H, d.W, d.Cin, d.Cout, d.R, d.stride, d.N)
            PROFILE.append((name, tag, e0, e1))
            if rc != 0:
                check(rc, name)
        setattr(self, name, wrapped)
        return wrapped


PROFILE = None
TRACE = None           # debugging hook: called with the entry-point name after every libsvk call (tests/diag_graph.py)
call = _Caller()


def profile_begin():
    global PROFILE
    PROFILE = []


def profile_end():
    """-> list of (name, tag, milliseconds); synchronises the device."""
    global PROFILE
    import torch
    torch.cuda.synchronize()
    out = [(n, t, e0.elapsed_time(e1)) for n, t, e0, e1 in PROFILE]
    PROFILE = None
    return out
