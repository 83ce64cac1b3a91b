"""ctypes binding of libsdt_b200.so (the C ABI declared in include/sdt_b200.h).

The library is built in-tree by ``speechdrivestemplates_b200.build`` (nvcc, sm_100a).  There is NO fallback:
if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdt_b200.so")

c_f32p = C.c_void_p      # device pointers are passed as integers (tensor.data_ptr())
c_ptr = C.c_void_p
i32, i64, f32, f64 = C.c_int32, C.c_int64, C.c_float, C.c_double


class ConvDesc(C.Structure):
    """Mirror of ``sdt_conv_desc`` (include/sdt_b200.h)."""
    _fields_ = [
        ("src", c_ptr), ("wt", c_ptr), ("wt_nk", c_ptr), ("bias", c_ptr), ("xf_scale", c_ptr), ("xf_shift", c_ptr),
        ("dst", c_ptr), ("stat_partial", c_ptr), ("dy", c_ptr), ("wpart", c_ptr),
        ("B", i32), ("SH", i32), ("SW", i32), ("C", i32),
        ("GH", i32), ("GW", i32), ("TH", i32), ("TW", i32),
        ("y_mul", i32), ("ty_mul", i32), ("y_off", i32), ("x_mul", i32), ("tx_mul", i32), ("x_off", i32),
        ("N", i32),
        ("DH", i32), ("DW", i32), ("dy_mul", i32), ("dy_off", i32), ("dx_mul", i32), ("dx_off", i32),
        ("xf_bstride", i32), ("xf_slope", f32),
        ("accumulate", i32), ("per_image_tiles", i32), ("splits", i32), ("math", i32),
        ("rn_act", c_ptr), ("rn_mean", c_ptr), ("rn_rstd", c_ptr), ("rn_eps", f32), ("rn_slope", f32), ("rn_out_tf32", i32),
    ]


_P = C.POINTER(ConvDesc)

# name -> argtypes (restype is int status unless listed in _RESTYPES)
SIGNATURES = {
    "sdt_last_error": [],
    "sdt_version": [],
    "sdt_set_conv_math": [i32],
    "sdt_get_conv_math": [],
    "sdt_tc_launches": [],
    "sdt_mel_fwd": [c_ptr, i32, i32, c_ptr, c_ptr, c_ptr, c_ptr, i32, c_ptr, c_ptr],
    "sdt_conv_row_tiles": [_P],
    "sdt_conv_gemm": [_P, c_ptr],
    "sdt_conv_rownorm_ok": [_P],
    "sdt_conv_plan": [_P, c_ptr],
    "sdt_conv_gemm_multi": [_P, i32, c_ptr, c_ptr],
    "sdt_conv_wgrad": [_P, c_ptr],
    "sdt_conv_wgrad_reduce": [c_ptr, i32, i32, i32, i32, c_ptr, i32, c_ptr],
    "sdt_weight_prep": [c_ptr, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, c_ptr, c_ptr],
    "sdt_weight_prep_batch": [c_ptr, i32, C.c_longlong, c_ptr],
    "sdt_conv_wgrad_reduce_batch": [c_ptr, i32, i32, i32, c_ptr],
    "sdt_p2p_allreduce": [c_ptr, C.c_uint64, i32, i32, C.c_longlong, c_ptr, c_ptr, i32, c_ptr],
    "sdt_chan_stats": [c_ptr, i32, i32, i32, i32, i32, i32, i32, c_ptr, i32, c_ptr],
    "sdt_norm_finalize": [c_ptr, i32, i32, i32, f64, c_ptr, c_ptr, f32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, f32, c_ptr],
    "sdt_bn_eval_scale_shift": [c_ptr, c_ptr, c_ptr, c_ptr, f32, i32, c_ptr, c_ptr, c_ptr],
    "sdt_norm_bwd_reduce": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, i32, i32, f32, c_ptr, i32, c_ptr],
    "sdt_norm_bwd_finalize": [c_ptr, i32, i32, i32, f64, c_ptr, c_ptr, c_ptr, c_ptr, i32, c_ptr],
    "sdt_norm_bwd_apply": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, i32, i32, f32, i32, c_ptr],
    "sdt_rownorm_act_fwd": [c_ptr, i32, i32, f32, f32, c_ptr, c_ptr, c_ptr, i32, c_ptr],
    "sdt_rownorm_act_bwd": [c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, f32, c_ptr, i32, c_ptr],
    "sdt_scale_shift_act": [c_ptr, c_ptr, c_ptr, i32, i32, i32, i32, f32, c_ptr, i32, c_ptr],
    "sdt_first_layer_units": [i32, i32],
    "sdt_first_layer_act": [c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, i32, i32, f32, c_ptr, i32, c_ptr],
    "sdt_first_layer_fwd": [c_ptr, c_ptr, i32, i32, i32, i32, f32, f32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, i32, c_ptr],
    "sdt_first_layer_bwd": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, i32, i32, f32, c_ptr, c_ptr, c_ptr],
    "sdt_enc_to_seq_fwd": [c_ptr, c_ptr, c_ptr, i32, f32, i32, i32, i32, i32, c_ptr, i32, i32, c_ptr, i32, c_ptr],
    "sdt_enc_to_seq_bwd": [c_ptr, i32, i32, i32, i32, i32, i32, c_ptr, c_ptr, c_ptr],
    "sdt_upsample_add_fwd": [c_ptr, c_ptr, i32, i32, i32, i32, c_ptr, i32, c_ptr],
    "sdt_upsample_bwd": [c_ptr, i32, i32, i32, i32, c_ptr, i32, c_ptr],
    "sdt_l1_loss": [c_ptr, c_ptr, i64, f32, c_ptr, c_ptr, c_ptr, c_ptr],
    "sdt_code_gather_kl": [c_ptr, c_ptr, i32, i32, f32, c_ptr, c_ptr, c_ptr, c_ptr],
    "sdt_code_scatter_grad": [c_ptr, c_ptr, c_ptr, i32, i32, c_ptr, c_ptr],
    "sdt_code_store_rows": [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, i32, i32, c_ptr],
    "sdt_colsum": [c_ptr, i32, i32, c_ptr, i32, c_ptr],
    "sdt_mse_const_loss": [c_ptr, i64, f32, f32, c_ptr, c_ptr, c_ptr],
    "sdt_motion_diff_fwd": [c_ptr, i32, i32, i32, c_ptr, c_ptr],
    "sdt_motion_diff_bwd": [c_ptr, i32, i32, i32, c_ptr, i32, c_ptr],
    "sdt_pose_head_fwd": [c_ptr, c_ptr, c_ptr, f32, i32, i32, i32, c_ptr, c_ptr, c_ptr],
    "sdt_vae_reparam_kl": [c_ptr, c_ptr, c_ptr, i32, f32, c_ptr, c_ptr, c_ptr],
    "sdt_pose_head_bwd": [c_ptr, c_ptr, i32, i32, i32, c_ptr, c_ptr],
    "sdt_vae_reparam_kl_bwd": [c_ptr, c_ptr, c_ptr, c_ptr, i32, f32, c_ptr, c_ptr, c_ptr],
    "sdt_pose_preprocess": [c_ptr, i32, c_ptr, c_ptr, i32, c_ptr, c_ptr],
    "sdt_pose_final_results": [c_ptr, i32, i32, c_ptr, c_ptr, c_ptr, i32, c_ptr, c_ptr],
    "sdt_pose_parted2global": [c_ptr, i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr],
    "sdt_pose_metrics": [c_ptr, c_ptr, i32, i32, c_ptr, c_ptr, c_ptr],
    "sdt_adam_advance": [c_ptr, f32, f64, f64, c_ptr],
    "sdt_adam_flat": [c_ptr, c_ptr, c_ptr, c_ptr, i64, c_ptr, f64, f64, f64, f32, f32, c_ptr],
}
_RESTYPES = {"sdt_last_error": C.c_char_p, "sdt_tc_launches": C.c_int64}
# entry points whose int return value is a result, not a status
_NOT_STATUS = {"sdt_last_error", "sdt_version", "sdt_get_conv_math", "sdt_conv_row_tiles", "sdt_tc_launches",
               "sdt_first_layer_units", "sdt_conv_rownorm_ok"}

_lib = None
launch_count = 0     # number of CUDA kernels launched through this binding (bench.py's gpu_launches)
# entry points that launch more than one kernel
_KERNELS_PER_CALL = {"sdt_l1_loss": 2, "sdt_pose_metrics": 2, "sdt_enc_to_seq_bwd": 2, "sdt_first_layer_fwd": 3,
                     "sdt_first_layer_bwd": 2}
# optional (pre, post) callables invoked around every kernel-launching call: bench.py brackets calls with CUDA events
hooks = None


class SdtError(RuntimeError):
    pass


def load(path=None):
    """Load the shared library (once) and declare every prototype of include/sdt_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise SdtError("libsdt_b200.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the library does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def last_error():
    return load().sdt_last_error().decode()


def call(name, *args):
    """Call a status-returning entry point; raise SdtError with sdt_last_error() on failure."""
    global launch_count
    fn = getattr(load(), name)
    if name in _NOT_STATUS:
        return fn(*args)
    if hooks is not None:
        token = hooks[0](name, args)
        rc = fn(*args)
        hooks[1](token)
    else:
        rc = fn(*args)
    if rc != 0:
        raise SdtError("%s failed (%d): %s" % (name, rc, last_error()))
    launch_count += _KERNELS_PER_CALL.get(name, 1)
    return 0
