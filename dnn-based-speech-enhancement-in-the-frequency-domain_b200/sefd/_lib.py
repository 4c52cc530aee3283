"""ctypes binding of libsefd.so (the C ABI declared in include/sefd.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEFD_LIB") or os.path.join(_HERE, "libsefd.so")   # SEFD_LIB: A/B-test another build

_vp, _i, _f, _ll, _sz = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t

# name -> (restype, argtypes); kept in the order of include/sefd.h
SIGNATURES = {
    "sefd_abi_version": (_i, []),
    "sefd_last_error": (C.c_char_p, []),
    "sefd_stale_cuda_errors": (_i, []),
    "sefd_last_stale_cuda_error": (C.c_char_p, []),
    "sefd_stft_forward": (_i, [_vp, _vp, _i, _i, _vp]),
    "sefd_istft_forward": (_i, [_vp, _vp, _i, _i, _vp]),
    "sefd_istft_backward": (_i, [_vp, _vp, _i, _i, _vp]),
    "sefd_mask_istft_forward": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "sefd_stft_forward_n": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "sefd_mask_istft_forward_n": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "sefd_stft_mask_istft_fused": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "sefd_mask_istft_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "sefd_loss_forward": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "sefd_loss_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "sefd_cconv_workspace_bytes": (_sz, [_i, _i]),
    "sefd_cconv2d_forward": (_i, [_vp] * 6 + [_i] * 5 + [_vp, _vp]),
    "sefd_cconv2d_backward": (_i, [_vp] * 9 + [_i] * 5 + [_vp, _vp]),
    "sefd_cconvT2d_forward": (_i, [_vp] * 7 + [_i] * 5 + [_vp, _vp]),
    "sefd_cconvT2d_backward": (_i, [_vp] * 11 + [_i] * 5 + [_vp, _vp]),
    "sefd_bn_prelu_forward": (_i, [_vp, _vp, _ll, _i] + [_vp] * 8),
    "sefd_bn_prelu_backward": (_i, [_vp, _vp, _vp, _ll, _i] + [_vp] * 9),
    "sefd_cbn_prelu_forward": (_i, [_vp, _vp, _ll, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "sefd_cbn_prelu_backward": (_i, [_vp, _vp, _vp, _ll, _i] + [_vp] * 10),
    "sefd_lstm_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "sefd_lstm_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "sefd_adam_step": (_i, [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _i, _f, _vp]),
    "sefd_axpby": (_i, [_vp, _vp, _f, _f, _ll, _vp]),
    "sefd_counters_inc": (_i, [_vp, _i, _ll, _vp]),
    "sefd_adam_step_dev": (_i, [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _vp, _vp, _f, _vp]),
    "sefd_dccrn_plan_create": (_vp, [_i, _i, _i]),
    "sefd_dccrn_plan_create_ex": (_vp, [_i, _i, _i, _i]),
    "sefd_dccrn_plan_destroy": (None, [_vp]),
    "sefd_dccrn_workspace_bytes": (_sz, [_vp]),
    "sefd_dccrn_param_floats": (_ll, [_vp]),
    "sefd_dccrn_buffer_floats": (_ll, [_vp]),
    "sefd_dccrn_num_params": (_i, [_vp]),
    "sefd_dccrn_num_buffers": (_i, [_vp]),
    "sefd_dccrn_entry_info": (_i, [_vp, _i, _i, C.c_char_p, _i, C.POINTER(_ll), C.POINTER(_ll), C.POINTER(_i),
                                   C.POINTER(_ll)]),
    "sefd_dccrn_tensor_info": (_i, [_vp, C.c_char_p, C.POINTER(_ll), C.POINTER(_i), C.POINTER(_ll)]),
    "sefd_dccrn_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sefd_dccrn_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sefd_dccrn_backward_spec": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sefd_dccrn_loss": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "sefd_lms_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "sefd_fsn_frames": (_i, [_i]),
    "sefd_fsn_features": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "sefd_fsn_stft": (_i, [_vp, _i, _i, _vp, _vp]),
    "sefd_fsn_mag_phase": (_i, [_vp, C.c_longlong, _vp, _vp, _vp]),
    "sefd_fsn_cirm": (_i, [_vp, _vp, C.c_longlong, _vp, _vp]),
    "sefd_fsn_decompress_cirm": (_i, [_vp, C.c_longlong, _vp, _vp]),
    "sefd_fsn_compress_cirm": (_i, [_vp, C.c_longlong, _vp, _vp]),
    "sefd_fsn_istft": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "sefd_pmsqe_table_floats": (_i, []),
    "sefd_pmsqe_workspace_bytes": (C.c_size_t, [_i, _i]),
    "sefd_pmsqe_forward": (_i, [_vp, _vp, _i, _i, _vp, _vp, C.c_size_t, _vp, _vp]),
    "sefd_pmsqe_backward": (_i, [_vp, _i, _i, _vp, _vp, C.c_size_t, _vp, _vp]),
    "sefd_lms_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "sefd_crn_plan_create": (_vp, [_i, _i]),
    "sefd_crn_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sefd_crn_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sefd_crn_backward_spec": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sefd_fsn_plan_create": (_vp, [_i, _i]),
    "sefd_fsn_forward": (_i, [_vp, _vp, _vp, _i, _f, _vp, _vp, C.c_ulonglong, _vp, _vp, _sz, _vp]),
    "sefd_fsn_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sefd_dropout_forward": (_i, [_vp, _vp, _ll, _f, _vp, C.c_ulonglong, C.c_uint, _vp]),
    "sefd_dccrn_grad_split": (_ll, [_vp]),
    "sefd_dccrn_backward_overlap": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "sefd_nccl_unique_id_bytes": (_i, []),
    "sefd_nccl_unique_id": (_i, [_vp]),
    "sefd_nccl_init": (_vp, [_i, _i, _vp]),
    "sefd_nccl_allreduce": (_i, [_vp, _vp, _ll, _vp]),
    "sefd_nccl_world": (_i, [_vp]),
    "sefd_nccl_destroy": (None, [_vp]),
    "sefd_set_engine": (_i, [_i]),
    "sefd_get_engine": (_i, []),
    "sefd_launch_count": (_ll, []),
    "sefd_prof_enable": (_i, [_i]),
    "sefd_prof_reset": (_i, []),
    "sefd_prof_dump": (_i, [C.c_char_p]),
    "sefd_prof_get": (_i, [_i, C.POINTER(C.c_double), C.POINTER(_ll), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load libsefd.so and attach signatures. Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the sefd CUDA library has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` or csrc/build.sh). "
            "There is no CPU / PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError here means header and library are out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().sefd_last_error().decode(errors="replace")
        raise RuntimeError(f"sefd {what} failed (rc={rc}): {msg}")
