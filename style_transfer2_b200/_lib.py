"""ctypes binding of libst2.so (include/st2.h).  No torch types cross this boundary: only raw
device/host pointers, sizes and a CUDA stream handle.  There is no CPU fallback: importing this
module builds/loads the CUDA library and fails loudly if it cannot."""
import ctypes as C
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))

NUM_BLOBS = 22
NUM_CONVS = 16
SCAL_PER_BLOB = 24
SCAL_GLOBAL_BASE = NUM_BLOBS * SCAL_PER_BLOB
SCAL_TOTAL = SCAL_GLOBAL_BASE + 32
PREC_FP32, PREC_FP16 = 0, 1
PROF_CATS = 12
PROF_NAMES = ('conv_tc', 'conv_first', 'pool', 'gram', 'style_grad', 'loss_elementwise', 'pixel_terms',
              'optimizer', 'conv_exact', 'halo', 'gram_finalize')
RESAMPLE_LANCZOS, RESAMPLE_BILINEAR = 0, 1
# per-blob scalar fields (st2_common.cuh)
(SB_C_SUMSQ, SB_S_GRAMSQ, SB_S_RAWSQ, SB_D_SUMSQ, SB_C_NORM, SB_S_NORM, SB_D_NORM, SB_C_VALID, SB_S_VALID,
 SB_D_VALID, SB_C_COEF, SB_S_COEF, SB_D_COEF, SB_C_LOSS, SB_C_GRAD, SB_S_LOSS, SB_S_GRAD, SB_D_LOSS,
 SB_D_GRAD, SB_S_DSCALE) = range(20)
# global scalar fields (include/st2.h)
(G_SCD_LOSS, G_TV_NORM, G_P_NORM, G_SCD_GRAD_SQ, G_T_GRAD_SQ, G_P_GRAD_SQ, G_GRAD_SQ, G_T_LOSS, G_P_LOSS,
  G_LOSS, G_SCD_GRAD, G_T_GRAD, G_P_GRAD, G_GRAD, G_HALO_TIMEOUT, G_PROTOCOL_ERROR) = range(16)
IPC_HANDLE_BYTES = 64

_vp, _i, _ll, _f, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double
_ip, _dp = C.POINTER(C.c_int), C.POINTER(C.c_double)

_SIGS = {
    'st2_ctx_create': (_i, [_i, C.POINTER(_vp)]),
    'st2_ctx_destroy': (None, [_vp]),
    'st2_last_error': (C.c_char_p, [_vp]),
    'st2_set_stream': (_i, [_vp, _vp]),
    'st2_launch_count': (_ll, [_vp]),
    'st2_profile': (_i, [_vp, _i]),
    'st2_profile_read': (_i, [_vp, _dp, C.POINTER(C.c_longlong)]),
    'st2_debug_flags': (_i, [_vp, _i]),
    'st2_bench_layer': (_i, [_vp, _i, _i, _i, C.POINTER(C.c_float)]),
    'st2_set_conv_weights': (_i, [_vp, _i, _vp, _vp, _i, _i]),
    'st2_blob_count': (_i, []),
    'st2_blob_name': (C.c_char_p, [_i]),
    'st2_blob_channels': (_i, [_i]),
    'st2_blob_kind': (_i, [_i]),
    'st2_plan_create': (_i, [_vp, _i, _i, _i, C.POINTER(_vp)]),
    'st2_plan_destroy': (None, [_vp]),
    'st2_plan_blob_dims': (_i, [_vp, _i, _ip, _ip, _ip]),
    'st2_forward': (_i, [_vp, _vp, _i]),
    'st2_blob_export': (_i, [_vp, _i, _vp]),
    'st2_backward': (_i, [_vp, _i, _ip, C.POINTER(_vp), _vp]),
    'st2_capture_content': (_i, [_vp, _i]),
    'st2_gram': (_i, [_vp, _i, _vp]),
    'st2_set_style_gram': (_i, [_vp, _i, _vp]),
    'st2_set_blob_weights': (_i, [_vp, _i, _f, _f, _f]),
    'st2_set_eval_order': (_i, [_vp, _i, _ip]),
    'st2_set_params': (_i, [_vp, _f, _f, _f, _f]),
    'st2_reset_norms': (_i, [_vp]),
    'st2_set_norm': (_i, [_vp, _i, _i, _d]),
    'st2_eval': (_i, [_vp, _vp, _vp, _i]),
    'st2_eval_begin': (_i, [_vp, _vp, _i]),
    'st2_eval_mid': (_i, [_vp]),
    'st2_eval_end': (_i, [_vp, _vp]),
    'st2_eval_final': (_i, [_vp]),
    'st2_strip_plan_create': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, C.POINTER(_vp)]),
    'st2_strip_ipc_handle': (_i, [_vp, _vp]),
    'st2_strip_attach': (_i, [_vp, _i, _vp, _vp, _i]),
    'st2_strip_set_fold': (_i, [_vp, _i]),
    'st2_strip_set_deferred': (_i, [_vp, _i]),
    'st2_strip_reduce_block': (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(C.c_longlong)]),
    'st2_strip_halo_error': (_i, [_vp, _ip]),
    'st2_read_scalars': (_i, [_vp, _dp]),
    'st2_copy_scalars_async': (_i, [_vp, _vp]),
    'st2_scalars_dev': (_vp, [_vp]),
    'st2_gram_nchw': (_i, [_vp, _vp, _i, _ll, _vp]),
    'st2_pixel_terms': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _f, _f, _vp]),
    'st2_preprocess_u8': (_i, [_vp, _vp, _vp, _i, _i]),
    'st2_preprocess_f32': (_i, [_vp, _vp, _vp, _i, _i]),
    'st2_deprocess': (_i, [_vp, _vp, _vp, _i, _i]),
    'st2_dot': (_i, [_vp, _vp, _vp, _ll, _dp]),
    'st2_axpy': (_i, [_vp, _f, _vp, _vp, _ll]),
    'st2_sumsq': (_i, [_vp, _vp, _ll, _dp]),
    'st2_lbfgs_create': (_i, [_vp, _ll, _i, C.POINTER(_vp)]),
    'st2_lbfgs_destroy': (None, [_vp]),
    'st2_lbfgs_reset': (_i, [_vp]),
    'st2_lbfgs_advance': (_i, [_vp, _vp, _vp, _f]),
    'st2_lbfgs_commit': (_i, [_vp, _vp, _vp]),
    'st2_lbfgs_advance_begin': (_i, [_vp, _vp]),
    'st2_lbfgs_advance_end': (_i, [_vp, _vp, _vp, _f]),
    'st2_lbfgs_commit_begin': (_i, [_vp, _vp, _vp]),
    'st2_lbfgs_commit_end': (_i, [_vp]),
    'st2_lbfgs_sums_dev': (_vp, [_vp]),
    'st2_lbfgs_sums_count': (_i, []),
    'st2_lbfgs_set_global_length': (_i, [_vp, _d]),
    'st2_lbfgs_load': (_i, [_vp, _i, _vp, _vp, _dp]),
    'st2_lbfgs_export': (_i, [_vp, _ip, _vp, _vp, _dp]),
    'st2_adam_step': (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _f, _d, _d, _i, _i]),
    'st2_resample': (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _i]),
}

EXPORTS = tuple(sorted(_SIGS))


class St2Error(RuntimeError):
    pass


def load():
    """Build (if nvcc is here and sources changed) and dlopen libst2.so; bind every export."""
    path = _build.build()
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)          # AttributeError = header/library mismatch: fail loudly
        fn.restype, fn.argtypes = res, args
    return lib


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB


def check(ctx, rc, what=''):
    if rc != 0:
        msg = lib().st2_last_error(ctx)
        raise St2Error('%s failed (%d): %s' % (what or 'libst2 call', rc, (msg or b'').decode()))
