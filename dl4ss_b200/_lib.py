"""ctypes binding of libdl4ss_b200.so (the C ABI declared in include/dl4ss_b200.h).

There is no CPU fallback: if the shared library cannot be loaded (or built with nvcc), importing
any compute entry point raises.  Every wrapper checks that its tensors are CUDA, contiguous and of
the expected dtype, passes raw device pointers plus the current torch stream, and turns a non-zero
status into a RuntimeError carrying dl4ss_last_error().
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdl4ss_b200.so')

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_f = ctypes.c_float
c_sz = ctypes.c_size_t
c_ll = ctypes.c_longlong

# name -> (restype, argtypes); mirrors include/dl4ss_b200.h one to one
SIGNATURES = {
    'dl4ss_version': (c_i, []),
    'dl4ss_last_error': (ctypes.c_char_p, []),
    'dl4ss_launch_count': (ctypes.c_uint64, []),
    'dl4ss_stft_feat': (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_f, c_i, c_p, c_p, c_p]),
    'dl4ss_mask_istft': (c_i, [c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    'dl4ss_linear_fwd': (c_i, [c_p, c_i, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    'dl4ss_rnn_workspace_bytes': (c_sz, [c_i, c_i, c_i, c_i]),
    'dl4ss_rnn_layer_fwd': (c_i, [c_i, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_sz, c_p]),
    'dl4ss_rnn_mma_supported': (c_i, [c_i, c_i]),
    'dl4ss_rnn_mma_workspace_bytes': (c_sz, [c_i, c_i, c_i, c_i]),
    'dl4ss_rnn_layer_mma_fwd': (c_i, [c_i, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_sz, c_p]),
    'dl4ss_rnn_tc_supported': (c_i, [c_i, c_i]),
    'dl4ss_rnn_tc_set_trace': (None, [c_p, c_i]),
    'dl4ss_rnn_tc_set_tiles_per_cta': (None, [c_i]),
    'dl4ss_gemm_tc_set_max_ctas': (None, [c_i]),
    'dl4ss_gemm_tc_set_two_cta': (None, [c_i]),
    'dl4ss_rnn_tc_set_cluster_pairs': (None, [c_i]),
    'dl4ss_rnn_tc_whh_bytes': (c_sz, [c_i]),
    'dl4ss_rnn_tc_pack_whh': (c_i, [c_i, c_p, c_i, c_p, c_p]),
    'dl4ss_rnn_tc_workspace_bytes': (c_sz, [c_i, c_i, c_i, c_i]),
    'dl4ss_rnn_layer_tc_fwd': (c_i, [c_i, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_sz, c_p]),
    'dl4ss_emb_attn_mask_workspace_bytes': (c_sz, [c_i, c_i, c_i, c_i]),
    'dl4ss_emb_attn_mask_fwd': (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_p,
                                      c_p, c_sz, c_p]),
    'dl4ss_split_bf16_bytes': (c_sz, [c_ll, c_i]),
    'dl4ss_split_bf16': (c_i, [c_p, c_i, c_ll, c_i, c_p, c_p]),
    'dl4ss_split_bf16_t': (c_i, [c_p, c_ll, c_i, c_i, c_p, c_p]),
    'dl4ss_linear_tc_fwd': (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    'dl4ss_linear_tc_splitk_fwd': (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p]),
    'dl4ss_linear_tc_lda_fwd': (c_i, [c_p, c_i, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p]),
    'dl4ss_linear_tc_tn_splitk_fwd': (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    'dl4ss_emb_attn_mask_tc_fwd': (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_p]),
    'dl4ss_attn_dot_fwd': (c_i, [c_p, c_ll, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_p]),
    'dl4ss_speaker_query_fwd': (c_i, [c_p, c_i, c_i, c_i, c_p, c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p]),
    'dl4ss_premix_fwd': (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p]),
    'dl4ss_premix_shift_fwd': (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p]),
    'dl4ss_xcorr_f64': (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    'dl4ss_conv3x3s2_relu_fwd': (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    'dl4ss_rowdot_sigmoid_fwd': (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_p]),
    'dl4ss_mask_loss_bwd': (c_i, [c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_f, c_f, c_p, c_p]),
    'dl4ss_attn_dot_bwd': (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_p]),
    'dl4ss_attn_dot_bwd_planes': (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_i, c_p, c_p]),
    'dl4ss_rnn_bwd_step': (c_i, [c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p]),
    'dl4ss_rnn_bwd_supported': (c_i, [c_i, c_i]),
    'dl4ss_rnn_bwd_set_trace': (None, [c_p, c_i]),
    'dl4ss_rnn_bwd_workspace_bytes': (c_sz, [c_i, c_i, c_i, c_i]),
    'dl4ss_rnn_layer_bwd': (c_i, [c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_sz, c_p]),
    'dl4ss_rnn_bwd_tc_supported': (c_i, [c_i, c_i]),
    'dl4ss_rnn_bwd_tc_xplanes_bytes': (c_sz, [c_i, c_i, c_i, c_i]),
    'dl4ss_rnn_layer_bwd_tc': (c_i, [c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_sz, c_p]),
    'dl4ss_mask_pair_loss_fwd': (c_i, [c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    'dl4ss_mask_loss_fwd': (c_i, [c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
}

# enums (include/dl4ss_b200.h)
FEAT_NONE, FEAT_ABS, FEAT_LOG = 0, 1, 2
WAV_F32, WAV_F64 = 0, 1
MASK_NONE, MASK_REAL, MASK_COMPLEX = 0, 1, 2
CELL_LSTM, CELL_GRU = 0, 1
ACT_NONE, ACT_TANH, ACT_SIGMOID = 0, 1, 2
ATT_DOT, ATT_DOT_CRM = 0, 1

_lib = None


def load(build_if_missing=True):
    """Load (building first when absent and nvcc is available) and return the CDLL."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and build_if_missing:
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('dl4ss_b200: %s is missing and could not be built; there is no CPU fallback' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().dl4ss_last_error().decode('utf-8', 'replace')


def launch_count():
    return int(load().dl4ss_launch_count())


def check(rc, what):
    if rc != 0:
        raise RuntimeError('%s failed (%d): %s' % (what, rc, last_error()))


def ptr(t, dtype=torch.float32, name='tensor'):
    """Raw device pointer of a CUDA, contiguous tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('dl4ss_b200: %s must be a CUDA tensor (no CPU path)' % name)
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError('dl4ss_b200: %s must be %s, got %s' % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError('dl4ss_b200: %s must be contiguous' % name)
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
