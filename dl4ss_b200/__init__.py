"""dl4ss_b200 -- B200-native (sm_100a) implementation of the DL4SS separation hot path.

Public surface (names follow the reference, shincling/DL4SS):
    config                                   the globals the modules read
    MIX_SPEECH, ATTENTION, SPEECH_EMBEDDING, ADDJUST, top_k_mask      (modules.py)
    stft_features, mask_istft, prepare_batch                         (features.py)
    Separator, mask_loss, pit_mask_loss                              (pipeline.py)
    TrainStep, allreduce_gradients, shard_range                      (training.py)
    prepare_data, bss_eval, bss_eval_cRM, eval_bss, bss_test.cal      (compat.py: the reference's loop-level entry points)
All compute goes through libdl4ss_b200.so (C ABI, include/dl4ss_b200.h); no CPU fallback.
"""
from . import config  # noqa: F401
from ._lib import load as load_library, launch_count  # noqa: F401
from .features import stft_features, mask_istft, prepare_batch, premix, window_tensor  # noqa: F401
from .modules import (MIX_SPEECH, MIX_SPEECH_classifier, Discriminator, gan_loss_terms, ATTENTION, SPEECH_EMBEDDING, ADDJUST, top_k_mask,  # noqa: F401
                      DeferredEmbedding, linear_fwd, linear_tc, split_bf16, weight_planes, rnn_forward,
                      emb_attn_mask, crm_decompress)
from .pipeline import Separator, GraphedSeparator, PipelinedSeparator, HostPipeline, sm_sharing, mask_loss, pit_mask_loss  # noqa: F401
from .training import TrainStep, allreduce_gradients, shard_range  # noqa: F401
from . import compat  # noqa: F401
from .compat import prepare_data, prepare_datasize, bss_eval, bss_eval_cRM, eval_bss, multi_label_vector  # noqa: F401
from . import bss_test, predata_fromList  # noqa: F401

__version__ = '0.1.0'
