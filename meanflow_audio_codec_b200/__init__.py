"""B200-native (sm_100a) implementation of the meanflow_audio_codec hot path.

MDCT analysis -> iMF training step on MDCT tokens -> few-NFE sampling -> IMDCT
overlap-add, as hand-written CUDA behind the reference's Python signatures and a
C ABI (``include/mfac.h``, ``libmfac.so``).  See DESIGN.md.
"""
from ._lib import LIB_PATH, MfacError  # noqa: F401
from .mdct import MDCTConfig, MDCTLayer, IMDCTLayer, imdct, mdct  # noqa: F401
from .tokenization import (MDCTTokenization, compute_token_shape, compute_tokenized_dimension,  # noqa: F401
                           create_tokenization_strategy)
from .mlp_flow import AdamW, ConditionalFlow, TrainState, adamw  # noqa: F401
from .loss_strategies import (FlowMatchingLoss, ImprovedMeanFlowLoss, LinearNoiseSchedule,  # noqa: F401
                              LogitNormalTimeSampling, LossStrategy, MeanFlowLoss, MeanFlowTimeSampling,
                              UniformNoiseSchedule, UniformTimeSampling, create_loss_strategy, train_step)
from .sampling import sample, sample_mean_flow  # noqa: F401
from .graphs import GraphedCodec, GraphedTrainStep  # noqa: F401
from .flows import ConditionalConvFlow, ConditionalMLPMixerFlow, create_flow_model  # noqa: F401
from .checkpoint import load_checkpoint, save_checkpoint  # noqa: F401
from .codec import MeanFlowCodec  # noqa: F401
from .input_pipeline import HostBatchStager, LazyTokens  # noqa: F401
