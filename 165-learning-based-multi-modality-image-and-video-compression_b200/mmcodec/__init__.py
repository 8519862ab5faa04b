"""mmcodec -- B200-native (sm_100a) hot path of the CompressAI-based multi-modality learned codec
(SZU-AdvTech-2022/165).  The public surface mirrors the reference's modules:

    mmcodec.layers           GDN, conv, deconv, LowerBound, NonNegativeParametrizer
    mmcodec.entropy_models   EntropyModel, EntropyBottleneck, GaussianConditional
    mmcodec.models           CompressionModel, FactorizedPrior, ScaleHyperprior, MeanScaleHyperprior
    mmcodec.models_mm        JointAutoregressiveHierarchicalPriors_R / _D (RGB + depth two-branch codec), ESA, MaskedConv2d
    mmcodec.models_master    Master_compresser (RGB-T reproduction: feature codecs, channel aligner, window cross-attention decoder)
    mmcodec.models_video     ScaleSpaceFlow (ssf2020 video codec: keyframe / inter-frame forward, compress, decompress)
    mmcodec.autograd         training path: autograd Functions over the forward / backward kernels
    mmcodec.training         RateDistortionLoss, configure_optimizers, GradBucketReducer (NCCL), TrainStep
    mmcodec.transforms_functional   rgb2ycbcr, ycbcr2rgb, yuv_444_to_420, yuv_420_to_444 (compressai.transforms.functional)
    mmcodec.ops              functional access to every entry point of include/mmcodec.h
    mmcodec.library          torch.library ops (mmcodec::gdn, mmcodec::lower_bound) + the stand-ins torch.jit.script compiles

All compute runs in libmmcodec.so (hand-written CUDA for sm_100a).  No CPU fallback.
"""
from . import _lib, compress_pipeline, entropy_models, graphs, host_pipeline, layers, library, models, models_master, models_mm, models_video, ops, training, transforms, transforms_functional  # noqa: F401
from .accelerate import accelerate  # noqa: F401
from .graphs import GraphedForward  # noqa: F401
from .transforms import precision  # noqa: F401
from .host_pipeline import HostPipeline  # noqa: F401
from .compress_pipeline import CompressPipeline  # noqa: F401
from ._lib import MmcodecError, build  # noqa: F401
from .entropy_models import EntropyBottleneck, EntropyModel, GaussianConditional  # noqa: F401
from .layers import GDN, LowerBound, NonNegativeParametrizer, conv, deconv  # noqa: F401
from .models import (CompressionModel, FactorizedPrior, MeanScaleHyperprior, ScaleHyperprior,  # noqa: F401
                     build_model, get_scale_table, set_entropy_coder)

from .models_mm import (ESA, Guided_compresser, JointAutoregressiveHierarchicalPriors, JointAutoregressiveHierarchicalPriors_D,  # noqa: F401
                        JointAutoregressiveHierarchicalPriors_R, MaskedConv2d)
from .models_master import Master_compresser  # noqa: F401
from .models_video import ScaleSpaceFlow  # noqa: F401
from .training import GradBucketReducer, GraphedTrainStep, RateDistortionLoss, TrainStep, configure_optimizers  # noqa: F401

__version__ = "0.1.0"
