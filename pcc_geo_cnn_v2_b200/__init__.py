"""pcc_geo_cnn_v2_b200 -- B200 (sm_100a) implementation of the pcc_geo_cnn_v2 hot path: the per-block 3D-conv
analysis/synthesis/hyper transforms and learned entropy models, behind the reference's own Python call
signatures.  All compute runs in libpccgeo.so (hand-written CUDA; C ABI in include/pccgeo.h)."""
from .model_transforms import (TransformType, set_seed, set_precision, get_precision, relu)  # noqa: F401
from .model_types import ModelType, CompressionModelV1, CompressionModelV2  # noqa: F401
from .model_configs import ModelConfig, ModelConfigType, PAPER_CONFIGS  # noqa: F401
from .entropy_models import EntropyBottleneck, GaussianConditional  # noqa: F401
from .focal_loss import focal_loss  # noqa: F401
