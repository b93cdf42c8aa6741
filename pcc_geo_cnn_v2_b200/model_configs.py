"""Config table: host mirror of reference src/model_configs.py:7-49 (network configs c1, c2, c3, c3p) plus the
paper's experiment labels c1..c6 from src/ev_experiment.yml:10-46 (which network / alpha / threshold mode each uses)."""
from enum import Enum

from .model_transforms import TransformType
from .model_types import ModelType


class ModelConfig:
    def __init__(self, model_type: ModelType, model_params):
        self.model_type = model_type
        self.model_params = model_params

    def build(self, **overrides):
        return self.model_type.value(**{**self.model_params, **overrides})


class ModelConfigType(Enum):
    c1 = ModelConfig(ModelType.v1, {
        'num_filters': 32,
        'analysis_transform_type': TransformType.AnalysisTransformV1,
        'synthesis_transform_type': TransformType.SynthesisTransformV1})
    c2 = ModelConfig(ModelType.v2, {
        'num_filters': 32,
        'analysis_transform_type': TransformType.AnalysisTransformV1,
        'synthesis_transform_type': TransformType.SynthesisTransformV1,
        'hyper_analysis_transform_type': TransformType.HyperAnalysisTransform,
        'hyper_synthesis_transform_type': TransformType.HyperSynthesisTransform})
    c3 = ModelConfig(ModelType.v2, {
        'num_filters': 32,
        'analysis_transform_type': TransformType.AnalysisTransformV2,
        'synthesis_transform_type': TransformType.SynthesisTransformV2,
        'hyper_analysis_transform_type': TransformType.HyperAnalysisTransform,
        'hyper_synthesis_transform_type': TransformType.HyperSynthesisTransform})
    c3p = ModelConfig(ModelType.v2, {
        'num_filters': 64,
        'analysis_transform_type': TransformType.AnalysisTransformProgressiveV2,
        'synthesis_transform_type': TransformType.SynthesisTransformProgressiveV2,
        'hyper_analysis_transform_type': TransformType.HyperAnalysisTransform,
        'hyper_synthesis_transform_type': TransformType.HyperSynthesisTransform})

    @staticmethod
    def keys():
        return ModelConfigType.__members__.keys()

    def build(self, **overrides):
        return self.value.build(**overrides)


# paper label -> (network config, focal-loss alpha, fixed_threshold, train_mode); src/ev_experiment.yml:10-53
PAPER_CONFIGS = {
    'c1': dict(model_config='c1', alpha=0.9, fixed_threshold=True, train_mode='independent'),
    'c2': dict(model_config='c2', alpha=0.9, fixed_threshold=True, train_mode='independent'),
    'c3': dict(model_config='c3p', alpha=0.9, fixed_threshold=True, train_mode='independent'),
    'c4': dict(model_config='c3p', alpha=0.75, fixed_threshold=True, train_mode='independent'),
    'c5': dict(model_config='c3p', alpha=0.75, fixed_threshold=False, train_mode='independent'),
    'c6': dict(model_config='c3p', alpha=0.75, fixed_threshold=False, train_mode='warm_seq'),
}
