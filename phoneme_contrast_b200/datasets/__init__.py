from .dataset import DeviceFrontendLoader, PhonemeContrastiveDataset
from .features import FeatureExtractor, MFCCExtractor, MelSpectrogramExtractor, build_feature_extractor
from .transforms import (BaseTransform, Compose, FrequencyMask, GaussianNoise, TimeMask, TimeStretch,
                         build_augmentation_pipeline, build_view_descriptors, pack_view_descs)

__all__ = ["FeatureExtractor", "MFCCExtractor", "MelSpectrogramExtractor", "build_feature_extractor", "BaseTransform",
           "Compose", "FrequencyMask", "GaussianNoise", "TimeMask", "TimeStretch", "build_augmentation_pipeline",
           "build_view_descriptors", "pack_view_descs", "PhonemeContrastiveDataset", "DeviceFrontendLoader"]
