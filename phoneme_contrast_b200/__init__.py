"""B200-native (sm_100a) implementation of the phoneme_contrast data-parallel training hot path:
MFCC front end + view augmentation, PhonemeNet / PhonemeNetDeep forward+backward, SupCon loss, and the
clip+Adam step, behind the reference's own Python interfaces.

Sub-packages mirror the reference's module layout (src/datasets, src/models, src/training).
"""
__all__ = ["datasets", "models", "training"]
