from .base import BaseModel
from .phoneme_cnn import PhonemeNet, PhonemeNetDeep
from .registry import model_registry

__all__ = ["BaseModel", "model_registry", "PhonemeNet", "PhonemeNetDeep"]
