"""Model base class with the reference's contract (src/models/base.py:9-31): constructed from a plain
config mapping, forward(x[B,C,F,T]) -> [B, embedding_dim], get_embedding_dim()."""
from __future__ import annotations

from typing import Any, Mapping

import torch
import torch.nn as nn


class BaseModel(nn.Module):
    def __init__(self, config: Mapping[str, Any]):
        super().__init__()
        self.config = config

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # pragma: no cover - abstract
        raise NotImplementedError

    def get_embedding_dim(self) -> int:
        return self.config.get("embedding_dim", 128)
