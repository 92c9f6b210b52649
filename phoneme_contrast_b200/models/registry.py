"""Name -> model class registry with the reference's behaviour (src/models/registry.py:7-41):
register(name) decorator (duplicate -> ValueError), get / create (unknown -> ValueError), list()."""
from __future__ import annotations

from typing import Callable, Dict, List, Mapping, Type

from .base import BaseModel


class ModelRegistry:
    def __init__(self) -> None:
        self._models: Dict[str, Type[BaseModel]] = {}

    def register(self, name: str) -> Callable[[Type[BaseModel]], Type[BaseModel]]:
        def wrap(cls: Type[BaseModel]) -> Type[BaseModel]:
            if name in self._models:
                raise ValueError(f"Model {name} already registered")
            self._models[name] = cls
            return cls

        return wrap

    def get(self, name: str) -> Type[BaseModel]:
        try:
            return self._models[name]
        except KeyError:
            raise ValueError(f"Model {name} not found. Available: {list(self._models)}") from None

    def create(self, name: str, config: Mapping) -> BaseModel:
        return self.get(name)(config)

    def list(self) -> List[str]:
        return list(self._models)


model_registry = ModelRegistry()
