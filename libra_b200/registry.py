"""Name -> class tables with the reference registry's interface for the two kinds this path uses
(libra/common/registry.py:9-19 the mapping, :54-75 `register_model`, :78-101 `register_processor`, :205-211 the getters):
train.py resolves `registry.get_model_class(model_config.arch).from_config(model_config)` (train.py:28-30), the wrapper
registers itself as "libra_train_wrapper" (libra/models/libra/modeling_libra.py:1292) and the dataset builders look up
"libra_image" / "libra_image_eval" (libra/data/processors/libra_processor.py:65, 96)."""
from __future__ import annotations

from typing import Callable, Dict, Optional

_KINDS = ("model", "processor")


def _decorator_for(table: Dict[str, type], kind: str) -> Callable[[str], Callable[[type], type]]:
    def register(name: str):
        def wrap(cls: type) -> type:
            taken = table.get(name)
            if taken is not None:
                raise KeyError(f"{kind} name '{name}' already registered for {taken}.")
            table[name] = cls
            return cls
        return wrap
    return register


class Registry:
    # same layout as the reference's class attribute, so code that peeks at `registry.mapping[...]` keeps working
    mapping: Dict[str, dict] = {f"{k}_name_mapping": {} for k in _KINDS}
    mapping.update({"state": {}, "paths": {}})

    register_model = staticmethod(_decorator_for(mapping["model_name_mapping"], "model"))
    register_processor = staticmethod(_decorator_for(mapping["processor_name_mapping"], "processor"))

    @classmethod
    def _lookup(cls, kind: str, name: str) -> Optional[type]:
        return cls.mapping[f"{kind}_name_mapping"].get(name)

    @classmethod
    def get_model_class(cls, name: str) -> Optional[type]:
        return cls._lookup("model", name)

    @classmethod
    def get_processor_class(cls, name: str) -> Optional[type]:
        return cls._lookup("processor", name)

    @classmethod
    def list_models(cls):
        return sorted(cls.mapping["model_name_mapping"])

    @classmethod
    def list_processors(cls):
        return sorted(cls.mapping["processor_name_mapping"])


registry = Registry()
