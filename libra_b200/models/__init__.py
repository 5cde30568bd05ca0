"""Host-side mirror of the reference's `libra.models` package surface (libra/models/__init__.py,
libra/models/libra/__init__.py): same class names, running on the libra_b200 CUDA kernels."""
from .configuration_libra import LibraConfig
from .modeling_libra import (LibraCausalLMOutputWithPast, LibraDecoderLayer, LibraForCausalLM, LibraLinear, LibraModel, LibraTrainWrapper,
                             LibraPreTrainedModel)

from .tokenization_libra import LibraTokenizer, SimpleTextTokenizer, VisionTokenizer
from .modeling_clip import CLIPVisionConfig, CLIPVisionModel
from .vq_decoder import VQDecoder

__all__ = ["VQDecoder", "LibraTokenizer", "SimpleTextTokenizer", "VisionTokenizer", "CLIPVisionConfig", "CLIPVisionModel", "LibraConfig", "LibraForCausalLM", "LibraTrainWrapper", "LibraModel", "LibraDecoderLayer", "LibraLinear", "LibraPreTrainedModel",
           "LibraCausalLMOutputWithPast"]
