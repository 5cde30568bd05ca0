"""Import-path compatibility with the reference: `libra.models`, `libra.models.libra`, `libra.models.clip`,
`libra.common.registry` resolve to the libra_b200 classes, so the reference's train.py / notebook imports
(`from libra.models import *`, `from libra.models.libra import LibraForCausalLM, LibraTokenizer, LibraConfig`,
`from libra.common.registry import registry`, `from libra.models.clip import CLIPVisionModel, CLIPImageProcessor,
CLIPVisionConfig`; libra/models/__init__.py:1-5, libra/models/libra/__init__.py:1-3, libra/models/clip/__init__.py:1-3)
run against the CUDA path without edits:

    import libra_b200.compat; libra_b200.compat.install()          # before the first `import libra...`
    python -m libra_b200.compat train.py --cfg-path ...             # or: run a script with the aliases installed

The aliases are sys.modules entries, not a `libra/` directory: a same-named package in this repository would shadow the
reference wherever both are importable (the parity tests import the real one).  install() refuses to replace a `libra`
that is already imported from elsewhere unless force=True."""
from __future__ import annotations

import runpy
import sys
import types


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__libra_b200_alias__ = True
    return m


def install(force: bool = False):
    cur = sys.modules.get("libra")
    if cur is not None and not getattr(cur, "__libra_b200_alias__", False) and not force:
        raise RuntimeError("a different `libra` package is already imported; call install(force=True) to replace it")
    from . import models, processors, registry as reg
    from .models import configuration_libra, modeling_clip, modeling_libra, tokenization_libra
    from transformers import CLIPImageProcessor          # the reference's image_processing_clip.py is HF's class verbatim

    m_libra = _module("libra.models.libra", LibraForCausalLM=models.LibraForCausalLM, LibraTokenizer=tokenization_libra.LibraTokenizer,
                      LibraConfig=models.LibraConfig, modeling_libra=modeling_libra, tokenization_libra=tokenization_libra,
                      configuration_libra=configuration_libra)
    m_clip = _module("libra.models.clip", CLIPVisionModel=modeling_clip.CLIPVisionModel, CLIPImageProcessor=CLIPImageProcessor,
                     CLIPVisionConfig=modeling_clip.CLIPVisionConfig, modeling_clip=modeling_clip,
                     CLIPImageProcessorCUDA=processors.CLIPImageProcessor)     # same pipeline on the GPU, bit-exact (N4)
    m_models = _module("libra.models", LibraTrainWrapper=models.LibraTrainWrapper, libra=m_libra, clip=m_clip,
                       __all__=["LibraTrainWrapper"])
    m_reg = _module("libra.common.registry", registry=reg.registry, Registry=reg.Registry)
    m_common = _module("libra.common", registry=m_reg)
    # libra/data/processors/libra_processor.py:65-111 ("libra_image", "libra_image_eval"): the GPU processors
    m_lp = _module("libra.data.processors.libra_processor", LibraImageProcessor=processors.LibraImageProcessor,
                   LibraEvalImageProcessor=processors.LibraEvalImageProcessor)
    m_procs = _module("libra.data.processors", libra_processor=m_lp, LibraImageProcessor=processors.LibraImageProcessor,
                      LibraEvalImageProcessor=processors.LibraEvalImageProcessor)
    m_data = _module("libra.data", processors=m_procs)
    root = _module("libra", models=m_models, common=m_common, data=m_data)
    root.__path__ = []                                    # a package: `import libra.models.libra` walks sys.modules
    for m in (m_models, m_common, m_data, m_procs):
        m.__path__ = []
    sys.modules.update({
        "libra": root, "libra.models": m_models, "libra.models.libra": m_libra, "libra.models.clip": m_clip,
        "libra.models.libra.modeling_libra": modeling_libra, "libra.models.libra.tokenization_libra": tokenization_libra,
        "libra.models.libra.configuration_libra": configuration_libra, "libra.models.clip.modeling_clip": modeling_clip,
        "libra.common": m_common, "libra.common.registry": m_reg,
        "libra.data": m_data, "libra.data.processors": m_procs, "libra.data.processors.libra_processor": m_lp,
    })
    return root


def uninstall():
    for k in [k for k, v in sys.modules.items() if k == "libra" or k.startswith("libra.")]:
        if getattr(sys.modules[k], "__libra_b200_alias__", False) or getattr(sys.modules[k], "__name__", "").startswith("libra_b200"):
            del sys.modules[k]


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m libra_b200.compat <script.py> [args...]")
    install()
    sys.argv = sys.argv[1:]
    runpy.run_path(sys.argv[0], run_name="__main__")
