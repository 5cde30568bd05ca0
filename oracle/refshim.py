"""TEST INFRASTRUCTURE ONLY -- import shim for the *live* reference.

Makes ``/root/reference`` (pinned to transformers==4.38.2, omegaconf, ...)
importable under this image's transformers 5.x so that ``oracle/`` can be
validated against the reference itself and golden vectors can be generated
(see ``oracle/make_golden.py``).  Nothing in the product (``libra_b200/``)
imports this module, and it is never used on the GPU box (``/root/reference``
does not exist there).

The shims are the ones SURVEY.md section 8(c) lists; each only supplies a name
the reference imports at module scope and never calls on the hot path.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LIBRA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "libra", "models"))


def _stub_module(name: str, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def _placeholder(name):
    return type(name, (), {"__doc__": "placeholder injected by oracle/refshim.py"})


_installed = False


def install() -> None:
    """Idempotently install the shims and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    import transformers  # noqa: F401
    # Pre-import: otherwise a later lazy import re-creates the aliased
    # transformers.tokenization_utils module and drops the injected names.
    from transformers import CLIPImageProcessor  # noqa: F401
    import transformers.tokenization_utils as tu
    import re as _re

    # (1) transformers.onnx.OnnxConfig (configuration_clip.py:28)
    if "transformers.onnx" not in sys.modules:
        try:
            import transformers.onnx  # noqa: F401
        except Exception:
            onnx = _stub_module("transformers.onnx", OnnxConfig=_placeholder("OnnxConfig"))
            transformers.onnx = onnx

    # (2) names tokenization_libra.py:12 pulls from tokenization_utils
    for nm, val in (("TextInput", str), ("re", _re)):
        if not hasattr(tu, nm):
            setattr(tu, nm, val)
    if not hasattr(tu, "logger"):
        import logging
        tu.logger = logging.getLogger("transformers.tokenization_utils")
    if not hasattr(tu, "AddedToken"):
        from tokenizers import AddedToken
        tu.AddedToken = AddedToken

    # (3) generation.beam_constraints / beam_search (modeling_libra_utils.py:12-13)
    import transformers.generation as gen
    for modname, names in (
        ("beam_constraints", ["DisjunctiveConstraint", "PhrasalConstraint"]),
        ("beam_search", ["BeamScorer", "BeamSearchScorer", "ConstrainedBeamSearchScorer"]),
    ):
        full = f"transformers.generation.{modname}"
        try:
            __import__(full)
            mod = sys.modules[full]
        except Exception:
            mod = _stub_module(full)
            setattr(gen, modname, mod)
        for nm in names:
            if not hasattr(mod, nm):
                setattr(mod, nm, _placeholder(nm))

    # (4) logits processors removed in 5.x (modeling_libra_utils.py:15-40)
    import transformers.generation.logits_process as lp
    import ast
    src = open(os.path.join(REFERENCE_ROOT, "libra/models/libra/modeling_libra_utils.py")).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("transformers"):
            try:
                __import__(node.module)
                mod = sys.modules[node.module]
            except Exception:
                mod = _stub_module(node.module)
            for alias in node.names:
                if not hasattr(mod, alias.name):
                    setattr(mod, alias.name, _placeholder(alias.name))
    del lp

    # (6) omegaconf (tokenization_libra.py:10); only .load/.create exist and
    # neither is called on the synthetic path.
    if "omegaconf" not in sys.modules:
        try:
            import omegaconf  # noqa: F401
        except Exception:
            class _OmegaConf:
                @staticmethod
                def load(path):
                    raise RuntimeError("omegaconf stub: load() unavailable")

                @staticmethod
                def create(obj=None):
                    return obj
            _stub_module("omegaconf", OmegaConf=_OmegaConf)

    # (7) the reference itself
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load_reference_class(relpath: str, class_name: str, namespace: dict):
    """Execute ONE class definition of a reference file from the source where it lies (read at run time, never copied), for
    modules whose unrelated top-level imports are unavailable here (libra/data/processors/libra_processor.py pulls in
    omegaconf, torchvision and -- through the registry -- timm just to define `Expand2Square`).  `namespace` provides the
    names the class body needs."""
    import ast
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    path = os.path.join(REFERENCE_ROOT, relpath)
    tree = ast.parse(open(path).read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = dict(namespace)
            exec(compile(mod, path, "exec"), ns)
            return ns[class_name]
    raise KeyError(f"{class_name} not found in {relpath}")


def import_reference():
    """Return the reference modules the oracle is validated against."""
    install()
    import importlib
    mods = types.SimpleNamespace()
    mods.modeling_libra = importlib.import_module("libra.models.libra.modeling_libra")
    mods.modeling_clip = importlib.import_module("libra.models.clip.modeling_clip")
    mods.configuration_clip = importlib.import_module("libra.models.clip.configuration_clip")
    mods.configuration_libra = importlib.import_module("libra.models.libra.configuration_libra")
    mods.lfq = importlib.import_module(
        "libra.models.libra.taming.modules.quantization.lookup_free_quantization")
    mods.modeling_llama = importlib.import_module("libra.models.llama.modeling_llama")
    return mods
