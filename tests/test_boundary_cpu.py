"""The reference-facing surface that needs no GPU: import-path aliases + registry (libra/models/__init__.py:1-5,
modeling_libra.py:1292, train.py:28-30), LibraTokenizer.forward's assembly against the oracle's restatement of
tokenization_libra.py:167-316, checkpoint-directory construction of LibraTrainWrapper, and the optimizer policy
(trainer.py:27-85, libra_pretrain.yaml:81-85)."""
import json
import math
import os
import subprocess
import sys

import pytest
import torch

from oracle import libra_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_compat_aliases_and_registry_in_a_clean_process():
    code = (
        "import libra_b200.compat as c; c.install()\n"
        "from libra.models import *\n"
        "from libra.models.libra import LibraForCausalLM, LibraTokenizer, LibraConfig\n"
        "from libra.models.clip import CLIPVisionModel, CLIPImageProcessor, CLIPVisionConfig\n"
        "from libra.common.registry import registry\n"
        "import libra.models.libra.modeling_libra as m\n"
        "cls = registry.get_model_class('libra_train_wrapper')\n"
        "assert cls is LibraTrainWrapper and cls.__module__.startswith('libra_b200'), cls\n"
        "assert m.LibraForCausalLM is LibraForCausalLM and hasattr(cls, 'from_config') and hasattr(cls, 'get_optimizer_parameters')\n"
        "from libra.data.processors.libra_processor import LibraImageProcessor, LibraEvalImageProcessor\n"
        "assert registry.get_processor_class('libra_image') is LibraImageProcessor\n"
        "assert registry.get_processor_class('libra_image_eval') is LibraEvalImageProcessor\n"
        "print('ALIASES_OK')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert "ALIASES_OK" in r.stdout, r.stderr[-2000:]


class _FakeImageTokenizer(torch.nn.Module):
    """ids / features by formula: what LibraTokenizer.forward needs from an image tokenizer, without the CLIP tower."""

    def __init__(self, L=10, S=6, offset=320):
        super().__init__()
        self.max_vision_token_length, self.num_codebook, self.S = L, 2, S
        self.boi_token_id, self.eoi_token_id = offset + 512, offset + 513
        self.offset = offset
        self.w = torch.nn.Parameter(torch.zeros(1))

    device = property(lambda self: self.w.device)
    dtype = property(lambda self: self.w.dtype)

    def forward(self, images):
        n, L = images.shape[0], self.max_vision_token_length
        g = torch.Generator().manual_seed(int(images.sum().item() * 1000) % 100000)
        ids = torch.randint(0, 512, (2, n, L), generator=g) + self.offset
        ids[:, :, 0], ids[:, :, -1] = self.boi_token_id, self.eoi_token_id
        return {"input_ids": ids, "encoder_feat": torch.randn(n, L - 2, self.S, generator=g)}


def _samples(L):
    ph = " ".join(["<img_ph>"] * L)
    return {"language": [f"a cute dog {ph} and a cat {ph} . I like them", f"look {ph} nice", "no picture here at all"],
            "vision": [torch.full((3, 4, 4), 0.1), torch.full((3, 4, 4), 0.2), torch.full((3, 4, 4), 0.3)],
            "label_mask_position_map": [[[1, 3]], [[2, 4]], []]}


@pytest.mark.parametrize("side", ["right", "left"])
def test_libra_tokenizer_forward_matches_oracle_assembly(side):
    from libra_b200.models.tokenization_libra import LibraTokenizer, SimpleTextTokenizer
    L = 10
    tt = SimpleTextTokenizer(vocab_size=320, padding_side=side)
    it = _FakeImageTokenizer(L=L)
    tok = LibraTokenizer(text_tokenizer=tt, image_tokenizer=it)
    assert tt.img_ph_token_id == 320 and tt.img_gen_token_id == 321 and tt.pad_token_id == tt.unk_token_id == 0
    s = _samples(L)
    out = tok(s, return_tensors="pt", padding="longest", max_length=2048, truncation=True)
    assert set(out.keys()) == {"input_ids", "attention_mask", "vision_indices", "coninous_signal"}
    # oracle: the reference's scatter formulation on the same tokenised text and the same image tokens
    text = tt(s["language"], return_tensors="pt", padding="longest")
    enc = it(torch.stack(s["vision"]))
    want = O.assemble_inputs(text["input_ids"], text["attention_mask"], 320, enc["input_ids"], enc["encoder_feat"], max_vision_token_length=L)
    for k in ("input_ids", "attention_mask", "vision_indices", "coninous_signal"):
        assert torch.equal(out[k], want[k]), k
    assert out["input_ids"].shape[0] == 2 and (out["input_ids"][0] == it.boi_token_id).sum() == 3
    # truncation and list-of-samples input
    short = tok([{"language": s["language"][:2], "vision": s["vision"][:2]}, {"language": s["language"][2]}, {"vision": s["vision"][2]}],
                padding="longest", truncation=True, max_length=17)
    assert short["input_ids"].shape[2] == 17
    # generation prompt: <img_gen> becomes BOI with vision index 0 (tokenization_libra.py:249-251, 274-275)
    gen = tok({"language": ["draw a dog <img_gen>"]}, padding="longest")
    assert gen["input_ids"][0, 0, -1] == it.boi_token_id and gen["vision_indices"][0, -1] == 0 and gen["coninous_signal"] is None
    assert tok.batch_decode(text["input_ids"][2:3])[0].startswith("no picture")
    with pytest.raises(ValueError):
        tok({"language": ["x"]}, return_tensors="np")


def test_train_wrapper_from_checkpoint_directory(tmp_path):
    """The reference's construction path: LibraTrainWrapper.from_config(model_cfg) with model_cfg.pretrained -> config.json,
    weights, HF text tokenizer files and vision_tokenizer_config.yaml (modeling_libra.py:1294-1304, tokenization_libra.py:141-165)."""
    import yaml
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    from libra_b200.models import LibraConfig, LibraForCausalLM, LibraTrainWrapper
    from libra_b200.models.modeling_clip import CLIPVisionConfig, CLIPVisionModel
    from libra_b200.registry import registry
    d = str(tmp_path)
    cfg = LibraConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2, vocab_size=320,
                      contiguous_signal_size=64, max_vision_token_length=18, image_feature_resolution=4)
    torch.manual_seed(0)
    LibraForCausalLM(cfg).save_pretrained(d)
    vocab = {"<unk>": 0, "<s>": 1, "</s>": 2, **{f"w{i}": 3 + i for i in range(317)}}
    t = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    t.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    PreTrainedTokenizerFast(tokenizer_object=t, unk_token="<unk>", bos_token="<s>", eos_token="</s>", model_max_length=128).save_pretrained(d)
    clip_cfg = CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=2, image_size=56, patch_size=14)
    CLIPVisionModel(clip_cfg).save_pretrained(os.path.join(d, "clip_tower"))
    yaml.safe_dump({"max_vision_token_length": 18, "freeze": True,
                    "params": {"embed_dim": 18, "codebook_size": 512, "num_codebook": 2, "ckpt_path": "vq_f14_.ckpt",
                               "ddconfig": {"encoder_name": "clip_tower", "select_layer": [-2, -3], "z_channels": 32}}},
                   open(os.path.join(d, "vision_tokenizer_config.yaml"), "w"))
    model_cfg = {"arch": "libra_train_wrapper", "pretrained": d, "model_kwargs": {"frozen_language": True}}
    w = registry.get_model_class(model_cfg["arch"]).from_config(model_cfg)
    assert isinstance(w, LibraTrainWrapper) and w.config.hidden_size == 64
    assert w.tokenizer.text_tokenizer.img_ph_token_id == 320 and w.tokenizer.image_tokenizer.max_vision_token_length == 18
    assert w.tokenizer.image_tokenizer.boi_token_id == 320 + 512
    emb = w.module.get_input_embeddings().weight
    assert torch.equal(emb[w.tokenizer.text_tokenizer.pad_token_id], emb[2])              # change_pad_token_to_eos (:1390-1395)
    trainable = [n for n, p in w.module.named_parameters() if p.requires_grad]
    assert trainable and all("vision" in n for n in trainable)                           # frozen_language (:1342-1346)
    groups = w.get_optimizer_parameters()
    assert len(groups) == 2 and groups[0]["use_weight_decay"] and not groups[1]["use_weight_decay"]
    # no-decay group = the norm weights; the only 1-D tensor that decays is vision_hidden_placeholder (not inside a norm module)
    assert all(p.ndim == 1 for p in groups[1]["params"]) and sum(p.ndim < 2 for p in groups[0]["params"]) == 1


def test_optimizer_policy_matches_the_reference_recipe():
    from transformers import get_cosine_schedule_with_warmup
    from transformers.trainer_pt_utils import get_parameter_names
    from libra_b200.models import LibraConfig, LibraForCausalLM
    from libra_b200.models.modeling_libra import LlamaRMSNorm
    from libra_b200.optim import cosine_with_warmup, decay_parameter_names
    m = LibraForCausalLM(LibraConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2, vocab_size=320,
                                     contiguous_signal_size=64))
    want = [n for n in get_parameter_names(m, [torch.nn.LayerNorm, LlamaRMSNorm]) if "bias" not in n]      # trainer.py:27-37
    assert sorted(decay_parameter_names(m)) == sorted(want)
    assert not any("norm" in n for n in want) and any("weight_A" in n for n in want)
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    sch = get_cosine_schedule_with_warmup(opt, num_warmup_steps=5, num_training_steps=100)
    for s in range(100):
        assert math.isclose(cosine_with_warmup(s, 100, 5), sch.get_last_lr()[0], rel_tol=1e-6, abs_tol=1e-9), s
        opt.step(); sch.step()
