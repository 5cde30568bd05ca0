"""The public training call on the GPU: LibraTrainWrapper.forward(samples) -- LibraTokenizer (text ids, CLIP tower, quant_conv,
LFQ pack, sync-free assembly) -> get_labels -> LibraForCausalLM -- against the oracle's restatement of the same pipeline
(modeling_libra.py:1397-1433, tokenization_libra.py:167-316, image_tokenizer.py:75-95), and the optimizer recipe
(trainer.py:27-85, libra_pretrain.yaml:81-85,116) against torch.optim.AdamW + clip_grad_norm_ + the HF cosine schedule."""
import math

import pytest
import torch

from gpu_util import need_gpu, rel_err
from oracle import libra_oracle as O

pytestmark = pytest.mark.gpu
dev = "cuda"


def _build(golden):
    from libra_b200.models import LibraConfig, LibraForCausalLM, LibraTokenizer, LibraTrainWrapper, SimpleTextTokenizer, VisionTokenizer
    from libra_b200.models.modeling_clip import CLIPVisionConfig
    gd, gc = golden("decoder_tiny"), golden("clip_tiny")
    cfg = dict(gd["config"])
    N = (gc["config"]["image_size"] // gc["config"]["patch_size"]) ** 2           # tokens per image of the tiny tower
    cfg.update(max_vision_token_length=N + 2, image_feature_resolution=int(math.isqrt(N)), contiguous_signal_size=2 * gc["config"]["hidden_size"])
    torch.manual_seed(3)
    module = LibraForCausalLM(LibraConfig(**cfg))
    from libra_b200 import synthetic
    synthetic.randomize_for_bench(module, seed=3, std=0.05)
    vt = VisionTokenizer(CLIPVisionConfig(**gc["config"]), select_layer=(-2, -3), embed_dim=18, token_offset=cfg["vocab_size"])
    vt.encoder.load_state_dict(gc["state_dict"], strict=False)
    torch.nn.init.normal_(vt.quant_conv.weight, std=0.3)
    torch.nn.init.normal_(vt.quant_conv.bias, std=0.1)
    vt = vt.to(torch.bfloat16).to(dev)
    tok = LibraTokenizer(text_tokenizer=SimpleTextTokenizer(vocab_size=cfg["vocab_size"], model_max_length=512), image_tokenizer=vt)
    w = LibraTrainWrapper(LibraConfig(**cfg), module=module.to(torch.bfloat16).to(dev), tokenizer=tok).train()
    return w, gc, cfg, N


def test_train_wrapper_forward_from_pixels_vs_oracle(golden):
    need_gpu()
    w, gc, cfg, N = _build(golden)
    px = gc["pixel_values"].to(dev)
    B = px.shape[0]
    ph = " ".join(["<img_ph>"] * (N + 2))
    texts = [f"describe this {ph} it shows a small dog on the grass", f"{ph} two cats"][:B] + ["plain text sample without any picture"] * max(0, B - 2)
    n_img = min(B, 2)
    samples = {"language": texts, "vision": [px[i] for i in range(n_img)],
               "label_mask_position_map": [[[3 + N + 2, 4 + N + 2]], [[1 + N + 2, 2 + N + 2]]][:B] + [[]] * max(0, B - 2)}
    out = w(samples, return_loss=True)
    assert out.loss is not None and torch.isfinite(out.loss)
    out.loss.backward()
    assert w.module.model.layers[0].mlp.vision_gate_proj.weight_A.grad is not None
    # ---- oracle pipeline on the tokenizer's own image ids / features (the bf16 tower's sign flips are measured elsewhere)
    tok = w.tokenizer
    inputs = tok(samples, return_tensors="pt", padding="longest", max_length=512, truncation=True)
    labels = w.get_labels(inputs, samples["label_mask_position_map"])
    text = tok.text_tokenizer(texts, return_tensors="pt", padding="longest")
    enc = tok.image_tokenizer(px[:n_img].to(torch.bfloat16))
    want_in = O.assemble_inputs(text["input_ids"], text["attention_mask"], tok.text_tokenizer.img_ph_token_id, enc["input_ids"].cpu(),
                                enc["encoder_feat"].cpu(), max_vision_token_length=N + 2)
    for k in ("input_ids", "attention_mask", "vision_indices", "coninous_signal"):
        assert torch.equal(inputs[k].cpu(), want_in[k]), k
    want_lab = O.get_labels(want_in["input_ids"], want_in["attention_mask"], tok.image_tokenizer.boi_token_id, 1, samples["label_mask_position_map"])
    assert torch.equal(labels.cpu(), want_lab)
    sd = {k: v.detach().float() for k, v in w.module.state_dict().items()}
    d = O.LibraDims.from_config(cfg)
    o32 = O.libra_forward(sd, d, want_in["input_ids"].to(dev), want_in["vision_indices"].to(dev), attention_mask=want_in["attention_mask"].to(dev),
                          contiguous_signal=want_in["coninous_signal"].to(dev).float(), labels=want_lab.to(dev))
    sd16 = {k: (v.bfloat16() if v.is_floating_point() else v) for k, v in sd.items()}
    o16 = O.libra_forward(sd16, d, want_in["input_ids"].to(dev), want_in["vision_indices"].to(dev), attention_mask=want_in["attention_mask"].to(dev),
                          contiguous_signal=want_in["coninous_signal"].to(dev).bfloat16(), labels=want_lab.to(dev))
    e_ours, e_orc = abs(float(out.loss) - float(o32["loss"])), abs(float(o16["loss"]) - float(o32["loss"]))
    assert e_ours <= 2.0 * e_orc + 2e-2, (float(out.loss), float(o32["loss"]), float(o16["loss"]))


def test_flat_adamw_recipe_vs_torch(golden):
    """decay exclusions, clip 1.0, cosine + warm-up: three optimizer steps of FlatAdamW.for_buffer == torch.optim.AdamW with
    the reference's parameter groups + clip_grad_norm_ + get_cosine_schedule_with_warmup (bf16 parameters/states)."""
    need_gpu()
    from transformers import get_cosine_schedule_with_warmup
    from libra_b200.dist import FlatGradBuffer
    from libra_b200.models import LibraConfig, LibraForCausalLM
    from libra_b200.optim import FlatAdamW, decay_parameter_names
    g = golden("decoder_tiny")
    torch.manual_seed(0)
    m1 = LibraForCausalLM(LibraConfig(**g["config"])).to(torch.bfloat16).to(dev)
    m2 = LibraForCausalLM(LibraConfig(**g["config"])).to(torch.bfloat16).to(dev)
    m2.load_state_dict(m1.state_dict())
    buf = FlatGradBuffer(m1.named_parameters(), flatten_weights=True, fused=False)
    opt1 = FlatAdamW.for_buffer(buf, m1, lr=1e-2, betas=(0.9, 0.99), weight_decay=0.01, max_grad_norm=1.0, total_steps=40, warmup_ratio=0.05)
    decay = set(decay_parameter_names(m2))
    opt2 = torch.optim.AdamW([{"params": [p for n, p in m2.named_parameters() if n in decay], "weight_decay": 0.01},
                              {"params": [p for n, p in m2.named_parameters() if n not in decay], "weight_decay": 0.0}],
                             lr=1e-2, betas=(0.9, 0.99), eps=1e-8)
    sch = get_cosine_schedule_with_warmup(opt2, num_warmup_steps=2, num_training_steps=40)
    gen = torch.Generator(device=dev).manual_seed(1)
    for step in range(3):
        for (n, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
            gr = (torch.randn(p1.shape, device=dev, generator=gen) * 0.05).to(torch.bfloat16)
            p1.grad.copy_(gr)
            p2.grad = gr.clone()
        norm2 = torch.nn.utils.clip_grad_norm_(m2.parameters(), 1.0)
        opt2.step(); sch.step()
        opt1.step()
        gn = opt1.grad_norm()
        assert abs(float(gn[0]) - float(norm2)) < 2e-2 * float(norm2)
        assert math.isclose(opt1.last_lr, [1e-2 * 0 / 2, 1e-2 * 1 / 2, 1e-2][step], rel_tol=1e-6, abs_tol=1e-12)
    for (n, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        # bf16 parameters: both sides round every step; compare within bf16 resolution of the accumulated update
        assert rel_err(p1, p2) < 1e-2, (n, rel_err(p1, p2))
        assert (p1.float() - p2.float()).abs().max() < 4e-2, n
