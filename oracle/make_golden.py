"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt by RUNNING THE REFERENCE.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Each fixture stores seeded inputs, the weights (reference state-dict key
names) and the outputs the *unmodified reference modules* produced on CPU in
fp32.  ``tests/test_golden.py`` checks the oracle against them everywhere
(including the GPU box, where /root/reference does not exist) and the GPU
tests check the CUDA path against them.  Fixtures are small on purpose.
"""
from __future__ import annotations

import os
import sys

import torch

from oracle import refshim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY_LIBRA = dict(hidden_size=128, intermediate_size=352, num_hidden_layers=2, num_attention_heads=1,
                  vocab_size=320, contiguous_signal_size=32, max_position_embeddings=2048)
ATTN_HD128 = dict(hidden_size=256, intermediate_size=352, num_hidden_layers=1, num_attention_heads=2,
                  vocab_size=320, contiguous_signal_size=32, max_position_embeddings=2048)
TINY_CLIP = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2,
                 image_size=56, patch_size=14, num_channels=3)


def randomize_like_bench(model, seed: int):
    """randomise weight_B (bridges are zero-init) and norm weights so
    routing, low-rank and bridge paths are numerically exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("weight_B"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            elif "norm" in n and p.ndim == 1:
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))


def make_libra_inputs(V: int, S: int, B: int, n_text: int, pad_last: int, seed: int, images_per_sample=1):
    g = torch.Generator().manual_seed(seed)
    T = 1 + 578 * images_per_sample + n_text
    ids = torch.randint(3, V, (B, T), generator=g)
    ids[:, 0] = 1
    input_ids = ids[None].repeat(2, 1, 1)
    vi = torch.full((B, T), 578)
    spans = []
    for b in range(B):
        sp = []
        pos = 1 + 2 * b
        for _ in range(images_per_sample):
            input_ids[:, b, pos] = V + 512
            input_ids[:, b, pos + 577] = V + 513
            input_ids[:, b, pos + 1:pos + 577] = torch.randint(0, 512, (2, 576), generator=g) + V
            vi[b, pos:pos + 578] = torch.arange(578)
            if pos + 578 < T:
                sp.append([pos + 578, pos + 579])
            pos += 578 + 3
        spans.append(sp)
    am = torch.ones(B, T, dtype=torch.long)
    if pad_last:
        am[-1, -pad_last:] = 0
    sig = torch.randn(B, T, S, generator=g)
    sig[(vi >= 578) | (vi == 0) | (vi == 577)] = 0
    labels = input_ids.clone()
    labels[:, am == 0] = -100
    labels[labels == V + 512] = -100
    labels[labels == 1] = -100
    for b, sp in enumerate(spans):
        for s, e in sp:
            labels[:, b, s:e] = -100
    return dict(input_ids=input_ids, attention_mask=am, vision_indices=vi, contiguous_signal=sig, labels=labels)


def golden_decoder(m):
    torch.manual_seed(0)
    cfg = m.configuration_libra.LibraConfig(**TINY_LIBRA)
    model = m.modeling_libra.LibraForCausalLM(cfg).eval()
    randomize_like_bench(model, 11)
    inp = make_libra_inputs(cfg.vocab_size, cfg.contiguous_signal_size, B=2, n_text=40, pad_last=7, seed=5)
    out = model(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"],
                vision_indices=inp["vision_indices"], contiguous_signal=inp["contiguous_signal"],
                labels=inp["labels"], use_cache=False, output_hidden_states=True)
    out.loss.backward()
    logits = out.logits.detach()
    sel = torch.tensor([0, 1, 2, 3, 300, 577, 578, 579, 580, 581, 600, logits.shape[2] - 8, logits.shape[2] - 1])
    grads = {n: p.grad.clone() for n, p in model.named_parameters()
             if p.grad is not None and (("layers.1" in n and ("bridge" in n or "vision_q_proj" in n or n.endswith("q_proj.weight")
                                                               or "layernorm" in n or "down_proj" in n))
                                        or n in ("model.vision_signal_norm.weight", "vision_lm_head.heads.1.weight"))}
    keep = {k: v.clone() for k, v in model.state_dict().items()
            if "placeholder" not in k and "inv_freq" not in k}
    return dict(config=TINY_LIBRA, state_dict=keep, inputs=inp, loss=out.loss.detach(),
                logits_positions=sel, logits_at=logits[:, :, sel].clone(),
                logits_lse=torch.logsumexp(logits, dim=-1),
                hidden_after_layer0=out.hidden_states[1].detach().clone(),
                last_hidden=out.hidden_states[-1].detach().clone(), grads=grads)


DECODE_STEPS = [([7, 320 + 512], [578, 0]), ([9, 320 + 3], [578, 1]), ([11, 320 + 200], [578, 2]),
                ([5, 320 + 513], [578, 577]), ([8, 13], [578, 578])]


def golden_decode(m):
    """N1: the reference's KV-cached generation path on the decoder_tiny model (same construction => same weights as
    golden_decoder): prompt with use_cache=True, then one-token steps -- text continuation on sample 0; <img>, two grid
    tokens, </img> (which must predict "newline") and a text token on sample 1.  Keeps the logits of the last prompt
    position and of every step (fp32)."""
    torch.manual_seed(0)
    cfg = m.configuration_libra.LibraConfig(**TINY_LIBRA)
    model = m.modeling_libra.LibraForCausalLM(cfg).eval()
    randomize_like_bench(model, 11)
    inp = make_libra_inputs(cfg.vocab_size, cfg.contiguous_signal_size, B=2, n_text=9, pad_last=0, seed=5)
    ids, vi, am = inp["input_ids"], inp["vision_indices"], inp["attention_mask"]
    g = torch.Generator().manual_seed(3)
    logits, tokens = [], []
    with torch.no_grad():
        r = model(input_ids=ids, attention_mask=am, vision_indices=vi, contiguous_signal=inp["contiguous_signal"], use_cache=True)
        logits.append(r.logits[:, :, -1].clone())
        past = r.past_key_values
        for tok, vidx in DECODE_STEPS:
            nid = torch.tensor(tok)[None, :, None].repeat(2, 1, 1)
            nid[1] = torch.where(nid[0] >= cfg.vocab_size, torch.randint(cfg.vocab_size, cfg.vocab_size + 512, nid[0].shape, generator=g), nid[0])
            nvi = torch.tensor(vidx)[:, None]
            am = torch.cat([am, am.new_ones(am.shape[0], 1)], dim=1)
            pos = (am.long().cumsum(-1) - 1)[:, -1:]
            r = model(input_ids=nid, attention_mask=am, vision_indices=nvi, position_ids=pos, past_key_values=past, use_cache=True)
            past = r.past_key_values
            logits.append(r.logits[:, :, -1].clone())
            tokens.append((nid.clone(), nvi.clone()))
    return dict(config=TINY_LIBRA, inputs={k: inp[k] for k in ("input_ids", "attention_mask", "vision_indices", "contiguous_signal")},
                step_input_ids=torch.stack([t[0] for t in tokens]), step_vision_indices=torch.stack([t[1] for t in tokens]),
                logits=torch.stack(logits), k_for_language_l1=past[1][0][1][:, :, -8:].clone(),
                k_for_vision_l1=past[1][0][0][:, :, -8:].clone())


TINY_VQ_DECODER = dict(ch=32, out_ch=3, ch_mult=(1, 2, 2), num_res_blocks=1, attn_resolutions=(6,), dropout=0.0, in_channels=3,
                       resolution=48, z_channels=32, initial_resolution=6, num_attn_head=1)


def golden_vq_decode(m):
    """N2: ids -> pixels through the reference's taming Decoder + LFQ.indices_to_codes + post_quant_conv, chained as
    VQModel.decode_code does (taming/models/vqgan.py:122-130) behind ImageTokenizer.decode (image_tokenizer.py:97-124).
    embed_dim 24 so that the LFQ's project_out Linear is exercised."""
    import importlib
    dm = importlib.import_module("libra.models.libra.taming.modules.diffusionmodules.model")
    lfq = importlib.import_module("libra.models.libra.taming.modules.quantization.lookup_free_quantization")
    torch.manual_seed(7)
    dec = dm.Decoder(**TINY_VQ_DECODER).eval()
    quant = lfq.LFQ(dim=24, codebook_size=512, num_codebooks=2, entropy_loss_weight=0.1, commitment_loss_weight=1.,
                    diversity_gamma=2.5).eval()
    pqc = torch.nn.Conv2d(24, TINY_VQ_DECODER["z_channels"], 1)
    offset, boi = 32000, 32512
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, 512, (2, 2, 36), generator=g) + offset
    ids = torch.cat([torch.full((2, 2, 1), boi), ids, torch.full((2, 2, 1), boi + 1)], dim=2)
    with torch.no_grad():
        code = (ids[:, :, 1:-1].reshape(2, 2, 6, 6).permute(1, 2, 3, 0) - offset)
        z = quant.indices_to_codes(code)
        pixels = dec(pqc(z))
    sd = {f"decoder.{k}": v.clone() for k, v in dec.state_dict().items()}
    sd.update({f"post_quant_conv.{k}": v.clone() for k, v in pqc.state_dict().items()})
    sd.update({f"quantize.{k}": v.clone() for k, v in quant.state_dict().items() if k.startswith("project_out")})
    return dict(config=TINY_VQ_DECODER, state_dict=sd, ids=ids, token_offset=offset, boi_token_id=boi, codebook_size=512,
                codes=z.clone(), pixels=pixels.clone())


def golden_attention(m):
    """One LibraAttention module at the production head_dim (128), 2 heads."""
    torch.manual_seed(1)
    cfg = m.configuration_libra.LibraConfig(**ATTN_HD128)
    attn = m.modeling_libra.LibraAttention(cfg).eval()
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for n, p in attn.named_parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.05 if n.endswith("weight_B") or "weight_A" in n else 0.06))
    B, T = 2, 192
    x = torch.randn(B, T, cfg.hidden_size, generator=g)
    flag = torch.zeros(B, T, dtype=torch.bool)
    flag[0, 1:120] = True
    flag[1, 30:150] = True
    am = torch.ones(B, T, dtype=torch.long)
    am[1, -20:] = 0
    mask = m.modeling_llama._make_causal_mask((B, T), x.dtype, x.device) + m.modeling_llama._expand_mask(am, x.dtype, T)
    pos = torch.arange(T)[None].expand(B, T)
    x.requires_grad_(True)
    y, _, _ = attn(hidden_states=x, attention_mask=mask, position_ids=pos, vision_flag=flag)
    gy = torch.randn(y.shape, generator=g)
    (y * gy).sum().backward()
    sd = {"self_attn." + k: v.detach().clone() for k, v in attn.state_dict().items() if "inv_freq" not in k}
    return dict(config=ATTN_HD128, state_dict=sd, x=x.detach(), flag=flag, attention_mask=am, grad_out=gy,
                out=y.detach(), grad_x=x.grad.clone(),
                grad_kbridge_B=attn.vision_k_bridge_on_language.weight_B.grad.clone(),
                grad_vbridge_A=attn.vision_v_bridge_on_vision.weight_A.grad.clone(),
                grad_vq_A=attn.vision_q_proj.weight_A.grad.clone())


def golden_clip(m):
    torch.manual_seed(2)
    cfg = m.configuration_clip.CLIPVisionConfig(**TINY_CLIP)
    model = m.modeling_clip.CLIPVisionModel(cfg).eval()
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "norm" in n or n.endswith("bias"):
                p.copy_((1.0 if n.endswith("weight") else 0.0) + 0.1 * torch.randn(p.shape, generator=g))
    px = torch.rand(3, 3, 56, 56, generator=g)
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
    px = (px - mean) / std
    with torch.no_grad():
        out = model(px, output_hidden_states=True)
    return dict(config=TINY_CLIP, state_dict={k: v.clone() for k, v in model.state_dict().items() if "position_ids" not in k},
                pixel_values=px, hidden_states=[h.clone() for h in out.hidden_states])


def golden_lfq(m):
    lfq = m.lfq.LFQ(dim=18, codebook_size=512, num_codebooks=2).eval()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 18, 24, 24, generator=g)
    x[0, :, 0, 0] = 0.0                        # exact zeros quantise to bit 0 (x > 0 is strict)
    x[0, 3, 1, 1] = -0.0
    with torch.no_grad():
        quant, _, idx = lfq(x)
        codes = lfq.indices_to_codes(idx)
    lfq_p = m.lfq.LFQ(dim=32, codebook_size=512, num_codebooks=2).eval()
    xp = torch.randn(2, 32, 6, 6, generator=g)
    with torch.no_grad():
        _, _, idx_p = lfq_p(xp)
    return dict(x=x, indices=idx, quantized=quant, codes=codes,
                xp=xp, proj_w=lfq_p.project_in.weight.detach().clone(), proj_b=lfq_p.project_in.bias.detach().clone(),
                indices_p=idx_p)


def golden_norms_rope(m):
    g = torch.Generator().manual_seed(13)
    x = torch.randn(37, 256, generator=g) * 3
    w = 1 + 0.1 * torch.randn(256, generator=g)
    norm = m.modeling_llama.LlamaRMSNorm(256, eps=1e-6)
    with torch.no_grad():
        norm.weight.copy_(w)
    y32 = norm(x).detach()
    ybf = norm(x.bfloat16()).detach()
    rot = m.modeling_llama.LlamaRotaryEmbedding(128, max_position_embeddings=2048)
    q = torch.randn(1, 2, 40, 128, generator=g)
    k = torch.randn(1, 2, 40, 128, generator=g)
    cos, sin = rot(q, seq_len=40)
    pos = torch.arange(40)[None]
    qe, ke = m.modeling_libra.apply_rotary_pos_emb(q, [k, k * 2], cos, sin, pos)
    return dict(x=x, w=w, y32=y32, ybf=ybf, q=q, k=k, q_rot=qe, k_rot=ke[0], k2_rot=ke[1])


def golden_clip_preprocess(m):
    """The reference's own image processors (libra/models/clip/image_processing_clip.py CLIPImageProcessor; libra/data/
    processors/libra_processor.py:44-60 Expand2Square with the mean colour) on small synthetic uint8 images: plain and
    padded-to-square pixel_values [3,336,336] float32.  Small sources keep the fixture small; up- and down-scaling, portrait
    and landscape, smooth and noisy content."""
    import importlib
    import numpy as np
    from PIL import Image
    ip = importlib.import_module("libra.models.clip.image_processing_clip")
    Expand2Square = refshim.load_reference_class("libra/data/processors/libra_processor.py", "Expand2Square",
                                                  {"torch": torch, "Image": Image})
    P = ip.CLIPImageProcessor(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
    bg = tuple(int(x * 255) for x in P.image_mean)
    # the processor's float32 output takes 256 values per channel (a function of the uint8 resize result): store it as the
    # table the REFERENCE produces on a 0..255 ramp plus uint8 indices, after checking that the pair reproduces the
    # reference's pixel_values exactly -- 0.3 MB per image instead of 1.4 MB
    ramp = np.repeat(np.arange(256, dtype=np.uint8)[:, None, None], 3, axis=2)                      # [256, 1, 3]
    lut = P.normalize(P.rescale(ramp, scale=P.rescale_factor), mean=P.image_mean, std=P.image_std)   # [256, 1, 3] float32
    lut = np.ascontiguousarray(lut[:, 0, :].T).astype(np.float32)                                    # [3, 256]

    def pack(pv):
        idx = np.stack([np.abs(pv[c][..., None] - lut[c][None, None, :]).argmin(-1) for c in range(3)], -1).astype(np.uint8)
        back = np.stack([lut[c][idx[..., c]] for c in range(3)])
        assert np.array_equal(back, pv), "the table form must reproduce the reference output bit for bit"
        return torch.from_numpy(idx)

    rng = np.random.default_rng(5)
    images, pv, pvs = [], [], []
    for i, (h, w) in enumerate([(61, 90), (120, 75), (400, 523)]):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if i == 1:
            yy, xx = np.mgrid[0:h, 0:w]
            img = np.stack([yy * 255 // (h - 1), xx * 255 // (w - 1), (yy * 3 + xx * 5) % 256], -1).astype(np.uint8)
        images.append(torch.from_numpy(img))
        pv.append(pack(P(img, return_tensors="np")["pixel_values"][0]))
        sq = Expand2Square(bg)(Image.fromarray(img))
        pvs.append(pack(P(sq, return_tensors="np")["pixel_values"][0]))
    return dict(images=images, index=pv, index_square=pvs, lut=torch.from_numpy(lut), background=list(bg))


def main():
    m = refshim.import_reference()
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, fn in (("decoder_tiny", golden_decoder), ("decode_tiny", golden_decode), ("vq_decode_tiny", golden_vq_decode),
                     ("attention_hd128", golden_attention),
                     ("clip_tiny", golden_clip), ("lfq", golden_lfq), ("norms_rope", golden_norms_rope),
                     ("clip_preprocess", golden_clip_preprocess)):
        if only and name not in only:
            continue
        obj = fn(m)
        path = os.path.join(OUT, name + ".pt")
        torch.save(obj, path)
        print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    sys.exit(main())
