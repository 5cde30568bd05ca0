"""Measured parity of the CUDA path against the oracle at BASELINE widths -> JSON (committed as profiles/r02_parity.json).

One FULL-WIDTH Libra-11B decoder layer (hidden 4096, 32 heads x 128, intermediate 11008, vocab 32000 + 514, bridge rank 8)
inside LibraForCausalLM (embeddings, final norms, all three heads, loss), T = 2048, B = 1, one image + right padding;
forward and backward.  Three runs on the same weights / inputs:
    ours  = libra_b200 CUDA path, bf16 (the product)
    o16   = oracle/libra_oracle.py in bf16 on the GPU (the arithmetic of the reference's training dtype; noise floor)
    o32   = the same oracle in fp32 (the reference value)
Reported per tensor: e_ours = |ours - o32| / |o32| (Frobenius), e_o16 likewise, max-abs errors, and for the logits the
north-star statistic: the fraction of finite logits with |ours - o32| <= 1e-3 * max(|o32|, 1) next to the same fraction for o16.

    python scripts/parity_report.py [--out gpurun_out/parity.json] [--seq 2048]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libra_b200 import synthetic  # noqa: E402
from libra_b200.models import LibraConfig, LibraForCausalLM  # noqa: E402
from oracle import libra_oracle as O  # noqa: E402   (the checker; never on the product path)

dev = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def full_width_layer(seq: int, layers: int = 1, pad: int = 37, seed: int = 0):
    cfg = LibraConfig(num_hidden_layers=layers)
    torch.manual_seed(seed)
    model = LibraForCausalLM(cfg).to(dev)
    synthetic.randomize_for_bench(model, seed=seed)
    # all three runs share the same, bf16-representable weights: the errors below are arithmetic, not weight rounding
    sd32 = {k: (v.detach().bfloat16().float() if v.is_floating_point() else v.detach().clone()) for k, v in model.state_dict().items()}
    model = model.to(torch.bfloat16).train()
    inp = synthetic.libra_batch(1, seq, 1, vocab=cfg.vocab_size, signal=cfg.contiguous_signal_size, seed=77, device=dev,
                                signal_dtype=torch.float32)
    am = inp["attention_mask"].clone()
    am[:, seq - pad:] = 0                                   # right padding: ragged batch as the parity layout
    inp["attention_mask"] = am
    inp["labels"][:, am == 0] = -100
    return cfg, model, sd32, inp


def run_ours(model, inp, with_logits):
    model.zero_grad(set_to_none=True)
    kw = dict(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], vision_indices=inp["vision_indices"],
              contiguous_signal=inp["contiguous_signal"], labels=inp["labels"])
    out = model(**kw)
    out.loss.backward()
    grads = {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None}
    logits = None
    if with_logits:
        with torch.no_grad():
            logits = model(**kw, return_logits=True).logits
    return float(out.loss), logits, grads


def run_oracle(sd32, cfg, inp, dtype):
    sd = {k: (v.to(dtype).requires_grad_(True) if v.is_floating_point() else v) for k, v in sd32.items()}
    d = O.LibraDims.from_config(cfg)
    o = O.libra_forward(sd, d, inp["input_ids"], inp["vision_indices"], attention_mask=inp["attention_mask"],
                        contiguous_signal=inp["contiguous_signal"].to(dtype), labels=inp["labels"])
    o["loss"].backward()
    grads = {k: v.grad.detach().float() for k, v in sd.items() if v.is_floating_point() and v.grad is not None}
    return float(o["loss"]), o["logits"].detach(), grads


def report(seq: int):
    cfg, model, sd32, inp = full_width_layer(seq)
    l_ours, lg_ours, g_ours = run_ours(model, inp, True)
    l16, lg16, g16 = run_oracle(sd32, cfg, inp, torch.bfloat16)
    l32, lg32, g32 = run_oracle(sd32, cfg, inp, torch.float32)
    valid = inp["attention_mask"].bool()[None, :, :, None]
    fin = torch.isfinite(lg32) & valid
    same_inf = bool(torch.equal(torch.isfinite(lg32) | ~valid, torch.isfinite(lg_ours) | ~valid))
    a, b16, c = lg_ours.float()[fin], lg16.float()[fin], lg32[fin]
    tol = 1e-3 * c.abs().clamp(min=1.0)
    out = {
        "config": {"hidden": cfg.hidden_size, "heads": cfg.num_attention_heads, "intermediate": cfg.intermediate_size,
                   "layers": cfg.num_hidden_layers, "vocab": cfg.vocab_size + cfg.vision_vocab_size, "seq": seq, "batch": 1,
                   "images": 1, "right_padding": 37},
        "loss": {"ours_bf16": l_ours, "oracle_bf16": l16, "oracle_fp32": l32, "abs_err_ours": abs(l_ours - l32),
                 "abs_err_oracle_bf16": abs(l16 - l32)},
        "logits": {"inf_pattern_matches": same_inf, "n": int(c.numel()),
                   "rel_fro_ours_vs_fp32": rel(a, c), "rel_fro_oracle_bf16_vs_fp32": rel(b16, c),
                   "max_abs_ours": (a - c).abs().max().item(), "max_abs_oracle_bf16": (b16 - c).abs().max().item(),
                   "frac_within_1e-3_rel_ours": ((a - c).abs() <= tol).float().mean().item(),
                   "frac_within_1e-3_rel_oracle_bf16": ((b16 - c).abs() <= tol).float().mean().item(),
                   "rms_logit": c.pow(2).mean().sqrt().item()},
        "grads": {},
    }
    worst = (None, 0.0)
    for n, want in g32.items():
        if n not in g_ours:
            continue
        e_o, e_16 = rel(g_ours[n], want), rel(g16[n], want)
        out["grads"][n] = {"rel_fro_ours_vs_fp32": e_o, "rel_fro_oracle_bf16_vs_fp32": e_16}
        if e_o > worst[1]:
            worst = (n, e_o)
    out["grads_worst"] = {"name": worst[0], "rel_fro_ours_vs_fp32": worst[1]}
    out["grads_median_ratio_ours_over_oracle_bf16"] = float(torch.tensor(
        [v["rel_fro_ours_vs_fp32"] / max(v["rel_fro_oracle_bf16_vs_fp32"], 1e-12) for v in out["grads"].values()]).median())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--seq", type=int, default=2048)
    args = ap.parse_args()
    r = report(args.seq)
    r["gpu"] = torch.cuda.get_device_name(0)
    txt = json.dumps(r, indent=1)
    print(txt)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt)


if __name__ == "__main__":
    main()
