"""Repeat forward+backward of the tiny golden decoder (with and without gradient checkpointing, frozen language) and report
every parameter whose gradient is not bit-identical between two runs -- hunting nondeterminism (races, atomics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libra_b200.models import LibraConfig, LibraForCausalLM

dev = "cuda"
g = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "decoder_tiny.pt"), map_location="cpu", weights_only=True)
inp = {k: v.to(dev) for k, v in g["inputs"].items()}


def build():
    m = LibraForCausalLM(LibraConfig(**g["config"]))
    m.load_state_dict(g["state_dict"], strict=False)
    m = m.to(torch.bfloat16).to(dev).train()
    for n, p in m.named_parameters():
        p.requires_grad = "vision" in n
    return m


m1, m2 = build(), build()
m2.gradient_checkpointing_enable()
kw = dict(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], vision_indices=inp["vision_indices"],
          contiguous_signal=inp["contiguous_signal"], labels=inp["labels"])
bad = {}
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ref = None
for it in range(N):
    for m in (m1, m2):
        m.zero_grad(set_to_none=True)
        m(**kw).loss.backward()
    torch.cuda.synchronize()
    cur = {n: p.grad.clone() for n, p in m1.named_parameters() if p.grad is not None}
    for (n, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        if p1.grad is None:
            continue
        if not torch.equal(p1.grad, p2.grad):
            d = (p1.grad.float() - p2.grad.float()).abs()
            bad.setdefault(("ckpt", n), []).append((it, float(d.max()), int((d > 0).sum())))
        if ref is not None and not torch.equal(ref[n], cur[n]):
            d = (ref[n].float() - cur[n].float()).abs()
            bad.setdefault(("rerun", n), []).append((it, float(d.max()), int((d > 0).sum())))
    if ref is None:
        ref = cur
print(f"{N} iterations; parameters with a non-identical gradient: {len(bad)}")
for k, v in sorted(bad.items()):
    print(k, v[:4], "...", len(v))
