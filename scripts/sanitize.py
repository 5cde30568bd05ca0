"""Every CUDA kernel of the library once at tiny shapes, for compute-sanitizer:

    compute-sanitizer --tool memcheck  --error-exitcode 3 python scripts/sanitize.py
    compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize.py
    compute-sanitizer --tool synccheck --error-exitcode 3 python scripts/sanitize.py

Covers: grouped GEMM (all layouts, epilogues, chain, segments, both CTA-group modes through LB_GEMM_CG), the decoder forward +
backward (norms, SwiGLU, attention prologue, attention forward / dQ / dK,dV, embeddings, cross-entropy), the ViT + vision
tokenizer (patch embed, LayerNorm, non-causal attention, LFQ pack), KV-cached decoding, AdamW + clip."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libra_b200 import ops, synthetic  # noqa: E402
from libra_b200.dist import FlatGradBuffer  # noqa: E402
from libra_b200.models import LibraConfig, LibraForCausalLM, VisionTokenizer  # noqa: E402
from libra_b200.models.modeling_clip import CLIPVisionConfig  # noqa: E402
from libra_b200.optim import FlatAdamW  # noqa: E402

dev, BF16 = "cuda", torch.bfloat16
torch.manual_seed(0)
rnd = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).to(BF16)
G = ops.gp

# ---- grouped GEMM
for ta, tb in ((0, 0), (0, 1), (1, 0), (1, 1)):
    M, N, K = 264, 200, 136
    a = rnd(K, M) if ta else rnd(M, K)
    b = rnd(K, N) if tb else rnd(N, K)
    c = torch.empty(M, N, device=dev, dtype=BF16)
    ops.gemm_grouped([G(a, b, c, ta=bool(ta), tb=bool(tb))])
x, W, A, Bw, res = rnd(300, 128), rnd(128, 128, sc=0.1), rnd(32, 128, sc=0.1), rnd(128, 32, sc=0.1), rnd(300, 128)
y, mid = torch.empty(300, 128, device=dev, dtype=BF16), torch.empty(120, 32, device=dev, dtype=BF16)
ops.gemm_grouped([G(x[180:], A, mid), G(x[:180], W, y[:180], d=res[:180]), G(mid, Bw, y[180:], d=res[180:], wait_on=0)])
h, g, u = (torch.empty(300, 128, device=dev, dtype=BF16) for _ in range(3))
ops.gemm_grouped([G(x, W, h, b2=W, epi=ops.EPI_SWIGLU, g=g, u=u)])
ops.gemm_grouped([G(x, W, h, bias=rnd(128), epi=ops.EPI_QGELU, g=g)])
ops.gemm_grouped([G(x, W, h, tb=True), G(x, W, h, tb=True, acc_prev=True)])
t8 = torch.empty(300, 8, device=dev, dtype=BF16)
ops.gemm_grouped([G(x, rnd(8, 128), t8)])
torch.cuda.synchronize()
print("gemm ok")

# ---- decoder forward + backward + optimizer
cfg = LibraConfig(hidden_size=256, intermediate_size=704, num_hidden_layers=2, num_attention_heads=2, vocab_size=512, contiguous_signal_size=64)
model = LibraForCausalLM(cfg).to(BF16).to(dev).train()
synthetic.randomize_for_bench(model, seed=1, std=0.05)
inp = synthetic.libra_batch(2, 700, 1, vocab=cfg.vocab_size, signal=cfg.contiguous_signal_size, seed=3, device=dev)
buf = FlatGradBuffer(model.named_parameters(), flatten_weights=True)
opt = FlatAdamW.for_buffer(buf, model, lr=1e-3, max_grad_norm=1.0)
for step in range(2):
    buf.begin_step()
    out = model(input_ids=inp["input_ids"], vision_indices=inp["vision_indices"], contiguous_signal=inp["contiguous_signal"], labels=inp["labels"])
    out.loss.backward()
    opt.step()
torch.cuda.synchronize()
print("decoder ok", float(out.loss))

# ---- ViT + vision tokenizer
cc = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, image_size=56, patch_size=14)
vt = VisionTokenizer(cc, select_layer=(-2, -3), embed_dim=18, token_offset=512).to(BF16).to(dev)
enc = vt.encode(torch.randn(2, 3, 56, 56, device=dev).to(BF16))
torch.cuda.synchronize()
print("vision tokenizer ok", tuple(enc["input_ids"].shape))

# ---- KV-cached decoding
model.eval()
with torch.no_grad():
    ids = model.generate(inp["input_ids"][:, :, :600], attention_mask=torch.ones(2, 600, dtype=torch.long, device=dev),
                         vision_indices=inp["vision_indices"][:, :600], contiguous_signal=inp["contiguous_signal"][:, :600],
                         max_new_tokens=3, cuda_graph=False)
torch.cuda.synchronize()
print("decode ok", tuple(ids.shape))
