#!/usr/bin/env python
"""BASELINE.json configs[1]: CLIP ViT-L/14-336 vision encoder only, batch 64 synthetic images, fwd+bwd on 1 GPU
(loss = mean of hidden_states[-2], SURVEY.md section 8(d) cfg 2).  Prints one JSON line (a parity/perf case, not the
bench.py headline)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from libra_b200 import _lib
from libra_b200.models.modeling_clip import CLIPVisionConfig, CLIPVisionModel


def main(batch=64, steps=5, warmup=3):
    _lib.require_device()
    dev = "cuda"
    torch.manual_seed(0)
    model = CLIPVisionModel(CLIPVisionConfig.vit_l_14_336()).to(torch.bfloat16).to(dev).train()
    g = torch.Generator(device=dev).manual_seed(1234)
    px = torch.rand(batch, 3, 336, 336, device=dev, generator=g)
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073], device=dev).view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711], device=dev).view(1, 3, 1, 1)
    px = ((px - mean) / std).bfloat16()

    def step():
        for p in model.parameters():
            p.grad = None
        out = model(px, output_hidden_states=True)
        out.hidden_states[-2].float().mean().backward()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.reset_launch_counts()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flops = 3 * 381.9e9 * batch * (23 / 24)        # hidden_states[-2]: the last layer is not on the graph
    print(json.dumps({"workload": f"ViT-L/14-336 fwd+bwd, batch {batch}", "ms_per_step": ms, "images_per_s": batch / ms * 1e3,
                      "model_tflops": flops / (ms * 1e-3) / 1e12, "gpu_launches": _lib.total_launches(),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
