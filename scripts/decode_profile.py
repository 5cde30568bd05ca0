#!/usr/bin/env python
"""Kernel-time table of ONE eager one-token decode step of Libra-11B (torch.profiler CUDA activity; kernel durations are valid
although the eager step is launch-bound in wall time): which kernels the 8.4 ms graph-replayed step consists of.
    python scripts/decode_profile.py [--batch 8] [--prompt 1024] [--layers 32]"""
import argparse, collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from libra_b200 import synthetic
from libra_b200.models import LibraConfig, LibraForCausalLM

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--prompt", type=int, default=1024)
ap.add_argument("--layers", type=int, default=32)
a = ap.parse_args()
dev = "cuda"
cfg = LibraConfig(num_hidden_layers=a.layers)
torch.manual_seed(0)
torch.set_default_dtype(torch.bfloat16)
with torch.device(dev):
    model = LibraForCausalLM(cfg)
torch.set_default_dtype(torch.float32)
model = model.to(torch.bfloat16).eval()
synthetic.randomize_for_bench(model, seed=0)
inp = synthetic.libra_batch(a.batch, a.prompt, 1, vocab=cfg.vocab_size, signal=cfg.contiguous_signal_size, seed=7, device=dev)
with torch.no_grad():
    out = model(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], vision_indices=inp["vision_indices"],
                contiguous_signal=inp["contiguous_signal"], use_cache=True)
    am = inp["attention_mask"]
    nxt = out.logits[:, :, -1].float().argmax(-1)
    vi = torch.full((a.batch, 1), 578, device=dev)

    def step():
        global am, out, nxt
        am = torch.cat([am, am.new_ones(a.batch, 1)], dim=1)
        pos = (am.cumsum(-1) - 1)[:, -1:]
        out = model(input_ids=nxt[:, :, None], attention_mask=am, position_ids=pos, vision_indices=vi,
                    past_key_values=out.past_key_values, use_cache=True)
        nxt = out.logits[:, :, -1].float().argmax(-1)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        tot[e.name][0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
        tot[e.name][1] += 1
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
mid = len(evs) // 2
print("kernels in launch order around the middle of the step (one decoder layer):")
for e in evs[mid:mid + 16]:
    print(f"   {(e.device_time if hasattr(e, 'device_time') else e.cuda_time):8.2f} us  {e.name[:100]}")
rows = sorted(tot.items(), key=lambda kv: -kv[1][0])
total = sum(v[0] for _, v in rows)
print(f"one decode step: {sum(v[1] for _, v in rows)} kernels, {total / 1e3:.3f} ms summed device time")
for name, (us, n) in rows[:40]:
    print(f"{us / total * 100:6.2f}%  {us:9.1f} us  {n:5d} x {us / n:8.2f} us  {name[:110]}")
