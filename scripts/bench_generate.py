#!/usr/bin/env python
"""N1 benchmark: Libra-11B KV-cached greedy decoding on one B200 (random-init bf16 weights, synthetic prompt of 1 image +
text).  Prints one JSON line: prefill tokens/s, decode tokens/s (all samples), ms per decode step, and the decode-attention
kernel's achieved HBM bandwidth measured with CUDA events around every launch inside the timed steps.
    python scripts/bench_generate.py [--batch 8] [--prompt 1024] [--new 32] [--layers 32]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libra_b200 import ops, synthetic
from libra_b200.models import LibraConfig, LibraForCausalLM

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--prompt", type=int, default=1024)
ap.add_argument("--new", type=int, default=32)
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
ev = lambda: torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    out = None
    for rep in range(2):                                  # first pass = warm-up (cuBLAS heuristics, caches)
        out = None                                        # drop the previous cache first: its blocks are reused, no cudaMalloc in the timing
        s0, s1 = ev(), ev()
        s0.record()
        out = model(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], vision_indices=inp["vision_indices"],
                    contiguous_signal=inp["contiguous_signal"], use_cache=True)
        s1.record()
        torch.cuda.synchronize()
        prefill_ms = s0.elapsed_time(s1)
        am = inp["attention_mask"]
        nxt = out.logits[:, :, -1].float().argmax(-1)
        vi = torch.full((a.batch, 1), 578, device=dev)
        steps = a.new if rep else 4
        if rep:
            ops.enable_timing(names=("lb_attn_decode",))
        t0, t1 = ev(), ev()
        t0.record()
        for _ in range(steps):
            am = torch.cat([am, am.new_ones(a.batch, 1)], dim=1)
            pos = (am.cumsum(-1) - 1)[:, -1:]
            out = model(input_ids=nxt[:, :, None], attention_mask=am, position_ids=pos, vision_indices=vi,
                        past_key_values=out.past_key_values, use_cache=True)
            nxt = out.logits[:, :, -1].float().argmax(-1)
        t1.record()
        torch.cuda.synchronize()
        decode_ms = t0.elapsed_time(t1) / steps
kt = ops.disable_timing() or {}
# ---- the same decoding through generate(): the one-token step captured in a CUDA graph and replayed
model.generate(inp["input_ids"], attention_mask=inp["attention_mask"], vision_indices=inp["vision_indices"],
               contiguous_signal=inp["contiguous_signal"], max_new_tokens=8)                       # warm-up
model.generate(inp["input_ids"], attention_mask=inp["attention_mask"], vision_indices=inp["vision_indices"],
               contiguous_signal=inp["contiguous_signal"], max_new_tokens=a.new + 2)
torch.cuda.synchronize()
n_rep, e0, e1 = model.last_graph_decode                     # CUDA events around the replays only
graph_ms = e0.elapsed_time(e1) / n_rep
evs = kt.get("lb_attn_decode", [])
attn_ms = sum(s.elapsed_time(e) for s, e in evs) / max(len(evs), 1)
kv = a.prompt + a.new / 2
C = cfg.hidden_size
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
by = a.batch * kv * C * 2 * 2
n_param = sum(p.numel() for p in model.parameters())
n_lang = sum(p.numel() for n, p in model.named_parameters() if "vision" not in n)      # what a text token's step reads
kv_gb = a.layers * by / 1e9
print(json.dumps({
    "workload": f"Libra-11B greedy decoding, B={a.batch}, prompt {a.prompt} (1 image), {a.new} new tokens, {a.layers} layers, bf16, random init",
    "prefill_tokens_per_s": a.batch * a.prompt / (prefill_ms * 1e-3), "prefill_ms": prefill_ms,
    "decode_tokens_per_s": a.batch / (graph_ms * 1e-3), "decode_ms_per_step": graph_ms,
    "decode_eager_tokens_per_s": a.batch / (decode_ms * 1e-3), "decode_eager_ms_per_step": decode_ms,
    "weights_gb": n_param * 2 / 1e9, "language_weights_gb_per_step": n_lang * 2 / 1e9, "kv_cache_gb_per_step": kv_gb,
    "hbm_floor_ms": (n_lang * 2 / 1e9 + kv_gb) / peaks["hbm_gbs"] * 1e3,
    "decode_frac_of_hbm_peak": (n_lang * 2 / 1e9 + kv_gb) / peaks["hbm_gbs"] * 1e3 / graph_ms,
    "attn_decode": {"avg_launch_us": attn_ms * 1e3, "achieved_gbs": by / (attn_ms * 1e-3) / 1e9 if attn_ms else None,
                    "peak_gbs": peaks["hbm_gbs"], "frac": by / (attn_ms * 1e-3) / 1e9 / peaks["hbm_gbs"] if attn_ms else None},
    "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
