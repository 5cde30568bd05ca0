#!/usr/bin/env python
"""bench.py -- Libra-11B training throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W      (CPU reference arm)

A "step" is one optimizer step of the whole hot path on a synthetic batch per GPU:
pixel_values [B,3,336,336] -> CLIP ViT-L/14-336 tower -> quant_conv -> LFQ ids (VisionTokenizer.encode) -> tensor
assembly + labels (assemble_inputs / get_labels) -> micro-batched forward + backward of LibraForCausalLM (Libra-11B shapes,
random init, bf16) -> gradient all-reduce of the flat bf16 buffer when N > 1 (issued in pieces during the last micro-batch's
backward, libra_b200.dist.GradSync) -> the reference's optimizer recipe (clip 1.0 + AdamW, one fused pass).
Every dense product runs on this repository's grouped tcgen05 GEMM; nothing on the path calls cuBLAS.

Workloads (BASELINE.json configs):
  cfg3 (default, the headline `value`): B=8 samples x T=2048 per GPU as 2 micro-batches of 4 (<s> + one 578-token image +
        1469 text tokens), all parameters trainable, no checkpointing.  Weak scaling: per-GPU batch fixed.
  cfg4 (reported beside it in the same line, key "cfg4"; `--workload cfg4` makes it the headline): the instruction-tuning
        step the "@1/2/4/8" metric is quoted on -- global batch 128 = N GPUs x micro-batch 2 x accumulation 64/N, T=2048,
        gradient checkpointing on, all parameters trainable.
`value` = tokens of all ranks / max-over-ranks device time with the inputs resident in HBM; `e2e` repeats the measurement
with the step's inputs (pixels, text ids, mask) in pinned HOST memory, copied in inside the timed region, and the loss read back.

Only this file's `cpu_baseline` / `gpu_eager_baseline` legs and `--impl reference` execute anything under oracle/ (the
checker), never the measured CUDA path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

METRIC = "libra11b_train_tokens_per_sec"
UNIT = "tokens/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="libra_b200", choices=["libra_b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4", "cfg5", "custom"],
                    help="BASELINE.json config the headline value is measured on (custom: take the shape flags below)")
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU per optimizer step")
    ap.add_argument("--micro-batch", type=int, default=None)
    ap.add_argument("--seq", type=int, default=None)
    ap.add_argument("--images", type=int, default=None)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--ckpt", type=int, default=None, help="gradient checkpointing per decoder layer")
    ap.add_argument("--frozen-language", type=int, default=None)
    ap.add_argument("--optimizer", default="adamw", choices=["adamw", "none"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the secondary cfg 4 block")
    ap.add_argument("--no-tokenizer", action="store_true", help="synthetic ids/signal instead of pixels through the vision tokenizer")
    ap.add_argument("--tiny", action="store_true", help="tiny model for plumbing checks (NOT a bench value)")
    ap.add_argument("--overlap-allreduce", type=int, default=1,
                    help="N>1: reduce the ready prefix of the flat gradient buffer in pieces during the last micro-batch's backward")
    ap.add_argument("--allreduce-min-mb", type=int, default=1024, help="smallest piece of the overlapped all-reduce, MiB")
    ap.add_argument("--busy-trace", action="store_true",
                    help="after the timed region run ONE more step under torch.profiler (CUDA activity) and report which share of "
                         "the step the GPU had a kernel running, plus device time per kernel family (key 'busy'; not a bench value)")
    ap.add_argument("--cuda-profiler-range", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop (use with ncu --profile-from-start off)")
    return ap.parse_args()


WORKLOADS = {
    # name: (BASELINE.json index, samples per GPU per step (None: 128 / world), micro-batch, T, images, ckpt, frozen language)
    "cfg3": (2, 8, 4, 2048, 1, 0, 0),
    "cfg4": (3, None, 2, 2048, 1, 1, 0),
    "cfg5": (4, 8, 8, 4096, 4, 1, 1),
}


def resolve_workload(args, world, name=None):
    name = name or args.workload
    if name == "custom":
        idx, B, MB, T, img, ck, fr = None, args.batch or 8, args.micro_batch or 4, args.seq or 2048, args.images or 1, args.ckpt or 0, args.frozen_language or 0
    else:
        idx, B, MB, T, img, ck, fr = WORKLOADS[name]
        if B is None:
            B = max(MB, 128 // world)
        over = lambda v, d: d if v is None else v
        B, MB, T, img = over(args.batch, B), over(args.micro_batch, MB), over(args.seq, T), over(args.images, img)
        ck, fr = over(args.ckpt, ck), over(args.frozen_language, fr)
        if (args.batch, args.micro_batch, args.seq, args.images, args.ckpt, args.frozen_language) != (None,) * 6:
            idx = None                                   # a modified shape is not the BASELINE config any more
    return dict(name=name, baseline_index=idx, B=B, MB=MB, T=T, images=img, ckpt=int(ck), frozen=int(fr))


def attn_traffic(mb, T, cfg):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the attention forward kernel from the committed
    `ncu --set full` capture (profiles/attn_fwd_traffic.json); only valid for the shape it was captured on."""
    p = os.path.join(ROOT, "profiles", "attn_fwd_traffic.json")
    try:
        d = json.load(open(p))
        if d["shape"].startswith(f"B={mb},T={T},H={cfg.num_attention_heads},"):
            return d["dram_bytes_read"] + d["dram_bytes_write"]
    except Exception:
        pass
    return None


def gemm_traffic():
    """dram bytes of representative grouped-GEMM launches from the committed `ncu --set full` capture
    (profiles/r02_gemm_traffic.json), with the algorithmic operand bytes of the same launches beside them."""
    p = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs"), tflops_burst=d.get("bf16_tflops"),
                    tflops_sustained=d.get("bf16_tflops_sustained"), source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm = [float(r[0]) for r in rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.strip().lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# --------------------------------------------------------------------------------------- CPU reference arm
class CpuReference:
    """The oracle (CPU restatement of the reference's LibraForCausalLM path, oracle/libra_oracle.py) on the host cores, on
    BASELINE.json configs[0]'s sample (1 image + 32 text tokens, T=611, random init): the WHOLE model structure -- embeddings
    (token + vision + signal path), `n_layers` FULL-WIDTH decoder layers, final norms, the three heads and the loss -- with the
    decoder-layer time scaled by 32 / n_layers (the layers are identical).  fp32 by default: bf16 GEMMs are far slower than fp32
    on host CPUs without AMX (measured on the GPU box: 89 s for 2 layers in bf16)."""

    def __init__(self, n_layers: int = 1, threads: int = 0, dtype=torch.float32):
        from oracle import libra_oracle as O
        self.O = O
        # MKL/oneDNN GEMMs of this size stop scaling (and start thrashing) well before 128 threads: cap at 64
        self.threads = threads or min(os.cpu_count() or 1, 64)
        torch.set_num_threads(self.threads)
        self.n_layers, self.dtype = n_layers, dtype
        self.d = d = O.LibraDims(num_hidden_layers=n_layers)
        g = torch.Generator().manual_seed(0)
        H, I, R, V, Vv, S = d.hidden_size, d.intermediate_size, d.bridge_rank, d.vocab_size, d.vision_vocab_size, d.contiguous_signal_size
        sd = {}

        def w(*s):
            return (torch.randn(*s, generator=g) * 0.02).to(dtype).requires_grad_(True)
        for i in range(n_layers):
            p = f"model.layers.{i}"
            for n in "qkvo":
                sd[f"{p}.self_attn.{n}_proj.weight"] = w(H, H)
                sd[f"{p}.self_attn.vision_{n}_proj.weight_A"] = w(H // 4, H)
                sd[f"{p}.self_attn.vision_{n}_proj.weight_B"] = w(H, H // 4)
            for n in "kv":
                for m in ("language", "vision"):
                    sd[f"{p}.self_attn.vision_{n}_bridge_on_{m}.weight_A"] = w(R, H)
                    sd[f"{p}.self_attn.vision_{n}_bridge_on_{m}.weight_B"] = w(H, R)
            for n, (i_, o_) in dict(gate=(H, I), up=(H, I), down=(I, H)).items():
                sd[f"{p}.mlp.{n}_proj.weight"] = w(o_, i_)
                sd[f"{p}.mlp.vision_{n}_proj.weight_A"] = w(o_ // 4, i_)
                sd[f"{p}.mlp.vision_{n}_proj.weight_B"] = w(o_, o_ // 4)
            for n in ("input_layernorm", "post_attention_layernorm", "vision_input_layernorm", "vision_post_attention_layernorm"):
                sd[f"{p}.{n}.weight"] = torch.ones(H, dtype=dtype).requires_grad_(True)
        sd["model.embed_tokens.weight"] = w(V, H)
        for c in range(2):
            sd[f"model.vision_embed_tokens.{c}.weight"] = w(Vv, H // 2)
            sd[f"vision_lm_head.heads.{c}.weight"] = w(Vv, H)
        sd["model.vision_signal_norm.weight"] = torch.ones(H + S, dtype=dtype).requires_grad_(True)
        sd["model.vision_contiguous_signal_processor.weight"] = w(H, H + S)
        sd["model.norm.weight"] = torch.ones(H, dtype=dtype).requires_grad_(True)
        sd["model.vision_norm.weight"] = torch.ones(H, dtype=dtype).requires_grad_(True)
        sd["lm_head.weight"] = w(V, H)
        self.sd = sd
        from libra_b200 import synthetic
        self.T = 611
        self.inp = synthetic.libra_batch(1, self.T, 1, vocab=V, signal=S, seed=5, signal_dtype=dtype)

    def _run(self, backward: bool):
        """(seconds in the decoder layers, seconds in everything else) of one sample"""
        O, d, sd, inp = self.O, self.d, self.sd, self.inp
        for t in sd.values():
            t.grad = None
        flag = inp["vision_indices"] < d.max_vision_token_length
        pos = torch.arange(self.T)[None]
        t0 = time.perf_counter()
        h = O.embed_tokens(sd, d, inp["input_ids"], flag, inp["contiguous_signal"])
        t1 = time.perf_counter()
        hl = h
        if backward:
            hl = h.detach().requires_grad_(True)
        x = hl
        for i in range(self.n_layers):
            x = O.decoder_layer(sd, i, d, x, flag, pos, None)
        t2 = time.perf_counter()
        xo = x.detach().requires_grad_(True) if backward else x
        hn = O.route(xo, flag, lambda r: O.rmsnorm(r, sd["model.norm.weight"], d.rms_norm_eps),
                     lambda r: O.rmsnorm(r, sd["model.vision_norm.weight"], d.rms_norm_eps))
        logits = O.vl_logits(sd, d, hn, flag)
        loss = O.causal_lm_loss(logits, inp["labels"]) if backward else None
        t3 = time.perf_counter()
        t_layers, t_rest = t2 - t1, (t1 - t0) + (t3 - t2)
        if backward:
            loss.backward()
            t4 = time.perf_counter()
            x.backward(xo.grad)
            t5 = time.perf_counter()
            h.backward(hl.grad)
            t6 = time.perf_counter()
            t_layers += t5 - t4
            t_rest += (t4 - t3) + (t6 - t5)
        return t_layers, t_rest

    def step(self) -> float:
        """seconds of one forward+backward of the sample, the layer part scaled to 32 layers"""
        tl, tr = self._run(True)
        return tl * 32.0 / self.n_layers + tr

    def forward_seconds(self) -> float:
        with torch.no_grad():
            tl, tr = self._run(False)
        return tl * 32.0 / self.n_layers + tr

    def result(self, seconds: float) -> dict:
        return dict(value=self.T / seconds, unit=UNIT, cores=self.threads, kind="port",
                    sample=f"oracle fwd+bwd of the whole Libra-11B structure (embeddings, {self.n_layers} full-width decoder layer(s) "
                           f"scaled x{32 // self.n_layers} to 32, final norms, heads, loss), {str(self.dtype).split('.')[-1]}, B=1 T=611 "
                           f"(BASELINE.json configs[0]: 1 image + 32 text tokens), {seconds:.2f} s per sample-equivalent")


def cpu_forward_cfg1(budget_s: float = 60.0) -> dict:
    """the reference path's FORWARD on configs[0] (T = 611, no labels) in fp32 and bf16, 4 layers x 8
    (stated extrapolation), plus the CLIP ViT-L/14-336 forward of one image, on the host cores."""
    from oracle import libra_oracle as O
    out = {"layers_timed": 4, "extrapolation": "decoder-layer time x 8 to 32 layers; embeddings and heads timed as they are",
           "cpu_model": next((l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "unknown")}
    t_start = time.perf_counter()
    ref = CpuReference(n_layers=4)
    ref.forward_seconds()
    out["fp32_forward_s"] = statistics.median(ref.forward_seconds() for _ in range(3))
    out["fp32_forward_tokens_per_s"] = ref.T / out["fp32_forward_s"]
    del ref
    if time.perf_counter() - t_start < budget_s:
        r16 = CpuReference(n_layers=1, dtype=torch.bfloat16)         # bf16 host GEMMs are slow without AMX: one layer, one sample
        out["bf16_forward_s"] = r16.forward_seconds()
        out["bf16_layers_timed"] = 1
        del r16
    c = O.ClipDims()
    g = torch.Generator().manual_seed(0)
    sd = {}
    Hc, Ic, Lc = c.hidden_size, c.intermediate_size, c.num_hidden_layers
    rn = lambda *s: torch.randn(*s, generator=g) * 0.02
    sd["vision_model.embeddings.class_embedding"] = rn(Hc)
    sd["vision_model.embeddings.patch_embedding.weight"] = rn(Hc, 3, c.patch_size, c.patch_size)
    sd["vision_model.embeddings.position_embedding.weight"] = rn(c.grid ** 2 + 1, Hc)
    for nm in ("pre_layrnorm", "post_layernorm"):
        sd[f"vision_model.{nm}.weight"], sd[f"vision_model.{nm}.bias"] = torch.ones(Hc), torch.zeros(Hc)
    for i in range(Lc):
        p = f"vision_model.encoder.layers.{i}"
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[f"{p}.self_attn.{nm}.weight"], sd[f"{p}.self_attn.{nm}.bias"] = rn(Hc, Hc), torch.zeros(Hc)
        sd[f"{p}.mlp.fc1.weight"], sd[f"{p}.mlp.fc1.bias"] = rn(Ic, Hc), torch.zeros(Ic)
        sd[f"{p}.mlp.fc2.weight"], sd[f"{p}.mlp.fc2.bias"] = rn(Hc, Ic), torch.zeros(Hc)
        for nm in ("layer_norm1", "layer_norm2"):
            sd[f"{p}.{nm}.weight"], sd[f"{p}.{nm}.bias"] = torch.ones(Hc), torch.zeros(Hc)
    px = torch.randn(1, 3, c.image_size, c.image_size, generator=g)
    with torch.no_grad():
        O.clip_vision_hidden_states(sd, c, px)
        t0 = time.perf_counter()
        O.clip_vision_hidden_states(sd, c, px)
        out["clip_vit_l_336_forward_1_image_fp32_s"] = time.perf_counter() - t0
    return out


def cpu_reference_sample(n_layers: int = 1) -> dict:
    ref = CpuReference(n_layers)
    ref.step()                       # warm-up (thread pool, allocator)
    r = ref.result(ref.step())
    del ref
    try:
        r["forward_cfg1"] = cpu_forward_cfg1()
    except Exception as ex:          # the checker must never take the measured path down
        r["forward_cfg1"] = {"error": repr(ex)}
    return r


def gpu_eager_baseline(dev, layers: int = 1, B: int = 2, T: int = 2048) -> dict:
    """The reference's modules in eager bf16 on this same GPU (last bullet): the oracle restatement of
    LibraDecoderLayer (boolean-mask routing, materialised [B,32,T,T] scores, fp32 softmax -- the reference's arithmetic and its
    kernel sequence through PyTorch/cuBLAS), `layers` full-width layer(s), forward + backward, scaled x32/layers.  Embeddings
    and heads excluded (stated)."""
    from oracle import libra_oracle as O
    d = O.LibraDims(num_hidden_layers=layers)
    g = torch.Generator(device=dev).manual_seed(0)
    H, I, R = d.hidden_size, d.intermediate_size, d.bridge_rank
    sd = {}
    w = lambda *s: (torch.randn(*s, generator=g, device=dev) * 0.02).to(torch.bfloat16).requires_grad_(True)
    for i in range(layers):
        p = f"model.layers.{i}"
        for n in "qkvo":
            sd[f"{p}.self_attn.{n}_proj.weight"] = w(H, H)
            sd[f"{p}.self_attn.vision_{n}_proj.weight_A"] = w(H // 4, H)
            sd[f"{p}.self_attn.vision_{n}_proj.weight_B"] = w(H, H // 4)
        for n in "kv":
            for m in ("language", "vision"):
                sd[f"{p}.self_attn.vision_{n}_bridge_on_{m}.weight_A"] = w(R, H)
                sd[f"{p}.self_attn.vision_{n}_bridge_on_{m}.weight_B"] = w(H, R)
        for n, (i_, o_) in dict(gate=(H, I), up=(H, I), down=(I, H)).items():
            sd[f"{p}.mlp.{n}_proj.weight"] = w(o_, i_)
            sd[f"{p}.mlp.vision_{n}_proj.weight_A"] = w(o_ // 4, i_)
            sd[f"{p}.mlp.vision_{n}_proj.weight_B"] = w(o_, o_ // 4)
        for n in ("input_layernorm", "post_attention_layernorm", "vision_input_layernorm", "vision_post_attention_layernorm"):
            sd[f"{p}.{n}.weight"] = torch.ones(H, dtype=torch.bfloat16, device=dev).requires_grad_(True)
    flag = torch.zeros(B, T, dtype=torch.bool, device=dev)
    flag[:, 1:579] = True
    pos = torch.arange(T, device=dev)[None].expand(B, T)
    h0 = torch.randn(B, T, H, generator=g, device=dev).to(torch.bfloat16)

    def once():
        for t in sd.values():
            t.grad = None
        x = h0.clone().requires_grad_(True)
        for i in range(layers):
            x = O.decoder_layer(sd, i, d, x, flag, pos, None)
        x.float().pow(2).mean().backward()
    once()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    n = 3
    for _ in range(n):
        once()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    del sd, h0
    torch.cuda.empty_cache()
    return {"value": B * T / (ms * 1e-3 * 32.0 / layers), "unit": UNIT, "kind": "oracle (PyTorch eager restatement of the reference's "
            "LibraDecoderLayer: boolean-mask routing, materialised scores, fp32 softmax), bf16, cuBLAS GEMMs",
            "sample": f"{layers} full-width decoder layer(s) fwd+bwd, B={B} T={T} (1 image per sample), {ms:.1f} ms, scaled x{32 // layers} "
                      f"to 32 layers; embeddings, heads, vision tokenizer and optimizer excluded", "ms_per_layer_fwd_bwd": ms / layers}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = resolve_workload(args, max(1, world))
    ref = CpuReference(n_layers=1)
    secs = []
    t_start = time.perf_counter()
    budget_s = 150.0                      # the whole arm must end within a few minutes on any host
    done_w = 0
    for i in range(args.warmup + args.steps):
        t = ref.step()
        if i >= args.warmup:
            secs.append(t)
        else:
            done_w += 1
        if time.perf_counter() - t_start > budget_s and (secs or i + 1 >= args.warmup):
            if not secs:
                secs.append(t)            # budget exhausted during warm-up: report the last warm-up sample
            break
    cb = ref.result(statistics.median(secs)) if secs else None
    if cb is not None and len(secs) < args.steps:
        cb["sample"] += f"; time budget {budget_s:.0f} s reached after {done_w} warm-up + {len(secs)} timed samples"
    v = cb["value"] if cb else float("nan")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": (statistics.median(secs) * 1e3 if secs else None), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(wl, 32, True, world) + "; the CPU arm times a bounded sample of it (see cpu_baseline.sample)"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_string(wl, layers, with_opt, world):
    base = "" if wl["baseline_index"] is None else f" -- BASELINE.json configs[{wl['baseline_index']}]"
    return (f"Libra-11B train step (vision tokenizer + fwd+bwd{'+clip+AdamW' if with_opt else ''}{'+grad all-reduce' if world > 1 else ''}), "
            f"B={wl['B']} per GPU (micro-batch {wl['MB']} x {wl['B'] // wl['MB']}) x T={wl['T']}, {wl['images']} image(s)/sample, {layers} layers, "
            f"all params trainable={not wl['frozen']}, ckpt={bool(wl['ckpt'])}{base}")


# --------------------------------------------------------------------------------------- CUDA arm
class Workload:
    """One BASELINE workload on this rank: synthetic host batch, the step function, timing."""

    def __init__(self, wl, args, cfg, model, vtok, buf, sync, opt, dev, rank, world):
        from libra_b200 import synthetic
        from libra_b200.models.tokenization_libra import assemble_inputs, get_labels
        global attach_host_layout
        from libra_b200.schedule import attach_host_layout
        self.wl, self.args, self.cfg, self.model, self.vtok = wl, args, cfg, model, vtok
        self.buf, self.sync, self.opt, self.dev, self.world = buf, sync, opt, dev, world
        self.assemble_inputs, self.get_labels = assemble_inputs, get_labels
        B, T = wl["B"], wl["T"]
        V = cfg.vocab_size
        self.PH = V                                    # <img_ph> id of the reference's text tokenizer (tokenization_libra.py:145-147)
        g = torch.Generator().manual_seed(1234 + rank)
        text = torch.randint(3, V, (B, T), generator=g)
        text[:, 0] = 1
        L = synthetic.IMG if vtok is None else vtok.max_vision_token_length
        spans = []
        pos = 1
        for _ in range(wl["images"]):
            text[:, pos:pos + L] = self.PH
            pos += L
        for b in range(B):
            sp, p = [], 1
            for _ in range(wl["images"]):
                p += L
                if p < T:
                    sp.append([p, p + 1])              # "the nearest text token after an image" (laion_dataset.py:231-239)
            spans.append(sp)
        self.spans = spans
        self.n_img = B * wl["images"]
        if vtok is not None:
            S = vtok.encoder.config.image_size
            px = torch.rand(self.n_img, 3, S, S, generator=g)
            mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
            std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
            host = {"pixel_values": ((px - mean) / std).to(torch.bfloat16), "text_ids": text,
                    "attention_mask": torch.ones(B, T, dtype=torch.long)}
        else:
            sb = synthetic.libra_batch(B, T, wl["images"], vocab=V, signal=cfg.contiguous_signal_size, seed=1234 + rank)
            host = {"input_ids": sb["input_ids"], "attention_mask": sb["attention_mask"], "vision_indices": sb["vision_indices"],
                    "contiguous_signal": sb["contiguous_signal"], "labels": sb["labels"]}
        self.host = {k: v.pin_memory() for k, v in host.items()}
        self.resident = {k: v.to(dev) for k, v in self.host.items()}
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        self.opt_events = []
        self.ar_events = []

    def prepare(self, inp):
        """A1-A8 + labels: pixels -> ids / signal -> model inputs (device tensors, no host round trip besides the span mask)."""
        if self.vtok is None:
            return inp
        enc = self.vtok.encode(inp["pixel_values"])
        out = self.assemble_inputs(inp["text_ids"], inp["attention_mask"], self.PH, enc["input_ids"], enc["encoder_feat"],
                                   max_vision_token_length=self.vtok.max_vision_token_length)
        labels = self.get_labels(out["input_ids"], out["attention_mask"], self.vtok.boi_token_id, 1, self.spans)
        return {"input_ids": out["input_ids"], "vision_indices": out["vision_indices"], "contiguous_signal": out["coninous_signal"],
                "labels": labels, "flag_cpu": self.host["text_ids"] == self.PH}

    def step(self, from_host: bool):
        wl = self.wl
        inp = self.host if from_host else self.resident
        if from_host:
            inp = {k: v.to(self.dev, non_blocking=True) for k, v in inp.items()}
        inp = self.prepare(inp)
        B, MB = wl["B"], wl["MB"]
        n_micro = B // MB
        self.buf.begin_step()
        total = None
        for i in range(n_micro):
            sl = slice(i * MB, (i + 1) * MB)
            self.sync.arm(last=(i == n_micro - 1))
            vi = inp["vision_indices"][sl]
            if inp.get("flag_cpu") is not None:         # the layout is known on the host: no device->host copy in the step
                attach_host_layout(vi, inp["flag_cpu"][sl])
            out = self.model(input_ids=inp["input_ids"][:, sl], attention_mask=None, vision_indices=vi,
                             contiguous_signal=inp["contiguous_signal"][sl], labels=inp["labels"][:, sl])
            loss = out.loss * (MB / B) / self.world
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        if self.world > 1:
            # device time between "backward fully enqueued" and "every piece of the all-reduce done" = the EXPOSED part of the
            # gradient reduction (pieces issued from the layer hooks overlap the backward that is still running)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if self.args.overlap_allreduce:
                self.sync.finish()
            else:
                dist.all_reduce(self.buf.flat)
            e1.record()
            self.ar_events.append((e0, e1))
        if self.opt is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.opt.step()
            e1.record()
            self.opt_events.append((e0, e1))
        return float(total.item()) if from_host else total

    def timed(self, k, from_host):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        last = None
        for _ in range(k):
            last = self.step(from_host)
        ev[1].record()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1])
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last


def busy_trace(W) -> dict:
    """One step under torch.profiler: the union of the kernel intervals against the step's device span (how much of the step is
    launch gaps), and device time per kernel family.  Profiler overhead inflates the span, so this is evidence about
    composition, never a throughput number."""
    import collections
    import re
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        W.step(False)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    if not evs:
        return {"error": "no CUDA events"}
    evs.sort(key=lambda e: e.time_range.start)
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    union, cur_s, cur_e = 0.0, evs[0].time_range.start, evs[0].time_range.end
    fam = collections.defaultdict(lambda: [0.0, 0])
    short = lambda n: re.sub(r"<.*", "", re.sub(r"^void ", "", n)).split("(")[0]
    gaps, prev_name, hist = [], short(evs[0].name), collections.Counter()
    for e in evs:
        s_, e_ = e.time_range.start, e.time_range.end
        if s_ > cur_e:
            union += cur_e - cur_s
            gaps.append((s_ - cur_e, prev_name, short(e.name)))
            hist["<2us" if s_ - cur_e < 2 else "<5us" if s_ - cur_e < 5 else "<20us" if s_ - cur_e < 20 else "<100us" if s_ - cur_e < 100 else ">=100us"] += s_ - cur_e
            cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
        prev_name = short(e.name)
        name = short(e.name)
        fam[name][0] += e_ - s_
        fam[name][1] += 1
    union += cur_e - cur_s
    top = sorted(fam.items(), key=lambda kv: -kv[1][0])[:24]
    return {"span_ms": (t1 - t0) / 1e3, "busy_ms": union / 1e3, "busy_frac": union / (t1 - t0), "kernels": len(evs),
            "summed_kernel_ms": sum(v[0] for v in fam.values()) / 1e3,
            "idle_ms_by_gap_length": {k: v / 1e3 for k, v in hist.items()},
            "largest_gaps": [{"us": g, "after": a_, "before": b_} for g, a_, b_ in sorted(gaps, key=lambda t: -t[0])[:12]],
            "top": [{"kernel": k, "ms": v[0] / 1e3, "launches": v[1]} for k, v in top]}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and not (world == 1 and args.gpus == 1):
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    from libra_b200 import _lib, ops, synthetic
    from libra_b200.dist import FlatGradBuffer, GradSync
    from libra_b200.models import LibraConfig, LibraForCausalLM, VisionTokenizer
    from libra_b200.models.modeling_clip import CLIPVisionConfig
    from libra_b200.optim import FlatAdamW

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device()        # no fallback: fail loudly without the CUDA library / an sm_100 device
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = LibraConfig(num_hidden_layers=args.layers)
    clip_cfg = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                                image_size=336, patch_size=14)                       # ViT-L/14-336
    if args.tiny:
        cfg = LibraConfig(hidden_size=256, intermediate_size=704, num_hidden_layers=2, num_attention_heads=2, vocab_size=1024,
                          contiguous_signal_size=256)
        clip_cfg = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2,
                                    image_size=336, patch_size=14)
    torch.manual_seed(0)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        model = LibraForCausalLM(cfg)
        vtok = None
        if not args.no_tokenizer:
            # select_layer / embed_dim live in the un-shipped vision_tokenizer_config.yaml: two tower layers (2 x 1024 = the 2048-wide
            # contiguous signal of configuration_libra.py:17) and 18 = 2 codebooks x 9 bits
            vtok = VisionTokenizer(clip_cfg, select_layer=(-2, -6) if not args.tiny else (-2, -3), embed_dim=18, token_offset=cfg.vocab_size)
    torch.set_default_dtype(prev)
    model = model.to(torch.bfloat16).train()
    synthetic.randomize_for_bench(model, seed=0)
    if vtok is not None:
        vtok = vtok.to(torch.bfloat16)
        torch.nn.init.normal_(vtok.quant_conv.weight, std=0.05)

    main_wl = resolve_workload(args, world)
    if main_wl["frozen"]:
        for n, p in model.named_parameters():
            p.requires_grad = "vision" in n
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    # Flat storage, laid out in the order gradients become final in backward: one bf16 buffer for the trainable weights and one
    # for their gradients; every parameter (and its .grad) is a view.  => the gradient all-reduce is a growing prefix of one
    # buffer, the optimizer a few launches, and weight gradients are written by the GEMM epilogues (first micro-batch of a step
    # overwrites, later ones add): no zeroing and no autograd accumulation pass over 22 GB.
    buf = FlatGradBuffer(model.named_parameters(), flatten_weights=True)
    sync = GradSync(buf, model.model, min_bytes=args.allreduce_min_mb << 20, n_layers=cfg.num_hidden_layers)
    opt = None
    if args.optimizer == "adamw":
        # the reference recipe (libra_pretrain.yaml:81-91,116; trainer.py:27-85): AdamW 0.9/0.99, wd 0.01 outside norms/biases, clip 1.0
        opt = FlatAdamW.for_buffer(buf, model, lr=1e-5, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.01, max_grad_norm=1.0)

    def set_ckpt(on):
        (model.gradient_checkpointing_enable if on else model.gradient_checkpointing_disable)()

    set_ckpt(main_wl["ckpt"])
    W = Workload(main_wl, args, cfg, model, vtok, buf, sync, opt, dev, rank, world)
    for _ in range(args.warmup):
        W.step(False)
    torch.cuda.synchronize()
    mem_gb = torch.cuda.max_memory_allocated() / 2 ** 30

    _lib.reset_launch_counts()
    W.opt_events.clear()
    W.ar_events.clear()
    ops.enable_timing(("lb_attn_fwd", "lb_attn_fwd_stream", "lb_attn_bwd_dq", "lb_attn_bwd_dq_stream", "lb_attn_bwd_dkv",
                       "lb_attn_bwd_dkv_stream", "lb_gemm_grouped"))
    sampler = ClockSampler(local) if rank == 0 else None
    if args.cuda_profiler_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ms, last_loss = W.timed(args.steps, False)
    if args.cuda_profiler_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if sampler else None
    kt = ops.disable_timing()
    launches = _lib.total_launches()
    B, MB, T = main_wl["B"], main_wl["MB"], main_wl["T"]
    tokens_step = B * T * world
    value = tokens_step * args.steps / (ms / 1e3)
    opt_ms = (sum(a.elapsed_time(b) for a, b in W.opt_events) / len(W.opt_events)) if W.opt_events else None
    allreduce = None
    if world > 1 and W.ar_events:
        ex = [a.elapsed_time(b) for a, b in W.ar_events[-args.steps:]]
        allreduce = {"bytes": buf.numel * buf.flat.element_size(), "pieces": [[lo, hi] for lo, hi in sync.pieces],
                     "n_pieces": len(sync.pieces), "exposed_ms_per_step": sum(ex) / len(ex), "overlapped": bool(args.overlap_allreduce),
                     "note": "exposed = device time from the end of the last backward to the completion of the last piece (CUDA events); "
                             "the other pieces run under the last micro-batch's backward"}

    busy = None
    if args.busy_trace and rank == 0:
        busy = busy_trace(W)

    e2e = None
    if not args.no_e2e:
        W.step(True)
        ms_e, _ = W.timed(args.steps, True)
        e2e = {"value": tokens_step * args.steps / (ms_e / 1e3), "unit": UNIT, "h2d_bytes_per_step": W.h2d_bytes,
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e / args.steps,
               "inputs": "pinned host pixel_values + text ids + mask" if vtok is not None else "pinned host ids/signal/labels (no tokenizer)"}

    # ---- secondary block: BASELINE.json configs[3] (instruction-tuning step, micro-batch 2, checkpointing, global batch 128)
    cfg4 = None
    if main_wl["name"] == "cfg3" and not args.no_cfg4 and not args.tiny and main_wl["baseline_index"] is not None:
        del W
        torch.cuda.empty_cache()
        wl4 = resolve_workload(args, world, "cfg4")
        set_ckpt(True)
        W4 = Workload(wl4, args, cfg, model, vtok, buf, sync, opt, dev, rank, world)
        small = dict(wl4, B=2 * wl4["MB"])                    # warm-up on two micro-batches of the same shape (allocator, caches)
        Ww = Workload(small, args, cfg, model, vtok, buf, sync, opt, dev, rank, world)
        Ww.step(False)
        del Ww
        ms4, _ = W4.timed(1, False)
        tok4 = wl4["B"] * wl4["T"] * world
        cfg4 = {"workload": workload_string(wl4, cfg.num_hidden_layers, opt is not None, world), "value": tok4 / (ms4 / 1e3),
                "unit": UNIT, "ms_per_step": ms4, "steps": 1, "warmup_micro_batches": 2, "global_batch": wl4["B"] * world,
                "grad_accum": wl4["B"] // wl4["MB"], "scaling": "strong (global batch fixed at 128)",
                "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        del W4
        set_ckpt(main_wl["ckpt"])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- rooflines: the dominant kernel (grouped GEMM, ~3/4 of the step) and the bridge-attention kernel BASELINE.json names
    peaks = load_peaks()
    fl = synthetic.decoder_flops(cfg, MB, T, main_wl["images"] * synthetic.IMG)
    n_micro = B // MB
    kern, count = {}, {}
    for name, evs in (kt or {}).items():
        if evs:
            kern[name] = sum(s.elapsed_time(e) for s, e in evs) / len(evs)
            count[name] = len(evs)
    roof = None
    if "lb_gemm_grouped" in kern:
        # algorithmic GEMM FLOPs of the step: 3 x forward (dgrad + wgrad) (+1 forward of recompute FLOPs under checkpointing is
        # hardware work, not counted), decoder + heads + signal projection; ViT projections of the frozen tower: forward only
        vit_fl = 0.0
        if vtok is not None and not args.tiny:
            vit_fl = B * main_wl["images"] * 381.9e9 * (23.0 / 24.0) * 0.857          # GEMM share of the tower forward (attention core excluded)
        gemm_fl = (3.0 if not main_wl["frozen"] else 2.5) * fl["gemm"] * n_micro + vit_fl
        t_total_ms = kern["lb_gemm_grouped"] * count["lb_gemm_grouped"] / args.steps
        ach = gemm_fl / (t_total_ms * 1e-3) / 1e12
        roof = {"kernel": "gemm_grouped_kernel<2> (persistent 2-CTA tcgen05 grouped GEMM: every dense product of the step)",
                "bound": "tensor", "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["tflops_sustained"],
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}); kernel timed inside a long step",
                "traffic": None, "traffic_profiled": gemm_traffic(),
                "avg_launch_ms": kern["lb_gemm_grouped"], "launches_per_step": count["lb_gemm_grouped"] / args.steps,
                "algorithmic_flops_per_launch": gemm_fl / (count["lb_gemm_grouped"] / args.steps),
                "share_of_step": t_total_ms / (ms / args.steps)}
    attn_roof = None
    fwd_names = {"lb_attn_fwd_stream": "attn_fwd_stream_kernel<128,causal> (bridge attention forward, persistent)",
                 "lb_attn_fwd": "attn_fwd_kernel<128,causal> (bridge attention forward)"}
    fwd_key = next((k for k in fwd_names if k in kern), None)
    BWD_KEYS = ("lb_attn_bwd_dq", "lb_attn_bwd_dq_stream", "lb_attn_bwd_dkv", "lb_attn_bwd_dkv_stream")   # dq family, dkv family
    if fwd_key:
        t_ms = kern[fwd_key]
        ach = fl["attn_per_layer_fwd"] / (t_ms * 1e-3) / 1e12
        attn_roof = {"kernel": fwd_names[fwd_key], "bound": "tensor", "achieved": ach,
                     "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tflops_sustained"],
                     "traffic": attn_traffic(MB, T, cfg), "avg_launch_ms": t_ms,
                     "algorithmic_flops_per_launch": fl["attn_per_layer_fwd"],
                     "bwd_ms": {k: kern[k] for k in BWD_KEYS if k in kern},
                     "bwd_achieved_tflops": (2.5 * fl["attn_per_layer_fwd"] / (sum(kern.get(k, 0) for k in BWD_KEYS) * 1e-3) / 1e12)
                     if (any(k in kern for k in BWD_KEYS[:2]) and any(k in kern for k in BWD_KEYS[2:])) else None}
    model_flops_step = 3.0 * fl["total"] * n_micro
    gb = None
    if not args.no_gpu_baseline and world == 1 and not args.tiny:
        try:
            gb = gpu_eager_baseline(dev)
        except Exception as ex:
            gb = {"error": repr(ex)}
    cb = None
    if not args.no_cpu_baseline and world == 1:
        try:
            cb = cpu_reference_sample(n_layers=1)
        except Exception as ex:      # the checker must never take the measured path down
            cb = {"error": repr(ex)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if main_wl["name"] != "cfg4" else "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_string(main_wl, cfg.num_hidden_layers, opt is not None, world),
                   "global_batch": B * world, "seq_len": T, "parallelism": f"dp{world}", "trainable_params": n_train,
                   "vision_tokenizer_in_step": vtok is not None,
                   "l2": "working set (>= 22 GB of weights per step) far exceeds the 126 MB L2; no explicit flush",
                   "tiny": bool(args.tiny)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "attention_roofline": attn_roof,
        "cpu_baseline": cb, "gpu_eager_baseline": gb, "cfg4": cfg4, "optimizer_ms": opt_ms, "allreduce": allreduce,
        "loss": float(last_loss.item()) * world if last_loss is not None else None,      # rank 0's mean loss over its samples
        "model_tflops_per_gpu": model_flops_step / (ms / args.steps * 1e-3) / 1e12, "peak_mem_gb": mem_gb,
    }
    if busy is not None:
        line["busy"] = busy
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
