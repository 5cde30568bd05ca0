#!/usr/bin/env python
"""bench.py -- Libra-11B training throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W      (CPU reference arm)

A "step" is one optimizer step of LibraForCausalLM (Libra-11B shapes, random init, bf16) on a synthetic batch of
B=8 samples x T=2048 tokens per GPU (<s> + one 578-token image + 1469 text tokens: BASELINE.json configs[2]):
micro-batched forward + backward through the libra_b200 CUDA path, one NCCL all-reduce of the flat bf16 gradient
buffer when N > 1, fused AdamW.  `value` = tokens of all ranks / max-over-ranks device time with the inputs resident
in HBM; `e2e` repeats the measurement through the public model API with pinned HOST inputs copied in, and the loss read
back, inside the timed region.  Weak scaling: per-GPU batch is fixed.

Only this file's `cpu_baseline` leg and `--impl reference` execute anything under oracle/ (the CPU checker), never
the measured CUDA path.
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
    ap.add_argument("--batch", type=int, default=8, help="samples per GPU per optimizer step")
    ap.add_argument("--micro-batch", type=int, default=4)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--images", type=int, default=1)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--ckpt", type=int, default=0, help="gradient checkpointing per decoder layer")
    ap.add_argument("--frozen-language", type=int, default=0)
    ap.add_argument("--optimizer", default="adamw", choices=["adamw", "none"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tiny", action="store_true", help="tiny model for plumbing checks (NOT a bench value)")
    ap.add_argument("--overlap-allreduce", type=int, default=1,
                    help="N>1: reduce the flat gradient buffer in a few large reverse-layer chunks during the last micro-batch's backward")
    ap.add_argument("--allreduce-chunks", type=int, default=8)
    ap.add_argument("--cuda-profiler-range", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop (use with ncu --profile-from-start off)")
    return ap.parse_args()


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
    """The oracle (CPU restatement of the reference's LibraForCausalLM path, oracle/libra_oracle.py) on the host cores:
    forward+backward of `n_layers` FULL-WIDTH Libra-11B decoder layers on BASELINE.json configs[0]'s sequence (1 image +
    32 text tokens, T=611), scaled by 32/n_layers to a whole-model tokens/s figure.  fp32 by default: bf16 GEMMs are far
    slower than fp32 on host CPUs without AMX (measured on the GPU box: 89 s for 2 layers in bf16)."""

    def __init__(self, n_layers: int = 1, threads: int = 0, dtype=torch.float32):
        from oracle import libra_oracle as O
        self.O = O
        # MKL/oneDNN GEMMs of this size stop scaling (and start thrashing) well before 128 threads: cap at 64
        self.threads = threads or min(os.cpu_count() or 1, 64)
        torch.set_num_threads(self.threads)
        self.n_layers, self.dtype = n_layers, dtype
        self.d = d = O.LibraDims()
        g = torch.Generator().manual_seed(0)
        H, I, R = d.hidden_size, d.intermediate_size, d.bridge_rank
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
        self.sd = sd
        self.T = 611
        self.flag = torch.zeros(1, self.T, dtype=torch.bool)
        self.flag[0, 1:579] = True
        self.pos = torch.arange(self.T)[None]
        self.h = torch.randn(1, self.T, H, generator=g).to(dtype)

    def step(self) -> float:
        """seconds for one forward+backward of the sample"""
        for t in self.sd.values():
            t.grad = None
        h = self.h.clone().requires_grad_(True)
        t0 = time.perf_counter()
        x = h
        for i in range(self.n_layers):
            x = self.O.decoder_layer(self.sd, i, self.d, x, self.flag, self.pos, None)
        x.float().pow(2).mean().backward()
        return time.perf_counter() - t0

    def result(self, seconds: float) -> dict:
        value = self.T / (seconds * 32.0 / self.n_layers)
        return dict(value=value, unit=UNIT, cores=self.threads, kind="port",
                    sample=f"oracle fwd+bwd of {self.n_layers} full-width Libra-11B decoder layer(s), {str(self.dtype).split('.')[-1]}, "
                           f"B=1 T=611 (1 image + 32 text), {seconds:.2f} s, scaled x{32 // self.n_layers} to 32 layers "
                           f"(embeddings/heads excluded)")


def cpu_reference_sample(n_layers: int = 1) -> dict:
    ref = CpuReference(n_layers)
    ref.step()                       # warm-up (thread pool, allocator)
    return ref.result(ref.step())


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
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
            "config": {"workload": "Libra-11B train step B=8 T=2048 per GPU (BASELINE.json configs[2]); the CPU arm times a bounded "
                                   "sample (see cpu_baseline.sample)"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------- CUDA arm
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
    from libra_b200 import functional as LF
    from libra_b200.models import LibraConfig, LibraForCausalLM

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device()        # no fallback: fail loudly without the CUDA library / an sm_100 device
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = LibraConfig(num_hidden_layers=args.layers)
    if args.tiny:
        cfg = LibraConfig(hidden_size=256, intermediate_size=704, num_hidden_layers=2, num_attention_heads=2, vocab_size=1024,
                          contiguous_signal_size=64)
    torch.manual_seed(0)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    with torch.device(dev):
        model = LibraForCausalLM(cfg)
    torch.set_default_dtype(prev)
    model = model.to(torch.bfloat16).train()
    synthetic.randomize_for_bench(model, seed=0)
    if args.ckpt:
        model.gradient_checkpointing_enable()
    if args.frozen_language:
        for n, p in model.named_parameters():
            p.requires_grad = "vision" in n
    params = [p for p in model.parameters() if p.requires_grad]
    n_train = sum(p.numel() for p in params)
    # Flat storage: one bf16 buffer for the trainable weights and one for their gradients; every parameter (and its
    # .grad) is a view.  => a single NCCL all-reduce per step and a single-tensor fused AdamW launch.
    flat_w = torch.empty(n_train, dtype=torch.bfloat16, device=dev)
    flat = torch.zeros(n_train, dtype=torch.bfloat16, device=dev)
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            flat_w[off:off + n].copy_(p.reshape(-1))
            p.data = flat_w[off:off + n].view_as(p)
            p.grad = flat[off:off + n].view_as(p)
            off += n
    # the flat buffer is owned here: weight gradients are written into it by the GEMM epilogues (the first micro-batch of a
    # step overwrites, later ones add) -- no zeroing pass and no autograd accumulation pass over 22 GB
    LF.mark_fused_grad(params)
    opt = None
    if args.optimizer == "adamw":
        from libra_b200.optim import FlatAdamW
        opt = FlatAdamW(flat_w, flat, lr=1e-5, betas=(0.9, 0.95), weight_decay=0.0)     # one fused kernel per step

    B, T, MB = args.batch, args.seq, args.micro_batch
    assert B % MB == 0
    host = synthetic.libra_batch(B, T, args.images, vocab=cfg.vocab_size, signal=cfg.contiguous_signal_size, seed=1234 + rank, pin=True)
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def micro(inp, i):
        sl = slice(i * MB, (i + 1) * MB)
        out = model(input_ids=inp["input_ids"][:, sl], attention_mask=None, vision_indices=inp["vision_indices"][sl],
                    contiguous_signal=inp["contiguous_signal"][sl], labels=inp["labels"][:, sl])
        return out.loss

    # ---- gradient all-reduce plan (N > 1): the flat buffer is [embeddings | layer 0 | ... | layer L-1 | norms+heads] in
    # registration order; chunk c covers a contiguous group of layers and is reduced as soon as the LAST micro-batch's
    # backward has passed the group's first layer.  Still one logical all-reduce of the buffer, issued as a few large pieces.
    offsets = {}
    off = 0
    for n_, p_ in model.named_parameters():
        if p_.requires_grad:
            offsets[n_] = (off, off + p_.numel())
            off += p_.numel()
    L = cfg.num_hidden_layers
    layer_lo = [min(v[0] for k, v in offsets.items() if k.startswith(f"model.layers.{i}.")) for i in range(L)]
    n_chunks = max(1, min(args.allreduce_chunks, L))
    group = (L + n_chunks - 1) // n_chunks
    first_layers = list(range(0, L, group))                     # chunk c starts at layer first_layers[c]
    pending = []
    state = {"armed": False}

    def on_layer_grad_ready(li):
        if not state["armed"] or li not in first_layers:
            return
        c = first_layers.index(li)
        lo = layer_lo[li] if c > 0 else 0                        # chunk 0 also carries the embeddings
        hi = layer_lo[first_layers[c + 1]] if c + 1 < len(first_layers) else layer_lo[L - 1] + sum(
            p_.numel() for k, p_ in model.named_parameters() if p_.requires_grad and k.startswith(f"model.layers.{L - 1}."))
        if c == 0:
            return                                                # reduced after backward together with the tail (embeddings finish last)
        pending.append(dist.all_reduce(flat[lo:hi], async_op=True))
        state.setdefault("done_hi", []).append((lo, hi))

    if world > 1 and args.overlap_allreduce:
        model.model.layer_grad_ready_hook = on_layer_grad_ready

    def reduce_gradients():
        if world == 1:
            return
        if not args.overlap_allreduce:
            dist.all_reduce(flat)
            return
        done = sorted(state.pop("done_hi", []))
        # everything not yet issued: [0, first issued lo) and [last issued hi, end)
        lo_issued = done[0][0] if done else n_train
        hi_issued = done[-1][1] if done else n_train
        if lo_issued > 0:
            pending.append(dist.all_reduce(flat[:lo_issued], async_op=True))
        if hi_issued < n_train:
            pending.append(dist.all_reduce(flat[hi_issued:], async_op=True))
        for w_ in pending:
            w_.wait()
        pending.clear()

    def step(inp, from_host: bool):
        if from_host:
            inp = {k: v.to(dev, non_blocking=True) for k, v in inp.items()}
        LF.begin_grad_step(params)
        total = None
        n_micro = B // MB
        for i in range(n_micro):
            state["armed"] = (i == n_micro - 1)
            loss = micro(inp, i) * (MB / B) / world
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        state["armed"] = False
        if world > 1:
            reduce_gradients()
        if opt is not None:
            opt.step()
        return float(total.item()) if from_host else total

    def timed(k, from_host):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        last = None
        for _ in range(k):
            last = step(host if from_host else resident, from_host)
        ev[1].record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1])
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    for _ in range(args.warmup):
        step(resident, False)
    torch.cuda.synchronize()
    mem_gb = torch.cuda.max_memory_allocated() / 2 ** 30

    _lib.reset_launch_counts()
    ops.enable_timing()
    sampler = ClockSampler(local) if rank == 0 else None
    if args.cuda_profiler_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ms, last_loss = timed(args.steps, False)
    if args.cuda_profiler_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if sampler else None
    kt = ops.disable_timing()
    launches = _lib.total_launches()
    tokens_step = B * T * world
    value = tokens_step * args.steps / (ms / 1e3)

    e2e = None
    if not args.no_e2e:
        for _ in range(1):
            step(host, True)
        ms_e, _ = timed(args.steps, True)
        e2e = {"value": tokens_step * args.steps / (ms_e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": 4, "ms_per_step": ms_e / args.steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the bridge-attention forward kernel (the kernel BASELINE.json's second metric names)
    peaks = load_peaks()
    fl = synthetic.decoder_flops(cfg, MB, T, args.images * synthetic.IMG)
    kern = {}
    for name, evs in (kt or {}).items():
        if evs:
            kern[name] = sum(s.elapsed_time(e) for s, e in evs) / len(evs)
    roof = None
    fwd_names = {"lb_attn_fwd_stream": "attn_fwd_stream_kernel<128,causal> (bridge attention forward, persistent)",
                 "lb_attn_fwd": "attn_fwd_kernel<128,causal> (bridge attention forward)"}
    fwd_key = next((k for k in fwd_names if k in kern), None)
    if fwd_key:
        t_ms = kern[fwd_key]
        ach = fl["attn_per_layer_fwd"] / (t_ms * 1e-3) / 1e12
        roof = {"kernel": fwd_names[fwd_key], "bound": "tensor", "achieved": ach,
                "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tflops_sustained"],
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}); kernel timed inside a long step",
                "traffic": attn_traffic(MB, T, cfg), "avg_launch_ms": t_ms,
                "algorithmic_flops_per_launch": fl["attn_per_layer_fwd"],
                "other_kernels_ms": {k: v for k, v in kern.items() if k != fwd_key},
                "bwd_achieved_tflops": (2.5 * fl["attn_per_layer_fwd"] / ((kern.get("lb_attn_bwd_dq", 0) + kern.get("lb_attn_bwd_dkv", 0)) * 1e-3) / 1e12)
                if kern.get("lb_attn_bwd_dq") else None}
    model_flops_step = 3.0 * fl["total"] * (B // MB)
    cb = None
    if not args.no_cpu_baseline and world == 1:
        try:
            cb = cpu_reference_sample(n_layers=1)
        except Exception as ex:      # the checker must never take the measured path down
            cb = {"error": repr(ex)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"Libra-11B train step (fwd+bwd{'+AdamW' if opt else ''}{'+grad all-reduce' if world > 1 else ''}), "
                               f"B={B} (micro-batch {MB}) x T={T} per GPU, {args.images} image(s)/sample, {cfg.num_hidden_layers} layers, "
                               f"all params trainable={not args.frozen_language}, ckpt={bool(args.ckpt)} -- BASELINE.json configs[2]",
                   "global_batch": B * world, "seq_len": T, "parallelism": f"dp{world}", "trainable_params": n_train,
                   "l2": "working set (>= 22 GB of weights per step) far exceeds the 126 MB L2; no explicit flush",
                   "tiny": bool(args.tiny)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cb,
        "loss": float(last_loss.item()) * world if last_loss is not None else None,      # rank 0's mean loss over its samples
        "model_tflops_per_gpu": model_flops_step / (ms / args.steps * 1e-3) / 1e12, "peak_mem_gb": mem_gb,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
