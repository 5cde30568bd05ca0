#!/usr/bin/env python
"""Data-parallel parity on real GPUs (run under torchrun, world_size >= 2):
gradients after the flat-buffer NCCL all-reduce == single-GPU gradients on the concatenated batch (bf16 tolerance).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dp_parity.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from libra_b200 import _lib
from libra_b200.dist import FlatGradBuffer, GradSync, shard_batch
from libra_b200.models import LibraConfig, LibraForCausalLM
from libra_b200.synthetic import libra_batch, randomize_for_bench


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device()
    dist.init_process_group("nccl", device_id=dev)
    cfg = LibraConfig(hidden_size=256, intermediate_size=704, num_hidden_layers=2, num_attention_heads=2, vocab_size=512,
                      contiguous_signal_size=64)
    torch.manual_seed(0)
    model = LibraForCausalLM(cfg).to(torch.bfloat16).to(dev).train()
    randomize_for_bench(model, seed=1, std=0.05)
    B = 2 * world
    full = libra_batch(B, 700, 1, vocab=cfg.vocab_size, signal=cfg.contiguous_signal_size, seed=3, device=dev)
    mine = shard_batch({k: full[k] for k in ("input_ids", "vision_indices", "contiguous_signal", "labels")}, rank, world)
    buf = FlatGradBuffer(model.named_parameters())
    sync = GradSync(buf, model.model, min_bytes=1 << 16)
    buf.begin_step()
    sync.arm(last=True)
    out = model(input_ids=mine["input_ids"], vision_indices=mine["vision_indices"], contiguous_signal=mine["contiguous_signal"],
                labels=mine["labels"])
    (out.loss / world).backward()
    sync.finish()
    dp = buf.flat.float().clone()
    if rank == 0:
        print(f"[dp_parity] all-reduce pieces: {len(sync.pieces)} over {buf.numel} elements")
    # single-process reference on the whole batch: mean over ranks of per-shard mean losses == loss of equal-size shards
    buf.begin_step()
    tot = 0.0
    for r in range(world):
        sh = shard_batch({k: full[k] for k in ("input_ids", "vision_indices", "contiguous_signal", "labels")}, r, world)
        l = model(input_ids=sh["input_ids"], vision_indices=sh["vision_indices"], contiguous_signal=sh["contiguous_signal"],
                  labels=sh["labels"]).loss / world
        l.backward()
        tot += float(l)
    ref = buf.flat.float()
    rel = ((dp - ref).norm() / ref.norm()).item()
    ok = rel < 2e-2
    if rank == 0:
        print(f"[dp_parity] world={world} rel_fro(dp_grads, single_gpu_grads)={rel:.3e} loss={tot:.4f} -> {'OK' if ok else 'FAIL'}")
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
