"""N>1 host logic on CPU with gloo (world_size 2): flat-buffer gradient all-reduce == single-process gradients on the
concatenated batch; batch sharding.  The model here is a small torch module (the CUDA path cannot run on CPU): what is
under test is libra_b200.dist, which is device agnostic."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from libra_b200.dist import FlatGradBuffer, shard_batch
    m = _model()
    buf = FlatGradBuffer(m.parameters())
    g = torch.Generator().manual_seed(7)
    batch = {"x": torch.randn(8, 16, generator=g), "input_ids": torch.randint(0, 9, (2, 8, 5), generator=g), "none": None}
    sh = shard_batch(batch, rank, world)
    assert sh["x"].shape[0] == 4 and sh["input_ids"].shape == (2, 4, 5) and sh["none"] is None
    buf.zero()
    for mb in range(2):                                  # two micro-batches accumulate in place into the flat buffer
        x = sh["x"][mb * 2:(mb + 1) * 2]
        (m(x).pow(2).mean() / 2).backward()
    buf.all_reduce_mean(chunks=3 if rank >= 0 else 1)
    if rank == 0:
        ret.put(buf.flat.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_allreduce_equals_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    got = ret.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    m = _model()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 16, generator=g)
    # mean over ranks of (mean over each rank's 2 micro-batches) == mean over 4 groups of 2 samples
    loss = sum(m(x[i * 2:(i + 1) * 2]).pow(2).mean() for i in range(4)) / 4
    loss.backward()
    want = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    assert torch.allclose(got, want, atol=1e-6), (got - want).abs().max()


def test_flat_buffer_views_accumulate_in_place():
    from libra_b200.dist import FlatGradBuffer
    m = _model()
    buf = FlatGradBuffer(m.parameters())
    ptrs = [p.grad.data_ptr() for p in m.parameters()]
    m(torch.randn(3, 16)).sum().backward()
    m(torch.randn(3, 16)).sum().backward()
    assert [p.grad.data_ptr() for p in m.parameters()] == ptrs        # autograd kept accumulating into the views
    assert buf.flat.abs().sum() > 0
    off = 0
    for p in m.parameters():
        assert torch.equal(buf.flat[off:off + p.numel()].view_as(p), p.grad)
        off += p.numel()
