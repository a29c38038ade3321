"""World-size-2 CPU (gloo) test of the data-parallel wrapper's host logic (calm_ddp.DataParallel): rank-0 broadcast at
construction, bucketing in reverse registration order, gradient-ready hooks, averaging, .grad aliasing the buckets —
i.e. DDP's contract at distributed_trainer_cls.py:55. The NCCL/side-stream path is the same code with CUDA streams."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "calm-vit-dte_b200"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _net(seed):
    torch.manual_seed(seed)
    net = torch.nn.Sequential(torch.nn.Linear(12, 64), torch.nn.GELU(), torch.nn.Linear(64, 64), torch.nn.GELU(),
                              torch.nn.Linear(64, 5))
    net.register_buffer("u", torch.randn(7))
    return net


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from calm_ddp import DataParallel
    net = _net(100 + rank)                                     # different init per rank: ctor must broadcast rank 0's
    dp = DataParallel(net, bucket_mb=0.01)                     # ~2.6k floats per bucket -> more than one bucket
    assert len(dp.buckets) >= 2
    assert dp.check_buffers()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 12, generator=g)[rank * 4:(rank + 1) * 4]
    for step in range(2):                                      # second step: .grad was reset, hooks re-arm
        for p in net.parameters():
            p.grad = None
        dp(x).pow(2).mean().backward()
    torch.save({k: p.grad.clone() for k, p in net.named_parameters()}, out % rank)
    torch.save(net.state_dict(), (out % rank) + ".sd")
    dist.destroy_process_group()


def test_data_parallel_matches_full_batch(tmp_path):
    world, port = 2, _free_port()
    out = str(tmp_path / "g%d.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    g0, g1 = torch.load(out % 0), torch.load(out % 1)
    net = _net(100)                                            # rank 0's weights
    sd0 = torch.load((out % 0) + ".sd")
    assert all(torch.equal(sd0[k], v) for k, v in net.state_dict().items())
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 12, generator=g)
    # average over ranks of per-rank mean losses == mean over the two half-batches
    (0.5 * (net(x[:4]).pow(2).mean() + net(x[4:]).pow(2).mean())).backward()
    for k, p in net.named_parameters():
        assert torch.allclose(g0[k], p.grad, rtol=1e-5, atol=1e-7), k
        assert torch.equal(g0[k], g1[k]), k
