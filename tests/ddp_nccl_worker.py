"""Worker of tests/test_ddp_nccl_gpu.py — launched with torch.distributed.run, one rank per GPU (NCCL).

Checks calm_ddp.DataParallel's contract (DDP at distributed_trainer_cls.py:55) on the real NCCL / side-stream path:
  gradients after the wrapped backward == mean over ranks of the single-GPU gradients of each rank's own batch,
both for eager launches and for the step captured into a CUDA graph (calm_trainer.GraphedStep), and identical on all ranks.
Prints one line "DDP_NCCL_OK ..." on rank 0 when everything holds; any failure raises (non-zero exit).
"""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (os.path.join(ROOT, "calm-vit-dte_b200"), ROOT, os.path.join(HERE, "golden")):
    sys.path.insert(0, p)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import synth
    import CALM_ViT_V2 as rvh
    import calm_trainer
    from calm_ddp import DataParallel
    cfg = synth.CONFIGS["small_cls"]
    kw = {k: v for k, v in cfg.items() if k != "batch"}

    def make():
        m = rvh.ViT(dev, type=8, force_reduce=False, **kw).to(dev)
        st = synth.synth_state({k: tuple(v.shape) for k, v in m.state_dict().items()})
        m.load_state_dict(st)
        m.train()
        return m, {k: v.to(dev) for k, v in st.items()}

    x, y = synth.synth_input(dict(cfg, batch=2), seed=10 + rank)          # every rank its own batch
    x, y = x.to(dev), y.to(dev)

    def fwd_bwd(model):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out, _kl = model(x)
        loss, _acc = calm_trainer.soft_target_cross_entropy(out.squeeze(), y)
        loss.backward()

    def reset(model, state):
        model.load_state_dict(state)             # in place: u / v back to their start, parameter storage unchanged
        for p in model.parameters():
            p.grad = None
        torch.manual_seed(1234 + rank)           # per-rank latent noise stream, same for every arm

    # ---- expected: single-GPU gradients of this rank's batch, averaged over ranks by the test's own plain all_reduce
    solo, state = make()
    reset(solo, state)
    fwd_bwd(solo)
    expected = {}
    for k, p in solo.named_parameters():
        g = p.grad.detach().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        expected[k] = g / world
    del solo

    def compare(model, what):
        worst = 0.0
        for k, p in model.named_parameters():
            e = expected[k]
            err = ((p.grad - e).double().norm() / e.double().norm().clamp_min(1e-30)).item()
            worst = max(worst, err)
            assert err < 1e-5, (what, k, err)
            other = p.grad.detach().clone()
            dist.broadcast(other, src=0)
            assert torch.equal(other, p.grad), (what, k, "differs between ranks")
        return worst

    # ---- eager
    model, state = make()
    dp = DataParallel(model, bucket_mb=1.0)
    assert len(dp.buckets) >= 3
    reset(model, state)
    fwd_bwd(dp)
    torch.cuda.synchronize()
    w_eager = compare(model, "eager")
    assert dp.check_buffers()
    # ---- the same step captured into a CUDA graph together with its all-reduces. The captured step drops the gradients first
    # (optimizer.zero_grad() of the loop): backward then allocates them from the graph's pool and DataParallel leaves every
    # .grad aliasing its bucket slice, which each replay refills — so .grad must not be touched between capture and replay.
    def graph_step():
        for p in model.parameters():
            p.grad = None
        fwd_bwd(dp)

    reset(model, state)
    graphed = calm_trainer.GraphedStep(graph_step, warmup=2, capture=True)
    assert graphed.graph is not None
    model.load_state_dict(state)
    torch.manual_seed(1234 + rank)
    graphed()
    torch.cuda.synchronize()
    w_graph = compare(model, "graph")
    assert dp.check_buffers()
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("DDP_NCCL_OK world=%d buckets=%d worst_rel_err eager=%.2e graph=%.2e" % (world, len(dp.buckets), w_eager, w_graph), flush=True)
    calm_trainer.release_graphs()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
