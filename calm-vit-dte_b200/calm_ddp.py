"""Single-node data parallelism for the CALM-ViT hot path: one process per GPU, bucketed gradient all-reduce over
NCCL/NVLink launched from gradient-ready hooks on a side stream so that it overlaps the rest of backward.

Replaces `DDP(model, device_ids=[local_rank])` of the reference's per-rank loop (distributed_trainer_cls.py:55; SURVEY
§2c): parameters and buffers are broadcast from rank 0 at construction, gradients are averaged over ranks (sum / world)
in ~25 MB buckets filled in reverse registration order (the order backward produces them). The spectral-norm vectors
u/v are NOT re-broadcast every forward (the reference's DDP does, `broadcast_buffers=True`): every rank applies the same
deterministic power iteration to identical weights, so they stay identical — `check_buffers()` verifies it.
The path has no other exchange step: the model shards along the batch only.

`comm_ctas=N > 0` moves the all-reduces to a communicator of their own whose kernels are capped at N thread blocks (NCCL's
`max_ctas`), the knob for keeping NCCL's channel CTAs off the SMs of the persistent GEMM / attention grids they overlap with.
Measured on 2 B200s (trainer config, 170 MB of gradients in 7 buckets, the step captured with and without its all-reduces,
profiles/r02d_scale2_ctas.txt): the exchange costs 0.51 ms of a 49.6 ms step with NCCL's own choice, 0.37 ms at 8 CTAs,
1.8 ms at 4 and 2.7 ms at 2 — the cost is the exposed all-reduce of the LAST buckets (the first Block's weight gradients all
become ready when its spectral-norm bank runs backward, after which only the optimizer is left), which a cap slows down, not
SM contention. The default is therefore 0 = the default process group and NCCL's own channel count.
"""
import os

import torch
import torch.distributed as dist

DEFAULT_COMM_CTAS = 0


def _capped_nccl_group(max_ctas):
    """A process group over all ranks whose NCCL kernels use at most `max_ctas` CTAs (every rank must call this)."""
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.min_ctas = 1
    opts.config.max_ctas = int(max_ctas)
    return dist.new_group(ranks=list(range(dist.get_world_size())), backend="nccl", pg_options=opts)


class DataParallel(torch.nn.Module):
    def __init__(self, module, bucket_mb=25.0, process_group=None, comm_ctas=None):
        super().__init__()
        self.module = module
        self.pg = process_group
        self.comm_ctas = 0
        if process_group is None and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl":
            if comm_ctas is None:
                comm_ctas = int(os.environ.get("CALM_DDP_CTAS", DEFAULT_COMM_CTAS))
            if comm_ctas > 0:
                self.pg = _capped_nccl_group(comm_ctas)
                self.comm_ctas = int(comm_ctas)
        self.world = dist.get_world_size(self.pg) if dist.is_initialized() else 1
        self.backend = dist.get_backend(self.pg) if dist.is_initialized() else None
        params = [p for p in module.parameters() if p.requires_grad]
        self.device = params[0].device
        self.on_cuda = self.device.type == "cuda"
        self.comm = torch.cuda.Stream(device=self.device) if self.on_cuda else None
        if self.world > 1:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t, src=0, group=self.pg)
        # buckets in reverse registration order
        cap = int(bucket_mb * 1024 * 1024 / 4)
        self.buckets, cur, n = [], [], 0
        for p in reversed(params):
            cur.append(p)
            n += p.numel()
            if n >= cap:
                self.buckets.append(cur)
                cur, n = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat, self.views, self.bucket_of = [], [], {}
        for bi, bucket in enumerate(self.buckets):
            # slices start on 16-byte boundaries (4 floats) so that the optimizer kernels keep their 128-bit accesses
            flat = torch.zeros(sum(-(-p.numel() // 4) * 4 for p in bucket), dtype=torch.float32, device=self.device)
            views, off = [], 0
            for p in bucket:
                views.append(flat[off: off + p.numel()].view_as(p))
                off += -(-p.numel() // 4) * 4
                self.bucket_of[p] = bi
            self.flat.append(flat)
            self.views.append(views)
        self.pending = [0] * len(self.buckets)
        self.works = []
        self.callback_queued = False
        self.require_sync = True
        if self.world > 1:
            for p in params:
                p.register_post_accumulate_grad_hook(self._on_grad_ready)

    def forward(self, *args, **kw):
        return self.module(*args, **kw)

    # ---------------------------------------------------------------------------------------------------------------
    def _on_grad_ready(self, p):
        if not self.require_sync:
            return
        if not self.callback_queued:
            torch.autograd.Variable._execution_engine.queue_callback(self._finalize)
            self.callback_queued = True
        bi = self.bucket_of[p]
        self.pending[bi] += 1
        if self.pending[bi] == len(self.buckets[bi]):
            self._reduce_bucket(bi)

    def _reduce_bucket(self, bi):
        grads = [p.grad for p in self.buckets[bi]]
        flat = self.flat[bi]
        if self.on_cuda:
            ready = torch.cuda.current_stream().record_event()
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ready)
                torch._foreach_copy_(self.views[bi], grads)
                self.works.append(self._allreduce(flat))
        else:
            torch._foreach_copy_(self.views[bi], grads)
            self.works.append(self._allreduce(flat))

    def _allreduce(self, flat):
        if self.backend == "nccl":
            return dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.pg, async_op=True)
        w = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        w.wait()
        flat.div_(self.world)
        return None

    def _finalize(self):
        """End of backward: wait for the buckets, then let every .grad alias its (averaged) bucket slice."""
        # once any hook has fired, EVERY bucket must be complete: a bucket without gradients would otherwise hand last step's
        # averaged gradient back through the aliased .grad
        if any(n != len(b) for n, b in zip(self.pending, self.buckets)):
            raise RuntimeError("a parameter produced no gradient this step; data-parallel buckets cannot be completed "
                               "(the reference runs DDP without find_unused_parameters as well)")
        if self.on_cuda:
            with torch.cuda.stream(self.comm):
                for w in self.works:
                    if w is not None:
                        w.wait()
            torch.cuda.current_stream().wait_stream(self.comm)
        else:
            for w in self.works:
                if w is not None:
                    w.wait()
        for bucket, views in zip(self.buckets, self.views):
            for p, v in zip(bucket, views):
                p.grad = v
        self.pending = [0] * len(self.buckets)
        self.works = []
        self.callback_queued = False

    def check_buffers(self):
        """True if every buffer (u/v power-iteration vectors) is identical on all ranks."""
        if self.world == 1:
            return True
        ok = True
        for b in self.module.buffers():
            ref = b.detach().clone()
            dist.broadcast(ref, src=0, group=self.pg)
            ok = ok and bool(torch.equal(ref, b))
        flag = torch.tensor([1.0 if ok else 0.0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.pg)
        return bool(flag.item() == 1.0)
