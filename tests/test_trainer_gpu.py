"""GPU parity of the training-step glue against the torch objects the reference loop uses:
CrossEntropyLoss with soft targets, HuberLoss + 0.1 kl, GradScaler + clip_grad_norm_ + AdamW
(distributed_trainer_cls.py:63-64,84-102,158; distributed_trainer_reg.py:76-98). fp32 arithmetic: tolerance 1e-5 relative
(different summation orders), skip/scale bookkeeping exact."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


@pytest.mark.parametrize("B,C", [(256, 1000), (8, 1000), (3, 17)])
def test_soft_target_cross_entropy_matches_torch(B, C):
    import calm_trainer
    g = torch.Generator(device="cuda").manual_seed(B * 31 + C)
    x = (torch.randn(B, C, device="cuda", generator=g) * 3).requires_grad_()
    t = torch.softmax(4 * torch.randn(B, C, device="cuda", generator=g), dim=1)
    ref = torch.nn.functional.cross_entropy(x, t)
    (gref,) = torch.autograd.grad(ref * 65536.0, x)
    x2 = x.detach().clone().requires_grad_()
    loss, acc = calm_trainer.soft_target_cross_entropy(x2, t)
    loss.backward(gradient=torch.tensor(65536.0, device="cuda"))
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert _rel(x2.grad, gref) < 1e-5
    want_acc = (x.argmax(1) == t.argmax(1)).float().mean().item()
    assert acc.item() == pytest.approx(want_acc, abs=1e-7)


def test_soft_target_cross_entropy_class_indices_and_bf16_logits():
    import calm_trainer
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(8, 1000, device="cuda", generator=g).bfloat16().requires_grad_()
    y = torch.randint(0, 1000, (8,), device="cuda", generator=g)
    ref = torch.nn.functional.cross_entropy(x.float(), y)
    (gref,) = torch.autograd.grad(ref, x)
    x2 = x.detach().clone().requires_grad_()
    loss, _ = calm_trainer.soft_target_cross_entropy(x2, y)
    loss.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert x2.grad.dtype == torch.bfloat16 and _rel(x2.grad, gref) < 1e-2


@pytest.mark.parametrize("B,S", [(2, 224), (3, 20)])
def test_huber_kl_matches_torch(B, S):
    import calm_trainer
    g = torch.Generator(device="cuda").manual_seed(S)
    y_hat = (torch.randn(B, S, 3 * S, device="cuda", generator=g) * 1.5).requires_grad_()
    img = torch.randn(B, 3, S, S, device="cuda", generator=g)
    kl = torch.tensor(0.37, device="cuda", requires_grad=True)
    ref = torch.nn.functional.huber_loss(y_hat.reshape(-1, S, S, 3).permute(0, 3, 1, 2), img, delta=1.0) + kl * 0.1
    gy, gk = torch.autograd.grad(ref * 1024.0, (y_hat, kl))
    y2, k2 = y_hat.detach().clone().requires_grad_(), kl.detach().clone().requires_grad_()
    loss, hub = calm_trainer.huber_kl_loss(y2, img, k2, 0.1, 1.0)
    loss.backward(gradient=torch.tensor(1024.0, device="cuda"))
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert abs(hub.item() - (ref.item() - 0.037)) <= 1e-5
    assert _rel(y2.grad, gy) < 1e-6
    assert k2.grad.item() == pytest.approx(gk.item(), rel=1e-6)


def _make_params(seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    shapes = [(672, 672), (1344, 672), (672,), (3,), (32, 3, 1, 1), (28,), (1000, 1344), (1,), (8193,), (176, 224)]
    return [torch.nn.Parameter(torch.randn(*s, device="cuda", generator=g) * 0.05) for s in shapes], g


def _torch_loop(params, grads_per_step, growth_interval):
    from torch.amp import GradScaler
    opt = torch.optim.AdamW(params, lr=3.1e-3, weight_decay=0.02, betas=(0.9, 0.98), fused=True)
    scaler = GradScaler(enabled=True, growth_interval=growth_interval)
    norms, scales = [], []
    for grads in grads_per_step:
        scale = scaler.get_scale()
        for p, gr in zip(params, grads):
            p.grad = (gr * scale).clone()
        # GradScaler only records the scale it must divide by once scale() has been called
        scaler.scale(torch.zeros((), device="cuda"))
        scaler.unscale_(opt)
        norms.append(torch.nn.utils.clip_grad_norm_(params, max_norm=1, error_if_nonfinite=False).item())
        scaler.step(opt)
        scaler.update()
        opt.zero_grad()
        scales.append(scaler.get_scale())
    return opt, norms, scales


def test_trainer_step_matches_gradscaler_clip_adamw():
    import calm_trainer
    ref_params, g = _make_params(11)
    our_params = [torch.nn.Parameter(p.detach().clone()) for p in ref_params]
    steps = []
    for k in range(6):
        grads = [torch.randn(p.shape, device="cuda", generator=g) * (0.02 if k != 4 else 1e-4) for p in ref_params]
        if k == 2:
            grads[3][1] = float("inf")            # this step must be skipped and the scale halved
        steps.append(grads)
    opt, norms, scales = _torch_loop(ref_params, steps, growth_interval=2)
    ts = calm_trainer.TrainerStep(our_params, lr=3.1e-3, weight_decay=0.02, betas=(0.9, 0.98), max_norm=1.0, growth_interval=2)
    for k, grads in enumerate(steps):
        scale = ts.get_scale()
        for p, gr in zip(our_params, grads):
            p.grad = (gr * scale).clone()
        ts.step()
        ts.zero_grad()
        assert ts.get_scale() == scales[k], (k, ts.get_scale(), scales[k])
        if k == 2:
            assert ts.found_inf.item() == 1.0
        else:
            assert ts.found_inf.item() == 0.0
            assert ts.grad_norm.item() == pytest.approx(norms[k], rel=1e-5)
    assert ts.step_count.item() == 5.0
    for i, (a, b) in enumerate(zip(our_params, ref_params)):
        assert _rel(a, b) < 1e-5, i
        m, v = ts.moments(i)
        st = opt.state[b]
        assert _rel(m, st["exp_avg"]) < 1e-5 and _rel(v, st["exp_avg_sq"]) < 1e-5, i


def test_trainer_step_in_cuda_graph_and_lr_update():
    import calm_trainer
    params, g = _make_params(3)
    twin = [torch.nn.Parameter(p.detach().clone()) for p in params]
    def loss_of(ps):
        return sum((p ** 2).sum() for p in ps) * 0.5

    def one(ts, ps):
        ts.backward(loss_of(ps))
        ts.step()
        ts.zero_grad()

    eager = calm_trainer.TrainerStep(twin, lr=1e-2, weight_decay=0.02, betas=(0.9, 0.98))
    graphed = calm_trainer.TrainerStep(params, lr=1e-2, weight_decay=0.02, betas=(0.9, 0.98))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        one(graphed, params)
    torch.cuda.current_stream().wait_stream(s)
    one(eager, twin)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        one(graphed, params)
    one(eager, twin)                                  # capture does not execute; replay below runs step 2
    graph.replay()
    for ts in (eager, graphed):
        ts.set_lr(5e-3)
    one(eager, twin)
    graph.replay()
    torch.cuda.synchronize()
    assert graphed.step_count.item() == 3.0 and eager.step_count.item() == 3.0
    for a, b in zip(params, twin):
        assert _rel(a, b) < 1e-6


def test_state_dict_round_trip():
    import calm_trainer
    params, g = _make_params(9)
    ts = calm_trainer.TrainerStep(params, lr=1e-3)
    for p in params:
        p.grad = torch.randn(p.shape, device="cuda", generator=g) * ts.get_scale()
    ts.step()
    sd = ts.state_dict()
    other = calm_trainer.TrainerStep([torch.nn.Parameter(p.detach().clone()) for p in params], lr=7.0)
    other.load_state_dict(sd)
    assert torch.equal(other.exp_avg, ts.exp_avg) and torch.equal(other.exp_avg_sq, ts.exp_avg_sq)
    assert other.get_lr() == pytest.approx(1e-3) and other.step_count.item() == 1.0 and other.get_scale() == ts.get_scale()


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 5, 8])
@pytest.mark.parametrize("shape", [(6, 3, 224, 224), (3, 3, 30, 42)])
def test_mix_batch_matches_torchvision_semantics(seed, shape):
    """Device CutMix / MixUp vs the restatement that tests/test_input_mix.py pins on torchvision itself (bit-exact)."""
    import calm_trainer
    from test_input_mix import mix_reference
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g)
    y = torch.randint(0, 1000, (shape[0],), generator=g)
    mb = calm_trainer.MixBatch(num_classes=1000)
    torch.manual_seed(seed)
    params = mb.draw(shape[2], shape[3])
    want_x, want_y = mix_reference(x, y, params, 1000)
    got_x, got_y = mb(x.cuda(), y.cuda(), params)
    assert torch.equal(got_x.cpu(), want_x)
    assert torch.equal(got_y.cpu(), want_y)
