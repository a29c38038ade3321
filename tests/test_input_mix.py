"""CPU test: calm_trainer.MixBatch.draw samples what torchvision's RandomChoice([CutMix, MixUp]) samples (same generator,
same draw order), and the plain-torch restatement used as the GPU test's checker reproduces torchvision's outputs bit for bit
(distributed_trainer_cls.py:58-61)."""
import pytest
import torch
from torchvision.transforms import v2 as transforms

import calm_trainer


def mix_reference(images, labels, params, num_classes):
    """torchvision.transforms.v2 MixUp.transform / CutMix.transform / _mixup_label for explicit parameters."""
    onehot = torch.nn.functional.one_hot(labels, num_classes=num_classes).float()
    lam_t = params["lam_labels"]
    soft = onehot.roll(1, 0).mul_(1.0 - lam_t).add_(onehot.mul(lam_t))
    if params["mode"] == 0:
        lam = params["lam"]
        return images.roll(1, 0).mul_(1.0 - lam).add_(images.mul(lam)), soft
    x1, y1, x2, y2 = params["box"]
    out = images.clone()
    out[..., y1:y2, x1:x2] = images.roll(1, 0)[..., y1:y2, x1:x2]
    return out, soft


@pytest.mark.parametrize("seed", list(range(12)))
def test_draw_and_restatement_match_torchvision(seed):
    g = torch.Generator().manual_seed(100 + seed)
    x = torch.randn(5, 3, 32, 48, generator=g)
    y = torch.randint(0, 10, (5,), generator=g)
    cut_mix = transforms.CutMix(num_classes=10, alpha=1.0)
    mix_up = transforms.MixUp(num_classes=10, alpha=0.8)
    mix_both = transforms.RandomChoice([cut_mix, mix_up])
    torch.manual_seed(seed)
    want_x, want_y = mix_both(x, y)
    torch.manual_seed(seed)
    params = calm_trainer.MixBatch(num_classes=10, cutmix_alpha=1.0, mixup_alpha=0.8).draw(32, 48)
    got_x, got_y = mix_reference(x, y, params, 10)
    assert torch.equal(got_x, want_x)
    assert torch.equal(got_y, want_y)


def test_both_modes_are_drawn():
    mb = calm_trainer.MixBatch(num_classes=10)
    torch.manual_seed(0)
    modes = {mb.draw(32, 32)["mode"] for _ in range(40)}
    assert modes == {0, 1}


def test_cosine_annealing_lr_matches_torch_scheduler():
    """calm_trainer.cosine_annealing_lr vs torch's CosineAnnealingLR stepped once per epoch (distributed_trainer_cls.py:52,108-109)."""
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=3.1e-3)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=65, eta_min=1e-6)
    for epoch in range(65):
        assert calm_trainer.cosine_annealing_lr(epoch, 3.1e-3, 65, 1e-6) == pytest.approx(opt.param_groups[0]["lr"], rel=1e-9)
        opt.step()
        sched.step()
    assert calm_trainer.cosine_annealing_lr(65, 3.1e-3, 65, 1e-6) == pytest.approx(1e-6, rel=1e-9)
