"""Functional check of the higher-resolution configs (BASELINE configs[3]): 384^2 and 512^2 classification, fwd+bwd."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "calm-vit-dte_b200"))
import torch
import CALM_ViT_V2 as rvh
dev = torch.device("cuda:0")
for S, B, R, M in ((384, 8, 80, 240), (512, 4, 80, 240), (384, 8, 144, 416)):
    torch.manual_seed(0)
    m = rvh.ViT(dev, type=8, heads=12, seq_length=S, in_features=3 * S, dim_step=48, mean_var_hidden=M, seq_len_step=16,
                seq_len_reduce=R, out_features=1000, force_reduce=False, generate=False).to(dev)
    m.train()
    x = torch.randn(B, 3, S, S, device=dev)
    y = torch.softmax(torch.randn(B, 1000, device=dev) * 4, -1)
    for it in range(3):
        t0 = time.time()
        out, kl = m(x)
        loss = torch.nn.functional.cross_entropy(out, y)
        loss.backward()
        torch.cuda.synchronize()
        dt = time.time() - t0
        ok = all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
        for p in m.parameters(): p.grad = None
    print("S=%d latent (%d,%d) B=%d params %.1fM loss %.4f kl %.4f grads finite %s  step %.1f ms (eager, 3rd iter)" %
          (S, R, M, B, sum(p.numel() for p in m.parameters()) / 1e6, loss.item(), float(kl), ok, dt * 1e3), flush=True)
    del m
    torch.cuda.empty_cache()
