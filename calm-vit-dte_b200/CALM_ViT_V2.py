"""B200-native drop-in for the reference's model assembly (CALM-ViT/CALM_ViT_V2.py:21-118).

`import CALM_ViT_V2 as rvh` (distributed_trainer_cls.py:8) keeps working: same `ViT` constructor, `forward(q) ->
(x, kl_loss)`, state_dict keys and `save_samples`. The dataset / augmentation / dev-script parts of the reference file
(CALM_ViT_V2.py:86-111,120-240) are data-pipeline code outside the hot path and are not reproduced.
"""
import os

import torch
from torch.nn.utils import spectral_norm as sn

import Vi_Tools_CNN_less_V2 as vt
import calm_ops as ops
from calm_ops import GroupSpec

parent_dir = "/config"


class ViT(torch.nn.Module):
    def __init__(self, device, type=8, heads=12, seq_length=256, in_features=768, dim_step=48, mean_var_hidden=192,
                 seq_len_step=16, seq_len_reduce=128, out_features=1000, force_reduce=False, generate=True):
        super().__init__()
        self.device = device
        self.generate = generate
        self.num_classes = out_features
        self.seq_length = seq_length
        if type == 8:
            self.autoencoder = vt.EncoderDecoder_8(
                heads=heads, dim1=in_features, dim_step=dim_step, mean_var_hidden=mean_var_hidden, seq_length=seq_length,
                seq_len_step=seq_len_step, seq_len_reduce=seq_len_reduce, out_features_override=None,
                force_reduce=force_reduce).to(device)
        if not generate:
            self.pool = torch.nn.AdaptiveAvgPool1d(1).to(device)
            self.head = torch.nn.Sequential(
                sn(torch.nn.Linear(in_features, in_features * 2, bias=False)).to(device),
                torch.nn.GELU().to(device),
                sn(torch.nn.Linear(in_features * 2, out_features, bias=False)).to(device),
            ).to(device)
        else:
            self.proj = vt._cnn(32)

    def _sn_groups(self):
        if self.generate:
            return vt._cnn_groups(self.proj)
        return [GroupSpec([self.head[0]]), GroupSpec([self.head[2]])]

    def forward(self, q):
        x, kl_loss = self.autoencoder(q)
        with ops.Scope(self, self._sn_groups, self.training) as sc:
            if not self.generate:
                # mean over the sequence axis (permute + AdaptiveAvgPool1d(1) + squeeze, :73-75), then the 2-layer head
                pooled = ops.SeqMeanFn.apply(x)
                x = ops.MlpFn.apply(pooled, sc.token, None, sc.bank, sc.bank.gid(self.head[0]), sc.bank.gid(self.head[2]), True)
            else:
                x = vt._cnn_apply(sc, self.proj, x)
        return x, kl_loss


def save_samples(imgs, mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]):
    """Writes sigmoid(imgs) as PNGs (reference :113-118). matplotlib is imported lazily: it is only needed here."""
    import matplotlib.pyplot as plt
    imgs = torch.sigmoid(imgs)
    os.makedirs(f"{parent_dir}/Codebase/samples", exist_ok=True)
    for i, img in enumerate(imgs):
        plt.imsave(f"{parent_dir}/Codebase/samples/sample_{i}.png", img.permute(1, 2, 0).detach().float().cpu().numpy())
