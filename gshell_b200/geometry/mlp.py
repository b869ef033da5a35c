"""SDF / mSDF field MLP (reference geometry/mlp.py + embedding.py).  OUT OF SCOPE for the CUDA work (SURVEY 2 #4:
upstream of the hot path, stays PyTorch/cuBLAS); provided so `FLAGS.use_sdf_mlp=True` keeps working."""
import torch
import torch.nn as nn


class Embedding(nn.Module):
    def __init__(self, in_channels, n_freqs):
        super().__init__()
        self.out_channels = in_channels * (2 * n_freqs + 1)
        self.register_buffer("freq_bands", 2 ** torch.linspace(0, n_freqs - 1, n_freqs), persistent=False)

    def forward(self, x):
        out = [x]
        for f in self.freq_bands:
            out += [torch.sin(f * x), torch.cos(f * x)]
        return torch.cat(out, -1)


class MLP(nn.Module):
    def __init__(self, n_freq=6, d_hidden=128, d_out=1, n_hidden=3, skip_in=(), use_float16=False):
        super().__init__()
        self.emb = Embedding(3, n_freq)
        self.skip_in = set(skip_in)
        self.first = nn.Linear(self.emb.out_channels, d_hidden)
        self.hidden = nn.ModuleList(
            [nn.Linear(d_hidden + (self.emb.out_channels if i in self.skip_in else 0), d_hidden) for i in range(n_hidden)])
        self.last = nn.Linear(d_hidden, d_out)
        self.act = nn.Softplus(beta=100)
        self.use_float16 = use_float16

    def forward(self, x):
        emb = self.emb(x)
        with torch.autocast("cuda", dtype=torch.float16, enabled=self.use_float16):
            h = self.act(self.first(emb))
            for i, lin in enumerate(self.hidden):
                h = self.act(lin(torch.cat([h, emb], -1) if i in self.skip_in else h))
            return self.last(h)
