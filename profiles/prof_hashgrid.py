"""Timing probe of the hash-grid encoding (csrc/hashgrid.cu) at the size of one benchmark batch: 8 x 1024^2 surface points,
reference configuration (16 levels x 2 features, 2^19 entries).  usage: python profiles/prof_hashgrid.py [n_points]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gshell_b200.render import mlptexture   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8 * 1024 * 1024
dev = torch.device("cuda:0")
enc = mlptexture.HashGridEncoding(device=dev)
x = torch.rand(n, 3, device=dev, requires_grad=True)
g = torch.randn(n, enc.n_output_dims, device=dev)
for it in range(3):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    enc.params.grad = None
    x.grad = None
    torch.cuda.synchronize()
    e0.record()
    out = enc(x)
    e1.record()
    out.backward(g)
    e2.record()
    torch.cuda.synchronize()
fwd, bwd = e0.elapsed_time(e1), e1.elapsed_time(e2)
gathers = n * 16 * 8 * 8
print(f"hashgrid n={n}: fwd {fwd:.3f} ms ({gathers / fwd / 1e6:.0f} GB/s of 8-byte gathers, {n * 128 / fwd / 1e6:.0f} GB/s written), "
      f"bwd(table+x) {bwd:.3f} ms")
tex = mlptexture.MLPTexture3D(torch.tensor([[-1.0] * 3, [1.0] * 3], device=dev), channels=6,
                              min_max=[torch.zeros(6, device=dev), torch.ones(6, device=dev)])
pos = torch.rand(8, 1024, 1024, 3, device=dev) * 2 - 1
for it in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = tex.sample(pos)
    out.sum().backward()
    e1.record()
    torch.cuda.synchronize()
print(f"MLPTexture3D.sample + backward on 8 x 1024^2 points: {e0.elapsed_time(e1):.2f} ms")
