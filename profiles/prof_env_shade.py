"""One traced forward (generate -> trace -> shade) + one backward launch of the MC integrator on a synthetic G-buffer, for
`ncu --set full`: the first three k_env_shade launches are GEN (mode 2), FWD (mode 0), BWD (mode 1).
usage (GPU box): ncu --set full --clock-control none --import-source on -k regex:k_env_shade -c 3 -o gpurun_out/env_shade python profiles/prof_env_shade.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gshell_b200.render import light, optixutils as ou   # noqa: E402

B, H, W, n = 2, 512, 512, int(os.environ.get("N_SAMPLES", "16"))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
nrm = torch.nn.functional.normalize(torch.randn(B, H, W, 3, generator=g), dim=-1)
nrm[..., 2] = nrm[..., 2].abs()
pos = (torch.rand(B, H, W, 3, generator=g) - 0.5).to(dev).requires_grad_()
nrm = nrm.to(dev).requires_grad_()
view = torch.tensor([0.0, 0.0, 3.0], device=dev).view(1, 1, 1, 3).expand(B, 1, 1, 3)
kd = torch.rand(B, H, W, 3, generator=g).to(dev).requires_grad_()
ks = torch.stack([torch.zeros(B, H, W), 0.08 + 0.9 * torch.rand(B, H, W, generator=g), torch.rand(B, H, W, generator=g)], -1).to(dev).requires_grad_()
mask = torch.ones(B, H, W, device=dev)
lgt = light.create_trainable_env_rnd(256, device=dev)
# a far-away occluder: every sample in the upper hemisphere gets a shadow ray (GEN runs), almost none hits
ctx = ou.OptiXContext()
ou.optix_build_bvh(ctx, torch.tensor([[50.0, 50.0, 60.0], [51.0, 50.0, 60.0], [50.0, 51.0, 60.0]], device=dev),
                   torch.tensor([[0, 1, 2]], dtype=torch.int32, device=dev), 1)
for it in range(int(os.environ.get("REPS", "1"))):
    d, s = ou.optix_env_shade(ctx, mask, pos.detach(), pos, nrm, view, kd, ks, lgt.base, lgt._pdf, lgt.rows[:, 0], lgt.cols,
                              BSDF="pbr", n_samples_x=n, rnd_seed=it, shadow_scale=1.0)
    (d.sum() + s.sum()).backward()
torch.cuda.synchronize()
print("done", float(d.mean()), float(s.mean()))
