"""Diagnostic: where do sampled directions differ between the CUDA samplers and the oracle?"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from iris_b200 import core
from oracle import estimators as E
dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(9)
N = 200000
u = torch.rand(N, 3, generator=gen)
nn = torch.nn.functional.normalize(torch.randn(N, 3, generator=gen), dim=-1)
ww = torch.nn.functional.normalize(torch.randn(N, 3, generator=gen), dim=-1)
rr = torch.rand(N, 1, generator=gen) * 0.98 + 0.02
mat = torch.cat([torch.rand(N, 3, generator=gen), rr, torch.rand(N, 1, generator=gen)], 1)
wi = core.bsdf_sample(0, u[:, :2].contiguous().to(dev), None, nn.to(dev))[0].cpu()
ref = E.diffuse_sampler(u[:, :2], nn)
print("diffuse equal frac", (wi == ref).all(1).float().mean().item(), "max abs", (wi - ref).abs().max().item())
wi = core.bsdf_sample(1, u[:, :2].contiguous().to(dev), ww.to(dev), nn.to(dev), mat=mat.to(dev))[0].cpu()
ref = E.specular_sampler(u[:, :2], rr, ww, nn)
ok = ((wi == ref) | (wi.isnan() & ref.isnan())).all(1)
print("specular equal frac", ok.float().mean().item(), "max abs", torch.nan_to_num(wi - ref).abs().max().item())
