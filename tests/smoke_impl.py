"""__graft_entry__.smoke(): one tiny invocation of the hot path on cuda:0, checked against the oracle."""
import os

import torch

from tests.conftest import load_golden
from tests.helpers import oracle_model, product_model, rel_l2


def run():
    fx = load_golden("c2")
    m = product_model(fx)
    om = oracle_model(fx)
    g = torch.Generator().manual_seed(0)
    B, K = 4, 3
    x_T = torch.randn(B, 1, 32, 32, generator=g)
    noise = torch.randn(K, B, 1, 32, 32, generator=g)
    classes = torch.randint(0, 10, (B,), generator=g)
    ref = om.sample(x_T, [noise[i] for i in range(K)], ctx={"classes": classes}, num_sampling_steps=K)
    out, _ = m.sample(context={"classes": classes.cuda()}, num_samples=B, num_sampling_steps=K,
                      initial_noise=x_T.cuda(), noise=noise.cuda())
    err = rel_l2(out, ref)
    print(f"smoke: DiT B={B} {K}-step sampling loop on cuda:0, rel L2 vs oracle = {err:.3e}")
    assert err < 2e-2, err
