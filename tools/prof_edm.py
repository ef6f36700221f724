"""EDM (configs/image/mnist/edm.yaml: DDPM++ network, 18-step Heun sampler = 35 network evaluations) through
GaussianDiffusion_EDM.sample() at a few batch sizes: images/s, ms per network evaluation, kernels launched per evaluation."""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from xdiffusion_b200 import ops  # noqa: E402
from xdiffusion_b200.diffusion.edm import GaussianDiffusion_EDM  # noqa: E402
from xdiffusion_b200.utils import DotConfig  # noqa: E402

dev = "cuda"
cfg = torch.load(os.path.join("tests", "golden", "edm_net.pt"), weights_only=False)["config"]
torch.manual_seed(0)
m = GaussianDiffusion_EDM(DotConfig(cfg))
g = torch.Generator().manual_seed(1)
with torch.no_grad():
    for name, p in m.named_parameters():                     # the reference initialises these to ~1e-5: re-draw
        if name.endswith(("conv1.weight", "proj.weight", "aux_conv.weight")):
            p.copy_(torch.randn(p.shape, generator=g) * 0.02)
m = m.to(dev).eval()
for B in (64, 256, 1024):
    x = torch.randn(B, 1, 32, 32, device=dev)
    m.sample(num_samples=B, initial_noise=x)
    ops.LAUNCHES = 0
    m.sample(num_samples=B, initial_noise=x)
    launches = ops.LAUNCHES
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    reps = 3
    for _ in range(reps):
        out, _ = m.sample(num_samples=B, initial_noise=x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    wall = (time.perf_counter() - t0) / reps * 1e3
    flop = 2 * 0  # (reported as time only)
    print(f"EDM B={B}: {B / ms * 1e3:.1f} images/s, {ms:.1f} ms per 18-step sample() (wall {wall:.1f}), "
          f"{ms / 35:.2f} ms per network evaluation, {launches / 35:.0f} launches per evaluation, finite={bool(torch.isfinite(out).all())}")
