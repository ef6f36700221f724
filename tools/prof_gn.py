"""GroupNorm(+SiLU) timing at the UNet C1 shapes (batch 64): 20 launches replayed from a CUDA graph."""
import sys

import torch

sys.path.insert(0, ".")
from xdiffusion_b200 import ops  # noqa: E402

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for (hw, c, ctot) in [(32, 128, 128), (32, 256, 256), (32, 384, 384), (16, 256, 256), (16, 512, 512), (8, 256, 256),
                      (8, 512, 512), (4, 256, 256), (32, 128, 384)]:
    buf = torch.randn(B, hw * hw, ctot, device=dev).bfloat16()
    x = buf[:, :, :c]
    gamma = torch.randn(c, device=dev)
    beta = torch.randn(c, device=dev)
    out = torch.empty(B, hw * hw, c, device=dev, dtype=torch.bfloat16)
    call = lambda: ops.groupnorm(x, gamma, beta, silu=True, out=out)
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            call()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    mb = B * hw * hw * c * 4 / 1e6
    print(f"GN {B}x{hw}x{hw}x{c} (ld {ctot}): {us:.1f} us  ({mb:.1f} MB r+w -> {mb / us:.2f} TB/s)")
