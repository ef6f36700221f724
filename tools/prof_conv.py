"""conv3x3 (implicit GEMM on tcgen05) at the UNet C1 shapes, batch 64, in-graph timing.  argv: [force_bn]"""
import sys

import torch

sys.path.insert(0, ".")
from xdiffusion_b200 import ops  # noqa: E402

dev = "cuda"
bn = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B = 64
for (hw, c, cs, co, res) in [(32, 128, 0, 128, True), (32, 256, 256, 128, False), (32, 384, 384, 128, False),
                             (16, 128, 128, 256, False), (16, 256, 0, 256, True), (16, 512, 512, 256, False),
                             (8, 256, 0, 256, True), (8, 512, 512, 256, False), (4, 256, 0, 256, True)]:
    x = torch.randn(B, hw, hw, c, device=dev).bfloat16()
    xs = torch.randn(B, hw, hw, cs, device=dev).bfloat16() if cs else None
    wp = (torch.randn(co, 9 * c + cs, device=dev) * (9 * c) ** -0.5).bfloat16()
    bias = torch.randn(co, device=dev)
    r = torch.randn(B, hw, hw, co, device=dev).bfloat16() if res else None
    out = torch.empty(B, hw, hw, co, device=dev, dtype=torch.bfloat16)
    call = lambda: ops.conv3x3(x, wp, bias, residual=r, xs=xs, out=out, force_bn=bn)
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
    fl = 2.0 * B * hw * hw * co * (9 * c + cs)
    print(f"conv {hw}x{hw} C={c}+{cs} -> {co} res={res}: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s")
