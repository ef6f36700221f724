"""In-kernel cycle accounting for one UNet conv shape (needs a -DXDB200_INSTRUMENT build).  argv: hw C Cs Cout"""
import os
import sys

os.environ.setdefault("XDB200_PROF", "1")
import torch

sys.path.insert(0, ".")
from xdiffusion_b200 import ops  # noqa: E402

hw, c, cs, co = (int(a) for a in sys.argv[1:5])
dev, B = "cuda", 64
x = torch.randn(B, hw, hw, c, device=dev).bfloat16()
xs = torch.randn(B, hw, hw, cs, device=dev).bfloat16() if cs else None
wp = (torch.randn(co, 9 * c + cs, device=dev) * (9 * c) ** -0.5).bfloat16()
bias = torch.randn(co, device=dev)
r = torch.randn(B, hw, hw, co, device=dev).bfloat16() if not cs else None
out = torch.empty(B, hw, hw, co, device=dev, dtype=torch.bfloat16)
for rep in range(3):
    if rep == 2:
        print("=====", file=sys.stderr, flush=True)
    ops.conv3x3(x, wp, bias, residual=r, xs=xs, out=out)
    torch.cuda.synchronize()
