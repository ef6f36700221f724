"""LayerNorm-modulate + GEMM, fused (xd_ln_gemm_bf16_tc) vs the two-kernel path, DiT shapes, in-graph timing."""
import sys

import torch

sys.path.insert(0, ".")
from xdiffusion_b200 import ops  # noqa: E402

dev = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
D = 384
x = torch.randn(M, D, device=dev)
mod = torch.randn(M // 16, 6 * D, device=dev) * 0.1
shift, scale = mod[:, :D], mod[:, D:2 * D]
for (n, name, act) in [(1152, "qkv", 0), (1536, "fc1", ops.ACT_GELU)]:
    w = (torch.randn(n, D, device=dev) * D ** -0.5).bfloat16()
    bias = torch.randn(n, device=dev)
    out = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
    for fused in (True, False):
        ops.LN_GEMM_FUSED = fused
        call = lambda: ops.ln_linear(x, shift, scale, 16, w, bias, act=act, out=out)
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
        print(f"{name} fused={fused}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
