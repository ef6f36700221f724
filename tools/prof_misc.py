"""In-graph timing of the DiT per-step kernels outside the transformer blocks (batch 1024)."""
import sys

import torch

sys.path.insert(0, ".")
from xdiffusion_b200 import ops  # noqa: E402

dev = "cuda"
B, T, D = 1024, 16, 384


def timeit(name, call, n=20):
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            call()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / n * 1e3:.1f} us")


silu_c = torch.randn(B, D, device=dev).bfloat16()
w_ada = (torch.randn(74 * D, D, device=dev) * 0.05).bfloat16()
b_ada = torch.randn(74 * D, device=dev)
mod = torch.empty(B, 74 * D, device=dev)
timeit("adaLN conditioning GEMM [1024 x 28416 x 384] fp32 out", lambda: ops.linear(silu_c, w_ada, b_ada, out=mod))
h = torch.randn(B * T, D, device=dev)
pos = torch.randn(T, D, device=dev)
h2 = torch.empty_like(h)
timeit("add_rows_periodic (pos embed)", lambda: torch.ops.xdb200.add_rows_periodic(h, pos, T, h2))
a = ops.layernorm_modulate(h, mod[:, :D], mod[:, D:2 * D], T)
timeit("layernorm_modulate", lambda: ops.layernorm_modulate(h, mod[:, :D], mod[:, D:2 * D], T))
qkv = torch.randn(B, T, 3, 6, 64, device=dev).bfloat16()
q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
timeit("attention T=16", lambda: ops.attention(q, k, v, 0.125))
