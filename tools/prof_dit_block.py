"""Device time of the fused DiT half-block kernels (csrc/dit_block.cu) at the benchmark shape, in-graph (20 launches
replayed from a CUDA graph, CUDA events), next to the unfused launches they replace.

    python tools/prof_dit_block.py [batch]
"""
import sys

import torch

sys.path.insert(0, ".")
from xdiffusion_b200 import ops  # noqa: E402

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T, D, Hd = 16, 384, 1536
M = B * T


def timed(call, reps=20):
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            call()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps * 1e3)
    return sorted(ts)[2]


bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
o, h = bf(M, D), torch.randn(M, D, device=dev)
wq, wp, w1, w2 = bf(3 * D, D) * D ** -0.5, bf(D, D) * D ** -0.5, bf(Hd, D) * D ** -0.5, bf(D, Hd) * Hd ** -0.5
bq, bp, b1, b2 = (torch.randn(n, device=dev) * 0.1 for n in (3 * D, D, Hd, D))
mod = torch.randn(B, 6 * D, device=dev) * 0.1
s1, sc1, g1, s2, sc2, g2 = (mod[:, i * D:(i + 1) * D] for i in range(6))
stats = torch.zeros(M, 2, device=dev)

flop_mlp = 2.0 * M * (D * D + 2 * D * Hd)
us = timed(lambda: torch.ops.xdb200.dit_proj_mlp(o, wp, bp, w1, b1, w2, b2, h, h, g1, s2, sc2, g2, T, 1e-6, stats, 1))
print(f"B={B}: fused proj+LN+fc1+GELU+fc2      {us:7.1f} us   {flop_mlp / us / 1e6:7.1f} TFLOP/s")
h2 = torch.empty_like(h)
for split in (2, 3, 4):
    if (M + 255) // 256 * 2 * split <= 148:
        us = timed(lambda: torch.ops.xdb200.dit_proj_mlp(o, wp, bp, w1, b1, w2, b2, h, h2, g1, s2, sc2, g2, T, 1e-6, stats, split))
        print(f"B={B}:   split over {split} CTA pairs per tile  {us:7.1f} us   {flop_mlp / us / 1e6:7.1f} TFLOP/s")


def unfused():
    ops.linear(o, wp, bp, gate=g1, gate_rows=T, residual=h, out=h)
    a = ops.layernorm_modulate(h, s2, sc2, T)
    u = ops.linear(a, w1, b1, act=ops.ACT_GELU)
    ops.linear(u, w2, b2, gate=g2, gate_rows=T, residual=h, out=h)


us = timed(unfused)
print(f"B={B}: proj, LN, fc1, fc2 (4 launches)  {us:7.1f} us   {flop_mlp / us / 1e6:7.1f} TFLOP/s")
if hasattr(torch.ops.xdb200, "dit_attn"):
    flop_attn = 2.0 * M * D * 3 * D + 4.0 * B * 6 * T * T * 64
    us = timed(lambda: torch.ops.xdb200.dit_attn(h, stats, s1, sc1, T, 1e-6, wq, bq, 6, 0.125, o))
    print(f"B={B}: fused LN+qkv+attention           {us:7.1f} us   {flop_attn / us / 1e6:7.1f} TFLOP/s")
    us = timed(lambda: torch.ops.xdb200.dit_attn(h, None, s1, sc1, T, 1e-6, wq, bq, 6, 0.125, o))
    print(f"B={B}: fused LN+qkv+attention, no stats {us:7.1f} us   {flop_attn / us / 1e6:7.1f} TFLOP/s")


def unfused_attn():
    a = ops.layernorm_modulate(h, s1, sc1, T)
    qkv = ops.linear(a, wq, bq).view(B, T, 3, 6, 64)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ops.attention(q, k, v, 0.125)


us = timed(unfused_attn)
print(f"B={B}: LN, qkv, attention (3 launches)  {us:7.1f} us")
