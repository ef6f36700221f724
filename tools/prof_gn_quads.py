"""GroupNorm from producer statistics (xd_groupnorm_apply_quads) against the single-pass cluster kernel, the cost of emitting
the statistics in the conv3x3 epilogue, and how many GroupNorms of a UNet C1 forward take the new path (batch 64, in-graph)."""
import sys

import torch

sys.path.insert(0, ".")
from xdiffusion_b200 import ops  # noqa: E402

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64


def timed(call, reps=20):
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            call()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (hw, c) in [(32, 128), (32, 256), (32, 384), (16, 256), (16, 384), (16, 512)]:
    x = torch.randn(B, hw * hw, c, device=dev).bfloat16()
    gamma, beta = torch.randn(c, device=dev), torch.randn(c, device=dev)
    out = torch.empty_like(x)
    q = torch.randn(B * hw * hw // 32, c // 4, 2, device=dev).abs()
    t0 = timed(lambda: ops.groupnorm(x, gamma, beta, silu=True, out=out))
    t1 = timed(lambda: torch.ops.xdb200.groupnorm_quads(x.view(-1, c), q, gamma, beta, None, 1, 1e-5, 1, B, out.view(-1, c)))
    mb = B * hw * hw * c * 4 / 1e6
    print(f"GN {B}x{hw}x{hw}x{c}: cluster kernel {t0:.1f} us, from quads {t1:.1f} us ({mb / t1:.2f} TB/s)")

for (hw, c, co, cs) in [(32, 128, 128, 0), (32, 256, 128, 256), (16, 256, 256, 0), (16, 512, 256, 512)]:
    x = torch.randn(B, hw, hw, c, device=dev).bfloat16()
    xs = torch.randn(B, hw, hw, cs, device=dev).bfloat16() if cs else None
    K = 9 * c + cs
    wp = (torch.randn(co, K, device=dev) * K ** -0.5).bfloat16()
    bias = torch.randn(co, device=dev)
    out = torch.empty(B, hw, hw, co, device=dev, dtype=torch.bfloat16)
    t0 = timed(lambda: ops.conv3x3(x, wp, bias, xs=xs, out=out))
    with ops.quad_stats():
        t1 = timed(lambda: ops.conv3x3(x, wp, bias, xs=xs, out=out, qstats=True))
    print(f"conv {B}x{hw}x{hw} {c}+{cs}->{co}: {t0:.1f} us, with statistics {t1:.1f} us")

x1 = torch.randn(B, 1, 32, 32, device=dev)
w_in, b_in = torch.randn(128, 1, 3, 3, device=dev), torch.randn(128, device=dev)
o_in = torch.empty(B, 32, 32, 128, device=dev, dtype=torch.bfloat16)
print(f"conv3x3_in  {B}x1x32x32 -> 128: {timed(lambda: torch.ops.xdb200.conv3x3_in(x1, w_in, b_in, o_in)):.1f} us")
for c in (128, 256):
    h = torch.randn(B, 32, 32, c, device=dev).bfloat16()
    w_out = torch.randn(1, c, 3, 3, device=dev)
    o_out = torch.empty(B, 1, 32, 32, device=dev)
    print(f"conv3x3_out {B}x32x32x{c} -> 1: {timed(lambda: torch.ops.xdb200.conv3x3_out(h, w_out, None, o_out)):.1f} us")

# which GroupNorms of a C1 forward take the quad path
sys.path.insert(0, "tests")
from bench import build_model  # noqa: E402
m = build_model("unet", dev)
counts = {}
real_q, real_g = torch.ops.xdb200.groupnorm_quads, torch.ops.xdb200.groupnorm
orig = ops.groupnorm


def spy(x, *a, **k):
    path = "quads" if ops._qs_lookup(x) is not None else "cluster"
    key = (tuple(x.shape), x.stride(1), path)
    counts[key] = counts.get(key, 0) + 1
    return orig(x, *a, **k)


ops.groupnorm = spy
import xdiffusion_b200.layers.resnet as R, xdiffusion_b200.layers.attention as A, xdiffusion_b200.score_networks.unet as U  # noqa
xb = torch.randn(B, 1, 32, 32, device=dev)
with torch.no_grad():
    m.predict_score(xb, context={"timestep": torch.full((B,), 500, device=dev)})
ops.groupnorm = orig
for k, v in sorted(counts.items()):
    print(v, "x", k)
