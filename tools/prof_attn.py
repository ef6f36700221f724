"""Device-time of the attention kernels at the benchmark shapes (CUDA-graph replay, 20 launches)."""
import sys

import torch

sys.path.insert(0, ".")
from xdiffusion_b200 import ops  # noqa: E402


def timed(call):
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
    return e0.elapsed_time(e1) / 20 * 1e3


def main():
    dev = "cuda"
    B, T, H = 1024, 16, 6                                   # DiT
    qkv = torch.randn(B * T, 3 * H * 64, device=dev).bfloat16().view(B, T, 3, H, 64)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    out = torch.empty(B, T, H, 64, device=dev, dtype=torch.bfloat16).permute(0, 2, 1, 3)
    us = timed(lambda: ops.attention(q, k, v, 0.125, out=out))
    print(f"DiT  T=16  B=1024 H=6 : {us:.1f} us  ({(qkv.numel() + out.numel()) * 2 / us / 1e3:.0f} GB/s)")
    B, T, H = 64, 256, 4                                    # UNet 16x16 level
    qkv = torch.randn(B, T, H, 3, 64, device=dev).bfloat16()
    q, k, v = (qkv[:, :, :, i].permute(0, 2, 1, 3) for i in range(3))
    out = torch.empty(B, T, H, 64, device=dev, dtype=torch.bfloat16).permute(0, 2, 1, 3)
    us = timed(lambda: ops.attention(q, k, v, 0.125, out=out))
    flop = 4.0 * B * H * T * T * 64
    print(f"UNet T=256 B=64  H=4 : {us:.1f} us  ({flop / us / 1e6:.1f} TFLOP/s)")


if __name__ == "__main__":
    main()
