"""Profiling driver: the four DiT-block contraction shapes at batch 1024 (M = 16384) through the C ABI.
Each shape is captured 20x in a CUDA graph and replayed, so the printed time is device time per launch
(host launch overhead excluded).  Also used under `ncu` (see profiles/README.md)."""
import sys

import torch

sys.path.insert(0, ".")
from xdiffusion_b200 import ops  # noqa: E402


def main():
    dev = "cuda"
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    bn0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    only = sys.argv[3].split(',') if len(sys.argv) > 3 else None
    torch.manual_seed(0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for (n, k, name, kw) in [(1152, 384, "qkv", {}), (384, 384, "proj", {"res": True}),
                             (1536, 384, "fc1", {"act": ops.ACT_GELU}), (384, 1536, "fc2", {"res": True}),
                             (1536, 384, "fc1_noact", {})]:
        if only and name not in only:
            continue
        bn = bn0 if not (bn0 >= 2000 and k > 384) else 0       # A-stationary tiles need K <= 384
        a = torch.randn(M, k, device=dev).bfloat16()
        w = torch.randn(n, k, device=dev).bfloat16()
        bias = torch.randn(n, device=dev)
        if kw.get("res"):
            out = torch.randn(M, n, device=dev)
            gate = torch.randn(M // 16, n, device=dev)
            call = lambda: ops.linear(a, w, bias, gate=gate, gate_rows=16, residual=out, out=out, force_bn=bn)
        else:
            out = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
            call = lambda: ops.linear(a, w, bias, act=kw.get("act", 0), out=out, force_bn=bn)
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
        # cold: one launch after an L2 flush
        flush.zero_()
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {us:.1f} us/launch in-graph ({2 * M * n * k / us / 1e6:.0f} TFLOP/s); "
              f"single eager launch incl. host latency {e0.elapsed_time(e1) * 1e3:.1f} us")


if __name__ == "__main__":
    main()
