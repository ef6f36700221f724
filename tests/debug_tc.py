"""Standalone tcgen05 bring-up diagnostics (not a pytest file): prints error patterns so that a wrong
descriptor / swizzle / TMEM mapping can be identified from one GPU run."""
import math
import sys

import torch

sys.path.insert(0, ".")
from xdiffusion_b200 import ops  # noqa: E402


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def main():
    dev = "cuda"
    torch.manual_seed(0)
    # 1. identity weights: out must equal A (reveals row/column permutations)
    for (M, K) in [(128, 64), (128, 128), (256, 256)]:
        a = torch.randn(M, K).bfloat16().to(dev)
        w = torch.eye(K).bfloat16().to(dev)
        out = ops.linear(a, w, out_dtype=torch.float32, force_bn=64 if K == 64 else 128)
        torch.cuda.synchronize()
        err = rel(out, a)
        print(f"identity M={M} K={K}: rel={err:.3e}")
        if err > 1e-3:
            o, af = out.cpu(), a.float().cpu()
            for r in (0, 1, 8, 33, 127):
                src = [int((af[r] - o[r, c]).abs().argmin()) if o[r, c] != 0 else -1 for c in range(min(K, 16))]
                print(f"  row {r}: out cols 0..15 come from A cols {src}")
            rowmatch = [int((af - o[r]).abs().sum(1).argmin()) for r in range(0, 16)]
            print("  out rows 0..15 best-match A rows", rowmatch)
    # 2. random problems, all tile widths
    for (M, N, K) in [(128, 128, 64), (128, 128, 384), (512, 256, 256), (2048, 1152, 384), (1000, 64, 384)]:
        for bn in (64, 128, 256):
            a = torch.randn(M, K).bfloat16().to(dev)
            w = (torch.randn(N, K) / math.sqrt(K)).bfloat16().to(dev)
            out = ops.linear(a, w, out_dtype=torch.float32, force_bn=bn)
            torch.cuda.synchronize()
            print(f"gemm M={M} N={N} K={K} bn={bn}: rel={rel(out, a.float() @ w.float().T):.3e}")
    # 3. timing of the DiT shapes (events, warm)
    for (M, N, K) in [(16384, 1152, 384), (16384, 384, 384), (16384, 1536, 384), (16384, 384, 1536)]:
        a = torch.randn(M, K).bfloat16().to(dev)
        w = (torch.randn(N, K) / math.sqrt(K)).bfloat16().to(dev)
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        for bn in (128, 256):
            for _ in range(3):
                ops.linear(a, w, out=out, force_bn=bn)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(20):
                ops.linear(a, w, out=out, force_bn=bn)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"time M={M} N={N} K={K} bn={bn}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
