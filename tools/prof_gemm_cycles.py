"""In-kernel cycle accounting of the tcgen05 GEMM (XDB200_PROF=1): where the producer / MMA / epilogue warps wait.
Usage: XDB200_PROF=1 python tools/prof_gemm_cycles.py [M] [force_bn] [shape,...]   (eager launches, prints to stderr)"""
import os
import sys

os.environ.setdefault("XDB200_PROF", "1")
import torch

sys.path.insert(0, ".")
from xdiffusion_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
dev = "cuda"
for (n, k, name, res, act) in [(1152, 384, "qkv", False, 0), (384, 384, "proj", True, 0), (1536, 384, "fc1", False, 2),
                               (384, 1536, "fc2", True, 0)]:
    if only and name not in only:
        continue
    a = torch.randn(M, k, device=dev).bfloat16()
    w = torch.randn(n, k, device=dev).bfloat16()
    bias = torch.randn(n, device=dev)
    for rep in range(3):
        if rep == 2:
            print(f"===== {name} M={M} N={n} K={k}", file=sys.stderr, flush=True)
        if res:
            out = torch.randn(M, n, device=dev)
            gate = torch.randn(M // 16, n, device=dev)
            ops.linear(a, w, bias, gate=gate, gate_rows=16, residual=out, out=out, force_bn=bn)
        else:
            out = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
            ops.linear(a, w, bias, act=act, out=out, force_bn=bn)
        torch.cuda.synchronize()
