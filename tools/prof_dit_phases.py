"""Phase timeline of dit_mlp_kernel (csrc/dit_block.cu) from in-kernel clock64() stamps.  Needs the instrumented build:

    NVCC_EXTRA=-DXDB200_INSTRUMENT bash xdiffusion_b200/csrc/build.sh      (never ship that .so)
    python tools/prof_dit_phases.py [batch] [split]

Prints, for a few CTAs, the cycle offsets of every phase relative to the CTA's own start (1 cycle ~ 0.52 ns at 1.9 GHz).
"""
import os
import sys

import torch

sys.path.insert(0, ".")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
split = int(sys.argv[2]) if len(sys.argv) > 2 else 1
T, D, Hd = 16, 384, 1536
M = B * T
dev = "cuda"
tiles = (M + 255) // 256
G = max(split, 1)
prof = torch.zeros(tiles * 2 * G * 64, dtype=torch.int64, device=dev)
os.environ["XDB200_DIT_PROF"] = str(prof.data_ptr())
from xdiffusion_b200 import ops  # noqa: E402,F401

bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
o, h, h2 = bf(M, D), torch.randn(M, D, device=dev), torch.empty(M, D, device=dev)
wp, w1, w2 = bf(D, D) * D ** -0.5, bf(Hd, D) * D ** -0.5, bf(D, Hd) * Hd ** -0.5
bp, b1, b2 = (torch.randn(n, device=dev) * 0.1 for n in (D, Hd, D))
mod = torch.randn(B, 6 * D, device=dev) * 0.1
s1, sc1, g1, s2, sc2, g2 = (mod[:, i * D:(i + 1) * D] for i in range(6))
stats = torch.zeros(M, 2, device=dev)
for _ in range(3):
    torch.ops.xdb200.dit_proj_mlp(o, wp, bp, w1, b1, w2, b2, h, h2, g1, s2, sc2, g2, T, 1e-6, stats, split)
torch.cuda.synchronize()
p = prof.view(-1, 64).cpu()
names = {63: "entry", 61: "setup done (tmem, barriers, cluster sync)", 62: "pdl_wait returned", 0: "epilogue warp starts",
         48: "mma warp starts", 49: "mma: first O k-block landed", 50: "mma: proj issued", 1: "epi: proj accumulator ready",
         2: "epi: pass 1 (gated residual -> h1) done", 3: "epi: LayerNorm panel written", 51: "mma: panel ready seen",
         52: "mma: everything issued", 40: "epi: fc2 accumulator ready", 41: "epi: final pass done",
         42: "cluster sync #1 passed", 43: "push done", 44: "cluster sync #2 passed", 45: "reduce done", 46: "cluster sync #3 passed",
         60: "exit"}
for c in range(8, 8 + 2 * (Hd // 128), 2):
    names[c] = f"epi: fc1 chunk {(c - 8) // 2} accumulator ready"
    names[c + 1] = f"epi: fc1 chunk {(c - 8) // 2} GELU -> hidden buffer done"
if len(sys.argv) > 3 and sys.argv[3] == "attn":
    prof.zero_()
    wq, bq = bf(3 * D, D) * D ** -0.5, torch.randn(3 * D, device=dev) * 0.1
    for _ in range(3):
        torch.ops.xdb200.dit_attn(h, stats, s1, sc1, T, 1e-6, wq, bq, 6, 0.125, o)
    torch.cuda.synchronize()
    p = prof.view(-1, 64).cpu()
    names = {63: "entry", 61: "setup done", 0: "epilogue warp starts (pdl_wait returned)", 1: "epi: LayerNorm panel written",
             48: "mma warp starts", 51: "mma: panel ready seen", 60: "exit"}
    for i in range(6):
        names[8 + 3 * i] = f"epi: head {i} accumulator ready"
        names[9 + 3 * i] = f"epi: head {i} q|k|v staged (both halves)"
        names[10 + 3 * i] = f"epi: head {i} attention done"
for cta in sorted({0, 1, (tiles * 2 * G) // 2, tiles * 2 * G - 2}):
    row = p[cta]
    t0 = int(row[63])
    ev = sorted((int(row[k]) - t0, names[k]) for k in names if int(row[k]) != 0)
    print(f"--- B={B} split={split} CTA {cta}: {len(ev)} stamps")
    prev = 0
    for t, n in ev:
        print(f"  {t:8d} clk  (+{t - prev:6d})  {n}")
        prev = t
