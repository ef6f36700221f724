"""Kernel-level parity on the B200: every C-ABI entry point against the oracle / plain fp32 torch math
on the same seeded inputs.  Integer / table / sampler-step work is bit-exact; bf16 tensor-core
contractions are compared on bf16-rounded inputs (exact products, fp32 accumulate) so the only
differences are summation order (1e-5) and the final bf16 store (2^-9 relative per element)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import nets as onets  # noqa: E402
from oracle import samplers as osamplers  # noqa: E402
from oracle import schedules as oschedules  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    from xdiffusion_b200 import ops as _ops
    return _ops


DEV = "cuda"


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def bf(x):
    return x.to(torch.bfloat16)


def _gemm_ref(a, w, bias, act, gate, gate_rows, residual):
    v = a.float() @ w.float().T
    if bias is not None:
        v = v + bias
    if act == 1:
        v = F.silu(v)
    elif act == 2:
        v = F.gelu(v, approximate="tanh")
    if gate is not None:
        v = v * gate.repeat_interleave(gate_rows, dim=0)[: v.shape[0]]
    if residual is not None:
        v = v + residual.float()
    return v


@pytest.mark.parametrize("backend", ["tc", "simt"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 384), (2048, 1152, 384), (1000, 64, 384),
                                    (154, 384, 768), (64, 2304, 384), (300, 200, 128), (4096, 384, 1536)])
def test_gemm_plain(ops, backend, M, N, K, monkeypatch):
    monkeypatch.setattr(ops, "MATMUL_BACKEND", backend)
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = bf(torch.randn(M, K, generator=g)).to(DEV)
    w = bf(torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    out = ops.linear(a, w, bias, out_dtype=torch.float32)
    ref = _gemm_ref(a, w, bias, 0, None, 1, None)
    assert rel_l2(out, ref) < 1e-5
    out = ops.linear(a, w, bias, act=ops.ACT_GELU, out_dtype=torch.bfloat16)
    assert rel_l2(out, _gemm_ref(a, w, bias, 2, None, 1, None)) < 3e-3


@pytest.mark.parametrize("bn", [64, 128, 192, 256, 1128, 1256])      # 1xxx = CTA-pair (cta_group::2) 256 x xxx tiles
def test_gemm_tile_widths_and_epilogues(ops, bn):
    g = torch.Generator().manual_seed(bn)
    M, N, K, rows = 512, 512, 256, 16
    a = bf(torch.randn(M, K, generator=g)).to(DEV)
    a2 = bf(torch.randn(M, 128, generator=g)).to(DEV)
    w = bf(torch.randn(N, K + 128, generator=g) / math.sqrt(K)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    gate_buf = torch.randn(M // rows, 3 * N, generator=g).to(DEV)
    gate = gate_buf[:, N:2 * N]                      # strided view, like an adaLN chunk
    res = torch.randn(M, N, generator=g).to(DEV)
    out = ops.linear(a, w, bias, act=ops.ACT_SILU, out_dtype=torch.float32, gate=gate, gate_rows=rows,
                     residual=res, a2=a2, force_bn=bn)
    ref = _gemm_ref(torch.cat([a, a2], 1), w, bias, 1, gate, rows, res)
    assert rel_l2(out, ref) < 1e-5
    # in-place residual (x += gate * f(x)) with a bf16 residual and bf16 output
    resb = bf(res)
    out = ops.linear(a, w[:, :K].contiguous(), None, out_dtype=torch.bfloat16, residual=resb, force_bn=bn)
    assert rel_l2(out, _gemm_ref(a, w[:, :K], None, 0, None, 1, resb)) < 3e-3


# A-stationary tiles (2xxx = 1-CTA, 3xxx = CTA pair): the 128 x K panel of A stays in shared memory while the CTA
# walks its n-tiles.  Shapes cover one item per CTA, several items per CTA with a ragged last m-tile (panel reuse
# barriers), and n-groups (few m-tiles).
@pytest.mark.parametrize("bn", [2192, 3192, 3256])
@pytest.mark.parametrize("M,N,K,K2", [(2048, 1152, 384, 0), (40000 - 24, 768, 256, 128), (300, 1536, 384, 0),
                                       (640, 384, 128, 0)])
def test_gemm_a_stationary(ops, bn, M, N, K, K2):
    g = torch.Generator().manual_seed(bn + M)
    rows = 4
    a = bf(torch.randn(M, K, generator=g)).to(DEV)
    a2 = bf(torch.randn(M, K2, generator=g)).to(DEV) if K2 else None
    w = bf(torch.randn(N, K + K2, generator=g) / math.sqrt(K + K2)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    full = a if a2 is None else torch.cat([a, a2], 1)
    out = ops.linear(a, w, bias, act=ops.ACT_GELU, out_dtype=torch.bfloat16, a2=a2, force_bn=bn)
    assert rel_l2(out, _gemm_ref(full, w, bias, 2, None, 1, None)) < 3e-3
    gate = torch.randn(M // rows, N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    ref = _gemm_ref(full, w, bias, 0, gate, rows, res)
    out = ops.linear(a, w, bias, gate=gate, gate_rows=rows, residual=res, out=res, a2=a2, force_bn=bn)
    assert out.data_ptr() == res.data_ptr()
    assert rel_l2(out, ref) < 1e-5


# Split-K (few output tiles, long K): partial tiles in the workspace + reduce kernel, against the unsplit result.
@pytest.mark.parametrize("M,N,K", [(256, 256, 4096), (512, 128, 2304), (100, 192, 8192)])
def test_gemm_split_k(ops, M, N, K, monkeypatch):
    g = torch.Generator().manual_seed(M + K)
    a = bf(torch.randn(M, K, generator=g)).to(DEV)
    w = bf(torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = bf(torch.randn(M, N, generator=g)).to(DEV)
    out = ops.linear(a, w, bias, act=ops.ACT_SILU, out_dtype=torch.bfloat16, residual=res)
    ref = _gemm_ref(a, w, bias, 1, None, 1, res)
    assert rel_l2(out, ref) < 3e-3
    gate = torch.randn(M // 4, N, generator=g).to(DEV)
    r32 = torch.randn(M, N, generator=g).to(DEV)
    ref = _gemm_ref(a, w, bias, 0, gate, 4, r32)
    out = ops.linear(a, w, bias, gate=gate, gate_rows=4, residual=r32, out=r32)
    assert rel_l2(out, ref) < 1e-5


# LayerNorm + modulate fused into the A operand (xd_ln_gemm_bf16_tc) against the two-kernel path and fp32 torch.
@pytest.mark.parametrize("M,N,D,rows,act", [(2048, 1152, 384, 16, 0), (300, 1536, 384, 4, 2), (40, 192, 256, 8, 0),
                                             (20000, 384, 384, 16, 2), (4096, 576, 128, 256, 0)])
def test_ln_gemm_fused(ops, M, N, D, rows, act, monkeypatch):
    g = torch.Generator().manual_seed(M + N)
    x = (torch.randn(M, D, generator=g) * 2 + 0.5).to(DEV)
    nb = (M + rows - 1) // rows
    mod = (torch.randn(nb, 3 * D, generator=g) * 0.3).to(DEV)
    shift, scale = mod[:, :D], mod[:, 2 * D:]                     # strided views, like the adaLN chunks
    w = bf(torch.randn(N, D, generator=g) / math.sqrt(D)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    monkeypatch.setattr(ops, "LN_GEMM_FUSED", True)
    fused = ops.ln_linear(x, shift, scale, rows, w, bias, act=act)
    monkeypatch.setattr(ops, "LN_GEMM_FUSED", False)
    split = ops.ln_linear(x, shift, scale, rows, w, bias, act=act)
    assert rel_l2(fused, split) < 1e-6                            # same operand bits, same K order
    ln = F.layer_norm(x, (D,), eps=1e-6)
    a = ln * (1 + scale.repeat_interleave(rows, 0)[:M]) + shift.repeat_interleave(rows, 0)[:M]
    ref = _gemm_ref(bf(a), w, bias, act, None, 1, None)
    assert rel_l2(fused, ref) < 4e-3


def _pack_conv(w, wskip=None):
    """[Cout,C,3,3] -> [Cout, 9*C (+Cs)] tap-major (same packing as the product's weight repack)."""
    co, c = w.shape[:2]
    p = w.permute(0, 2, 3, 1).reshape(co, 9 * c)
    if wskip is not None:
        p = torch.cat([p, wskip.reshape(co, -1)], 1)
    return bf(p).contiguous()


@pytest.mark.parametrize("backend", ["tc", "simt"])
@pytest.mark.parametrize("nimg,H,C,Co,Cs", [(2, 32, 128, 128, 0), (2, 16, 256, 256, 0), (3, 8, 256, 256, 0),
                                             (5, 4, 512, 256, 512), (2, 16, 384, 256, 384), (1, 32, 256, 128, 256),
                                             (9, 4, 256, 256, 0)])
def test_conv3x3(ops, backend, nimg, H, C, Co, Cs, monkeypatch):
    monkeypatch.setattr(ops, "MATMUL_BACKEND", backend)
    g = torch.Generator().manual_seed(nimg * 100 + H + C + Co)
    x = bf(torch.randn(nimg, C, H, H, generator=g))
    w = bf(torch.randn(Co, C, 3, 3, generator=g) / math.sqrt(9 * C))
    bias = torch.randn(Co, generator=g)
    ref = F.conv2d(x.float(), w.float(), bias, padding=1)
    xs = ws = None
    if Cs:
        xs = bf(torch.randn(nimg, Cs, H, H, generator=g))
        ws = bf(torch.randn(Co, Cs, 1, 1, generator=g) / math.sqrt(Cs))
        ref = ref + F.conv2d(xs.float(), ws.float())
    res = None
    if not Cs and C == Co:
        res = bf(torch.randn(nimg, Co, H, H, generator=g))
        ref = ref + res.float()
    nhwc = lambda t: None if t is None else t.permute(0, 2, 3, 1).contiguous().to(DEV)
    out = ops.conv3x3(nhwc(x), _pack_conv(w, ws).to(DEV), bias.to(DEV), residual=nhwc(res), xs=nhwc(xs))
    assert rel_l2(out.permute(0, 3, 1, 2), ref) < 3e-3


def test_conv3x3_strided_concat_slot(ops):
    """input and output living inside wider (concat) buffers"""
    g = torch.Generator().manual_seed(5)
    nimg, H, C, Co = 2, 16, 256, 256
    wide = bf(torch.randn(nimg, H, H, C + 128, generator=g)).to(DEV)
    x = wide[..., 128:]
    w = bf(torch.randn(Co, C, 3, 3, generator=g) / math.sqrt(9 * C))
    outw = torch.zeros(nimg, H, H, Co + 64, dtype=torch.bfloat16, device=DEV)
    ops.conv3x3(x, _pack_conv(w).to(DEV), out=outw[..., :Co])
    ref = F.conv2d(x.permute(0, 3, 1, 2).float().cpu(), w.float(), padding=1)
    assert rel_l2(outw[..., :Co].permute(0, 3, 1, 2), ref) < 3e-3
    assert float(outw[..., Co:].abs().max()) == 0.0


@pytest.mark.parametrize("nimg,H,C,Co,Cs,c0,wide,splitk", [
    (4, 32, 128, 128, 0, 0, 128, False), (2, 32, 128, 128, 128, 128, 384, False), (4, 16, 256, 256, 0, 256, 512, False),
    (8, 16, 128, 256, 128, 0, 256, False), (16, 8, 256, 256, 0, 0, 256, False), (2, 8, 256, 256, 0, 0, 256, True),
    (16, 32, 128, 128, 0, 0, 128, True), (64, 16, 256, 256, 0, 256, 512, True)])
def test_conv3x3_quad_stats_and_groupnorm_from_them(ops, nimg, H, C, Co, Cs, c0, wide, splitk):
    """GroupNorm statistics from the producer (include/xdb200.h: xd_conv3x3_bf16_tc_qstats + xd_groupnorm_apply_quads): the
    conv writes its slice of a concat buffer and emits per-(32 rows, 4 channels) sums; they equal the sums of the stored
    tensor (to the bf16 rounding of the stored values), and the GroupNorm computed from them equals the stand-alone
    GroupNorm and the fp32 torch reference.  Where the launch cannot emit (split-K at few tiles) the helper says so and
    the GroupNorm falls back."""
    g = torch.Generator().manual_seed(nimg + H + C + Co + c0)
    x = bf(torch.randn(nimg, H, H, C, generator=g)).to(DEV)
    w = bf(torch.randn(Co, C, 3, 3, generator=g) / math.sqrt(9 * C))
    xs = ws = None
    if Cs:
        xs = bf(torch.randn(nimg, H, H, Cs, generator=g)).to(DEV)
        ws = bf(torch.randn(Co, Cs, 1, 1, generator=g) / math.sqrt(Cs))
    bias = torch.randn(Co, generator=g).to(DEV)
    buf = bf(torch.randn(nimg, H, H, wide, generator=g)).to(DEV)                 # the other slices: someone else's data
    gamma, beta = (1 + 0.1 * torch.randn(wide, generator=g)).to(DEV), (0.1 * torch.randn(wide, generator=g)).to(DEV)
    ss = (torch.randn(nimg, 2 * wide, generator=g) * 0.3).to(DEV)
    wp = _pack_conv(w, ws).to(DEV)
    ops.set_split_k(splitk)
    try:
        _quad_stats_case(ops, x, wp, bias, xs, buf, c0, Co, wide, gamma, beta, ss, nimg, H, expect=not (splitk and nimg == 2))
    finally:
        ops.set_split_k(True)


def _quad_stats_case(ops, x, wp, bias, xs, buf, c0, Co, wide, gamma, beta, ss, nimg, H, expect):
    with ops.quad_stats():
        out = ops.conv3x3(x, wp, bias, xs=xs, out=buf[..., c0:c0 + Co], qstats=True)
        book = ops._qs_book[buf.untyped_storage().data_ptr()]
        emitted = (c0, c0 + Co) in book["cover"]
        rows = nimg * H * H
        assert emitted == expect, (emitted, rows)        # (2 images of 8 x 8 with split-K on: one tile, long K -> split -> none)
        if emitted:
            q = book["table"][:, c0 // 4:(c0 + Co) // 4].cpu()
            o = out.float().cpu().reshape(rows // 32, 32, Co // 4, 4)
            assert float((q[..., 0] - o.sum((1, 3))).abs().max()) < 0.15                   # 128 values of |x| ~ 1, bf16-rounded
            assert rel_l2(q[..., 1], (o * o).sum((1, 3))) < 3e-3
        # the other slices of the concat buffer come from a producer without statistics: full-width GroupNorm must not use them
        assert ops._qs_lookup(buf.view(nimg, H * H, wide)) is None or (c0 == 0 and Co == wide)
        sl = buf[..., c0:c0 + Co]
        view = sl.as_strided((nimg, H * H, Co), (H * H * wide, wide, 1), sl.storage_offset())
        assert (ops._qs_lookup(view) is not None) == emitted
        n0 = ops.LAUNCHES
        got = ops.groupnorm(view, gamma[:Co], beta[:Co], scale_shift=ss[:, :2 * Co], silu=True)
        assert ops.LAUNCHES - n0 == 1
    alone = ops.groupnorm(view, gamma[:Co], beta[:Co], scale_shift=ss[:, :2 * Co], silu=True)     # outside the scope
    ref = F.group_norm(view.float().cpu().permute(0, 2, 1), 32, gamma[:Co].cpu(), beta[:Co].cpu(), 1e-5)
    ref = F.silu(ref * (1 + ss[:, :Co, None].cpu()) + ss[:, Co:2 * Co, None].cpu())
    assert rel_l2(got.permute(0, 2, 1), ref) < 3e-3, rel_l2(got.permute(0, 2, 1), ref)
    assert rel_l2(got, alone) < 3e-3
    # conv result itself unchanged by the statistics epilogue
    plain = ops.conv3x3(x, wp, bias, xs=xs)
    assert torch.equal(plain, out)


def test_quad_stats_concat_of_two_producers_and_invalidation(ops):
    """A concat buffer whose two slices come from two contractions (a conv and a 1x1 projection with residual): GroupNorm over
    the full width uses the merged statistics (group boundaries straddle the seam: 384 channels = 32 groups of 12); a later
    write into the buffer by an op without statistics drops them."""
    g = torch.Generator().manual_seed(77)
    nimg, H, Ca, Cb = 4, 16, 256, 128
    buf = torch.zeros(nimg, H, H, Ca + Cb, dtype=torch.bfloat16, device=DEV)
    x = bf(torch.randn(nimg, H, H, Ca, generator=g)).to(DEV)
    w = _pack_conv(bf(torch.randn(Ca, Ca, 3, 3, generator=g) / math.sqrt(9 * Ca))).to(DEV)
    a = bf(torch.randn(nimg * H * H, 128, generator=g)).to(DEV)
    wl = bf(torch.randn(Cb, 128, generator=g) / math.sqrt(128)).to(DEV)
    res = bf(torch.randn(nimg * H * H, Cb, generator=g)).to(DEV)
    gamma, beta = (1 + 0.1 * torch.randn(Ca + Cb, generator=g)).to(DEV), (0.1 * torch.randn(Ca + Cb, generator=g)).to(DEV)
    full = buf.view(nimg, H * H, Ca + Cb)
    ops.set_split_k(False)                      # (4 images: the conv would otherwise be split over K and emit nothing)
    try:
        _two_producers_case(ops, buf, x, w, a, wl, res, gamma, beta, full, nimg, H, Ca, Cb, g)
    finally:
        ops.set_split_k(True)


def _two_producers_case(ops, buf, x, w, a, wl, res, gamma, beta, full, nimg, H, Ca, Cb, g):
    with ops.quad_stats():
        ops.conv3x3(x, w, out=buf[..., :Ca], qstats=True)
        assert ops._qs_lookup(full) is None                                        # second slice not produced yet
        right = buf[..., Ca:]
        ops.linear(a, wl, residual=res, out=right.as_strided((nimg * H * H, Cb), (Ca + Cb, 1), right.storage_offset()), qstats=True)
        assert ops._qs_lookup(full) is not None
        got = ops.groupnorm(full, gamma, beta, silu=True)
        torch.ops.xdb200.avgpool2x2(bf(torch.randn(nimg, 2 * H, 2 * H, Cb, generator=g)).to(DEV), right)    # foreign writer
        assert ops._qs_lookup(full) is None and ops._qs_lookup(buf[..., :Ca].as_strided((nimg, H * H, Ca), (H * H * (Ca + Cb), Ca + Cb, 1))) is not None
    with ops.quad_stats():                                                          # (buf was overwritten above: fresh pass)
        ops.conv3x3(x, w, out=buf[..., :Ca], qstats=True)
        ops.linear(a, wl, residual=res, out=right.as_strided((nimg * H * H, Cb), (Ca + Cb, 1), right.storage_offset()), qstats=True)
        n0 = ops.LAUNCHES
        got = ops.groupnorm(full, gamma, beta, silu=True)
        assert ops.LAUNCHES - n0 == 1
    ref = F.silu(F.group_norm(full.float().cpu().permute(0, 2, 1), 32, gamma.cpu(), beta.cpu(), 1e-5))
    assert rel_l2(got.permute(0, 2, 1), ref) < 3e-3, rel_l2(got.permute(0, 2, 1), ref)


@pytest.mark.parametrize("nimg,H,C,Co,samples,route", [(64, 8, 256, 256, 64, "split"), (64, 4, 256, 256, 64, "split"),
                                                       (16, 8, 512, 256, 16, "split"), (32, 8, 128, 128, 2, "split"),
                                                       (64, 16, 256, 256, 64, "quads"), (16, 32, 128, 128, 16, "quads"),
                                                       (2, 8, 256, 256, 2, "any"), (64, 4, 256, 512, 64, "any")])
def test_conv3x3_groupnorm_all_routes(ops, nimg, H, C, Co, samples, route):
    """xd_conv3x3_groupnorm_bf16_tc: GroupNorm32(conv3x3(x) + bias) * (1 + scale) + shift -> SiLU against fp32 torch, on the
    route the shape selects (split-K reduce that normalises / epilogue statistics / conv then GroupNorm); `samples` < nimg =
    several images share statistics (the frames of a clip)."""
    g = torch.Generator().manual_seed(nimg + H + C + Co)
    x = bf(torch.randn(nimg, C, H, H, generator=g))
    w = bf(torch.randn(Co, C, 3, 3, generator=g) / math.sqrt(9 * C))
    bias = torch.randn(Co, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(Co, generator=g), 0.1 * torch.randn(Co, generator=g)
    ss = torch.randn(samples, 2 * Co, generator=g) * 0.3
    conv = F.conv2d(x.float(), w.float(), bias, padding=1)                                 # [nimg, Co, H, H]
    per = nimg // samples
    cs = conv.view(samples, per, Co, H * H).permute(0, 2, 1, 3).reshape(samples, Co, per * H * H)
    ref = F.group_norm(cs, 32, gamma, beta, 1e-5) * (1 + ss[:, :Co, None]) + ss[:, Co:, None]
    ref = F.silu(ref).view(samples, Co, per, H, H).permute(0, 2, 1, 3, 4).reshape(nimg, Co, H, H)
    n0 = ops.LAUNCHES
    out = ops.conv3x3_groupnorm(x.permute(0, 2, 3, 1).contiguous().to(DEV), _pack_conv(w).to(DEV), bias.to(DEV), samples,
                                gamma.to(DEV), beta.to(DEV), scale_shift=ss.to(DEV))
    assert ops.LAUNCHES - n0 == 2
    assert rel_l2(out.permute(0, 3, 1, 2), ref) < 4e-3, (route, rel_l2(out.permute(0, 3, 1, 2), ref))


def test_conv_in_out(ops):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 1, 32, 32, generator=g)
    w = torch.randn(128, 1, 3, 3, generator=g)
    out = torch.empty(3, 32, 32, 128, dtype=torch.bfloat16, device=DEV)
    torch.ops.xdb200.conv3x3_in(x.to(DEV), w.to(DEV), None, out)
    assert rel_l2(out.permute(0, 3, 1, 2), F.conv2d(x, w, padding=1)) < 3e-3
    h = bf(torch.randn(3, 32, 32, 128, generator=g))
    w2 = torch.randn(1, 128, 3, 3, generator=g) / 30
    o2 = torch.empty(3, 1, 32, 32, device=DEV)
    torch.ops.xdb200.conv3x3_out(h.to(DEV), w2.to(DEV), None, o2)
    assert rel_l2(o2, F.conv2d(h.float().permute(0, 3, 1, 2), w2, padding=1)) < 1e-5


@pytest.mark.parametrize("ns,P,C", [(2, 1024, 128), (3, 256, 384), (2, 64, 512), (4, 16, 256), (1, 16384, 128)])
def test_groupnorm(ops, ns, P, C):
    g = torch.Generator().manual_seed(P + C)
    x = bf(torch.randn(ns, P, C, generator=g) * 2 + 0.5)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    ss = torch.randn(ns, 2 * C, generator=g) * 0.3
    xr = x.float().permute(0, 2, 1)
    ref = F.group_norm(xr, 32, gamma, beta, 1e-5)
    out = ops.groupnorm(x.to(DEV), gamma.to(DEV), beta.to(DEV), silu=True)
    assert rel_l2(out.permute(0, 2, 1), F.silu(ref)) < 3e-3
    ref2 = F.silu(ref * (1 + ss[:, :C, None]) + ss[:, C:, None])
    out = ops.groupnorm(x.to(DEV), gamma.to(DEV), beta.to(DEV), scale_shift=ss.to(DEV), silu=True)
    assert rel_l2(out.permute(0, 2, 1), ref2) < 3e-3


@pytest.mark.parametrize("B,F,HW,C", [(2, 16, 64, 128), (3, 16, 16, 256), (1, 8, 256, 128)])
def test_groupnorm_over_frames_split_output(ops, B, F, HW, C):
    """xd_groupnorm_frames_split (one warp per (clip, pixel) sample) against torch GroupNorm on the reference's "(b h w) c f"
    view; hi + lo reproduces the fp32 result to 2^-16, and the generic two-kernel path agrees."""
    from xdiffusion_b200 import _lib
    g = torch.Generator().manual_seed(B + F + HW + C)
    x = bf(torch.randn(B * F * HW, C, generator=g) * 1.5 + 0.3)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    xs = x.float().view(B, F, HW, C).permute(0, 2, 3, 1).reshape(B * HW, C, F)
    ref = F_gn(xs, gamma, beta).view(B, HW, C, F).permute(0, 3, 1, 2).reshape(B * F * HW, C)
    out = torch.empty(B * F * HW, 2 * C, dtype=torch.bfloat16, device=DEV)
    assert torch.ops.xdb200.groupnorm_frames_split(x.to(DEV), gamma.to(DEV), beta.to(DEV), 1e-5, B, F, HW, out) == 1
    got = out[:, :C].float() + out[:, C:].float()
    assert rel_l2(got, ref) < 2e-5, rel_l2(got, ref)
    two = torch.empty_like(out)
    stats = torch.empty(B * HW * 64 * _lib.lib().xd_groupnorm_slabs(B * HW, F, C), device=DEV)
    torch.ops.xdb200.groupnorm(x.to(DEV), gamma.to(DEV), beta.to(DEV), None, 1, 1e-5, 0, HW, B * HW, 1, stats, two)
    assert rel_l2(got, two[:, :C].float() + two[:, C:].float()) < 2e-5


def F_gn(x, gamma, beta):
    return F.group_norm(x, 32, gamma, beta, 1e-5)


def test_layernorm_modulate(ops):
    g = torch.Generator().manual_seed(3)
    B, T, D = 5, 16, 384
    x = torch.randn(B * T, D, generator=g) * 3 + 1
    mod = torch.randn(B, 6 * D, generator=g)
    shift, scale = mod[:, :D], mod[:, D:2 * D]
    ref = F.layer_norm(x, (D,), eps=1e-6).view(B, T, D) * (1 + scale[:, None]) + shift[:, None]
    md = mod.to(DEV)
    out = ops.layernorm_modulate(x.to(DEV), md[:, :D], md[:, D:2 * D], T)
    assert rel_l2(out, ref.view(B * T, D)) < 3e-3


def test_attention_layouts(ops):
    g = torch.Generator().manual_seed(9)
    # DiT: [Q|K|V]-major, T=16, 6 heads  (oracle.nets.mhsa core)
    B, T, H = 5, 16, 6
    qkv = bf(torch.randn(B * T, 3 * H * 64, generator=g))
    v5 = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = ((v5[0] * 64 ** -0.5) @ v5[1].transpose(-2, -1)).softmax(-1) @ v5[2]          # B,H,T,64
    d = qkv.to(DEV).view(B, T, 3, H, 64)
    q, k, v = (d[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    out = ops.attention(q, k, v, 64 ** -0.5)
    assert rel_l2(out, ref) < 4e-3
    # UNet: per-head interleaved, T=256, 4 heads  (oracle.nets.qkv_attention_interleaved)
    B, T, H = 3, 256, 4
    qkv = bf(torch.randn(B, T, 3 * H * 64, generator=g))
    ref = onets.qkv_attention_interleaved(qkv.float().permute(0, 2, 1), H)             # B, H*64, T
    d = qkv.to(DEV).view(B, T, H, 3, 64)
    q, k, v = (d[:, :, :, i].permute(0, 2, 1, 3) for i in range(3))
    out = ops.attention(q, k, v, 1 / 8.0)                                               # B,H,T,64 view of [B,T,H*64]
    assert rel_l2(out.permute(0, 1, 3, 2).reshape(B, H * 64, T), ref) < 4e-3
    # PixArt cross attention: 16 queries x 77 keys
    B, H = 4, 6
    qq = bf(torch.randn(B, 16, H, 64, generator=g))
    kv = bf(torch.randn(B, 77, 2, H, 64, generator=g))
    ref = ((qq.float().permute(0, 2, 1, 3) @ kv[:, :, 0].float().permute(0, 2, 3, 1)) * 64 ** -0.5).softmax(-1) \
        @ kv[:, :, 1].float().permute(0, 2, 1, 3)
    kd = kv.to(DEV)
    out = ops.attention(qq.to(DEV).permute(0, 2, 1, 3), kd[:, :, 0].permute(0, 2, 1, 3),
                        kd[:, :, 1].permute(0, 2, 1, 3), 64 ** -0.5)
    assert rel_l2(out, ref) < 4e-3


@pytest.mark.parametrize("Tk", [8, 24, 77, 100, 128])
def test_cross_attention_short_context(ops, Tk):
    """16 queries x Tk keys on the warp-level mma.sync kernel (key blocks of 16, keys >= Tk masked), odd (batch, head) count."""
    g = torch.Generator().manual_seed(Tk)
    B, H = 3, 5
    qq = bf(torch.randn(B, 16, H, 64, generator=g))
    kv = bf(torch.randn(B, Tk, 2, H, 64, generator=g))
    ref = ((qq.float().permute(0, 2, 1, 3) @ kv[:, :, 0].float().permute(0, 2, 3, 1)) * 0.125).softmax(-1) \
        @ kv[:, :, 1].float().permute(0, 2, 1, 3)
    kd = kv.to(DEV)
    out = ops.attention(qq.to(DEV).permute(0, 2, 1, 3), kd[:, :, 0].permute(0, 2, 1, 3), kd[:, :, 1].permute(0, 2, 1, 3), 0.125)
    assert rel_l2(out, ref) < 4e-3


@pytest.mark.parametrize("Tq,Tk,B,H", [(64, 64, 7, 4), (4, 81, 3, 4), (16, 93, 5, 4), (48, 48, 2, 2), (4, 4, 9, 4)])
def test_attention_query_blocks_on_tensor_cores(ops, Tq, Tk, B, H):
    """The warp-level mma.sync kernel with 16-row query blocks: the 8x8 level of the video UNet (64 x 64), the
    text-conditioned UNet (4 or 16 pixels against 77 text + own keys) and ragged sizes, per-head interleaved
    layout, against fp32 softmax attention on the same bf16 operands."""
    g = torch.Generator().manual_seed(1000 + Tq + Tk)
    q = bf(torch.randn(B, Tq, H, 64, generator=g))
    kv = bf(torch.randn(B, Tk, H, 2, 64, generator=g))
    ref = ((q.float().permute(0, 2, 1, 3) @ kv[:, :, :, 0].float().permute(0, 2, 3, 1)) / 8.0).softmax(-1) \
        @ kv[:, :, :, 1].float().permute(0, 2, 1, 3)                                      # B, H, Tq, 64
    kd = kv.to(DEV)
    out = ops.attention(q.to(DEV).permute(0, 2, 1, 3), kd[:, :, :, 0].permute(0, 2, 1, 3), kd[:, :, :, 1].permute(0, 2, 1, 3),
                        1 / 8.0)
    assert rel_l2(out, ref) < 4e-3, rel_l2(out, ref)


@pytest.mark.parametrize("T,B", [(256, 3), (64, 5), (100, 2)])
def test_attention_wide_single_head(ops, T, B):
    """Head dim 256, one head (the attention of the EDM DDPM++ network, layers/edm.py UNetBlock): the mma.sync flash kernel
    against fp32 softmax attention on the same bf16 operands, [Q | K | V]-major rows like the network produces them."""
    g = torch.Generator().manual_seed(500 + T)
    C = 256
    qkv = bf(torch.randn(B, T, 3 * C, generator=g))
    q, k, v = (qkv[:, :, i * C:(i + 1) * C].float() for i in range(3))
    ref = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(C), -1) @ v                  # B, T, C
    d = qkv.to(DEV).view(B, T, 3, 1, C)
    qd, kd, vd = (d[:, :, i].permute(0, 2, 1, 3) for i in range(3))                    # [B, 1, T, C] views
    out = ops.attention(qd, kd, vd, 1 / math.sqrt(C))
    assert rel_l2(out[:, 0], ref) < 4e-3, rel_l2(out[:, 0], ref)


def test_attention_relative_position_scrambled(ops):
    """TemporalSelfAttention core incl. the reference's raw reshape (oracle.nets.relpos_attention)."""
    g = torch.Generator().manual_seed(21)
    Bp, H, L = 32, 4, 16
    qkv = bf(torch.randn(Bp, L, 3 * H * 64, generator=g) * 0.3)
    ek = torch.randn(H, 2 * L - 1, 64, generator=g) / 8
    ref = onets.relpos_attention(qkv.float().permute(0, 2, 1), H, ek)                   # (Bp, H*64, L)
    d = qkv.to(DEV).view(Bp, L, H, 3, 64)
    q, k, v = (d[:, :, :, i].permute(0, 2, 1, 3) for i in range(3))
    out = torch.empty(Bp, L, H * 64, dtype=torch.bfloat16, device=DEV)                 # rows = positions
    ops.attention(q, k, v, 1.0, out=out.view(Bp, L, H, 64).permute(0, 2, 1, 3), relk=ek.to(DEV), scramble=True,
                  o_cs=1)
    assert rel_l2(out.permute(0, 2, 1), ref) < 4e-3
    # fp32 q/k/v (the split-precision temporal path, mma.sync with bf16 hi + lo operands): only the bf16 output rounding
    # remains -- also at unit gain, where the unscaled logits reach |30| and the softmax is nearly one-hot (bf16 operands
    # would be off by percents there)
    for gain in (0.3, 1.0):
        g32 = torch.randn(Bp, L, 3 * H * 64, generator=g) * gain
        ref = onets.relpos_attention(g32.double().permute(0, 2, 1), H, ek.double()).float()
        d = g32.to(DEV).view(Bp, L, H, 3, 64)
        q, k, v = (d[:, :, :, i].permute(0, 2, 1, 3) for i in range(3))
        ops.attention(q, k, v, 1.0, out=out.view(Bp, L, H, 64).permute(0, 2, 1, 3), relk=ek.to(DEV), scramble=True,
                      o_cs=1)
        assert rel_l2(out.permute(0, 2, 1), ref) < 3e-3, (gain, rel_l2(out.permute(0, 2, 1), ref))
    # grouped heads: all clips in ONE launch (rows ordered (clip, frame, pixel) as in the video UNet) must equal the
    # per-pixel-batch launch verified above, bit for bit
    Bc, HW, C = 2, 16, H * 64
    x = (torch.randn(Bc, L, HW, 3 * C, generator=g) * 0.3).to(DEV)
    pix = x.permute(0, 2, 1, 3).reshape(Bc * HW, L, H, 3, 64)                          # (clip, pixel) batch
    q, k, v = (pix[:, :, :, i].permute(0, 2, 1, 3) for i in range(3))
    want = torch.empty(Bc * HW, L, C, dtype=torch.bfloat16, device=DEV)
    ops.attention(q, k, v, 1.0, out=want.view(Bc * HW, L, H, 64).permute(0, 2, 1, 3), relk=ek.to(DEV), scramble=True,
                  o_cs=1)
    C3 = 3 * C
    q, k, v = (x.as_strided((Bc, HW * H, L, 64), (L * HW * C3, 192, HW * C3, 1), i * 64) for i in range(3))
    got = torch.empty(Bc, L, HW, C, dtype=torch.bfloat16, device=DEV)
    ops.attention(q, k, v, 1.0, out=got.as_strided((Bc, HW * H, L, 64), (L * HW * C, 64, HW * C, 1)), relk=ek.to(DEV),
                  scramble=True, o_cs=1, hpg=H, o_gs=C)
    assert torch.equal(got.permute(0, 2, 1, 3).reshape(Bc * HW, L, C), want)


def test_split_precision_gemm(ops):
    """x.W via bf16 hi/lo pairs on the tensor-core kernel: ~fp32 accuracy (used by temporal attention)."""
    g = torch.Generator().manual_seed(33)
    M, N, K = 512, 768, 256
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    xh = bf(x); xl = bf(x - xh.float())
    wh = bf(w); wl = bf(w - wh.float())
    n = torch.cat([xh, xl], 1).contiguous().to(DEV)
    w3 = torch.cat([wh, wh, wl], 1).contiguous().to(DEV)
    out = ops.linear(n, w3, out_dtype=torch.float32, a2=n[:, :K])
    ref = x @ w.T
    assert rel_l2(out, ref) < 2e-5
    plain = ops.linear(xh.to(DEV), wh.to(DEV), out_dtype=torch.float32)
    assert rel_l2(plain, ref) > 1e-3          # the single-bf16 product is ~100x less accurate


def test_timestep_embeddings(ops, golden):
    kat = golden("kat")
    half = 64
    freq = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    t = torch.tensor([0, 1, 500, 999])
    out = torch.empty(4, 128, device=DEV)
    torch.ops.xdb200.timestep_embed(t.to(DEV), freq.to(DEV), 1, 1000.0, 0.0, 0.0, 0, out)
    assert (out.cpu() - kat["sin_unet_1000"]).abs().max() < 2e-6
    tf = torch.tensor([0.001, 0.5, 0.999])
    out = torch.empty(3, 128, device=DEV)
    torch.ops.xdb200.timestep_embed(tf.to(DEV), freq.to(DEV), 1, 1.0, 0.0, 0.0, 0, out)
    assert (out.cpu() - kat["sin_unet_1"]).abs().max() < 2e-6
    fd = torch.exp(-math.log(10000) * torch.arange(0, 128, dtype=torch.float32) / 128)
    out = torch.empty(4, 256, device=DEV)
    torch.ops.xdb200.timestep_embed(t.to(DEV), fd.to(DEV), 0, 1.0, 0.0, 0.0, 1, out)
    assert (out.cpu() - kat["sin_dit"]).abs().max() < 2e-6


def test_patchify_unpatchify_pool(ops):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 1, 32, 32, generator=g)
    w = torch.randn(384, 1, 8, 8, generator=g)
    out = torch.empty(3 * 16, 64, dtype=torch.bfloat16, device=DEV)
    torch.ops.xdb200.patchify(x.to(DEV), 8, out)
    ref = F.conv2d(bf(x).float(), w, stride=8).flatten(2).transpose(1, 2).reshape(48, 384)
    assert rel_l2(out.float().cpu() @ w.view(384, 64).T, ref) < 1e-5
    y = torch.randn(48, 64, generator=g)
    img = torch.empty(3, 1, 32, 32, device=DEV)
    torch.ops.xdb200.unpatchify(y.to(DEV), 8, img)
    assert torch.equal(img.cpu(), onets.unpatchify(y.view(3, 16, 64), 8, 1))
    h = bf(torch.randn(2, 16, 16, 128, generator=g)).to(DEV)
    o = torch.empty(2, 8, 8, 128, dtype=torch.bfloat16, device=DEV)
    torch.ops.xdb200.avgpool2x2(h, o)
    assert rel_l2(o.permute(0, 3, 1, 2), F.avg_pool2d(h.float().permute(0, 3, 1, 2), 2)) < 3e-3
    u = torch.empty(2, 32, 32, 128, dtype=torch.bfloat16, device=DEV)
    torch.ops.xdb200.upsample2x(h, u)
    assert torch.equal(u.permute(0, 3, 1, 2).float(), F.interpolate(h.float().permute(0, 3, 1, 2), scale_factor=2))


def _coefs_discrete(tables, logvar, pred):
    T = tables["betas"].shape[0]
    c = torch.zeros(T, 8)
    if pred == "epsilon":
        c[:, 0], c[:, 1] = tables["sqrt_recip_alphas_cumprod"], tables["sqrt_recipm1_alphas_cumprod"]
    else:
        c[:, 0], c[:, 1] = tables["sqrt_alphas_cumprod"], tables["sqrt_one_minus_alphas_cumprod"]
    c[:, 2], c[:, 3] = tables["posterior_mean_coef1"], tables["posterior_mean_coef2"]
    c[:, 4] = torch.exp(0.5 * logvar)
    return c


@pytest.mark.parametrize("pred", ["epsilon", "v"])
@pytest.mark.parametrize("threshold", [False, True])
def test_sampler_step_discrete_bit_exact(ops, pred, threshold):
    tables = oschedules.discrete_tables(1000, "linear")
    logvar = oschedules.fixed_large_logvar(tables)
    coefs = _coefs_discrete(tables, logvar, pred).to(DEV)
    g = torch.Generator().manual_seed(17)
    B = 6
    for i in (999, 500, 37, 1, 0):
        x, o, z = (torch.randn(B, 1, 32, 32, generator=g) for _ in range(3))
        x = x * torch.linspace(0.2, 3.0, B)[:, None, None, None]
        ref = osamplers.ancestral_discrete(x, o, z, i, tables, logvar, pred, (0.99, 1.7) if threshold else None)
        ranks = torch.tensor(0.99, dtype=torch.float32) * (1024 - 1)
        out = torch.empty_like(x, device=DEV)
        torch.ops.xdb200.sampler_step(0, 0, 0, x.to(DEV), o.to(DEV), z.to(DEV), 0, out, coefs, None, i,
                                      int(threshold), int(ranks.floor()), float(ranks - ranks.floor()), 1.7, 0, None, 0)
        assert torch.equal(out.cpu(), ref), (i, float((out.cpu() - ref).abs().max()))
        # device-resident loop index (the CUDA-graph path)
        idx = torch.tensor([i], dtype=torch.int32, device=DEV)
        out2 = torch.empty_like(out)
        torch.ops.xdb200.sampler_step(0, 0, 0, x.to(DEV), o.to(DEV), z.to(DEV), 0, out2, coefs, idx, -1,
                                      int(threshold), int(ranks.floor()), float(ranks - ranks.floor()), 1.7, 0, None, 0)
        assert torch.equal(out2, out)


def test_sampler_step_euler_and_philox(ops):
    g = torch.Generator().manual_seed(4)
    x, o = torch.randn(4, 1, 32, 32, generator=g), torch.randn(4, 1, 32, 32, generator=g)
    coefs = torch.zeros(1000, 8)
    coefs[:, 0] = 1.0 / 1000
    out = torch.empty_like(x, device=DEV)
    torch.ops.xdb200.sampler_step(2, 0, 0, x.to(DEV), o.to(DEV), None, 0, out, coefs.to(DEV), None, 5, 0, 0, 0.0,
                                  0.0, 0, None, 0)
    assert torch.equal(out.cpu(), osamplers.euler_flow(x, o, 1000))
    # in-kernel noise: x0 == 0 => out = sigma * z ~ N(0, sigma^2), different per step and per seed
    n = 1 << 20
    zeros = torch.zeros(1, n, device=DEV)
    c = torch.zeros(4, 8, device=DEV)
    c[:, 4] = 1.0
    outs = []
    for step, seed in ((1, 1), (2, 1), (1, 2)):
        o_ = torch.empty_like(zeros)
        torch.ops.xdb200.sampler_step(0, 0, 0, zeros, zeros, None, 0, o_, c, None, step, 0, 0, 0.0, 0.0, seed, None, 0)
        outs.append(o_.cpu())
    for o_ in outs:
        assert abs(float(o_.mean())) < 5e-3 and abs(float(o_.std()) - 1) < 5e-3
        assert abs(float((o_ ** 4).mean()) - 3) < 0.05
    assert abs(float((outs[0] * outs[1]).mean())) < 5e-3 and abs(float((outs[0] * outs[2]).mean())) < 5e-3


def test_philox_offset_makes_shards_reproduce_the_unsharded_noise(ops):
    """The Philox counter is the GLOBAL element index: rows [lo, hi) computed with elem_offset = lo * n_per_sample draw
    the noise those rows get in the unsharded launch (xdiffusion_b200/dist.py), for both step kernels."""
    B, n = 8, 1024
    zeros = torch.zeros(B, 1, 32, 32, device=DEV)
    c = torch.zeros(4, 8, device=DEV)
    c[:, 4] = 1.0
    for threshold in (0, 1):
        full = torch.empty_like(zeros)
        torch.ops.xdb200.sampler_step(0, 0, 0, zeros, zeros, None, 0, full, c, None, 2, threshold, 1012, 0.77, 1.7, 5,
                                      None, 0)
        assert len(torch.unique(full.view(B, -1)[:, :16], dim=0)) == B          # rows differ
        for lo, hi in ((0, 3), (3, 8)):
            part = torch.empty_like(zeros[lo:hi])
            torch.ops.xdb200.sampler_step(0, 0, 0, zeros[lo:hi], zeros[lo:hi], None, 0, part, c, None, 2, threshold,
                                          1012, 0.77, 1.7, 5, None, lo * n)
            assert torch.equal(part, full[lo:hi])


@pytest.mark.parametrize("pred", ["v", "epsilon"])
@pytest.mark.parametrize("sampler", ["ancestral", "ddim"])
@pytest.mark.parametrize("N", [1024, 1000, 50])
def test_sampler_step_continuous_bit_exact(ops, pred, sampler, N):
    """MODE_DDIM, x0 form 1 (continuous epsilon) and the continuous ancestral rows: host-built coefficient rows
    (xdiffusion_b200/scheduler.py) + the fused kernel against the oracle's restatement of scheduler.py:414-494,524-544
    and samplers/ddim.py:43-123 -- bit for bit, at the first / middle / last loop indices incl. i = 0."""
    from xdiffusion_b200.scheduler import ContinuousNoiseScheduler
    sched = ContinuousNoiseScheduler(1024, "cosine", "l2", -20, 20)
    gammas = oschedules.cosine_logsnr_table(1024, -20, 20)
    assert torch.equal(sched.gammas, gammas)
    coefs, form = sched.step_coefficients(pred, N, sampler)
    assert form == (0 if pred == "v" else 1)
    coefs = coefs.to(DEV)
    g = torch.Generator().manual_seed(23 + N)
    B = 4
    mode = 0 if sampler == "ancestral" else 1
    for i in (N - 1, N // 2, 7, 1, 0):
        x, o, z = (torch.randn(B, 1, 32, 32, generator=g) for _ in range(3))
        x = x * torch.linspace(0.3, 2.5, B)[:, None, None, None]
        idx_s, idx_t = oschedules.continuous_indices(i, N, 1024)
        lam_s, lam_t = gammas[idx_s], gammas[idx_t]
        if sampler == "ancestral":
            ref = osamplers.ancestral_continuous(x, o, z, i, lam_s, lam_t, pred, None)
        else:
            ref = osamplers.ddim_continuous(x, o, i, lam_s, lam_t, pred, None)
        out = torch.empty_like(x, device=DEV)
        torch.ops.xdb200.sampler_step(mode, form, int(pred == "v"), x.to(DEV), o.to(DEV), z.to(DEV), 0, out, coefs,
                                      None, i, 0, 0, 0.0, 0.0, 0, None, 0)
        assert torch.equal(out.cpu(), ref), (sampler, pred, N, i, float((out.cpu() - ref).abs().max()))


def _ln(x, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), eps=eps)


@pytest.mark.parametrize("B,split", [(64, None), (13, None), (1024, None), (13, 2), (13, 3), (13, 4), (128, 2), (128, 3),
                                     (128, 4), (128, 0), (256, 0), (512, 0), (64, 1)])
def test_dit_proj_mlp_fused(ops, B, split):
    """xd_dit_proj_mlp_bf16_tc (proj + gated residual + LayerNorm-modulate + fc1 + GELU + fc2 + gated residual in one
    kernel) against fp32 torch math on the same bf16-rounded operands (score_networks/dit.py:46-59); B = 13 leaves a
    ragged last 256-row tile.  The emitted (mean, rstd) row statistics are checked against torch too.
    split = None: in place, one CTA pair per tile.  split = 2 / 3 / 4: separate output buffer, the hidden units of a tile
    spread over that many CTA pairs of a cluster and reduced through distributed shared memory (0 = the library chooses);
    the split kernels must also be deterministic (two launches, identical bits)."""
    g = torch.Generator().manual_seed(100 + B)
    T, D, Hd = 16, 384, 1536
    M = B * T
    o = bf(torch.randn(M, D, generator=g))
    h = torch.randn(M, D, generator=g) * 1.5 + 0.3
    wp = bf(torch.randn(D, D, generator=g) / math.sqrt(D))
    w1 = bf(torch.randn(Hd, D, generator=g) / math.sqrt(D))
    w2 = bf(torch.randn(D, Hd, generator=g) / math.sqrt(Hd))
    bp, b1, b2 = (torch.randn(n, generator=g) * 0.1 for n in (D, Hd, D))
    mod = torch.randn(B, 6 * D, generator=g) * 0.5
    g1, s2, sc2, g2 = (mod[:, i * D:(i + 1) * D] for i in (2, 3, 4, 5))
    rep = lambda t: t.repeat_interleave(T, dim=0)
    h1 = h + rep(g1) * (o.float() @ wp.float().T + bp)
    a = bf(_ln(h1) * (1 + rep(sc2)) + rep(s2)).float()
    u = bf(F.gelu(a @ w1.float().T + b1, approximate="tanh")).float()
    ref = h1 + rep(g2) * (u @ w2.float().T + b2)
    hd = h.to(DEV)
    md = mod.to(DEV)
    g1d, s2d, sc2d, g2d = (md[:, i * D:(i + 1) * D] for i in (2, 3, 4, 5))
    stats = torch.zeros(M, 2, device=DEV)
    dev_args = [t.to(DEV) for t in (o, wp, bp, w1, b1, w2, b2)]
    if split is None:
        torch.ops.xdb200.dit_proj_mlp(*dev_args, hd, hd, g1d, s2d, sc2d, g2d, T, 1e-6, stats, 0)
    else:
        h_in, hd = hd, torch.full_like(hd, float("nan"))
        torch.ops.xdb200.dit_proj_mlp(*dev_args, h_in, hd, g1d, s2d, sc2d, g2d, T, 1e-6, stats, split)
        again, stats2 = torch.full_like(hd, float("nan")), torch.zeros_like(stats)
        torch.ops.xdb200.dit_proj_mlp(*dev_args, h_in, again, g1d, s2d, sc2d, g2d, T, 1e-6, stats2, split)
        assert torch.equal(again, hd) and torch.equal(stats, stats2)
        assert torch.equal(h_in.cpu(), h)                        # the input buffer is left untouched
    assert rel_l2(hd, ref) < 3e-3, rel_l2(hd, ref)
    mean, var = ref.mean(1), ref.var(1, unbiased=False)
    assert float((stats[:, 0].cpu() - mean).abs().max()) < 2e-3
    assert rel_l2(stats[:, 1], torch.rsqrt(var + 1e-6)) < 2e-3


@pytest.mark.parametrize("B", [64, 13, 128, 1024])
@pytest.mark.parametrize("with_stats", [False, True])
def test_dit_ln_qkv_attention_fused(ops, B, with_stats):
    """xd_dit_ln_qkv_attn_bf16_tc (LayerNorm-modulate + per-head qkv projection + softmax attention in one kernel) against
    fp32 torch math on bf16-rounded operands (score_networks/dit.py:46-51, layers/attention.py:350-375).  B = 13: ragged
    last tile; B = 128: heads spread over several CTA pairs (the small-M work split); B = 1024: benchmark shape."""
    g = torch.Generator().manual_seed(200 + B)
    T, D, H = 16, 384, 6
    M = B * T
    h = torch.randn(M, D, generator=g) * 1.5 + 0.3
    wqkv = torch.randn(3 * D, D, generator=g) / math.sqrt(D)
    bqkv = torch.randn(3 * D, generator=g) * 0.1
    mod = torch.randn(B, 2 * D, generator=g) * 0.5
    sh, sc = mod[:, :D], mod[:, D:]
    rep = lambda t: t.repeat_interleave(T, dim=0)
    a = bf(_ln(h) * (1 + rep(sc)) + rep(sh)).float()
    qkv = bf(a @ bf(wqkv).float().T + bqkv).float().view(B, T, 3, H, 64)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).permute(0, 2, 1, 3).reshape(M, D)
    wh = bf(wqkv).view(3, H, 64, D).permute(1, 0, 2, 3).reshape(3 * D, D).contiguous().to(DEV)
    bh = bqkv.view(3, H, 64).permute(1, 0, 2).reshape(3 * D).contiguous().to(DEV)
    md = mod.to(DEV)
    stats = None
    if with_stats:
        stats = torch.stack([h.mean(1), torch.rsqrt(h.var(1, unbiased=False) + 1e-6)], 1).contiguous().to(DEV)
    out = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
    torch.ops.xdb200.dit_attn(h.to(DEV), stats, md[:, :D], md[:, D:], T, 1e-6, wh, bh, H, 0.125, out)
    assert rel_l2(out, ref) < 1e-2, rel_l2(out, ref)


def test_schedule_advance_and_misc(ops):
    idx = torch.zeros(1, dtype=torch.int32, device=DEV)
    tab = torch.arange(1000, dtype=torch.int64, device=DEV) * 3
    tf = torch.linspace(0, 1, 1000, device=DEV)
    oi = torch.empty(7, dtype=torch.int64, device=DEV)
    of = torch.empty(7, device=DEV)
    torch.ops.xdb200.schedule_advance(idx, 999, tab, tf, None, oi, of, None, 7)
    assert int(idx) == 999 and bool((oi == 2997).all()) and bool((of == tf[999]).all())
    torch.ops.xdb200.schedule_advance(idx, -1, tab, tf, None, oi, of, None, 7)
    assert int(idx) == 998 and bool((oi == 2994).all())
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn(4096, generator=g), torch.randn(4096, generator=g)
    out = torch.empty(4096, device=DEV)
    torch.ops.xdb200.cfg_combine(a.to(DEV), b.to(DEV), 2.0, out)
    assert torch.equal(out.cpu(), osamplers.cfg_combine(a, b, 2.0))
    torch.ops.xdb200.unnormalize((a * 2).to(DEV), out)
    assert torch.equal(out.cpu(), osamplers.unnormalize(a * 2))


def test_cpu_tensors_are_rejected(ops):
    a = torch.randn(128, 64).bfloat16()
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.xdb200.gemm(a, None, a, None, 0, None, 1, None, torch.empty(128, 128), 0)


def test_strided_conv_channel_bias_and_sr_input(ops):
    """The three small kernels of the cascade's super-resolution stage against torch: stride-2 conv3x3 as im2col + GEMM,
    per-sample channel bias, and the input assembly [x | a * low + c * z] (bit-exact: un-fused fp32 like q_sample)."""
    g = torch.Generator().manual_seed(4)
    n, H, W, C, Co = 3, 8, 8, 64, 128
    x = bf(torch.randn(n, C, H, W, generator=g))
    w = bf(torch.randn(Co, C, 3, 3, generator=g) / math.sqrt(9 * C))
    b = torch.randn(Co, generator=g) * 0.1
    ref = F.conv2d(x.float(), w.float(), b, stride=2, padding=1)                     # [n, Co, 4, 4]
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    cols = torch.empty(n * 16, 9 * C, dtype=torch.bfloat16, device=DEV)
    torch.ops.xdb200.im2col3x3_s2(xd, cols)
    wp = w.permute(0, 2, 3, 1).reshape(Co, 9 * C).contiguous().to(DEV)
    out = ops.linear(cols, wp, b.to(DEV)).view(n, 4, 4, Co)
    assert rel_l2(out.permute(0, 3, 1, 2), ref) < 4e-3
    e = torch.randn(n, C, generator=g)
    y = torch.empty_like(xd)
    torch.ops.xdb200.add_channel_bias(xd.view(n, H * W, C), e.to(DEV), y.view(n, H * W, C))
    assert torch.equal(y.cpu(), bf(xd.cpu().float() + e[:, None, None, :]))
    xx, low, z = (torch.randn(n, 1, 32, 32, generator=g) for _ in range(3))
    a, c = 0.8314696550369263, 0.5555702447891235
    want = torch.cat([xx, torch.tensor(a) * low + torch.tensor(c) * z], 1)
    got = torch.empty(n, 2, 32, 32, device=DEV)
    torch.ops.xdb200.sr_input(xx.to(DEV), low.to(DEV), z.to(DEV), 0, got, a, c, None, 0, 0, None, 0)
    assert torch.equal(got.cpu(), want)
    p1, p2 = torch.empty_like(got), torch.empty_like(got)                            # in-kernel noise: N(0, 1), keyed by step
    torch.ops.xdb200.sr_input(xx.to(DEV), torch.zeros_like(low).to(DEV), None, 0, p1, 1.0, 1.0, None, 5, 9, None, 0)
    torch.ops.xdb200.sr_input(xx.to(DEV), torch.zeros_like(low).to(DEV), None, 0, p2, 1.0, 1.0, None, 6, 9, None, 0)
    assert not torch.equal(p1[:, 1], p2[:, 1]) and abs(float(p1[:, 1].std()) - 1.0) < 0.1 and torch.equal(p1[:, 0].cpu(), xx[:, 0])
