"""torch custom ops (namespace ``xdb200``) over the C ABI in include/xdb200.h.

Every op is an out-variant over CUDA tensors that forwards raw pointers, sizes, strides and the
current CUDA stream to libxdb200.so.  They are registered for the CUDA dispatch key only: calling
one with CPU tensors raises (there is no CPU or eager-PyTorch fallback).  The functional helpers
at the bottom allocate outputs with torch's caching allocator, which keeps the whole forward
capturable in a CUDA graph.
"""
import ctypes
import os

import torch

from . import _lib

F32, BF16 = 0, 1
ACT_NONE, ACT_SILU, ACT_GELU = 0, 1, 2
MODE_ANCESTRAL, MODE_DDIM, MODE_EULER = 0, 1, 2

# "tc": tcgen05/TMEM/TMA kernels (product path).  "simt": CUDA-core twins (device cross-check only).
MATMUL_BACKEND = os.environ.get("XDB200_MATMUL", "tc")
LAUNCHES = 0          # kernels launched through this module (bench.py reports it as gpu_launches)
GROUPNORM_FUSED = os.environ.get("XDB200_GN_FUSED", "1") == "1"      # cluster/DSMEM single-pass GroupNorm


def _p(t):
    return None if t is None else t.data_ptr()


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"xdb200: unsupported dtype {t.dtype}")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.XdError("xdb200 ops take CUDA tensors only (no CPU fallback)")


# Split-K scratch (fp32 partial tiles of long contractions over few tiles): one caller-owned buffer per process.
# The library keeps process-global state (this workspace, lazily configured kernel attributes), so a process drives
# ONE device -- the deployment model is one process per GPU (dist.py); a second device raises instead of silently
# freeing a buffer that captured graphs still reference.
WORKSPACE_BYTES = int(os.environ.get("XDB200_WORKSPACE_MB", "64")) << 20
_workspace = {}


def _ensure_workspace(t):
    dev = t.device
    if _workspace.get("device") not in (None, dev):
        raise _lib.XdError(f"xdb200 drives one CUDA device per process (first used {_workspace['device']}, now {dev}); "
                           "launch one process per GPU (xdiffusion_b200.dist)")
    if _workspace.get("device") != dev:
        buf = torch.empty(WORKSPACE_BYTES, dtype=torch.uint8, device=dev) if WORKSPACE_BYTES else None
        _lib.check(_lib.lib().xd_set_workspace(_p(buf), WORKSPACE_BYTES), "xd_set_workspace")
        _workspace["device"], _workspace["buf"] = dev, buf


#: batch-size-dependent work splits / kernel choices allowed (split-K, hidden-split MLP clusters, the small-batch DiT path)
BATCH_DEPENDENT_PATHS = os.environ.get("XDB200_SPLITK", "1") != "0"


def set_split_k(enabled: bool):
    """Split-K for long contractions over few tiles, and every other choice that depends on the row count (default on).
    Off = bit-exact batch independence: a sub-batch reproduces the rows of the full batch."""
    global BATCH_DEPENDENT_PATHS
    BATCH_DEPENDENT_PATHS = bool(enabled)
    _lib.check(_lib.lib().xd_set_split_k(int(bool(enabled))), "xd_set_split_k")


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


_defs = []
# Ablation timing only (tools/ablate.py): XDB200_SKIP=groupnorm,conv3x3 turns those ops into no-ops so that
# the in-graph cost of a kernel class can be read off the step time.  Results are garbage when set.
_SKIP = set(filter(None, os.environ.get("XDB200_SKIP", "").split(",")))


# ------------------------------------------------------------------------------------ GroupNorm statistics from the producer
# Inside ``with quad_stats():`` (one network forward) the conv3x3 / linear helpers ask the tensor-core kernels to emit, next
# to their bf16 output, the per-(32 rows, 4 channels) sums the GroupNorm that consumes the output needs
# (xd_*_bf16_tc_qstats); ``groupnorm`` then runs the one-pass xd_groupnorm_apply_quads when every channel of its input is
# covered.  The book is keyed by the output's storage (a concat buffer collects the slices of several producers) and keeps
# that storage alive for the duration of the forward, so an address is never re-used under a stale entry; any other op of
# this module that writes into a registered storage drops the entry (``_op`` below).
QUAD_STATS = os.environ.get("XDB200_GN_QSTATS", "1") == "1"
_qs_book = None


class quad_stats:
    def __enter__(self):
        global _qs_book
        self._prev, _qs_book = _qs_book, ({} if QUAD_STATS and MATMUL_BACKEND == "tc" else None)
        return self

    def __exit__(self, *exc):
        global _qs_book
        _qs_book = self._prev


def _qs_geometry(t):
    """(rows, ld, c0, C) of a bf16 [.., C] view whose rows are uniformly strided and start in row 0 of its storage."""
    if t.dim() == 4:
        n, H, W, C = t.shape
        if not (t.stride(3) == 1 and t.stride(1) == W * t.stride(2) and t.stride(0) == H * t.stride(1)):
            return None
        rows, ld = n * H * W, t.stride(2)
    elif t.dim() == 3:
        n, P, C = t.shape
        if not (t.stride(2) == 1 and t.stride(0) == P * t.stride(1)):
            return None
        rows, ld = n * P, t.stride(1)
    elif t.dim() == 2:
        rows, C = t.shape
        if t.stride(1) != 1:
            return None
        ld = t.stride(0)
    else:
        return None
    c0 = t.storage_offset()
    if t.dtype != torch.bfloat16 or c0 + C > ld or rows % 32 or ld % 4 or c0 % 4 or C % 4:
        return None
    return rows, ld, c0, C


def _qs_slot(out):
    """The quad-statistics view for ``out`` (allocating the storage's table on first use), or None."""
    if _qs_book is None:
        return None
    g = _qs_geometry(out)
    if g is None:
        return None
    rows, ld, c0, C = g
    key = out.untyped_storage().data_ptr()
    e = _qs_book.get(key)
    if e is None or e["rows"] != rows or e["ld"] != ld:
        e = {"rows": rows, "ld": ld, "cover": [], "keep": out,
             "table": torch.empty((rows // 32, ld // 4, 2), device=out.device, dtype=torch.float32)}
        _qs_book[key] = e
    return e, e["table"][:, c0 // 4:(c0 + C) // 4], (c0, c0 + C)


def _qs_written(t, emitted_range=None):
    """``t`` was (re)written: forget the statistics of the channels it covers, then record the new ones if emitted."""
    if _qs_book is None or t is None:
        return
    e = _qs_book.get(t.untyped_storage().data_ptr())
    if e is None:
        return
    g = _qs_geometry(t)
    if g is None or g[0] != e["rows"] or g[1] != e["ld"]:
        e["cover"] = []
    else:
        lo, hi = g[2], g[2] + g[3]
        e["cover"] = [(a, b) for a, b in e["cover"] if b <= lo or a >= hi]
    if emitted_range is not None:
        e["cover"].append(emitted_range)


def _qs_lookup(x):
    """Quad-statistics view covering every channel of the GroupNorm input ``x`` [ns, P, C], or None."""
    if _qs_book is None:
        return None
    e = _qs_book.get(x.untyped_storage().data_ptr())
    g = _qs_geometry(x)
    if e is None or g is None or g[0] != e["rows"] or g[1] != e["ld"]:
        return None
    lo, hi = g[2], g[2] + g[3]
    pos = lo
    for a, b in sorted(e["cover"]):
        if a <= pos < b:
            pos = b
    if pos < hi:
        return None
    return e["table"][:, lo // 4:hi // 4]


def _op(schema):
    def deco(fn):
        name = schema.split("(", 1)[0]
        args = schema.split("(", 1)[1].rsplit(")", 1)[0]
        mutated = [i for i, a in enumerate(_split_args(args)) if "!" in a]

        def run(*a):
            r = fn(*a)
            if _qs_book is not None and not name.endswith("_qs"):
                for i in mutated:
                    _qs_written(a[i])
            return r
        if name in _SKIP:
            _defs.append((schema, lambda *a, **k: None))
        else:
            _defs.append((schema, run))
        return fn
    return deco


def _split_args(args):
    out, depth, cur = [], 0, ""
    for ch in args:
        depth += ch == "("
        depth -= ch == ")"
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    return out + [cur]


# ------------------------------------------------------------------------------------ contractions
@_op("ln_gemm(Tensor x, Tensor? shift, Tensor? scale, int rows_per_mod, float eps, Tensor w, Tensor? bias, int act, "
     "Tensor(a!) out) -> ()")
def _ln_gemm(x, shift, scale, rows_per_mod, eps, w, bias, act, out):
    _cuda(x, shift, scale, w, bias, out)
    M, K = x.shape
    N = w.shape[0]
    assert x.dtype == torch.float32 and w.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and w.shape[1] == K
    assert x.stride(1) == 1 and w.stride(1) == 1 and out.stride(1) == 1 and out.shape == (M, N)
    mod = shift if shift is not None else scale
    assert mod is None or (mod.stride(1) == 1 and (shift is None or scale is None or shift.stride(0) == scale.stride(0)))
    _lib.check(_lib.lib().xd_ln_gemm_bf16_tc(_p(x), x.stride(0), _p(shift), _p(scale), 0 if mod is None else mod.stride(0),
                                             rows_per_mod, eps, _p(w), w.stride(0), M, N, K, _p(bias), act, _p(out),
                                             out.stride(0), _stream()), "xd_ln_gemm_bf16_tc")


@_op("dit_attn(Tensor h, Tensor? stats, Tensor shift, Tensor scale, int rows_per_mod, float eps, Tensor wh, Tensor bias, "
     "int heads, float sm_scale, Tensor(a!) out) -> ()")
def _dit_attn(h, stats, shift, scale, rows_per_mod, eps, wh, bias, heads, sm_scale, out):
    """Fused LayerNorm-modulate + per-head QKV projection + softmax attention: h fp32 [M, 384] -> out bf16 [M, 384]."""
    _cuda(h, stats, shift, scale, wh, bias, out)
    M, D = h.shape
    assert h.dtype == torch.float32 and h.stride(1) == 1 and out.dtype == torch.bfloat16 and out.shape == (M, D) and out.stride(1) == 1
    assert wh.dtype == torch.bfloat16 and wh.is_contiguous() and wh.shape == (heads * 192, D) and bias.is_contiguous()
    assert shift.dtype == torch.float32 and shift.stride(1) == 1 and scale.stride(1) == 1 and shift.stride(0) == scale.stride(0)
    assert stats is None or (stats.dtype == torch.float32 and stats.is_contiguous() and stats.numel() >= 2 * M)
    _lib.check(_lib.lib().xd_dit_ln_qkv_attn_bf16_tc(_p(h), h.stride(0), _p(stats), _p(shift), _p(scale), shift.stride(0),
                                                     rows_per_mod, eps, _p(wh), _p(bias), heads, M, D, sm_scale, _p(out),
                                                     out.stride(0), _stream()), "xd_dit_ln_qkv_attn_bf16_tc")
    _count()


@_op("dit_proj_mlp(Tensor o, Tensor wp, Tensor bp, Tensor w1, Tensor b1, Tensor w2, Tensor b2, Tensor h_in, Tensor(a!) h, "
     "Tensor gate1, Tensor shift2, Tensor scale2, Tensor gate2, int rows_per_mod, float eps, Tensor(b!)? stats, int split) -> ()")
def _dit_proj_mlp(o, wp, bp, w1, b1, w2, b2, h_in, h, gate1, shift2, scale2, gate2, rows_per_mod, eps, stats, split):
    """Fused proj + gated residual + LayerNorm-modulate + fc1 + GELU + fc2 + gated residual: h_in -> h (fp32 [M, 384]; h may be
    h_in).  split: CTA pairs per 256-row tile (0 = chosen from M; > 1 needs distinct buffers)."""
    _cuda(o, wp, bp, w1, b1, w2, b2, h_in, h, gate1, shift2, scale2, gate2, stats)
    M, D = h.shape
    assert h_in.shape == h.shape and h_in.dtype == torch.float32 and h_in.stride() == h.stride()
    hidden = w1.shape[0]
    assert o.dtype == torch.bfloat16 and o.shape == (M, D) and o.stride(1) == 1 and h.dtype == torch.float32 and h.stride(1) == 1
    assert wp.is_contiguous() and w1.is_contiguous() and w2.is_contiguous() and wp.shape == (D, D)
    assert w1.shape == (hidden, D) and w2.shape == (D, hidden) and all(w.dtype == torch.bfloat16 for w in (wp, w1, w2))
    mods = (gate1, shift2, scale2, gate2)
    assert all(t.dtype == torch.float32 and t.stride(1) == 1 and t.stride(0) == gate1.stride(0) for t in mods)
    assert stats is None or (stats.dtype == torch.float32 and stats.is_contiguous() and stats.numel() >= 2 * M)
    _lib.check(_lib.lib().xd_dit_proj_mlp_bf16_tc(_p(o), o.stride(0), _p(wp), _p(bp), _p(w1), _p(b1), _p(w2), _p(b2), hidden,
                                                  _p(h_in), _p(h), h.stride(0), M, D, _p(gate1), _p(shift2), _p(scale2),
                                                  _p(gate2), gate1.stride(0), rows_per_mod, eps, _p(stats), split, _stream()),
               "xd_dit_proj_mlp_bf16_tc")
    _count()


@_op("dit_attn_rows(Tensor h, Tensor? stats, Tensor shift, Tensor scale, int rows_per_mod, float eps, Tensor wh, Tensor bias, "
     "int heads, float sm_scale, Tensor mod_rows, Tensor(a!) out) -> ()")
def _dit_attn_rows(h, stats, shift, scale, rows_per_mod, eps, wh, bias, heads, sm_scale, mod_rows, out):
    """dit_attn with the modulation row of image i = mod_rows[i] (rows of a per-loop (label, step) table)."""
    _cuda(h, stats, shift, scale, wh, bias, mod_rows, out)
    M, D = h.shape
    assert h.dtype == torch.float32 and h.stride(1) == 1 and out.dtype == torch.bfloat16 and out.shape == (M, D) and out.stride(1) == 1
    assert wh.dtype == torch.bfloat16 and wh.is_contiguous() and wh.shape == (heads * 192, D) and bias.is_contiguous()
    assert shift.dtype == torch.float32 and shift.stride(1) == 1 and scale.stride(1) == 1 and shift.stride(0) == scale.stride(0)
    assert stats is None or (stats.dtype == torch.float32 and stats.is_contiguous() and stats.numel() >= 2 * M)
    assert mod_rows.dtype == torch.int32 and mod_rows.is_contiguous() and mod_rows.numel() * rows_per_mod == M
    _lib.check(_lib.lib().xd_dit_ln_qkv_attn_bf16_tc_rows(_p(h), h.stride(0), _p(stats), _p(shift), _p(scale), shift.stride(0),
                                                          rows_per_mod, eps, _p(wh), _p(bias), heads, M, D, sm_scale, _p(out),
                                                          out.stride(0), _p(mod_rows), _stream()),
               "xd_dit_ln_qkv_attn_bf16_tc_rows")
    _count()


@_op("dit_proj_mlp_rows(Tensor o, Tensor wp, Tensor bp, Tensor w1, Tensor b1, Tensor w2, Tensor b2, Tensor h_in, Tensor(a!) h, "
     "Tensor gate1, Tensor shift2, Tensor scale2, Tensor gate2, int rows_per_mod, float eps, Tensor(b!)? stats, int split, "
     "Tensor mod_rows) -> ()")
def _dit_proj_mlp_rows(o, wp, bp, w1, b1, w2, b2, h_in, h, gate1, shift2, scale2, gate2, rows_per_mod, eps, stats, split, mod_rows):
    """dit_proj_mlp with the modulation row of image i = mod_rows[i]."""
    _cuda(o, wp, bp, w1, b1, w2, b2, h_in, h, gate1, shift2, scale2, gate2, stats, mod_rows)
    M, D = h.shape
    assert h_in.shape == h.shape and h_in.dtype == torch.float32 and h_in.stride() == h.stride()
    hidden = w1.shape[0]
    assert o.dtype == torch.bfloat16 and o.shape == (M, D) and o.stride(1) == 1 and h.dtype == torch.float32 and h.stride(1) == 1
    assert wp.is_contiguous() and w1.is_contiguous() and w2.is_contiguous() and wp.shape == (D, D)
    assert w1.shape == (hidden, D) and w2.shape == (D, hidden) and all(w.dtype == torch.bfloat16 for w in (wp, w1, w2))
    mods = (gate1, shift2, scale2, gate2)
    assert all(t.dtype == torch.float32 and t.stride(1) == 1 and t.stride(0) == gate1.stride(0) for t in mods)
    assert stats is None or (stats.dtype == torch.float32 and stats.is_contiguous() and stats.numel() >= 2 * M)
    assert mod_rows.dtype == torch.int32 and mod_rows.is_contiguous() and mod_rows.numel() * rows_per_mod == M
    _lib.check(_lib.lib().xd_dit_proj_mlp_bf16_tc_rows(_p(o), o.stride(0), _p(wp), _p(bp), _p(w1), _p(b1), _p(w2), _p(b2), hidden,
                                                       _p(h_in), _p(h), h.stride(0), M, D, _p(gate1), _p(shift2), _p(scale2),
                                                       _p(gate2), gate1.stride(0), rows_per_mod, eps, _p(stats), split,
                                                       _p(mod_rows), _stream()), "xd_dit_proj_mlp_bf16_tc_rows")
    _count()


@_op("gemm(Tensor a, Tensor? a2, Tensor w, Tensor? bias, int act, Tensor? gate, int gate_rows, "
     "Tensor? residual, Tensor(a!) out, int force_bn) -> ()")
def _gemm(a, a2, w, bias, act, gate, gate_rows, residual, out, force_bn):
    _cuda(a, a2, w, bias, gate, residual, out)
    _ensure_workspace(a)
    M, K = a.shape
    K2 = 0 if a2 is None else a2.shape[1]
    N = w.shape[0]
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.shape[1] == K + K2
    assert a.stride(1) == 1 and w.stride(1) == 1 and out.stride(1) == 1 and out.shape == (M, N)
    args = [_p(a), a.stride(0), _p(a2), 0 if a2 is None else a2.stride(0), K2, _p(w), w.stride(0), M, N, K,
            _p(bias), act, _p(gate), gate_rows, 0 if gate is None else gate.stride(0), _p(residual),
            0 if residual is None else _dt(residual), 0 if residual is None else residual.stride(0),
            _p(out), _dt(out), out.stride(0)]
    if MATMUL_BACKEND == "tc" and K % 64 == 0 and K2 % 64 == 0:
        _lib.check(_lib.lib().xd_gemm_bf16_tc(*args, force_bn, _stream()), "xd_gemm_bf16_tc")
    else:
        _lib.check(_lib.lib().xd_gemm_bf16_simt(*args, _stream()), "xd_gemm_bf16_simt")
    _count()


@_op("conv3x3(Tensor x, Tensor? xs, Tensor wp, Tensor? bias, int act, Tensor? residual, Tensor(a!) out, "
     "int force_bn) -> ()")
def _conv3x3(x, xs, wp, bias, act, residual, out, force_bn):
    """x [nimg,H,W,C] bf16 NHWC view (channel stride 1, dense pixels, pixel stride >= C)."""
    _cuda(x, xs, wp, bias, residual, out)
    _ensure_workspace(x)
    nimg, H, W, C = x.shape
    Cs = 0 if xs is None else xs.shape[3]
    Cout = wp.shape[0]
    assert wp.shape[1] == 9 * C + Cs and wp.is_contiguous()
    for t in (x, xs, out, residual):
        if t is not None:
            assert t.stride(3) == 1 and t.stride(1) == W * t.stride(2) and t.stride(0) == H * t.stride(1)
    args = [_p(x), x.stride(2), nimg, H, W, C, _p(xs), 0 if xs is None else xs.stride(2), Cs, _p(wp), Cout,
            _p(bias), act, _p(residual), 0 if residual is None else _dt(residual),
            0 if residual is None else residual.stride(2), _p(out), _dt(out), out.stride(2)]
    if MATMUL_BACKEND == "tc":
        _lib.check(_lib.lib().xd_conv3x3_bf16_tc(*args, force_bn, _stream()), "xd_conv3x3_bf16_tc")
    else:
        _lib.check(_lib.lib().xd_conv3x3_bf16_simt(*args, _stream()), "xd_conv3x3_bf16_simt")
    _count()


@_op("gemm_qs(Tensor a, Tensor? a2, Tensor w, Tensor? bias, int act, Tensor? gate, int gate_rows, "
     "Tensor? residual, Tensor(a!) out, int force_bn, Tensor(b!) qstats) -> int")
def _gemm_qs(a, a2, w, bias, act, gate, gate_rows, residual, out, force_bn, qstats):
    """gemm + quad statistics of ``out`` (see quad_stats above); returns 1 when the statistics were written."""
    _cuda(a, a2, w, bias, gate, residual, out, qstats)
    _ensure_workspace(a)
    M, K = a.shape
    K2 = 0 if a2 is None else a2.shape[1]
    N = w.shape[0]
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.shape[1] == K + K2 and K % 64 == 0 and K2 % 64 == 0
    assert a.stride(1) == 1 and w.stride(1) == 1 and out.stride(1) == 1 and out.shape == (M, N)
    assert qstats.dtype == torch.float32 and qstats.shape == (M // 32, N // 4, 2) and qstats.stride(2) == 1 and qstats.stride(1) == 2
    emitted = ctypes.c_int(0)
    _lib.check(_lib.lib().xd_gemm_bf16_tc_qstats(
        _p(a), a.stride(0), _p(a2), 0 if a2 is None else a2.stride(0), K2, _p(w), w.stride(0), M, N, K,
        _p(bias), act, _p(gate), gate_rows, 0 if gate is None else gate.stride(0), _p(residual),
        0 if residual is None else _dt(residual), 0 if residual is None else residual.stride(0),
        _p(out), _dt(out), out.stride(0), force_bn, _p(qstats), qstats.stride(0), ctypes.addressof(emitted), _stream()),
        "xd_gemm_bf16_tc_qstats")
    _count()
    return emitted.value


@_op("conv3x3_qs(Tensor x, Tensor? xs, Tensor wp, Tensor? bias, int act, Tensor? residual, Tensor(a!) out, "
     "int force_bn, Tensor(b!) qstats) -> int")
def _conv3x3_qs(x, xs, wp, bias, act, residual, out, force_bn, qstats):
    _cuda(x, xs, wp, bias, residual, out, qstats)
    _ensure_workspace(x)
    nimg, H, W, C = x.shape
    Cs = 0 if xs is None else xs.shape[3]
    Cout = wp.shape[0]
    assert wp.shape[1] == 9 * C + Cs and wp.is_contiguous()
    for t in (x, xs, out, residual):
        if t is not None:
            assert t.stride(3) == 1 and t.stride(1) == W * t.stride(2) and t.stride(0) == H * t.stride(1)
    assert qstats.dtype == torch.float32 and qstats.shape == (nimg * H * W // 32, Cout // 4, 2)
    assert qstats.stride(2) == 1 and qstats.stride(1) == 2
    emitted = ctypes.c_int(0)
    _lib.check(_lib.lib().xd_conv3x3_bf16_tc_qstats(
        _p(x), x.stride(2), nimg, H, W, C, _p(xs), 0 if xs is None else xs.stride(2), Cs, _p(wp), Cout,
        _p(bias), act, _p(residual), 0 if residual is None else _dt(residual),
        0 if residual is None else residual.stride(2), _p(out), _dt(out), out.stride(2), force_bn,
        _p(qstats), qstats.stride(0), ctypes.addressof(emitted), _stream()), "xd_conv3x3_bf16_tc_qstats")
    _count()
    return emitted.value


@_op("conv3x3_groupnorm(Tensor x, Tensor wp, Tensor? bias, int nsamples, Tensor gamma, Tensor beta, Tensor? scale_shift, "
     "int ss_div, float eps, int silu, Tensor(a!) tmp, Tensor(b!) scratch, Tensor(c!) out) -> ()")
def _conv3x3_groupnorm(x, wp, bias, nsamples, gamma, beta, scale_shift, ss_div, eps, silu, tmp, scratch, out):
    """out = GroupNorm32(conv3x3(x) + bias) [*(1 + scale) + shift] [SiLU]; see include/xdb200.h for the three routes."""
    _cuda(x, wp, bias, gamma, beta, scale_shift, tmp, scratch, out)
    _ensure_workspace(x)
    nimg, H, W, C = x.shape
    Cout = wp.shape[0]
    assert wp.shape[1] == 9 * C and wp.is_contiguous() and tmp.is_contiguous() and scratch.is_contiguous()
    assert tmp.dtype == torch.bfloat16 and tmp.numel() == nimg * H * W * Cout and scratch.dtype == torch.float32
    for t in (x, out):
        assert t.stride(3) == 1 and t.stride(1) == W * t.stride(2) and t.stride(0) == H * t.stride(1)
    _lib.check(_lib.lib().xd_conv3x3_groupnorm_bf16_tc(
        _p(x), x.stride(2), nimg, H, W, C, _p(wp), Cout, _p(bias), nsamples, _p(gamma), _p(beta), _p(scale_shift),
        0 if scale_shift is None else scale_shift.stride(0), ss_div, eps, silu, _p(tmp), _p(scratch), _p(out), out.stride(2),
        _stream()), "xd_conv3x3_groupnorm_bf16_tc")
    _count(2)


@_op("groupnorm_quads(Tensor x, Tensor qstats, Tensor gamma, Tensor beta, Tensor? scale_shift, int ss_div, float eps, "
     "int silu, int nsamples, Tensor(a!) out) -> ()")
def _groupnorm_quads(x, qstats, gamma, beta, scale_shift, ss_div, eps, silu, nsamples, out):
    """x [rows, C] bf16 view of nsamples samples of P consecutive rows; qstats [rows / 32, C / 4, 2] from the producers."""
    _cuda(x, qstats, gamma, beta, scale_shift, out)
    rows, C = x.shape
    P = rows // nsamples
    assert x.stride(1) == 1 and out.stride(1) == 1 and nsamples * P == rows and qstats.shape == (rows // 32, C // 4, 2)
    assert qstats.stride(2) == 1 and qstats.stride(1) == 2
    _lib.check(_lib.lib().xd_groupnorm_apply_quads(
        _p(x), x.stride(0), nsamples, P, C, 32, _p(qstats), qstats.stride(0), _p(gamma), _p(beta), _p(scale_shift),
        0 if scale_shift is None else scale_shift.stride(0), ss_div, eps, silu, _p(out), out.stride(0), _stream()),
        "xd_groupnorm_apply_quads")
    _count()


@_op("conv3x3_in(Tensor x, Tensor w, Tensor? bias, Tensor(a!) out) -> ()")
def _conv3x3_in(x, w, bias, out):
    _cuda(x, w, bias, out)
    nimg, Cin, H, W = x.shape
    assert x.is_contiguous() and w.is_contiguous() and x.dtype == torch.float32 and out.dtype == torch.bfloat16
    _lib.check(_lib.lib().xd_conv3x3_in_f32_nchw(_p(x), nimg, Cin, H, W, _p(w), _p(bias), w.shape[0], _p(out),
                                                 out.stride(2), _stream()), "xd_conv3x3_in_f32_nchw")
    _count()


@_op("conv3x3_out(Tensor x, Tensor w, Tensor? bias, Tensor(a!) out) -> ()")
def _conv3x3_out(x, w, bias, out):
    _cuda(x, w, bias, out)
    nimg, H, W, C = x.shape
    assert out.is_contiguous() and w.is_contiguous() and out.dtype == torch.float32
    _lib.check(_lib.lib().xd_conv3x3_out_f32_nchw(_p(x), x.stride(2), nimg, H, W, C, _p(w), _p(bias), w.shape[0],
                                                  _p(out), _stream()), "xd_conv3x3_out_f32_nchw")
    _count()


@_op("attention(Tensor q, Tensor k, Tensor v, Tensor(a!) o, float scale, Tensor? relk, int scramble, int o_cs, "
     "int hpg, int o_gs) -> ()")
def _attention(q, k, v, o, scale, relk, scramble, o_cs, hpg, o_gs):
    """q [B,H,Tq,64], k/v [B,H,Tk,64], o [B,H,Tq,64]: arbitrary-stride bf16 views (last stride 1)."""
    _cuda(q, k, v, o, relk)
    B, H, Tq, D = q.shape
    Tk = k.shape[2]
    for t in (q, k, v):
        assert t.dtype == q.dtype and t.dtype in (torch.bfloat16, torch.float32) and t.stride(3) == 1
    _lib.check(_lib.lib().xd_attention_bf16(
        _p(q), q.stride(0), q.stride(1), q.stride(2), _p(k), k.stride(0), k.stride(1), k.stride(2),
        _p(v), v.stride(0), v.stride(1), v.stride(2), _p(o), o.stride(0), o.stride(1), o.stride(2),
        B, H, Tq, Tk, D, scale, _p(relk), scramble, o_cs, _dt(q), hpg, o_gs, _stream()), "xd_attention_bf16")
    _count()


# ------------------------------------------------------------------------------------ norms
@_op("groupnorm(Tensor x, Tensor gamma, Tensor beta, Tensor? scale_shift, int ss_div, float eps, int silu, "
     "int inner, int nsamples, int split, Tensor(a!) stats, Tensor(b!) out) -> ()")
def _groupnorm(x, gamma, beta, scale_shift, ss_div, eps, silu, inner, nsamples, split, stats, out):
    """x [rows, C] bf16 view (channel stride 1) holding nsamples samples of P = rows / nsamples rows
    each; row(s, p) = (s / inner)*P*inner + s % inner + p*inner; 32 groups; stats fp32 [nsamples*64]."""
    _cuda(x, gamma, beta, scale_shift, stats, out)
    rows, C = x.shape
    ns, P = nsamples, rows // nsamples
    assert x.stride(1) == 1 and out.stride(1) == 1 and ns * P == rows and stats.dtype == torch.float32
    assert stats.numel() >= ns * 64 * _lib.lib().xd_groupnorm_slabs(ns, P, C) and stats.is_contiguous()
    l = _lib.lib()
    if inner == 1 and not split and GROUPNORM_FUSED:
        rc = l.xd_groupnorm_fused(_p(x), x.stride(0), ns, P, C, 32, _p(gamma), _p(beta), _p(scale_shift),
                                  0 if scale_shift is None else scale_shift.stride(0), ss_div, eps, silu, _p(out),
                                  out.stride(0), _stream())
        if rc != -1:
            _lib.check(rc, "xd_groupnorm_fused")
            _count()
            return
    _lib.check(l.xd_groupnorm_stats(_p(x), x.stride(0), ns, P, C, 32, inner, _p(stats), _stream()),
               "xd_groupnorm_stats")
    _lib.check(l.xd_groupnorm_apply(_p(x), x.stride(0), ns, P, C, 32, _p(stats), _p(gamma), _p(beta),
                                    _p(scale_shift), 0 if scale_shift is None else scale_shift.stride(0), ss_div,
                                    eps, silu, inner, split, _p(out), out.stride(0), _stream()), "xd_groupnorm_apply")
    _count(3)


@_op("groupnorm_frames_split(Tensor x, Tensor gamma, Tensor beta, float eps, int B, int F, int HW, Tensor(a!) out) -> int")
def _groupnorm_frames_split(x, gamma, beta, eps, B, F, HW, out):
    """x bf16 [B*F*HW, C] rows (clip, frame, pixel) -> out bf16 [B*F*HW, 2C] = [hi | lo] of the per-pixel GroupNorm over
    frames; returns 0 when the shape is not covered by the one-pass kernel (nothing written)."""
    _cuda(x, gamma, beta, out)
    rows, C = x.shape
    assert rows == B * F * HW and x.stride(1) == 1 and out.stride(1) == 1 and out.shape == (rows, 2 * C)
    rc = _lib.lib().xd_groupnorm_frames_split(_p(x), x.stride(0), B, F, HW, C, _p(gamma), _p(beta), eps, _p(out),
                                              out.stride(0), _stream())
    if rc == -1:
        return 0
    _lib.check(rc, "xd_groupnorm_frames_split")
    _count()
    return 1


@_op("layernorm_modulate(Tensor x, Tensor? shift, Tensor? scale, int rows_per_mod, float eps, Tensor(a!) out) -> ()")
def _layernorm_modulate(x, shift, scale, rows_per_mod, eps, out):
    _cuda(x, shift, scale, out)
    M, D = x.shape
    mod_ld = 0
    for t in (shift, scale):
        if t is not None:
            assert t.stride(-1) == 1 and t.dtype == torch.float32
            mod_ld = t.stride(0)
    if shift is not None and scale is not None:
        assert shift.stride(0) == scale.stride(0)
    assert x.dtype == torch.float32 and x.stride(1) == 1 and out.dtype == torch.bfloat16
    _lib.check(_lib.lib().xd_layernorm_modulate(_p(x), x.stride(0), M, D, _p(shift), _p(scale), mod_ld,
                                                rows_per_mod, eps, _p(out), out.stride(0), _stream()),
               "xd_layernorm_modulate")
    _count()


# ------------------------------------------------------------------------------------ small kernels
@_op("timestep_embed(Tensor t, Tensor freq, int mode, float max_time, float clip_lo, float clip_hi, int order, "
     "Tensor(a!) out) -> ()")
def _timestep_embed(t, freq, mode, max_time, clip_lo, clip_hi, order, out):
    _cuda(t, freq, out)
    assert t.dtype in (torch.int64, torch.float32) and t.is_contiguous() and out.is_contiguous()
    of, ob = (_p(out), None) if out.dtype == torch.float32 else (None, _p(out))
    _lib.check(_lib.lib().xd_timestep_embed(_p(t), int(t.dtype == torch.int64), t.shape[0], _p(freq),
                                            freq.shape[0], mode, max_time, clip_lo, clip_hi, order, of, ob,
                                            _stream()), "xd_timestep_embed")
    _count()


@_op("act_cast(Tensor x, int act, Tensor(a!) out) -> ()")
def _act_cast(x, act, out):
    _cuda(x, out)
    assert x.is_contiguous() and out.is_contiguous() and x.numel() == out.numel()
    _lib.check(_lib.lib().xd_act_cast(_p(x), _dt(x), _p(out), _dt(out), act, x.numel(), _stream()), "xd_act_cast")
    _count()


@_op("class_combine(Tensor? table, Tensor? labels, Tensor temb, Tensor(a!)? c_out, Tensor(b!)? silu_out) -> ()")
def _class_combine(table, labels, temb, c_out, silu_out):
    _cuda(table, labels, temb, c_out, silu_out)
    B, D = temb.shape
    assert temb.is_contiguous() and (labels is None or labels.dtype == torch.int64)
    _lib.check(_lib.lib().xd_class_combine(_p(table), _p(labels), _p(temb), B, D, _p(c_out), _p(silu_out),
                                           _stream()), "xd_class_combine")
    _count()


@_op("class_combine_step(Tensor? table, Tensor? labels, Tensor temb_table, Tensor idx, int B, Tensor(a!)? c_out, "
     "Tensor(b!)? silu_out, Tensor(c!)? rows_out) -> ()")
def _class_combine_step(table, labels, temb_table, idx, B, c_out, silu_out, rows_out):
    _cuda(table, labels, temb_table, idx, c_out, silu_out, rows_out)
    assert rows_out is None or (rows_out.dtype == torch.int32 and rows_out.is_contiguous() and rows_out.numel() == B)
    D = temb_table.shape[1]
    assert temb_table.is_contiguous() and temb_table.dtype == torch.float32 and idx.dtype == torch.int32
    assert labels is None or (labels.dtype == torch.int64 and labels.shape[0] == B)
    _lib.check(_lib.lib().xd_class_combine_step(_p(table), _p(labels), _p(temb_table), _p(idx), B, D, _p(c_out),
                                                _p(silu_out), _p(rows_out), temb_table.shape[0], _stream()),
               "xd_class_combine_step")
    _count()


@_op("gather_row(Tensor table, Tensor idx, Tensor(a!) out) -> ()")
def _gather_row(table, idx, out):
    """out fp32 [W] = table[idx[0]] (idx int32 on the device)."""
    _cuda(table, idx, out)
    assert table.dtype == torch.float32 and table.stride(1) == 1 and out.dtype == torch.float32 and out.is_contiguous()
    assert idx.dtype == torch.int32 and out.numel() == table.shape[1]
    _lib.check(_lib.lib().xd_gather_row_f32(_p(table), table.stride(0), _p(idx), table.shape[1], _p(out), _stream()),
               "xd_gather_row_f32")
    _count()


@_op("patchify(Tensor x, int p, Tensor(a!) out) -> ()")
def _patchify(x, p, out):
    _cuda(x, out)
    B, C, H, W = x.shape
    assert x.is_contiguous() and x.dtype == torch.float32 and out.is_contiguous() and out.dtype == torch.bfloat16
    _lib.check(_lib.lib().xd_patchify(_p(x), B, C, H, W, p, _p(out), _stream()), "xd_patchify")
    _count()


@_op("unpatchify(Tensor y, int p, Tensor(a!) out) -> ()")
def _unpatchify(y, p, out):
    _cuda(y, out)
    B, C, H, W = out.shape
    assert out.is_contiguous() and y.dtype == torch.float32 and y.stride(1) == 1
    _lib.check(_lib.lib().xd_unpatchify(_p(y), y.stride(0), B, C, H, W, p, _p(out), _stream()), "xd_unpatchify")
    _count()


@_op("add_rows_periodic(Tensor a, Tensor b, int period, Tensor(a!) out) -> ()")
def _add_rows_periodic(a, b, period, out):
    _cuda(a, b, out)
    assert a.is_contiguous() and b.is_contiguous() and out.is_contiguous()
    _lib.check(_lib.lib().xd_add_rows_periodic(_p(a), _p(b), a.shape[0], a.shape[1], period, _p(out), _stream()),
               "xd_add_rows_periodic")
    _count()


@_op("add_table(Tensor a, Tensor tab, Tensor(a!) out) -> ()")
def _add_table(a, tab, out):
    _cuda(a, tab, out)
    G, C = tab.shape
    R = a.shape[0]
    assert a.is_contiguous() and tab.is_contiguous() and out.is_contiguous() and a.shape[1] == C
    _lib.check(_lib.lib().xd_add_table(_p(a), _p(tab), G, R, C, _p(out), _stream()), "xd_add_table")
    _count()


@_op("avgpool2x2(Tensor x, Tensor(a!) out) -> ()")
def _avgpool2x2(x, out):
    _cuda(x, out)
    nimg, H, W, C = x.shape
    _lib.check(_lib.lib().xd_avgpool2x2_nhwc(_p(x), x.stride(2), nimg, H, W, C, _p(out), out.stride(2), _stream()),
               "xd_avgpool2x2_nhwc")
    _count()


@_op("upsample2x(Tensor x, Tensor(a!) out) -> ()")
def _upsample2x(x, out):
    _cuda(x, out)
    nimg, H, W, C = x.shape
    _lib.check(_lib.lib().xd_upsample2x_nhwc(_p(x), x.stride(2), nimg, H, W, C, _p(out), out.stride(2), _stream()),
               "xd_upsample2x_nhwc")
    _count()


@_op("copy_rows(Tensor x, Tensor(a!) out) -> ()")
def _copy_rows(x, out):
    """bf16 [batch, rows, C] -> [batch, rows, C], both strided (unit stride on C)."""
    _cuda(x, out)
    nb, rows, C = x.shape
    assert out.shape == x.shape and x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
    assert x.stride(2) == 1 and out.stride(2) == 1
    _lib.check(_lib.lib().xd_copy_rows_bf16(_p(x), x.stride(1), x.stride(0), nb * rows, rows, C, _p(out), out.stride(1),
                                            out.stride(0), _stream()), "xd_copy_rows_bf16")
    _count()


@_op("im2col3x3_s2(Tensor x, Tensor(a!) out) -> ()")
def _im2col3x3_s2(x, out):
    """x bf16 NHWC [n, H, W, C] (pixel stride x.stride(2)) -> out bf16 [n * H/2 * W/2, 9 * C] (stride-2, pad-1 taps)."""
    _cuda(x, out)
    n, H, W, C = x.shape
    assert x.dtype == torch.bfloat16 and x.stride(3) == 1 and x.stride(1) == W * x.stride(2) and x.stride(0) == H * x.stride(1)
    assert out.dtype == torch.bfloat16 and out.is_contiguous() and out.shape == (n * (H // 2) * (W // 2), 9 * C)
    _lib.check(_lib.lib().xd_im2col3x3_s2_nhwc(_p(x), x.stride(2), n, H, W, C, _p(out), _stream()), "xd_im2col3x3_s2_nhwc")
    _count()


@_op("add_channel_bias(Tensor x, Tensor bias, Tensor(a!) out) -> ()")
def _add_channel_bias(x, bias, out):
    """x bf16 [n, P, C] (uniform row stride) + bias fp32 [n, C] -> out bf16 [n, P, C]."""
    _cuda(x, bias, out)
    n, P, C = x.shape
    assert x.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and bias.dtype == torch.float32
    assert x.stride(2) == 1 and out.stride(2) == 1 and x.stride(0) == P * x.stride(1) and out.stride(0) == P * out.stride(1)
    assert bias.shape == (n, C) and bias.stride(1) == 1 and out.shape == x.shape
    _lib.check(_lib.lib().xd_add_channel_bias_nhwc(_p(x), x.stride(1), _p(bias), bias.stride(0), n, P, C, _p(out),
                                                   out.stride(1), _stream()), "xd_add_channel_bias_nhwc")
    _count()


@_op("sr_input(Tensor x, Tensor low, Tensor? z, int z_step_stride, Tensor(a!) out, float a, float c, Tensor? idx_dev, "
     "int idx_host, int seed, Tensor? seed_dev, int elem_offset) -> ()")
def _sr_input(x, low, z, z_step_stride, out, a, c, idx_dev, idx_host, seed, seed_dev, elem_offset):
    """out[b] = [x[b] | a * low[b] + c * z[b]] on the channel axis (fp32 NCHW), z injected or in-kernel Philox."""
    _cuda(x, low, z, out, idx_dev, seed_dev)
    B = x.shape[0]
    nx, nl = x[0].numel(), low[0].numel()
    assert x.is_contiguous() and low.is_contiguous() and out.is_contiguous() and out[0].numel() == nx + nl
    assert all(t.dtype == torch.float32 for t in (x, low, out)) and low.shape[0] == B and out.shape[0] == B
    assert z is None or (z.dtype == torch.float32 and z.is_contiguous())
    _lib.check(_lib.lib().xd_sr_input(_p(x), _p(low), _p(z), z_step_stride, _p(out), B, nx, nl, a, c, _p(idx_dev), idx_host,
                                      seed, _p(seed_dev), elem_offset, _stream()), "xd_sr_input")
    _count()


@_op("blend_frames(Tensor(a!) x, Tensor x0, Tensor mask) -> ()")
def _blend_frames(x, x0, mask):
    """x[b, c, f] = mask[b, f] ? x[b, c, f] : x0[b, c, f] in place (x, x0 fp32 [B, C, F, H, W]; mask bool / uint8 [B, F])."""
    _cuda(x, x0, mask)
    B, C, F, H, W = x.shape
    assert x.dtype == torch.float32 and x0.dtype == torch.float32 and x.is_contiguous() and x0.is_contiguous()
    assert x0.shape == x.shape and mask.shape == (B, F) and mask.is_contiguous() and mask.element_size() == 1
    _lib.check(_lib.lib().xd_blend_frames(_p(x), _p(x0), _p(mask), B, C, F, H * W, _stream()), "xd_blend_frames")
    _count()


@_op("cfg_combine(Tensor cond, Tensor uncond, float w, Tensor(a!) out) -> ()")
def _cfg_combine(cond, uncond, w, out):
    _cuda(cond, uncond, out)
    assert cond.is_contiguous() and uncond.is_contiguous() and out.is_contiguous()
    _lib.check(_lib.lib().xd_cfg_combine(_p(cond), _p(uncond), w, _p(out), cond.numel(), _stream()),
               "xd_cfg_combine")
    _count()


# ------------------------------------------------------------------------------------ sampler step
@_op("sampler_step(int mode, int form, int pred_v, Tensor x, Tensor o, Tensor? z, int z_step_stride, "
     "Tensor(a!) out, Tensor coefs, Tensor? idx_dev, int idx_host, int threshold, int thr_k, float thr_w, "
     "float thr_c, int seed, Tensor? seed_dev, int elem_offset) -> ()")
def _sampler_step(mode, form, pred_v, x, o, z, z_step_stride, out, coefs, idx_dev, idx_host, threshold, thr_k,
                  thr_w, thr_c, seed, seed_dev, elem_offset):
    _cuda(x, o, z, out, coefs, idx_dev, seed_dev)
    assert seed_dev is None or seed_dev.dtype == torch.int64
    assert x.is_contiguous() and o.is_contiguous() and out.is_contiguous() and coefs.is_contiguous()
    assert x.dtype == torch.float32 and o.dtype == torch.float32 and coefs.dtype == torch.float32
    assert idx_dev is None or idx_dev.dtype == torch.int32
    n = x.numel()
    if idx_dev is None and not 0 <= idx_host < coefs.shape[0]:
        raise IndexError(f"timestep_idx {idx_host} outside the {coefs.shape[0]}-row schedule table")
    _lib.check(_lib.lib().xd_sampler_step(mode, form, pred_v, _p(x), _p(o), _p(z), z_step_stride, _p(out),
                                          _p(coefs), _p(idx_dev), idx_host, n, n // x.shape[0], threshold, thr_k,
                                          thr_w, thr_c, seed, _p(seed_dev), elem_offset, _stream()), "xd_sampler_step")
    _count()


@_op("edm_step(int stage, Tensor x_hat, Tensor? x_mid, Tensor? d_in, Tensor f, Tensor(a!)? d_out, Tensor(b!)? x_out, "
     "Tensor(c!)? den_out, Tensor(d!)? xin_out, float t_div, float h, float c_skip, float c_out, float c_in_next) -> ()")
def _edm_step(stage, x_hat, x_mid, d_in, f, d_out, x_out, den_out, xin_out, t_div, h, c_skip, c_out, c_in_next):
    """EDM Euler (stage 0) / Heun (stage 1) update on the fp64 state, or the preconditioned denoiser output alone (stage 2)."""
    _cuda(x_hat, x_mid, d_in, f, d_out, x_out, den_out, xin_out)
    for t in (x_hat, x_mid, d_in, d_out, x_out):
        assert t is None or (t.dtype == torch.float64 and t.is_contiguous() and t.numel() == x_hat.numel())
    for t in (f, den_out, xin_out):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.numel() == x_hat.numel())
    _lib.check(_lib.lib().xd_edm_step(stage, _p(x_hat), _p(x_mid), _p(d_in), _p(f), _p(d_out), _p(x_out), _p(den_out),
                                      _p(xin_out), float(t_div), float(h), float(c_skip), float(c_out), float(c_in_next),
                                      x_hat.numel(), _stream()), "xd_edm_step")
    _count()


@_op("edm_prepare(Tensor x, Tensor? z, float c_noise, Tensor(a!)? x_hat, float c_in, Tensor(b!)? xin) -> ()")
def _edm_prepare(x, z, c_noise, x_hat, c_in, xin):
    _cuda(x, z, x_hat, xin)
    assert x.dtype == torch.float64 and x.is_contiguous()
    assert z is None or (z.dtype == torch.float64 and z.is_contiguous() and x_hat is not None and x_hat.dtype == torch.float64)
    assert xin is None or (xin.dtype == torch.float32 and xin.is_contiguous())
    _lib.check(_lib.lib().xd_edm_prepare(_p(x), _p(z), float(c_noise), _p(x_hat), float(c_in), _p(xin), x.numel(),
                                         _stream()), "xd_edm_prepare")
    _count()


@_op("schedule_advance(Tensor(a!) idx_dev, int set_to, Tensor? tab_i64, Tensor? tab_a, Tensor? tab_b, "
     "Tensor(b!)? out_i64, Tensor(c!)? out_a, Tensor(d!)? out_b, int B) -> ()")
def _schedule_advance(idx_dev, set_to, tab_i64, tab_a, tab_b, out_i64, out_a, out_b, B):
    _cuda(idx_dev, tab_i64, tab_a, tab_b, out_i64, out_a, out_b)
    _lib.check(_lib.lib().xd_schedule_advance(_p(idx_dev), set_to, _p(tab_i64), _p(tab_a), _p(tab_b), _p(out_i64),
                                              _p(out_a), _p(out_b), B, _stream()), "xd_schedule_advance")
    _count()


@_op("unnormalize(Tensor x, Tensor(a!) out) -> ()")
def _unnormalize(x, out):
    _cuda(x, out)
    assert x.is_contiguous() and out.is_contiguous()
    _lib.check(_lib.lib().xd_unnormalize(_p(x), _p(out), x.numel(), _stream()), "xd_unnormalize")
    _count()


# ------------------------------------------------------------------------------------ registration
_library = torch.library.Library("xdb200", "DEF")
for _schema, _fn in _defs:
    _library.define(_schema)
    _library.impl(_schema.split("(", 1)[0], _fn, "CUDA")
_ops = torch.ops.xdb200


# ------------------------------------------------------------------------------------ functional helpers
def linear(a, w, bias=None, act=ACT_NONE, out_dtype=torch.bfloat16, gate=None, gate_rows=1, residual=None,
           out=None, a2=None, force_bn=0, qstats=False):
    """out = epilogue(a @ w.T): a bf16 [M,K], w bf16 [N,K(+K2)].  ``qstats``: a GroupNorm consumes ``out`` (quad_stats)."""
    if out is None:
        out = torch.empty((a.shape[0], w.shape[0]), device=a.device, dtype=out_dtype)
    slot = _qs_slot(out) if qstats and act == ACT_NONE and a.shape[1] % 64 == 0 and (a2 is None or a2.shape[1] % 64 == 0) else None
    if slot is not None:
        emitted = _ops.gemm_qs(a, a2, w, bias, act, gate, gate_rows, residual, out, force_bn, slot[1])
        _qs_written(out, slot[2] if emitted else None)
        return out
    _ops.gemm(a, a2, w, bias, act, gate, gate_rows, residual, out, force_bn)
    return out


# Off by default: measured slower than layernorm_modulate + linear (30.1 vs 24.0 us on the qkv shape, see
# profiles/README.md); the kernel is kept as a tested opt-in.
LN_GEMM_FUSED = os.environ.get("XDB200_LN_FUSED", "0") == "1"


def ln_linear(x, shift, scale, rows_per_mod, w, bias=None, act=ACT_NONE, eps=1e-6, out=None):
    """act(LNmod(x) @ w.T + bias) -> bf16: x fp32 [M, D], shift/scale fp32 [M / rows_per_mod, D].  One kernel when the
    row fits a shared-memory A panel (D = 128/256/384, N % 192 == 0), else layernorm_modulate + linear."""
    M, D = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
    if (LN_GEMM_FUSED and MATMUL_BACKEND == "tc" and D in (128, 256, 384) and N % 192 == 0
            and act in (ACT_NONE, ACT_GELU) and x.stride(0) % 4 == 0):
        _ops.ln_gemm(x, shift, scale, rows_per_mod, eps, w, bias, act, out)
        return out
    a = layernorm_modulate(x, shift, scale, rows_per_mod, eps)
    return linear(a, w, bias, act=act, out=out)


def conv3x3(x, wp, bias=None, act=ACT_NONE, residual=None, xs=None, out=None, force_bn=0, qstats=False):
    if out is None:
        out = torch.empty(x.shape[:3] + (wp.shape[0],), device=x.device, dtype=torch.bfloat16)
    slot = _qs_slot(out) if qstats and act == ACT_NONE else None
    if slot is not None:
        emitted = _ops.conv3x3_qs(x, xs, wp, bias, act, residual, out, force_bn, slot[1])
        _qs_written(out, slot[2] if emitted else None)
        return out
    _ops.conv3x3(x, xs, wp, bias, act, residual, out, force_bn)
    return out


def upsample2x(x, out):
    """Nearest x2 over (H, W) of an NHWC view.  Inside ``quad_stats()``: an upsampled tensor has exactly four times the sums
    of its source, so when the source's producers emitted their quad statistics the table of ``out`` is filled by replicating
    every source block four times (only per-sample totals are ever read) and the GroupNorm over ``out`` needs no statistics
    pass either.  Images must be whole 32-row blocks."""
    _ops.upsample2x(x, out)
    if _qs_book is None or (x.shape[1] * x.shape[2]) % 32:
        return out
    src = _qs_lookup(x)
    slot = _qs_slot(out) if src is not None else None
    if slot is not None:
        e, view, rng = slot
        view.unflatten(0, (src.shape[0], 4)).copy_(src.unsqueeze(1).expand(-1, 4, -1, -1))
        _qs_written(out, rng)
    return out


def conv3x3_groupnorm(x, wp, bias, samples, gamma, beta, scale_shift=None, eps=1e-5, silu=True):
    """GroupNorm32(conv3x3(x) + bias) [modulated] [SiLU] -> bf16 NHWC; x bf16 NHWC view, ``samples`` statistics units."""
    nimg, H, W, _ = x.shape
    co, M = wp.shape[0], nimg * H * W
    out = torch.empty((nimg, H, W, co), device=x.device, dtype=torch.bfloat16)
    tmp = torch.empty((M, co), device=x.device, dtype=torch.bfloat16)
    need = max(M // 32 * (co // 2), samples * 64 * _lib.lib().xd_groupnorm_slabs(samples, M // samples, co))
    scratch = torch.empty(max(need, 1), device=x.device, dtype=torch.float32)
    _ops.conv3x3_groupnorm(x, wp, bias, samples, gamma, beta, scale_shift, 1, eps, int(silu), tmp, scratch, out)
    return out


def groupnorm(x, gamma, beta, scale_shift=None, ss_div=1, eps=1e-5, silu=False, out=None, inner=1):
    """x [nsamples, P, C] bf16 view with uniformly strided rows (x.stride(0) == P * x.stride(1));
    ``inner`` > 1 selects the interleaved sample layout described in include/xdb200.h."""
    ns, P, C = x.shape
    assert x.stride(2) == 1 and x.stride(0) == P * x.stride(1)
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    assert out.stride(2) == 1 and out.stride(0) == P * out.stride(1)
    x2 = x.as_strided((ns * P, C), (x.stride(1), 1))
    o2 = out.as_strided((ns * P, C), (out.stride(1), 1))
    q = _qs_lookup(x) if inner == 1 and C % 128 == 0 and P % 32 == 0 else None
    if q is not None:                      # statistics came with the input: one streaming pass
        _ops.groupnorm_quads(x2, q, gamma, beta, scale_shift, ss_div, eps, int(silu), ns, o2)
        return out
    stats = torch.empty(ns * 64 * _lib.lib().xd_groupnorm_slabs(ns, P, C), device=x.device, dtype=torch.float32)
    _ops.groupnorm(x2, gamma, beta, scale_shift, ss_div, eps, int(silu), inner, ns, 0, stats, o2)
    return out


def layernorm_modulate(x, shift, scale, rows_per_mod, eps=1e-6):
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _ops.layernorm_modulate(x, shift, scale, rows_per_mod, eps, out)
    return out


def attention(q, k, v, scale, out=None, relk=None, scramble=False, o_cs=0, hpg=0, o_gs=0):
    if out is None:
        B, H, Tq, D = q.shape
        out = torch.empty((B, Tq, H, D), device=q.device, dtype=torch.bfloat16).permute(0, 2, 1, 3)
    _ops.attention(q, k, v, out, float(scale), relk, int(scramble), o_cs, hpg, o_gs)
    return out
