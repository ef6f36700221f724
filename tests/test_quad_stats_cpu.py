"""Host-side bookkeeping of the GroupNorm-statistics-from-the-producer path (xdiffusion_b200/ops.py: quad_stats): which
tensors get a statistics table, how the channel slices of a concat buffer are merged, when an entry is dropped.  Pure index
logic on CPU tensors -- no kernel runs here (the kernels are covered by tests/test_kernels_gpu.py)."""
import torch

from xdiffusion_b200 import ops


def _buf(n=2, H=8, W=8, C=384):
    return torch.zeros(n, H, W, C, dtype=torch.bfloat16)


def test_geometry_of_views():
    b = _buf()
    assert ops._qs_geometry(b) == (2 * 64, 384, 0, 384)
    assert ops._qs_geometry(b[..., 256:]) == (128, 384, 256, 128)
    left = b[..., :256]
    assert ops._qs_geometry(left.as_strided((128, 256), (384, 1), left.storage_offset())) == (128, 384, 0, 256)
    assert ops._qs_geometry(b.view(2, 64, 384)) == (128, 384, 0, 384)
    assert ops._qs_geometry(b[1:]) is None                       # does not start in row 0 of its storage
    assert ops._qs_geometry(b[..., 2:130]) is None               # channel offset not a multiple of 4
    assert ops._qs_geometry(_buf(1, 4, 4, 128)) is None          # 16 rows: not whole 32-row blocks
    assert ops._qs_geometry(b.float()) is None                   # bf16 activations only


def test_book_is_scoped_and_merges_slices_of_a_concat_buffer():
    b = _buf()
    assert ops._qs_slot(b) is None                               # no scope: nothing is recorded
    with ops.quad_stats():
        if ops._qs_book is None:                                 # XDB200_GN_QSTATS=0 / CUDA-core backend
            return
        e, view, rng = ops._qs_slot(b[..., :256])
        assert view.shape == (128 // 32, 256 // 4, 2) and rng == (0, 256) and e["table"].shape == (4, 96, 2)
        ops._qs_written(b[..., :256], rng)
        full = b.view(2, 64, 384)
        assert ops._qs_lookup(full) is None                      # right slice still missing
        assert ops._qs_lookup(b[..., :256]) is not None
        e2, view2, rng2 = ops._qs_slot(b[..., 256:])
        assert e2 is e and rng2 == (256, 384) and view2.data_ptr() == e["table"][:, 64:].data_ptr()
        ops._qs_written(b[..., 256:], rng2)
        assert ops._qs_lookup(full).shape == (4, 96, 2)
        ops._qs_written(b[..., 128:192])                         # someone overwrote part of the left slice: its statistics go
        assert ops._qs_lookup(full) is None and ops._qs_lookup(b[..., 256:]) is not None
        assert ops._qs_lookup(b[..., :256]) is None
    assert ops._qs_book is None


def test_op_wrapper_invalidates_outputs_of_other_ops():
    """Every op of the module reports the tensors it mutates ((a!) in its schema): an entry whose storage is written by an op
    without statistics is dropped, so a stale table can never be paired with new data."""
    names = {schema.split("(", 1)[0]: schema for schema, _ in ops._defs}
    assert "Tensor(a!) out" in names["avgpool2x2"] and "Tensor(a!) out" in names["upsample2x"]
    mutated = [i for i, a in enumerate(ops._split_args(names["conv3x3"].split("(", 1)[1].rsplit(")", 1)[0])) if "!" in a]
    assert mutated == [6]
    with ops.quad_stats():
        if ops._qs_book is None:
            return
        b = _buf()
        _, _, rng = ops._qs_slot(b)
        ops._qs_written(b, rng)
        assert ops._qs_lookup(b.view(2, 64, 384)) is not None
        ops._qs_written(b)                                       # what the wrapper does after e.g. avgpool2x2(x, out=b)
        assert ops._qs_lookup(b.view(2, 64, 384)) is None
