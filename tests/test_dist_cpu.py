"""Multi-rank host logic on CPU: world_size-2 gloo, a stub model whose sample() is a pure per-row
function -- sharded sampling must reproduce the single-process result exactly, ragged splits included."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xdiffusion_b200.dist import gather_rows, sample_sharded, shard_bounds, shard_context


class StubModel:
    def sample(self, context=None, num_samples=1, initial_noise=None, noise=None, **kw):
        x = initial_noise.clone()
        for i in reversed(range(noise.shape[0])):
            x = 0.9 * x + 0.1 * noise[i] + context["classes"].view(-1, 1).float()
        return x, []


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    x0, z = torch.randn(total, 5, generator=g), torch.randn(3, total, 5, generator=g)
    ctx = {"classes": torch.arange(total), "text_prompts": [str(i) for i in range(total)], "scalar": 3}
    out, _ = sample_sharded(StubModel(), ctx, total, initial_noise=x0, noise=z)
    ref, _ = StubModel().sample(ctx, total, x0, z)
    q.put((rank, bool(torch.equal(out, ref))))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_sharded_sampling_gloo_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_bounds_partition():
    for total in (1, 7, 8, 1024):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_shard_context_slices_batch_major_entries_only():
    c = shard_context({"classes": torch.arange(8), "table": torch.zeros(3, 2), "p": list("abcdefgh"), "k": 1}, 8, 2, 5)
    assert c["classes"].tolist() == [2, 3, 4] and c["table"].shape == (3, 2) and c["p"] == list("cde") and c["k"] == 1
