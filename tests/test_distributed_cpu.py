"""N > 1 host logic on CPU: world_size 2 over gloo (127.0.0.1).  Videos are sharded by rank with no data-path
collective; the gathered logits must equal the single-process result, including ragged and empty shards."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ewvit.distributed import gather_logits, score_videos_sharded, shard_counts, shard_range


def test_shard_range_is_a_balanced_partition():
    for n in (0, 1, 5, 64, 65):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = shard_counts(n, world)
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _fake_video(i, frames=3):
    g = torch.Generator().manual_seed(42 + i)
    return torch.randn(frames, 3, 8, 8, generator=g)


def _make(ids):
    return torch.stack([_fake_video(i) for i in ids])


def _score(x):                      # stands in for model(x, bs, 'dynamic')['logits']
    return x.mean(dim=(1, 2, 3, 4), keepdim=False).unsqueeze(1) * 10.0


def test_single_process_path():
    out = score_videos_sharded(_score, _make, 5, 2)
    ref = torch.cat([_score(_make([i])).reshape(-1) for i in range(5)])
    assert torch.allclose(out, ref)


def _worker(rank, world, port, num_videos, per_call, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = score_videos_sharded(_score, _make, num_videos, per_call)
        lo, hi = shard_range(num_videos, rank, world)
        bad = False
        try:
            gather_logits(torch.zeros(hi - lo + 1), num_videos)
        except ValueError:
            bad = True
        q.put((rank, out.tolist(), bad))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("num_videos,per_call", [(5, 2), (1, 4), (8, 3)])
def test_world_size_2_gloo_matches_single_process(num_videos, per_call):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_videos, per_call, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = torch.cat([_score(_make([i])).reshape(-1) for i in range(num_videos)])
    for rank, out, bad in results:
        assert torch.allclose(torch.tensor(out), ref), f"rank {rank}"
        assert bad, "a block of the wrong length must be rejected"
