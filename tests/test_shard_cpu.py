"""Multi-rank path on CPU: world_size-2 gloo processes shard streams the way bench.py does under torchrun."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_streams, out_dir):
    sys.path.insert(0, ROOT)
    import plviwo_b200  # noqa: F401
    from plviwo_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.streams_of_rank(n_streams, rank, world)
    # every rank announces its streams; the union must be a partition of range(n_streams)
    owned = torch.full((n_streams,), -1, dtype=torch.int64)
    for s in mine:
        owned[s] = rank
    gathered = [torch.empty_like(owned) for _ in range(world)]
    dist.all_gather(gathered, owned)
    claims = torch.stack(gathered)
    assert ((claims >= 0).sum(0) == 1).all(), "a stream is owned by exactly one rank"
    for s in range(n_streams):
        assert int(claims[:, s].max()) == shard.rank_of_stream(s, world)
    # per-rank work: 100 frames per stream, rank r takes (r + 1) ms per frame -> the slowest rank sets the time
    frames = 100 * len(mine)
    ms = float(frames * (rank + 1))
    dist.barrier()
    fps = shard.distributed_throughput(dist, torch, frames, ms)
    all_frames = [100 * len(shard.streams_of_rank(n_streams, r, world)) for r in range(world)]
    all_ms = [float(f * (r + 1)) for r, f in enumerate(all_frames)]
    assert abs(fps - shard.aggregate(all_frames, all_ms)) < 1e-9
    if rank == 0:
        open(os.path.join(out_dir, "ok"), "w").write("%f" % fps)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [2, 7, 64])
def test_stream_sharding_world2_gloo(tmp_path, n_streams):
    port = 29500 + (os.getpid() + n_streams) % 2000
    mp.spawn(_worker, args=(2, port, n_streams, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")


def test_streams_of_rank_properties():
    sys.path.insert(0, ROOT)
    import plviwo_b200  # noqa: F401
    from plviwo_b200 import shard
    for world in (1, 2, 4, 8):
        for n in (0, 1, 8, 63, 64):
            parts = [shard.streams_of_rank(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert shard.seed_of_stream(3) == 1003
    with pytest.raises(ValueError):
        shard.streams_of_rank(4, 2, 2)
