"""Multi-rank host logic on CPU: world_size = 2 over gloo (the GPU path uses NCCL for the same call)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dbsgym_b200.sharding import gather_episode_stats, shard_bounds, shard_params


def test_shard_bounds_partition_exactly():
    for n, w in ((16384, 8), (4096, 3), (7, 4), (5, 8)):
        spans = [shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(16384, 3, 8) == (6144, 8192)          # BASELINE config 4: 2048 envs per GPU
    part, lo = shard_params(list(range(10)), 1, 3)
    assert part == [4, 5, 6] and lo == 4
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(n_global, rank, world)
    # per-env episode statistics of this shard: [global env id, episode return, beta power]
    ids = np.arange(lo, hi, dtype=np.float64)
    local = np.stack([ids, -35.0 - ids, 0.01 * (ids + 1)], axis=1)
    table = gather_episode_stats(local, n_global, rank, world)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), table)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_episode_stats_world_size_2(tmp_path):
    n_global = 7                                               # uneven shards: 4 + 3
    mp.spawn(_worker, args=(2, _free_port(), n_global, str(tmp_path)), nprocs=2, join=True)
    t0, t1 = (np.load(tmp_path / f"r{r}.npy") for r in range(2))
    assert np.array_equal(t0, t1)                              # identical on every rank
    ids = np.arange(n_global, dtype=np.float64)
    assert np.array_equal(t0, np.stack([ids, -35.0 - ids, 0.01 * (ids + 1)], axis=1))


def test_gather_without_process_group_is_identity():
    local = np.arange(12, dtype=np.float64).reshape(4, 3)
    assert np.array_equal(gather_episode_stats(local, 4, 0, 1), local)
