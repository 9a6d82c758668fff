"""Multi-GPU host logic on CPU: world_size-2 gloo processes run the same partitioning code bench.py and
./fora --gpus use (queries sharded, graph replicated, no data-path collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fora_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    queries = np.random.default_rng(43).integers(0, 10 ** 6, 1000).astype(np.int32)
    # per-step shards: disjoint across ranks, same on every run
    ids, keys = shard.step_query_ids(queries, step=3, batch=50, rank=rank, world=world)
    gathered = [None] * world
    dist.all_gather_object(gathered, keys.tolist())
    flat = sum(gathered, [])
    assert len(set(flat)) == len(flat) == 50 * world
    assert (queries[keys] == ids).all()
    # contiguous blocks cover the list exactly once
    lo, hi = shard.query_block(1000, rank, world)
    blocks = [None] * world
    dist.all_gather_object(blocks, (lo, hi))
    assert blocks[0][0] == 0 and blocks[-1][1] == 1000 and all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    # timing is the max over ranks
    t = shard.max_over_ranks(1.0 + rank)
    assert t == float(world)
    # sharded index build: each rank fills its source range, rank 0 concatenates -> identical to unsharded
    rng = np.random.default_rng(1)
    counts = rng.integers(0, 50, 5000).astype(np.uint64)
    offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.uint64)
    cuts = shard.balanced_source_ranges(offsets, counts, world)
    a, b = cuts[rank], cuts[rank + 1]
    mine = np.concatenate([np.full(int(counts[v]), v, np.int32) for v in range(a, b)]) if b > a else np.empty(0, np.int32)
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    full = np.concatenate(parts)
    assert np.array_equal(full, np.repeat(np.arange(5000, dtype=np.int32), counts.astype(np.int64)))
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 60  # balanced by walks
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")


def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


def test_balanced_ranges_edge_cases():
    counts = np.array([0, 0, 10, 0, 1000, 0, 3], np.uint64)
    offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.uint64)
    for w in (1, 2, 3, 8):
        cuts = shard.balanced_source_ranges(offsets, counts, w)
        assert cuts[0] == 0 and cuts[-1] == 7 and len(cuts) == w + 1 and all(cuts[i] <= cuts[i + 1] for i in range(w))
    ids, keys = shard.step_query_ids(np.arange(10), step=7, batch=4, rank=1, world=2)
    assert keys.tolist() == [(7 * 8 + 4 + j) % 10 for j in range(4)]


def test_numa_binding_helper_is_safe_without_a_gpu():
    from fora_b200 import shard
    assert shard._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert shard._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    node = shard.bind_to_gpu_numa_node(0)   # no NVML device here: must be a no-op
    if node is None:
        assert os.sched_getaffinity(0) == before
    else:
        assert os.sched_getaffinity(0) <= before
        os.sched_setaffinity(0, before)
