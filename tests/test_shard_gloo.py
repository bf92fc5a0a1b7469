"""The N>1 host logic on CPU: utterance sharding and the result gather, world_size 2, gloo."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gvn.shard import shard_bounds, shard_list, gather_rows, length_sorted_shard, wave_batches, chain_tiles


def test_shard_bounds_match_array_split():
    # scripts/evaluate_M1.py:203 uses np.array_split(file_paths, nb_devices)
    for n in (0, 1, 7, 64, 513, 1000):
        for world in (1, 2, 3, 8):
            ref = np.array_split(np.arange(n), world)
            for r in range(world):
                a, b = shard_bounds(n, world, r)
                assert list(range(a, b)) == list(ref[r])
    assert shard_list(list("abcde"), 2, 0) == ["a", "b", "c"] and shard_list(list("abcde"), 2, 1) == ["d", "e"]


def test_wave_batches_fill_whole_waves():
    """Batches cut for the chain kernel (one 128-frame tile per SM and wave): consecutive, complete, never above the
    budget of `waves` x 148 tiles unless a single utterance is, and fuller than fixed batches of 64."""
    lens = np.random.RandomState(5).randint(537, 749, size=1024)
    order = length_sorted_shard(lens, 1, 0)
    n = [int(lens[i]) for i in order]
    for waves in (1, 2, 4):
        cuts = wave_batches(n, 148, waves)
        assert cuts[0][0] == 0 and cuts[-1][1] == len(n) and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        tiles = [chain_tiles(n[a:b]) for a, b in cuts]
        assert max(tiles) <= waves * 148 and min(tiles[:-1]) > waves * 148 - 8       # full up to one utterance (<= 6 tiles)
        total = sum(-(-t // 148) for t in tiles)
        fixed = sum(-(-chain_tiles(n[k:k + 64]) // 148) for k in range(0, len(n), 64))
        assert total < fixed                                                          # 36 - 37 waves against 45
    assert wave_batches([5000], 148, 1, align=32) == [(0, 1)]                         # a single over-long utterance still goes
    assert wave_batches([], 148, 2) == []
    assert wave_batches([100] * 10, 148, 2, max_batch=4) == [(0, 4), (4, 8), (8, 10)]
    assert chain_tiles([251] * 64) == 128 and chain_tiles([1876] * 8) == 118


def test_length_sorted_snake_shards_are_balanced():
    """C5 / real file lists: utterances of 537-748 frames.  Sorted by length and dealt in a snake, the ranks' frame totals
    differ by less than one utterance; the reference's contiguous split of the same sorted list is off by tens of percent."""
    lens = np.random.RandomState(5).randint(537, 749, size=1000)
    for world in (2, 3, 8):
        shards = [length_sorted_shard(lens, world, r) for r in range(world)]
        assert sorted(i for s in shards for i in s) == list(range(1000))                 # a partition
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
        tot = [int(lens[s].sum()) for s in shards]
        assert max(tot) - min(tot) <= 748
        for s in shards:                                                                  # longest first inside a shard
            assert all(lens[a] >= lens[b] for a, b in zip(s, s[1:]))
        order = sorted(range(1000), key=lambda i: -int(lens[i]))
        contiguous = [int(lens[shard_list(order, world, r)].sum()) for r in range(world)]
        assert max(contiguous) - min(contiguous) > 20 * (max(tot) - min(tot))


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_bounds(n_total, world, rank)
    ids = torch.arange(a, b, dtype=torch.float64)
    rows = torch.stack([ids, ids * 10 + rank], 1)
    table = gather_rows(rows, n_total)
    q.put((rank, table.numpy()))
    dist.destroy_process_group()


def test_gather_rows_world2_gloo():
    world, n_total = 2, 7                       # ragged: shards of 4 and 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        t = res[r]
        assert t.shape == (n_total, 2)
        np.testing.assert_array_equal(t[:, 0], np.arange(n_total))
        owner = np.array([0, 0, 0, 0, 1, 1, 1])
        np.testing.assert_array_equal(t[:, 1], np.arange(n_total) * 10 + owner)
